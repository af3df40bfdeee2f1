"""Per-step latency of the on-chip forward kernel for tiny slabs: C=1 vs C=2 vs C=4 with the same rows per CTA."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wavetorch_b200 as wt
from wavetorch_b200 import _lib
def tm(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
T = 2000
for rows_per_cta, R in ((20, 4), (20, 2), (40, 4), (75, 5)):
    for C in (1, 2, 4):
        Nx, Ny = rows_per_cta * C, 100
        if Nx < 12: continue
        B = 128 // C
        N = 2
        geom = wt.WaveGeometryFreeForm((Nx, Ny), 1.0, 1.0, 0.6, abs_N=N, abs_sig=3.0, abs_p=3.0, beta=10.0, rho="half")
        m = wt.WaveRNN(wt.WaveCell(0.6, geom), [wt.WaveSource(4, 50)], [wt.WaveIntensityProbe(Nx - 4, 50)]).to("cuda")
        m.cluster, m.rows_per_thread, m.plan_flags = C, R, _lib.WT_F_FORCE_RESIDENT
        x = torch.randn(B, T, device="cuda") * 0.1
        try:
            def fwd():
                with torch.no_grad(): m(x)
            t = tm(fwd)
            print(f"rows/CTA={rows_per_cta:3d} R={R} C={C} B={B:3d}: {t*1e3/T:7.1f} ns/step = {t*1e3/T*1.965:7.0f} cycles/step  ({rows_per_cta*Ny} cells/CTA)", flush=True)
        except RuntimeError as e:
            print(f"rows/CTA={rows_per_cta} R={R} C={C}: infeasible {str(e)[-50:]}")
