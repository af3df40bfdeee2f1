"""How much do the warps with special duties cost?  Per-step time of the on-chip kernels with and without a source, C = 1 / 2."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wavetorch_b200 as wt
from wavetorch_b200 import _lib
def tm(fn, n=4):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
T = 1000
for C in (1, 2):
    for nsrc in (1, 0):
        for prow in (4, 70):
            Nx, Ny = 75 * C, 100
            B = 64
            geom = wt.WaveGeometryFreeForm((Nx, Ny), 1.0, 1.0, 0.6, abs_N=2, abs_sig=3.0, abs_p=3.0, beta=10.0, rho="half")
            m = wt.WaveRNN(wt.WaveCell(0.6, geom), [wt.WaveSource(40, 50)] if nsrc else [], [wt.WaveIntensityProbe(prow, 50)]).to("cuda")
            m.cluster, m.rows_per_thread, m.plan_flags = C, 5, _lib.WT_F_FORCE_RESIDENT
            x = torch.randn(B, T, device="cuda") * 0.1
            try:
                out = m(x)
            except Exception as e:
                print("C=%d nsrc=%d:" % (C, nsrc), e); continue
            g = torch.ones_like(out)
            def fwd():
                with torch.no_grad(): m(x)
            def fwdt(): return m(x)
            def full():
                o = m(x); o.backward(g); m.zero_grad(set_to_none=True)
            t0 = tm(fwd); tf = tm(fwdt); tb = tm(full)
            print(f"C={C} sources={nsrc} probe row {prow}: fwd {t0*1e3/T:5.2f}  fwd+tape {tf*1e3/T:5.2f}  adjoint {(tb-tf)*1e3/T:5.2f} us/step", flush=True)
