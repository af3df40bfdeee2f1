"""Per-step time of the on-chip forward (with tape) and adjoint kernels for C=1 / C=2 with the same rows per CTA."""
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
import os
for rows_per_cta, R in ((75, 5), (40, 4), (20, 2))[:int(os.environ.get("NCFG", 3))]:
    for C in (1, 2):
        for nprobe in (1, 0):
            Nx, Ny = rows_per_cta * C, 100
            B = 128 // C
            geom = wt.WaveGeometryFreeForm((Nx, Ny), 1.0, 1.0, 0.6, abs_N=2, abs_sig=3.0, abs_p=3.0, beta=10.0, rho="half")
            probes = [wt.WaveIntensityProbe(Nx - 4, 50)] if nprobe else [wt.WaveIntensityProbe(Nx - 4, 50)]
            m = wt.WaveRNN(wt.WaveCell(0.6, geom), [wt.WaveSource(4, 50)], probes).to("cuda")
            m.cluster, m.rows_per_thread, m.plan_flags = C, R, _lib.WT_F_FORCE_RESIDENT
            x = torch.randn(B, T, device="cuda") * 0.1
            out = m(x); g = torch.ones_like(out)
            def fwdt(): return m(x)
            def full():
                o = m(x); o.backward(g); m.zero_grad(set_to_none=True)
            tf = tm(fwdt); tb = tm(full)
            print(f"rows/CTA={rows_per_cta:3d} R={R} C={C}: fwd+tape {tf*1e3/T:6.2f} us/step, adjoint {(tb-tf)*1e3/T:6.2f} us/step", flush=True)
            break
