import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wavetorch_b200 as wt
from wavetorch_b200 import _lib
Nx, Ny, B, T = 20, 100, 128, 500
geom = wt.WaveGeometryFreeForm((Nx, Ny), 1.0, 1.0, 0.6, abs_N=2, abs_sig=3.0, abs_p=3.0, beta=10.0, rho="half")
m = wt.WaveRNN(wt.WaveCell(0.6, geom), [wt.WaveSource(4, 50)], [wt.WaveIntensityProbe(Nx - 4, 50)]).to("cuda")
m.cluster, m.rows_per_thread, m.plan_flags = 1, int(os.environ.get("PR", 2)), _lib.WT_F_FORCE_RESIDENT
x = torch.randn(B, T, device="cuda") * 0.1
for _ in range(3):
    with torch.no_grad(): m(x)
torch.cuda.synchronize(); print("done")
