"""Short run for ncu on the temporally blocked kernels: 4096x4096 window (env PN), B=8, T=8, one forward with tape and one adjoint."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wavetorch_b200 as wt
N = int(os.environ.get("PN", 4096)); B = int(os.environ.get("PB", 8)); T = int(os.environ.get("PT", 8))
ii = torch.arange(N, dtype=torch.float32)[:, None]; jj = torch.arange(N, dtype=torch.float32)[None, :]
rho = 0.5 + 0.5 * torch.sin(2 * math.pi * ii / 97) * torch.cos(2 * math.pi * jj / 61)
geom = wt.WaveGeometryFreeForm((N, N), 1.4283556979968262, 1.0, 0.5, abs_N=20, abs_sig=3.0, abs_p=4.0, rho=rho)
probes = [wt.WaveIntensityProbe(N - 60, N // 2 + 20 * k) for k in (-1, 0, 1)]
m = wt.WaveRNN(wt.WaveCell(1.0, geom), [wt.WaveSource(60, N // 2)], probes).to("cuda")
x = torch.randn(B, T, device="cuda") * 0.1
for _ in range(2):
    m(x).sum().backward()
    m.zero_grad(set_to_none=True)
torch.cuda.synchronize()
print("done")
