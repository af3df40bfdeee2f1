"""Short run for ncu: the temporally blocked forward on a thin [rows x 4096] grid (a slab of the 8-GPU decomposition) and on the full grid."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wavetorch_b200 as wt
for rows in (int(os.environ.get("PROWS", 544)), 4096):
    Ny, B, T = 4096, 8, 16
    ii = torch.arange(rows, dtype=torch.float32)[:, None]; jj = torch.arange(Ny, dtype=torch.float32)[None, :]
    rho = 0.5 + 0.5 * torch.sin(2 * math.pi * ii / 97) * torch.cos(2 * math.pi * jj / 61)
    geom = wt.WaveGeometryFreeForm((rows, Ny), 1.4283556979968262, 1.0, 0.5, abs_N=20, abs_sig=3.0, abs_p=4.0, rho=rho)
    probes = [wt.WaveIntensityProbe(rows - 60, Ny // 2 + 20 * k) for k in (-1, 0, 1)]
    m = wt.WaveRNN(wt.WaveCell(1.0, geom), [wt.WaveSource(60, Ny // 2)], probes).to("cuda")
    x = torch.randn(B, T, device="cuda") * 0.1
    for _ in range(2):
        with torch.no_grad():
            m(x)
    torch.cuda.synchronize()
print("done")
