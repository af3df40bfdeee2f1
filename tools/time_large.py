"""Time the HBM-streaming path on a large grid (BASELINE config 5 shape).  Usage: python tools/time_large.py N B T"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import math, torch
import wavetorch_b200 as wt
from wavetorch_b200 import _lib
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
T = int(sys.argv[3]) if len(sys.argv) > 3 else 20
ii = torch.arange(N, dtype=torch.float32)[:, None]; jj = torch.arange(N, dtype=torch.float32)[None, :]
rho = 0.5 + 0.5 * torch.sin(2 * math.pi * ii / 97) * torch.cos(2 * math.pi * jj / 61)
geom = wt.WaveGeometryFreeForm((N, N), 1.4283556979968262, 1.0, 0.5, abs_N=20, abs_sig=3.0, abs_p=4.0, rho=rho)
probes = [wt.WaveIntensityProbe(N - 60, N // 2 + 20 * k) for k in (-1, 0, 1)]
m = wt.WaveRNN(wt.WaveCell(1.0, geom), [wt.WaveSource(60, N // 2)], probes).to("cuda")
x = torch.randn(B, T, device="cuda") * 0.1
cells = B * T * N * N
def tm(fn, n=2):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
def fwd():
    with torch.no_grad(): m(x)
def full():
    o = m(x); o.sum().backward(); m.zero_grad(set_to_none=True)
tf = tm(fwd)
print(f"{N}x{N} B={B} T={T}: fwd {tf/T:.3f} ms/step {cells/tf/1e6:.1f} Gcell/s = {12*cells/tf/1e6:.0f} GB/s algorithmic (12 B/cell)", flush=True)
tb = tm(full)
print(f"   fwd+bwd {tb/T:.3f} ms/step {cells/tb/1e6:.1f} Gcell/s = {32*cells/tb/1e6:.0f} GB/s algorithmic (32 B/cell)  mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB", flush=True)
