"""What the ghost-row exchange costs in situ: a [rows x 4096] grid split over the ranks, forward and forward+backward,
with and without the exchange kernel (WT_SLAB_SKIP=1: wrong results, timing only).
   torchrun --nproc-per-node N tools/dd_probe.py [rows_total] [B] [T]"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import wavetorch_b200 as wt
from wavetorch_b200.domain import DomainDecomposedWaveRNN
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
T = int(sys.argv[3]) if len(sys.argv) > 3 else 128
Ny = 4096
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
ii = torch.arange(rows, dtype=torch.float32)[:, None]; jj = torch.arange(Ny, dtype=torch.float32)[None, :]
rho = 0.5 + 0.5 * torch.sin(2 * math.pi * ii / 97) * torch.cos(2 * math.pi * jj / 61)
geom = wt.WaveGeometryFreeForm((rows, Ny), 1.4283556979968262, 1.0, 0.5, abs_N=20, abs_sig=3.0, abs_p=4.0, rho=rho)
probes = [wt.WaveIntensityProbe(rows - 60, Ny // 2 + 20 * k) for k in (-1, 0, 1)]
m = wt.WaveRNN(wt.WaveCell(1.0, geom), [wt.WaveSource(60, Ny // 2)], probes).to(dev)
dd = DomainDecomposedWaveRNN(m, halo=16, checkpoint_every=64)
torch.manual_seed(0)
x = (0.1 * torch.randn(B, T)).to(dev); w = torch.randn(B, T, 3).to(dev)
def tm(fn, n=3):
    fn(); fn(); torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()
def fwd():
    with torch.no_grad(): dd(x)
def full():
    (dd(x) * w).sum().backward(); m.zero_grad(set_to_none=True)
tf, tb = tm(fwd), tm(full)
if rank == 0:
    cells = B * T * rows * Ny
    print(f"world={world} grid {rows}x{Ny} B={B} T={T} skip={os.environ.get('WT_SLAB_SKIP','0')}: fwd {tf:.2f} ms ({cells/tf/1e6:.0f} Gcell/s)  fwd+bwd {tb:.2f} ms ({cells/tb/1e6:.0f} Gcell/s)", flush=True)
dist.barrier(); torch.cuda.synchronize(); os._exit(0)
