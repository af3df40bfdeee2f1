"""Run under torchrun: checks the NCCL domain decomposition against the single-GPU streaming path and times it.
   torchrun --nproc-per-node 2 tools/dd_check.py [N] [B] [T] [halo]"""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import wavetorch_b200 as wt
from wavetorch_b200 import _lib
from wavetorch_b200.domain import DomainDecomposedWaveRNN
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
T = int(sys.argv[3]) if len(sys.argv) > 3 else 64
halo = int(sys.argv[4]) if len(sys.argv) > 4 else 16
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
def build():
    ii = torch.arange(N, dtype=torch.float32)[:, None]; jj = torch.arange(N, dtype=torch.float32)[None, :]
    rho = 0.5 + 0.5 * torch.sin(2 * math.pi * ii / 97) * torch.cos(2 * math.pi * jj / 61)
    geom = wt.WaveGeometryFreeForm((N, N), 1.4283556979968262, 1.0, 0.5, abs_N=20, abs_sig=3.0, abs_p=4.0, rho=rho)
    probes = [wt.WaveIntensityProbe(N // 2 + 10, N // 2 + 6 * k) for k in (-1, 0, 1)]
    return wt.WaveRNN(wt.WaveCell(1.0, geom), [wt.WaveSource(N // 2 - 10, N // 2)], probes).to(dev)
torch.manual_seed(0)
x = (0.1 * torch.randn(B, T)).to(dev)
w = torch.randn(B, T, 3).to(dev)
def tm(fn, n=2):
    fn(); torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()
ref = build(); ref.plan_flags = _lib.WT_F_FORCE_STREAM; ref.checkpoint_every = halo
out_ref = ref(x); (out_ref * w).sum().backward()
m = build(); dd = DomainDecomposedWaveRNN(m, halo=halo)
out = dd(x); (out * w).sum().backward()
rel = lambda a, b: ((a - b).norm() / b.norm()).item()
if rank == 0:
    print(f"world={world} {N}x{N} B={B} T={T} halo={halo}: out rel {rel(out, out_ref):.2e}  rho.grad rel {rel(m.cell.geom.rho.grad, ref.cell.geom.rho.grad):.2e}", flush=True)
cells = B * T * N * N
def f1():
    with torch.no_grad(): ref(x)
def fd():
    with torch.no_grad(): dd(x)
def b1():
    o = ref(x); (o * w).sum().backward(); ref.zero_grad(set_to_none=True)
def bd():
    o = dd(x); (o * w).sum().backward(); m.zero_grad(set_to_none=True)
t1, td = tm(f1), tm(fd)
tb1, tbd = tm(b1), tm(bd)
if rank == 0:
    print(f"  fwd: 1 GPU {t1:.2f} ms ({cells/t1/1e6:.1f} Gcell/s) | {world} GPUs decomposed {td:.2f} ms ({cells/td/1e6:.1f} Gcell/s) speed-up {t1/td:.2f}", flush=True)
    print(f"  fwd+bwd (checkpointed every {halo}): 1 GPU {tb1:.2f} ms ({cells/tb1/1e6:.1f}) | {world} GPUs {tbd:.2f} ms ({cells/tbd/1e6:.1f} Gcell/s) speed-up {tb1/tbd:.2f}", flush=True)
dist.barrier(); dist.destroy_process_group()
