"""Time the ghost-row exchange kernel alone (csrc/wt_slab.cu) on real peers.
   torchrun --nproc-per-node N tools/time_exchange.py [rows] [Ny] [B] [halo]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from wavetorch_b200 import _lib
from wavetorch_b200.domain import SlabContext
from wavetorch_b200.functional import LoopSpec
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
Ny = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
B = int(sys.argv[3]) if len(sys.argv) > 3 else 8
halo = int(sys.argv[4]) if len(sys.argv) > 4 else 16
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
spec = LoopSpec(src_ij=torch.zeros((1, 2), dtype=torch.int32, device=dev), prb_ij=torch.zeros((1, 2), dtype=torch.int32, device=dev),
                prb_sq=torch.zeros(1, dtype=torch.int32, device=dev), dt=1.0, h=1.0)
cx = SlabContext(rows, Ny, B, halo, spec, dev, None, 0)
s = cx.slabs[0]
lib = _lib.load()
def xchg():
    st = lib.wt_slab_exchange(ctypes.byref(s.desc_u), B, s.rows, Ny, _lib.ptr(s.u1), _lib.ptr(s.u2), local, _lib.stream_ptr(dev))
    _lib.check(st, "wt_slab_exchange")
for _ in range(5): xchg()
torch.cuda.synchronize(); dist.barrier()
for n in (1, 20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); dist.barrier()
    e0.record()
    for _ in range(n): xchg()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        nb = 2 * halo * Ny * B * 4
        print(f"world={world} slab rows={s.rows} Ny={Ny} B={B} halo={halo}: {t.item()*1e3:.1f} us per exchange ({n} back to back); "
              f"{nb/1e6:.1f} MB per neighbour and direction -> {nb/t.item()/1e6:.1f} GB/s per direction", flush=True)
dist.barrier(); torch.cuda.synchronize(); os._exit(0)
