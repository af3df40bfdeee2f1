"""torchrun --nproc-per-node N tools/peer_check.py : peer-memory gradient all-reduce vs NCCL (correctness, latency, graph replay)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from wavetorch_b200.peer import PeerGradReducer
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 15000
red = PeerGradReducer(2 * n, dev)
torch.manual_seed(rank)
ok = True
for it in range(20):
    m = n if it % 2 == 0 else 2 * n - 3
    x = torch.randn(m, device=dev)
    ref = x.clone() * 0.5
    dist.all_reduce(ref)
    out = red.all_reduce(x, 0.5)
    err = (out - ref).abs().max().item() / ref.abs().max().item()
    gathered = [torch.empty_like(out) for _ in range(world)]
    dist.all_gather(gathered, out)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    ok = ok and err < 1e-6 and same
    if rank == 0 and it < 3: print(f"iter {it}: n={m} rel err vs NCCL {err:.2e}, identical on all ranks: {same}", flush=True)
def tm(fn, iters=300):
    for _ in range(20): fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item() * 1e3
x = torch.randn(n, device=dev)
t_peer = tm(lambda: red.all_reduce(x))
y = x.clone()
t_nccl = tm(lambda: dist.all_reduce(y))
# graph replay
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3): out = red.all_reduce(x)
torch.cuda.current_stream().wait_stream(s)
with torch.cuda.graph(g):
    out = red.all_reduce(x)
ref = x.clone(); dist.all_reduce(ref)
for _ in range(5): g.replay()
torch.cuda.synchronize()
gerr = (out - ref).abs().max().item() / ref.abs().max().item()
t_graph = tm(lambda: g.replay())
if rank == 0:
    print(f"world {world}, n={n} floats: peer kernel {t_peer:.1f} us (graph replay {t_graph:.1f} us, err {gerr:.1e}), NCCL all_reduce {t_nccl:.1f} us; all checks {'OK' if ok and gerr < 1e-6 else 'FAILED'}", flush=True)
dist.barrier()
dist.destroy_process_group()
