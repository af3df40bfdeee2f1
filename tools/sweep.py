"""Time the on-chip kernels over (cluster, rows_per_thread) decompositions.  Usage: python tools/sweep.py [B] [T]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import wavetorch_b200 as wt
from wavetorch_b200 import _lib
from oracle import wave_oracle as wo
from test_gpu_parity import _vowel_model
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
combos = [(2, 8), (2, 6), (2, 5), (2, 4), (2, 3), (4, 8), (4, 5), (4, 4), (4, 3), (4, 2), (8, 5), (8, 4), (8, 2), (8, 1), (16, 2), (16, 1)]
if os.environ.get("SWEEP"):
    combos = [tuple(int(v) for v in c.split("x")) for c in os.environ["SWEEP"].split(",")]
m = _vowel_model()
x = torch.tensor(wo.synthetic_vowels(B, T), device="cuda")
labels = torch.arange(B, device="cuda") % 3
cells = B * T * 150 * 100
def tm(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for C, R in combos:
    m.cluster, m.rows_per_thread = C, R
    m.plan_flags = _lib.WT_F_FORCE_RESIDENT
    try:
        p = _lib.make_problem(150, 100, B, T, 1, 3, 1.0, 1.4283556979968262, flags=_lib.WT_F_FORCE_RESIDENT, cluster=C, rows_per_thread=R)
        plan = _lib.query_plan(p)
    except RuntimeError as e:
        print(C, R, "infeasible"); continue
    def fwd():
        with torch.no_grad(): m(x)
    def fwdtape():
        return m(x)
    out = m(x)
    loss = torch.nn.functional.cross_entropy(wt.utils.normalize_power(out.sum(1)), labels)
    (g,) = torch.autograd.grad(loss, out)
    t_f = tm(fwd); t_ft = tm(fwdtape)
    def full():
        o = m(x); o.backward(g); m.zero_grad(set_to_none=True)
    t_full = tm(full)
    print(f"C={C:2d} R={R} thr={plan.threads:4d} ncl={plan.n_clusters:3d} smem={plan.smem_fwd//1024:3d}/{plan.smem_bwd//1024:3d}KB | fwd {t_f:6.3f} ms {cells/t_f/1e6:7.1f} G/s | fwd+tape {t_ft:6.3f} | adj {t_full-t_ft:6.3f} | fwd+bwd {t_full:6.3f} ms {cells/t_full/1e6:7.1f} G/s", flush=True)
