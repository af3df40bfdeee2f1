"""Time config-4 variants (nonlinear) on the default plan and on the streaming path."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import wavetorch_b200 as wt
from wavetorch_b200 import _lib
from oracle import wave_oracle as wo
from test_gpu_parity import _vowel_model
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
x = torch.tensor(wo.synthetic_vowels(B, T), device="cuda")
labels = torch.arange(B, device="cuda") % 3
cells = B * T * 150 * 100
def tm(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for name, (b0, uth, cnl) in {"linear": (0, 0, 0), "satdamp": (0.1, 1.0, 0.0), "kerr": (0.0, 1.0, -30.0), "both": (0.1, 1.0, -30.0)}.items():
    for flags, pname in ((0, "auto"), (_lib.WT_F_FORCE_STREAM, "stream")):
        if pname == "stream" and T * B > 64 * 300 and name != "both": continue
        m = _vowel_model(b0, uth, cnl); m.plan_flags = flags
        p = _lib.make_problem(150, 100, B, T, 1, 3, 1.0, 1.4283556979968262, b0, uth, cnl, flags | _lib.WT_F_ZERO_INIT)
        plan = _lib.query_plan(p)
        def fwd():
            with torch.no_grad(): m(x)
        def full():
            o = m(x)
            loss = torch.nn.functional.cross_entropy(wt.utils.normalize_power(o.sum(1)), labels)
            loss.backward(); m.zero_grad(set_to_none=True)
        tf, tb = tm(fwd), tm(full)
        print(f"{name:8s} {pname:6s} path={plan.path} C={plan.cluster} R={plan.rows_per_thread} thr={plan.threads} ring={plan.reserved[0]} ncl={plan.n_clusters} | fwd {tf:8.3f} ms {cells/tf/1e6:7.1f} G/s | fwd+bwd {tb:8.3f} ms {cells/tb/1e6:7.1f} G/s", flush=True)
