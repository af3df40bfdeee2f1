"""Latency-bound cases: config 1/2 (B=1) and the 8-GPU strong-scaling shard (B=8) over cluster sizes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import wavetorch_b200 as wt
from wavetorch_b200 import _lib
from oracle import wave_oracle as wo
from test_gpu_parity import _vowel_model, _lens_model
def tm(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for name, mk, x, lab in (("lens B=1 T=500", lambda: _lens_model(0.5), torch.tensor(wo.propagate_waveform(500), device="cuda"), torch.tensor([2], device="cuda")),
                         ("vowel B=8 T=1000", _vowel_model, torch.tensor(wo.synthetic_vowels(8, 1000), device="cuda"), torch.arange(8, device="cuda") % 3)):
    for C, R in ((8, 1), (8, 2), (8, 3), (16, 1), (16, 2), (16, 3), (4, 4), (4, 5)):
        m = mk(); m.cluster, m.rows_per_thread, m.plan_flags = C, R, _lib.WT_F_FORCE_RESIDENT
        try:
            def fwd():
                with torch.no_grad(): m(x)
            def full():
                o = m(x); torch.nn.functional.cross_entropy(wt.utils.normalize_power(o.sum(1)), lab).backward(); m.zero_grad(set_to_none=True)
            tf, tb = tm(fwd), tm(full)
            print(f"{name} C={C:2d} R={R}: fwd {tf:.3f} ms  fwd+bwd {tb:.3f} ms", flush=True)
        except RuntimeError as e:
            print(f"{name} C={C} R={R}: infeasible ({str(e)[-60:]})")
