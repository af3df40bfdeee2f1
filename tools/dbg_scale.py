import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import wavetorch_b200 as wt
from wavetorch_b200 import _lib
from oracle import wave_oracle as wo
from test_gpu_parity import _vowel_model
B, T = 64, 1000
m = _vowel_model()
x = torch.tensor(wo.synthetic_vowels(B, T), device="cuda")
for (C, R) in [(0, 0), (2, 5), (4, 4), (8, 2)]:
    m.cluster, m.rows_per_thread = C, R
    with torch.no_grad():
        a = m(x); b = m(x); c4 = m(2.0 * x)
        f1 = m(x, output_fields=True) if False else None
    d = (c4 - 4 * a).abs()
    idx = torch.nonzero(d > 0)
    print("C,R", C, R, "repeat equal:", torch.equal(a, b), "scale maxdiff", d.max().item(), "n_diff", idx.shape[0],
          "first", idx[:3].tolist(), "rel", (d.norm() / (4 * a).norm()).item())
m.cluster = m.rows_per_thread = 0
m.plan_flags = _lib.WT_F_FORCE_STREAM
with torch.no_grad():
    a = m(x); c4 = m(2.0 * x)
d = (c4 - 4 * a).abs()
print("stream scale maxdiff", d.max().item(), (d.norm() / (4 * a).norm()).item())
# raw field check via fields on a small batch
m.plan_flags = 0
xs = x[:2, :300].contiguous()
with torch.no_grad():
    fa = m(xs, output_fields=True); fb = m(2.0 * xs, output_fields=True)
d = (fb - 2 * fa).abs()
print("fields scale maxdiff", d.max().item(), "max field", fa.abs().max().item(), "n_diff", int((d > 0).sum()))
idx = torch.nonzero(d > 0)
if idx.shape[0]:
    print("first diffs", idx[:5].tolist(), [fa[tuple(i)].item() for i in idx[:5]])
