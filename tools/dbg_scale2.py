import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import wavetorch_b200 as wt
from wavetorch_b200 import _lib
from oracle import wave_oracle as wo
from test_gpu_parity import _vowel_model
m = _vowel_model()
x = torch.tensor(wo.synthetic_vowels(1, 300), device="cuda")
for flags in (0, _lib.WT_F_FORCE_STREAM):
    m.plan_flags = flags
    with torch.no_grad():
        fa = m(x, output_fields=True)[0]; fb = m(2.0 * x, output_fields=True)[0]
    d = (fb - 2 * fa).abs()
    rel = d / (2 * fa.abs()).clamp_min(1e-30)
    big = (fa.abs() > 1e-20) & (d > 0)
    idx = torch.nonzero(big)
    print("flags", flags, "n big diffs", idx.shape[0])
    for i in idx[:12].tolist():
        t, r, c = i
        print("  t,r,c", i, "fa", fa[t, r, c].item(), "fb/2", fb[t, r, c].item() / 2, "rel", rel[t, r, c].item())
    # per time-step count
    cnt = big.reshape(300, -1).sum(1)
    nz = torch.nonzero(cnt)[:5].flatten().tolist()
    print("  first steps with diffs:", nz, [int(cnt[k]) for k in nz])
