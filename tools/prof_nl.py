"""Short run for ncu: BASELINE config 4 (saturable damping + Kerr), 150x100, B=64, T=300, two training iterations."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import wavetorch_b200 as wt
from oracle import wave_oracle as wo
from test_gpu_parity import _vowel_model
B, T = int(os.environ.get("PB", 64)), int(os.environ.get("PT", 300))
m = _vowel_model(0.1, 1.0, -30.0)
x = torch.tensor(0.05 * wo.synthetic_vowels(B, T), device="cuda")
lab = torch.arange(B, device="cuda") % 3
for _ in range(2):
    o = m(x); torch.nn.functional.cross_entropy(wt.utils.normalize_power(o.sum(1)), lab).backward(); m.zero_grad(set_to_none=True)
torch.cuda.synchronize()
print("done")
