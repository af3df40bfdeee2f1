"""Short run for ncu: config-3 geometry, B=64, T=200, one forward (with tape) and one adjoint."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import wavetorch_b200 as wt
from oracle import wave_oracle as wo
from test_gpu_parity import _vowel_model
B = int(os.environ.get("PB", 64)); T = int(os.environ.get("PT", 200))
m = _vowel_model()
m.cluster = int(os.environ.get("PC", 0)); m.rows_per_thread = int(os.environ.get("PR", 0))
x = torch.tensor(wo.synthetic_vowels(B, T), device="cuda")
labels = torch.arange(B, device="cuda") % 3
for _ in range(int(os.environ.get("PN", 2))):
    out = m(x)
    loss = torch.nn.functional.cross_entropy(wt.utils.normalize_power(out.sum(1)), labels)
    loss.backward()
    m.zero_grad(set_to_none=True)
torch.cuda.synchronize()
print("done")
