"""Short run for ncu: BASELINE config 2 (151x151, B=1, T=500), two training iterations."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import wavetorch_b200 as wt
from oracle import wave_oracle as wo
from test_gpu_parity import _lens_model
m = _lens_model(0.5); x = torch.tensor(wo.propagate_waveform(500), device="cuda")
for _ in range(3):
    o = m(x); torch.nn.functional.cross_entropy(wt.utils.normalize_power(o.sum(1)), torch.tensor([2], device="cuda")).backward(); m.zero_grad(set_to_none=True)
torch.cuda.synchronize()
print("done")
