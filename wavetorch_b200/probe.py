"""Probes: pixels whose field value (or its square) is the model output (wavetorch/probe.py)."""
import torch

from .utils import to_tensor


class WaveProbe(torch.nn.Module):
    """Reads field[:, x, y] (probe.py:14-15).  Inside WaveRNN the readout is fused into the CUDA loop."""

    squared = False

    def __init__(self, x, y):
        super().__init__()
        self.register_buffer('x', to_tensor(x, dtype=torch.int64))
        self.register_buffer('y', to_tensor(y, dtype=torch.int64))

    def pixels(self):
        return self.x.reshape(-1), self.y.reshape(-1)

    def forward(self, x):
        return x[:, self.x, self.y]


class WaveIntensityProbe(WaveProbe):
    """Reads field[:, x, y] ** 2 (probe.py:26-27)."""

    squared = True

    def forward(self, x):
        return super().forward(x).pow(2)
