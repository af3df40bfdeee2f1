"""wavetorch_b200: B200-native drop-in for the time loop of fancompute/wavetorch.

Same public names as wavetorch/__init__.py:2-10.  The WaveRNN forward/backward runs in hand-written sm_100a CUDA
kernels behind a C ABI (include/wavetorch_b200.h); there is no CPU fallback.
"""
from . import cell, geom, io, loss, probe, rnn, source, utils  # noqa: F401
from .cell import WaveCell
from .geom import WaveGeometryFreeForm, WaveGeometryHoley
from .loss import power_cross_entropy
from .probe import WaveIntensityProbe, WaveProbe
from .rnn import WaveRNN
from .source import WaveLineSource, WaveSource
from .train import train

__all__ = ["WaveCell", "WaveGeometryHoley", "WaveGeometryFreeForm", "WaveProbe", "WaveIntensityProbe", "WaveRNN",
           "WaveSource", "WaveLineSource"]

__version__ = "0.1.0"
