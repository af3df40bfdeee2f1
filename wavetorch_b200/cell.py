"""WaveCell: one leapfrog step of the damped scalar wave equation (API of wavetorch/cell.py).

`WaveRNN` does not call this module once per step any more -- the loop is fused into wt_forward -- but the
step-level API is kept: `WaveCell.forward(h1, h2, c_linear, rho) -> (y, h1)` runs the single-step CUDA kernel
through `functional.TimeStep`, with the nonlinear b(u), c(u) expressions in PyTorch as in cell.py:94-102.
"""
import math

import torch

from .functional import TimeStep, time_step  # noqa: F401  (TimeStep is part of the public surface)
from .utils import to_tensor


def saturable_damping(u, uth, b0):
    """b0 / (1 + |u/uth|^2)  (cell.py:8-9)."""
    return b0 / (1 + torch.abs(u / uth).pow(2))


class WaveCell(torch.nn.Module):
    """The recurrent cell implementing the scalar wave equation."""

    def __init__(self, dt: float, geometry, satdamp_b0: float = 0.0, satdamp_uth: float = 0.0, c_nl: float = 0.0):
        super().__init__()
        self.register_buffer("dt", to_tensor(dt))
        self.geom = geometry
        self.register_buffer("satdamp_b0", to_tensor(satdamp_b0))
        self.register_buffer("satdamp_uth", to_tensor(satdamp_uth))
        self.register_buffer("c_nl", to_tensor(c_nl))
        # host copies so that the fused loop never has to synchronise on .item()
        self._host = dict(dt=float(self.dt), b0=float(self.satdamp_b0), uth=float(self.satdamp_uth),
                          c_nl=float(self.c_nl))
        cmax = self.geom.cmax
        h = self.geom.h.item()
        if dt > 1 / cmax * h / math.sqrt(2):      # CFL condition (cell.py:70-73)
            raise ValueError(
                'The spatial discretization defined by the geometry `h = %f` and the temporal discretization defined '
                'by the model `dt = %f` do not satisfy the CFL stability criteria' % (h, dt))

    def host_scalars(self):
        """dt, b0, uth, c_nl as Python floats (float32-rounded like the reference's buffers)."""
        if getattr(self, "_host", None) is None:
            self._host = dict(dt=float(self.dt), b0=float(self.satdamp_b0), uth=float(self.satdamp_uth),
                              c_nl=float(self.c_nl))
        return self._host

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self._host = None

    def parameters(self, recursive=True):
        for param in self.geom.parameters():
            yield param

    def forward(self, h1, h2, c_linear, rho):
        """Advance one step: returns (u_{t+1}, u_t) given h1 = u_t, h2 = u_{t-1} (cell.py:79-107)."""
        s = self.host_scalars()
        if s["b0"] > 0:
            b = self.geom.b + rho * saturable_damping(h1, uth=self.satdamp_uth, b0=self.satdamp_b0)
        else:
            b = self.geom.b
        if s["c_nl"] != 0:
            c = c_linear + rho * self.c_nl * h1.pow(2)
        else:
            c = c_linear
        y = TimeStep.apply(b, c, h1, h2, s["dt"], self.geom.h.item())
        return y, h1
