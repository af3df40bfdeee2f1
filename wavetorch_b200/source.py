"""Sources: pixels that receive the input waveform after every step (wavetorch/source.py)."""
import numpy as np
import torch

from .utils import to_tensor


def line_pixels(r0, c0, r1, c1):
    """Bresenham line, both end points included: what skimage.draw.line returns at source.py:31."""
    r0, c0, r1, c1 = int(r0), int(c0), int(r1), int(c1)
    dr, dc = abs(r1 - r0), abs(c1 - c0)
    step_r = 1 if r1 >= r0 else -1
    step_c = 1 if c1 >= c0 else -1
    rows, cols = [], []
    r, c = r0, c0
    if dc >= dr:                      # shallow: one pixel per column
        err = dc // 2
        for _ in range(dc + 1):
            rows.append(r); cols.append(c)
            err -= dr
            if err < 0:
                r += step_r
                err += dc
            c += step_c
    else:                             # steep: one pixel per row
        err = dr // 2
        for _ in range(dr + 1):
            rows.append(r); cols.append(c)
            err -= dc
            if err < 0:
                c += step_c
                err += dr
            r += step_r
    return np.asarray(rows, dtype=np.int64), np.asarray(cols, dtype=np.int64)


class WaveSource(torch.nn.Module):
    """Point (or multi-pixel) source at rows `x`, columns `y` (int64 buffers, source.py:8-13).

    Inside WaveRNN the injection is fused into the CUDA time loop.  Called directly, `forward(Y, X, dt=1.0)` adds
    dt^2 * X[b] to Y[b, x, y] like source.py:15-22 (with dt defaulting to 1.0, which is what rnn.py:57 uses).
    For a multi-pixel source every pixel receives X[b]; the reference only defines that case for B == 1
    (SURVEY appendix B-2).
    """

    def __init__(self, x, y):
        super().__init__()
        self.register_buffer('x', to_tensor(x, dtype=torch.int64))
        self.register_buffer('y', to_tensor(y, dtype=torch.int64))

    def pixels(self):
        """Flat int64 views (rows, cols) of all pixels of this source."""
        return self.x.reshape(-1), self.y.reshape(-1)

    def forward(self, Y, X, dt=1.0):
        rows, cols = self.pixels()
        add = torch.zeros_like(Y)
        add[:, rows, cols] = (dt ** 2 * X).reshape(-1, 1).to(Y.dtype).expand(-1, rows.numel())
        return Y + add


class WaveLineSource(WaveSource):
    """All pixels on the segment (r0,c0)-(r1,c1) (source.py:29-37)."""

    def __init__(self, r0, c0, r1, c1):
        rows, cols = line_pixels(r0, c0, r1, c1)
        self.r0, self.c0, self.r1, self.c1 = r0, c0, r1, c1
        super().__init__(rows, cols)
