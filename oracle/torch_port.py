"""PyTorch-CPU port of the reference time loop, used as the timed CPU baseline ("kind": "port").

TEST / BENCHMARK INFRASTRUCTURE ONLY (see oracle/wave_oracle.py header).  The GPU box has no copy of the
reference, so bench.py times this port on the host cores instead.  It issues the same ATen work per step as
the reference does -- one 1-channel 3x3 conv2d for the Laplacian (operators.py:5-11), the elementwise update of
cell.py:12-17 behind a custom autograd.Function whose backward follows cell.py:27-44 (two more conv2d), a full
zero field + index_put + add per source (source.py:15-22), index (+pow) per probe (probe.py:15,27), a
torch.stack per step and one at the end (rnn.py:50-70) -- so its throughput tracks the reference's
(validated in the build container: tests/test_oracle_golden.py::test_torch_port_*).
Geometry helpers are shared with the numpy oracle.
"""
import numpy as np
import torch
from torch.nn.functional import conv2d


def _lap(u, h):
    k = h ** (-2) * torch.tensor([[[[0.0, 1.0, 0.0], [1.0, -4.0, 1.0], [0.0, 1.0, 0.0]]]], dtype=u.dtype)
    return conv2d(u.unsqueeze(1), k, padding=1).squeeze(1)


def _advance(b, c, u1, u2, dt, h):
    return torch.mul((dt ** -2 + b * dt ** -1).pow(-1),
                     (2 / dt ** 2 * u1 - torch.mul((dt ** -2 - b * dt ** -1), u2) + torch.mul(c.pow(2), _lap(u1, h))))


class _Step(torch.autograd.Function):
    @staticmethod
    def forward(ctx, b, c, u1, u2, dt, h):
        ctx.save_for_backward(b, c, u1, u2, dt, h)
        return _advance(b, c, u1, u2, dt, h)

    @staticmethod
    def backward(ctx, g):
        b, c, u1, u2, dt, h = ctx.saved_tensors
        gb = gc = g1 = g2 = None
        q = (b * dt + 1).pow(-1)
        if ctx.needs_input_grad[0]:
            gb = -q.pow(2) * dt * (c.pow(2) * dt ** 2 * _lap(u1, h) + 2 * u1 - 2 * u2) * g
        if ctx.needs_input_grad[1]:
            gc = q * (2 * c * dt ** 2 * _lap(u1, h)) * g
        if ctx.needs_input_grad[2]:
            g1 = dt ** 2 * _lap(q * c.pow(2) * g, h) + 2 * g * q
        if ctx.needs_input_grad[3]:
            g2 = (b * dt - 1) * q * g
        return gb, gc, g1, g2, None, None


def run_loop(c, b, rho, x, src, prb, intensity, dt, h, b0=0.0, uth=0.0, c_nl=0.0):
    """x [B,T] -> probe outputs [B,T,P]; c (and rho) may require grad.  Mirrors rnn.py:36-70."""
    B, T = x.shape
    dtype = x.dtype
    dt_t, h_t = torch.tensor(dt, dtype=dtype), torch.tensor(h, dtype=dtype)
    u1 = torch.zeros((B,) + tuple(c.shape), dtype=dtype)
    u2 = torch.zeros_like(u1)
    sx = [torch.tensor(int(i)) for i, _ in src]
    sy = [torch.tensor(int(j)) for _, j in src]
    outs = []
    for xi in x.chunk(T, dim=1):
        if b0 > 0:
            bb = b + rho * (b0 / (1 + torch.abs(u1 / uth).pow(2)))
        else:
            bb = b
        cc = c + rho * c_nl * u1.pow(2) if c_nl != 0 else c
        y = _Step.apply(bb, cc, u1, u2, dt_t, h_t)
        u2, u1 = u1, y
        for i, j in zip(sx, sy):
            add = torch.zeros(u1.size()).detach()
            add[:, i, j] = xi.squeeze(-1)
            u1 = u1 + 1.0 ** 2 * add
        vals = []
        for (i, j), sq in zip(prb, intensity):
            v = u1[:, int(i), int(j)]
            vals.append(v.pow(2) if sq else v)
        outs.append(torch.stack(vals, dim=-1))
    return torch.stack(outs, dim=1)


def geometry_c(rho, c0, c1, eta, beta, blur_kernel, blur_N):
    """rho -> c through blur and projection with autograd (geom.py:207-233)."""
    r = rho
    for _ in range(int(blur_N)):
        r = conv2d(r.unsqueeze(0).unsqueeze(0), blur_kernel, padding=blur_kernel.shape[-1] // 2).squeeze()
    proj = (np.tanh(beta * eta) + torch.tanh(beta * (r - eta))) / (np.tanh(beta * eta) + np.tanh(beta * (1 - eta)))
    return c0 + (c1 - c0) * proj


def training_step(cfg, x, labels, b0=0.0, uth=0.0, c_nl=0.0, with_grad=True):
    """One closure() of train.py:59-64 on an oracle config dict (wave_oracle.vowel_config / lens_config)."""
    from . import wave_oracle as wo
    dtype = x.dtype
    rho = torch.tensor(cfg["rho"], dtype=dtype, requires_grad=with_grad)
    k = torch.tensor(wo.disk_kernel(cfg["blur_radius"], np.float64), dtype=dtype)[None, None]
    b = torch.tensor(cfg["b"], dtype=dtype)
    with torch.set_grad_enabled(with_grad):
        c = geometry_c(rho, cfg["c0"], cfg["c1"], cfg["eta"], cfg["beta"], k, cfg["blur_N"])
        out = run_loop(c, b, rho, x, cfg["src"], cfg["prb"], cfg["intensity"], cfg["dt"], cfg["h"], b0, uth, c_nl)
        S = out.sum(dim=1)
        loss = torch.nn.functional.cross_entropy(S / S.sum(dim=1, keepdim=True), labels)
        if with_grad:
            loss.backward()
    return out.detach(), loss.detach(), (rho.grad if with_grad else None)


def time_cpu(cfg, B, T, fwd_bwd=True, repeats=1, threads=None):
    """Wall-clock seconds of one training step (or forward) of the port on the host; returns (seconds, cells)."""
    import time
    from . import wave_oracle as wo
    if threads:
        torch.set_num_threads(int(threads))
    x = torch.tensor(wo.synthetic_vowels(B, T, dtype=np.float32))
    labels = torch.arange(B) % 3
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        training_step(cfg, x, labels, with_grad=fwd_bwd)
        dtm = time.perf_counter() - t0
        best = dtm if best is None else min(best, dtm)
    return best, B * T * cfg["Nx"] * cfg["Ny"]
