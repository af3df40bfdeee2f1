"""autograd.Function wrappers around the C ABI: the whole WaveRNN time loop and the single TimeStep.

`wave_rnn` replaces the Python loop of wavetorch/rnn.py:50-70 with one call to wt_forward (and one call to
wt_backward when a gradient is requested).  `time_step` is the drop-in for TimeStep.apply (cell.py:20-44).
"""
import ctypes
import warnings
from dataclasses import dataclass

import torch

from . import _lib

_warned_f64 = False


def _f32(t, name):
    """float32 view of an input of the float32 paths (a float64 model takes _WaveLoop64 instead; what still arrives here in
    float64 -- e.g. through TimeStep.apply -- is cast with one warning)."""
    global _warned_f64
    if t is None:
        return None
    if t.dtype == torch.float64 and not _warned_f64:
        warnings.warn("wavetorch_b200: the CUDA time loop computes in float32; float64 %s is cast" % name)
        _warned_f64 = True
    return t.detach().to(torch.float32).contiguous()


def _require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(
            "wavetorch_b200: %s lives on %s. The wave-RNN hot path has no CPU fallback; move the model and the "
            "inputs to a CUDA device (model.to('cuda'), x.cuda())." % (what, t.device))


@dataclass
class LoopSpec:
    """Static description of one WaveRNN forward: everything that is not a differentiable tensor."""
    src_ij: torch.Tensor      # int32 [n_src, 2] on the device
    prb_ij: torch.Tensor      # int32 [n_prb, 2]
    prb_sq: torch.Tensor      # int32 [n_prb]
    dt: float
    h: float
    b0: float = 0.0
    uth: float = 0.0
    c_nl: float = 0.0
    output_fields: bool = False
    field_every: int = 1        # output_fields only: keep every field_every-th field (time-decimated snapshots, no gradient)
    flags: int = 0
    cluster: int = 0
    rows_per_thread: int = 0
    checkpoint_every: int = 0   # > 0: keep only (u_t, u_{t-1}) every S steps and recompute each segment's tape in backward
    batch_chunk: int = 0        # > 0: process the batch in chunks of this many waveforms (bounds the tape / checkpoints)
    track_grad: bool = True     # grad mode of the caller (Function.forward itself always runs with grad disabled)
    onchip_ckpt: int = 0        # set by wave_rnn: checkpoint interval handled inside the on-chip kernels (wt_problem.checkpoint_every)


def _call_forward(lib, prob, dev, c32, b32, rho32, x32, spec, u1, u2, probe_out, probe_raw, fields, hist, ws):
    with torch.cuda.device(dev):
        st = lib.wt_forward(ctypes.byref(prob), _lib.ptr(c32), _lib.ptr(b32), _lib.ptr(rho32), _lib.ptr(x32),
                            _lib.ptr(spec.src_ij), _lib.ptr(spec.prb_ij), _lib.ptr(spec.prb_sq), _lib.ptr(u1),
                            _lib.ptr(u2), _lib.ptr(probe_out), _lib.ptr(probe_raw), _lib.ptr(fields),
                            _lib.ptr(hist), hist.numel() if hist is not None else 0, _lib.ptr(ws), ws.numel(),
                            _lib.stream_ptr(dev))
    _lib.check(st, "wt_forward")


def _dev_index(dev):
    return dev.index if dev.index is not None else torch.cuda.current_device()


class _CheckpointedLoop(torch.autograd.Function):
    """Same contract as _WaveLoop, with O(T/S + S) instead of O(T) field storage: the forward keeps the pair
    (u_t, u_{t-1}) every S steps; the backward re-runs each segment with a tape and chains the adjoint state between
    segments through wt_backward's adj1/adj2 (checkpoint-and-recompute for long T / large grids).
    Runs on the streaming kernels; the batch can additionally be processed in chunks."""

    @staticmethod
    def _segments(T, S):
        return [(s0, min(s0 + S, T)) for s0 in range(0, T, S)]

    @staticmethod
    def _problem(spec, Nx, Ny, B, T, dev, zero_init, need_b):
        flags = (spec.flags | _lib.WT_F_FORCE_STREAM) & ~_lib.WT_F_ZERO_INIT
        if zero_init:
            flags |= _lib.WT_F_ZERO_INIT
        if need_b:
            flags |= _lib.WT_F_NEED_GRAD_B
        n_src, n_prb = spec.src_ij.shape[0], spec.prb_ij.shape[0]
        return _lib.make_problem(Nx, Ny, B, T, n_src, n_prb, spec.dt, spec.h, spec.b0, spec.uth, spec.c_nl, flags,
                                 _dev_index(dev))

    @staticmethod
    def forward(ctx, x, c, b, rho, spec):
        lib = _lib.load()
        _require_cuda(x, "the input waveform x")
        _require_cuda(c, "the wave speed c")
        if spec.output_fields:
            raise NotImplementedError("wavetorch_b200: output_fields is not available with checkpoint_every > 0")
        dev = x.device
        x32, c32, b32, rho32 = _f32(x, "x"), _f32(c, "c"), _f32(b, "b"), _f32(rho, "rho")
        B, T = x32.shape
        Nx, Ny = c32.shape
        n_prb = spec.prb_ij.shape[0]
        need = ctx.needs_input_grad
        nonlinear = spec.b0 > 0 or spec.c_nl != 0
        want_grad = spec.track_grad and T > 0 and (need[0] or need[1] or need[2] or (need[3] and nonlinear))
        segs = _CheckpointedLoop._segments(T, int(spec.checkpoint_every))
        bc = int(spec.batch_chunk) if spec.batch_chunk else B
        chunks = [(b0, min(b0 + bc, B)) for b0 in range(0, B, bc)]
        out = torch.empty((B, T, n_prb), device=dev, dtype=torch.float32)
        ckpts = {}
        # One chunk: its checkpoints are taken now.  Several chunks: keeping every chunk's checkpoints until the backward would
        # defeat the chunking, so the backward re-runs a chunk's forward to take them (one more forward sweep).
        keep_ck = want_grad and len(chunks) == 1
        for ci, (b0, b1) in enumerate(chunks):
            nb = b1 - b0
            u1 = torch.empty((nb, Nx, Ny), device=dev, dtype=torch.float32)
            u2 = torch.empty_like(u1)
            for k, (s0, s1) in enumerate(segs):
                if k > 0 and keep_ck:
                    ckpts[(ci, k)] = (u1.clone(), u2.clone())
                prob = _CheckpointedLoop._problem(spec, Nx, Ny, nb, s1 - s0, dev, k == 0, False)
                plan = _lib.query_plan(prob)
                ws = torch.empty(max(int(plan.workspace_fwd_bytes), 16), device=dev, dtype=torch.uint8)
                po = torch.empty((nb, s1 - s0, n_prb), device=dev, dtype=torch.float32)
                _call_forward(lib, prob, dev, c32, b32, rho32, x32[b0:b1, s0:s1].contiguous(), spec, u1, u2, po, None, None,
                              None, ws)
                _lib.count_launches(plan.launches_fwd)
                out[b0:b1, s0:s1] = po
        ctx.no_tape = not want_grad
        ctx.shape = (Nx, Ny)
        if want_grad:
            ctx.spec, ctx.segs, ctx.chunks, ctx.ckpts = spec, segs, chunks, ckpts
            # x32/c32/b32/rho32 may alias the caller's tensors: saved through autograd so that an in-place update between
            # forward and backward trips the version-counter check instead of silently changing the recomputation
            ctx.save_for_backward(x32, c32, b32, rho32)
            ctx.dtypes = (x.dtype, c.dtype, b.dtype, rho.dtype if rho is not None else None)
        return out.to(x.dtype)

    @staticmethod
    def backward(ctx, grad_out):
        if ctx.no_tape:
            need = ctx.needs_input_grad
            zero = torch.zeros(ctx.shape, device=grad_out.device, dtype=grad_out.dtype) if need[3] else None
            return None, None, None, zero, None
        lib = _lib.load()
        spec, segs, chunks, ckpts = ctx.spec, ctx.segs, ctx.chunks, ctx.ckpts
        if ckpts is None:
            ckpts = {}
        x32, c32, b32, rho32 = ctx.saved_tensors
        dev = c32.device
        B, T = x32.shape
        Nx, Ny = c32.shape
        n_prb = spec.prb_ij.shape[0]
        need = ctx.needs_input_grad
        nonlinear = spec.b0 > 0 or spec.c_nl != 0
        g = grad_out.detach().to(torch.float32)
        z = lambda: torch.zeros((Nx, Ny), device=dev, dtype=torch.float32)
        grad_c, grad_b, grad_rho = z(), (z() if need[2] else None), (z() if (need[3] and nonlinear) else None)
        tc, tb_, tr = z(), (z() if need[2] else None), (z() if (need[3] and nonlinear) else None)
        grad_x = torch.zeros((B, T), device=dev, dtype=torch.float32) if need[0] else None
        for ci, (b0, b1) in enumerate(chunks):
            nb = b1 - b0
            if len(chunks) > 1 and len(segs) > 1:   # this chunk's checkpoints: its forward once more, without tape or outputs
                u1 = torch.empty((nb, Nx, Ny), device=dev, dtype=torch.float32)
                u2 = torch.empty_like(u1)
                for k, (s0, s1) in enumerate(segs[:-1]):
                    prob = _CheckpointedLoop._problem(spec, Nx, Ny, nb, s1 - s0, dev, k == 0, False)
                    plan = _lib.query_plan(prob)
                    ws = torch.empty(max(int(plan.workspace_fwd_bytes), 16), device=dev, dtype=torch.uint8)
                    _call_forward(lib, prob, dev, c32, b32, rho32, x32[b0:b1, s0:s1].contiguous(), spec, u1, u2, None, None,
                                  None, None, ws)
                    _lib.count_launches(plan.launches_fwd)
                    ckpts[(ci, k + 1)] = (u1.clone(), u2.clone())
                del u1, u2
            adj1 = torch.zeros((nb, Nx, Ny), device=dev, dtype=torch.float32)
            adj2 = torch.zeros_like(adj1)
            for k in range(len(segs) - 1, -1, -1):
                s0, s1 = segs[k]
                prob = _CheckpointedLoop._problem(spec, Nx, Ny, nb, s1 - s0, dev, k == 0, bool(need[2]))
                plan = _lib.query_plan(prob)
                if k > 0:
                    if ckpts.get((ci, k)) is None:
                        raise RuntimeError("wavetorch_b200: the checkpoints of this graph were consumed by an earlier "
                                           "backward(); run the forward again")
                    u1, u2 = ckpts.pop((ci, k))      # advanced in place by the recomputation: used exactly once
                else:
                    u1 = torch.empty((nb, Nx, Ny), device=dev, dtype=torch.float32)
                    u2 = torch.empty_like(u1)
                ws = torch.empty(max(int(plan.workspace_fwd_bytes), int(plan.workspace_bwd_bytes), 16), device=dev,
                                 dtype=torch.uint8)
                hist = torch.empty(max(int(plan.history_bytes), 16), device=dev, dtype=torch.uint8)
                po = torch.empty((nb, s1 - s0, n_prb), device=dev, dtype=torch.float32)
                praw = torch.empty_like(po)
                _call_forward(lib, prob, dev, c32, b32, rho32, x32[b0:b1, s0:s1].contiguous(), spec, u1, u2, po, praw, None,
                              hist, ws)
                gseg = g[b0:b1, s0:s1].contiguous()
                gx = torch.empty((nb, s1 - s0), device=dev, dtype=torch.float32) if need[0] else None
                with torch.cuda.device(dev):
                    st = lib.wt_backward(ctypes.byref(prob), _lib.ptr(c32), _lib.ptr(b32), _lib.ptr(rho32),
                                         _lib.ptr(spec.src_ij), _lib.ptr(spec.prb_ij), _lib.ptr(spec.prb_sq),
                                         _lib.ptr(gseg), _lib.ptr(praw), None, _lib.ptr(hist), hist.numel(),
                                         _lib.ptr(adj1), _lib.ptr(adj2), _lib.ptr(tc), _lib.ptr(tb_), _lib.ptr(tr),
                                         _lib.ptr(gx), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev))
                _lib.check(st, "wt_backward")
                _lib.count_launches(plan.launches_fwd + plan.launches_bwd)
                grad_c += tc
                if grad_b is not None:
                    grad_b += tb_
                if grad_rho is not None:
                    grad_rho += tr
                if gx is not None:
                    grad_x[b0:b1, s0:s1] = gx
                del hist, u1, u2
        ctx.ckpts = None
        xd, cd, bd, rd = ctx.dtypes
        if need[3] and grad_rho is None:
            grad_rho = z()
        return (grad_x.to(xd) if need[0] else None, grad_c.to(cd) if need[1] else None,
                grad_b.to(bd) if need[2] else None, grad_rho.to(rd) if need[3] else None, None)


class _WaveLoop(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, c, b, rho, spec):
        lib = _lib.load()
        _require_cuda(x, "the input waveform x")
        _require_cuda(c, "the wave speed c")
        dev = x.device
        out_dtype = x.dtype
        x32, c32, b32, rho32 = _f32(x, "x"), _f32(c, "c"), _f32(b, "b"), _f32(rho, "rho")
        B, T = x32.shape
        Nx, Ny = c32.shape
        need = ctx.needs_input_grad
        nonlinear = spec.b0 > 0 or spec.c_nl != 0
        # rho enters the loop only through the nonlinear terms; without them (or with autograd off) no tape is written
        want_grad = spec.track_grad and T > 0 and (need[0] or need[1] or need[2] or (need[3] and nonlinear))
        flags = spec.flags | _lib.WT_F_ZERO_INIT
        if need[2]:
            flags |= _lib.WT_F_NEED_GRAD_B
        if spec.output_fields and want_grad:
            flags |= _lib.WT_F_FORCE_STREAM     # dLoss/dfields is implemented by the streaming adjoint
        n_src, n_prb = spec.src_ij.shape[0], spec.prb_ij.shape[0]
        prob = _lib.make_problem(Nx, Ny, B, T, n_src, n_prb, spec.dt, spec.h, spec.b0, spec.uth, spec.c_nl, flags,
                                 dev.index if dev.index is not None else torch.cuda.current_device(), spec.cluster,
                                 spec.rows_per_thread)
        if spec.onchip_ckpt and want_grad:
            prob.checkpoint_every = int(spec.onchip_ckpt)
        fe = int(spec.field_every) if spec.output_fields else 1
        if fe > 1:
            if want_grad:
                raise NotImplementedError("wavetorch_b200: time-decimated field output (field_every > 1) is a forward-only "
                                          "mode; call it under torch.no_grad()")
            prob.field_every = fe
        plan = _lib.query_plan(prob)
        u1 = torch.empty((B, Nx, Ny), device=dev, dtype=torch.float32)
        u2 = torch.empty((B, Nx, Ny), device=dev, dtype=torch.float32)
        probe_out = torch.empty((B, T, n_prb), device=dev, dtype=torch.float32)
        probe_raw = torch.empty((B, T, n_prb), device=dev, dtype=torch.float32) if want_grad else None
        fields = torch.empty((B, T // max(fe, 1), Nx, Ny), device=dev, dtype=torch.float32) if spec.output_fields else None
        ws = torch.empty(max(int(plan.workspace_fwd_bytes), 16), device=dev, dtype=torch.uint8)
        hist = torch.empty(max(int(plan.history_bytes), 16), device=dev, dtype=torch.uint8) if want_grad else None
        with torch.cuda.device(dev):
            st = lib.wt_forward(ctypes.byref(prob), _lib.ptr(c32), _lib.ptr(b32), _lib.ptr(rho32), _lib.ptr(x32),
                                _lib.ptr(spec.src_ij), _lib.ptr(spec.prb_ij), _lib.ptr(spec.prb_sq), _lib.ptr(u1),
                                _lib.ptr(u2), _lib.ptr(probe_out), _lib.ptr(probe_raw), _lib.ptr(fields),
                                _lib.ptr(hist), hist.numel() if hist is not None else 0, _lib.ptr(ws), ws.numel(),
                                _lib.stream_ptr(dev))
        _lib.check(st, "wt_forward")
        _lib.count_launches(plan.launches_fwd)
        ctx.no_tape = not want_grad
        ctx.shape = (Nx, Ny)
        if want_grad:
            ctx.prob, ctx.plan, ctx.spec = prob, plan, spec
            # c32/b32/rho32 may alias the caller's tensors (see _CheckpointedLoop); the tape and the raw probe samples
            # ride along so that autograd frees them with the graph (and keeps them under retain_graph=True)
            ctx.save_for_backward(c32, b32, rho32, probe_raw, hist)
            ctx.dtypes = (x.dtype, c.dtype, b.dtype, rho.dtype if rho is not None else None)
        result = fields if spec.output_fields else probe_out
        return result.to(out_dtype)

    @staticmethod
    def backward(ctx, grad_out):
        if ctx.no_tape:   # only reachable when rho requires grad in linear mode: it does not enter the loop
            need = ctx.needs_input_grad
            zero = torch.zeros(ctx.shape, device=grad_out.device, dtype=grad_out.dtype) if need[3] else None
            return None, None, None, zero, None
        lib = _lib.load()
        prob, plan, spec = ctx.prob, ctx.plan, ctx.spec
        c32, b32, rho32, probe_raw, hist = ctx.saved_tensors
        dev = c32.device
        B, T, Nx, Ny = prob.B, prob.T, prob.Nx, prob.Ny
        need = ctx.needs_input_grad
        g = grad_out.detach().to(torch.float32).contiguous()
        if spec.output_fields:
            grad_fields, grad_probe = g, torch.zeros((B, T, prob.n_prb), device=dev, dtype=torch.float32)
        else:
            grad_fields, grad_probe = None, g
        grad_c = torch.empty((Nx, Ny), device=dev, dtype=torch.float32)
        grad_b = torch.empty((Nx, Ny), device=dev, dtype=torch.float32) if need[2] else None
        grad_rho = torch.empty((Nx, Ny), device=dev, dtype=torch.float32) if (need[3] and plan.nonlinear) else None
        grad_x = torch.empty((B, T), device=dev, dtype=torch.float32) if need[0] else None
        ws = torch.empty(max(int(plan.workspace_bwd_bytes), 16), device=dev, dtype=torch.uint8)
        with torch.cuda.device(dev):
            st = lib.wt_backward(ctypes.byref(prob), _lib.ptr(c32), _lib.ptr(b32), _lib.ptr(rho32),
                                 _lib.ptr(spec.src_ij), _lib.ptr(spec.prb_ij), _lib.ptr(spec.prb_sq),
                                 _lib.ptr(grad_probe), _lib.ptr(probe_raw), _lib.ptr(grad_fields), _lib.ptr(hist),
                                 hist.numel(), None, None, _lib.ptr(grad_c), _lib.ptr(grad_b), _lib.ptr(grad_rho),
                                 _lib.ptr(grad_x), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev))
        _lib.check(st, "wt_backward")
        _lib.count_launches(plan.launches_bwd)
        xd, cd, bd, rd = ctx.dtypes
        if need[3] and grad_rho is None:
            grad_rho = torch.zeros((Nx, Ny), device=dev, dtype=torch.float32)   # linear: rho does not enter the loop
        return (grad_x.to(xd) if need[0] else None, grad_c.to(cd) if need[1] else None,
                grad_b.to(bd) if need[2] else None, grad_rho.to(rd) if need[3] else None, None)


class _WaveLoop64(torch.autograd.Function):
    """The time loop in float64 (the reference's utils.set_dtype('float64') mode, utils.py:14-20): wt_forward_f64 /
    wt_backward_f64, one launch per step on the streaming kernels.  Taken when x or the geometry is float64; no
    checkpointing and no dLoss/dfields in this mode (the float32 paths are the product, this one is for double-precision
    runs and cross-checks)."""

    @staticmethod
    def forward(ctx, x, c, b, rho, spec):
        lib = _lib.load()
        _require_cuda(x, "the input waveform x")
        _require_cuda(c, "the wave speed c")
        dev = x.device
        f64 = lambda t: None if t is None else t.detach().to(torch.float64).contiguous()
        x64, c64, b64, rho64 = f64(x), f64(c), f64(b), f64(rho)
        B, T = x64.shape
        Nx, Ny = c64.shape
        need = ctx.needs_input_grad
        nonlinear = spec.b0 > 0 or spec.c_nl != 0
        want_grad = spec.track_grad and T > 0 and (need[0] or need[1] or need[2] or (need[3] and nonlinear))
        if spec.output_fields and want_grad:
            raise NotImplementedError("wavetorch_b200: gradients through output_fields=True are not available in float64; "
                                      "use float32 or the probe outputs")
        n_src, n_prb = spec.src_ij.shape[0], spec.prb_ij.shape[0]
        prob = _lib.make_problem(Nx, Ny, B, T, n_src, n_prb, spec.dt, spec.h, spec.b0, spec.uth, spec.c_nl,
                                 spec.flags | _lib.WT_F_ZERO_INIT, _dev_index(dev))
        fe = max(int(spec.field_every), 1) if spec.output_fields else 1
        prob.field_every = fe
        plan = _lib.WtPlan()
        _lib.check(lib.wt_query_plan_f64(ctypes.byref(prob), ctypes.byref(plan)), "wt_query_plan_f64")
        u1 = torch.empty((B, Nx, Ny), device=dev, dtype=torch.float64)
        u2 = torch.empty_like(u1)
        probe_out = torch.empty((B, T, n_prb), device=dev, dtype=torch.float64)
        probe_raw = torch.empty_like(probe_out) if want_grad else None
        fields = torch.empty((B, T // fe, Nx, Ny), device=dev, dtype=torch.float64) if spec.output_fields else None
        hist = torch.empty(max(int(plan.history_bytes), 16), device=dev, dtype=torch.uint8) if want_grad else None
        ws = torch.empty(max(int(plan.workspace_fwd_bytes), 16), device=dev, dtype=torch.uint8)
        with torch.cuda.device(dev):
            st = lib.wt_forward_f64(ctypes.byref(prob), _lib.ptr(c64), _lib.ptr(b64), _lib.ptr(rho64), _lib.ptr(x64),
                                    _lib.ptr(spec.src_ij), _lib.ptr(spec.prb_ij), _lib.ptr(spec.prb_sq), _lib.ptr(u1),
                                    _lib.ptr(u2), _lib.ptr(probe_out), _lib.ptr(probe_raw), _lib.ptr(fields), _lib.ptr(hist),
                                    hist.numel() if hist is not None else 0, _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev))
        _lib.check(st, "wt_forward_f64")
        _lib.count_launches(plan.launches_fwd)
        ctx.no_tape = not want_grad
        ctx.shape = (Nx, Ny)
        if want_grad:
            ctx.prob, ctx.plan, ctx.spec = prob, plan, spec
            ctx.save_for_backward(c64, b64, rho64, probe_raw, hist)
            ctx.dtypes = (x.dtype, c.dtype, b.dtype, rho.dtype if rho is not None else None)
        return (fields if spec.output_fields else probe_out).to(x.dtype if x.dtype == torch.float64 else c.dtype)

    @staticmethod
    def backward(ctx, grad_out):
        if ctx.no_tape:
            need = ctx.needs_input_grad
            zero = torch.zeros(ctx.shape, device=grad_out.device, dtype=grad_out.dtype) if need[3] else None
            return None, None, None, zero, None
        lib = _lib.load()
        prob, plan, spec = ctx.prob, ctx.plan, ctx.spec
        c64, b64, rho64, probe_raw, hist = ctx.saved_tensors
        dev = c64.device
        B, T, Nx, Ny = prob.B, prob.T, prob.Nx, prob.Ny
        need = ctx.needs_input_grad
        g = grad_out.detach().to(torch.float64).contiguous()
        grad_c = torch.empty((Nx, Ny), device=dev, dtype=torch.float64)
        grad_b = torch.empty((Nx, Ny), device=dev, dtype=torch.float64) if need[2] else None
        grad_rho = torch.empty((Nx, Ny), device=dev, dtype=torch.float64) if need[3] else None
        grad_x = torch.empty((B, T), device=dev, dtype=torch.float64) if need[0] else None
        ws = torch.empty(max(int(plan.workspace_bwd_bytes), 16), device=dev, dtype=torch.uint8)
        with torch.cuda.device(dev):
            st = lib.wt_backward_f64(ctypes.byref(prob), _lib.ptr(c64), _lib.ptr(b64), _lib.ptr(rho64), _lib.ptr(spec.src_ij),
                                     _lib.ptr(spec.prb_ij), _lib.ptr(spec.prb_sq), _lib.ptr(g), _lib.ptr(probe_raw),
                                     _lib.ptr(hist), hist.numel(), _lib.ptr(grad_c), _lib.ptr(grad_b), _lib.ptr(grad_rho),
                                     _lib.ptr(grad_x), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev))
        _lib.check(st, "wt_backward_f64")
        _lib.count_launches(plan.launches_bwd)
        xd, cd, bd, rd = ctx.dtypes
        return (grad_x.to(xd) if need[0] else None, grad_c.to(cd) if need[1] else None,
                grad_b.to(bd) if need[2] else None, grad_rho.to(rd) if need[3] else None, None)


def wave_rnn(x, c, b, rho, spec):
    """Run the fused time loop.  x [B,T]; c, b, rho [Nx,Ny]; returns [B,T,n_prb] (or [B,T,Nx,Ny])."""
    spec.track_grad = torch.is_grad_enabled()
    if x.dtype == torch.float64 or c.dtype == torch.float64:
        return _WaveLoop64.apply(x, c, b, rho, spec)
    T = x.shape[1]
    chunked = bool(spec.batch_chunk) and 0 < spec.batch_chunk < x.shape[0]
    if spec.checkpoint_every and 0 < spec.checkpoint_every < T and not chunked and not spec.output_fields and c.is_cuda:
        # small grids: checkpoint-and-recompute inside the on-chip kernels (snapshots of the register patches, a tape that
        # lives for one segment) when the planner puts this problem on that path; otherwise the segment loop below
        prob = _lib.make_problem(c.shape[0], c.shape[1], x.shape[0], T, spec.src_ij.shape[0], spec.prb_ij.shape[0], spec.dt,
                                 spec.h, spec.b0, spec.uth, spec.c_nl, spec.flags | _lib.WT_F_ZERO_INIT,
                                 _dev_index(c.device), spec.cluster, spec.rows_per_thread)
        prob.checkpoint_every = int(spec.checkpoint_every)
        plan = _lib.query_plan(prob)
        if plan.path == _lib.WT_PATH_RESIDENT and plan.reserved[2] > 0:
            spec.onchip_ckpt = int(spec.checkpoint_every)
            return _WaveLoop.apply(x, c, b, rho, spec)
    if (spec.checkpoint_every and 0 < spec.checkpoint_every < T) or (chunked and T > 0 and not spec.output_fields):
        if not spec.checkpoint_every or spec.checkpoint_every >= T:
            spec.checkpoint_every = T      # batch chunks without time checkpoints: one segment per chunk
        return _CheckpointedLoop.apply(x, c, b, rho, spec)
    return _WaveLoop.apply(x, c, b, rho, spec)


class TimeStep(torch.autograd.Function):
    """Drop-in for wavetorch.cell.TimeStep (cell.py:20-44): y = TimeStep.apply(b, c, y1, y2, dt, h)."""

    @staticmethod
    def forward(ctx, b, c, y1, y2, dt, h):
        lib = _lib.load()
        _require_cuda(y1, "the field y1")
        dev = y1.device
        B, Nx, Ny = y1.shape
        b32, c32, y132, y232 = (_f32(t, n) for t, n in ((b, "b"), (c, "c"), (y1, "y1"), (y2, "y2")))
        for t, n in ((b32, "b"), (c32, "c")):
            if tuple(t.shape) not in ((Nx, Ny), (B, Nx, Ny)):
                raise ValueError("TimeStep: %s must be [Nx,Ny] or [B,Nx,Ny], got %s" % (n, tuple(t.shape)))
        prob = _lib.make_problem(Nx, Ny, B, 1, 0, 0, float(dt), float(h),
                                 device=dev.index if dev.index is not None else torch.cuda.current_device())
        y = torch.empty_like(y132)
        with torch.cuda.device(dev):
            st = lib.wt_step_forward(ctypes.byref(prob), _lib.ptr(b32), int(b32.dim() == 3), _lib.ptr(c32),
                                     int(c32.dim() == 3), _lib.ptr(y132), _lib.ptr(y232), _lib.ptr(y),
                                     _lib.stream_ptr(dev))
        _lib.check(st, "wt_step_forward")
        _lib.count_launches(1)
        ctx.prob = prob
        ctx.save_for_backward(b32, c32, y132, y232)
        ctx.dtypes = (b.dtype, c.dtype, y1.dtype, y2.dtype)
        return y.to(y1.dtype)

    @staticmethod
    def backward(ctx, grad_output):
        lib = _lib.load()
        b32, c32, y1, y2 = ctx.saved_tensors
        dev = y1.device
        need = ctx.needs_input_grad
        g = grad_output.detach().to(torch.float32).contiguous()
        outs = [torch.empty_like(y1) if need[i] else None for i in range(4)]
        with torch.cuda.device(dev):
            st = lib.wt_step_backward(ctypes.byref(ctx.prob), _lib.ptr(b32), int(b32.dim() == 3), _lib.ptr(c32),
                                      int(c32.dim() == 3), _lib.ptr(y1), _lib.ptr(y2), _lib.ptr(g),
                                      _lib.ptr(outs[0]), _lib.ptr(outs[1]), _lib.ptr(outs[2]), _lib.ptr(outs[3]),
                                      _lib.stream_ptr(dev))
        _lib.check(st, "wt_step_backward")
        _lib.count_launches(1)
        gb, gc, gy1, gy2 = outs
        # per-sample gradients are summed over the batch when the coefficient was shared (what autograd's
        # sum_to_size does for the reference)
        if gb is not None and b32.dim() == 2:
            gb = gb.sum(0)
        if gc is not None and c32.dim() == 2:
            gc = gc.sum(0)
        cast = lambda t, d: None if t is None else t.to(d)
        bd, cd, y1d, y2d = ctx.dtypes
        return cast(gb, bd), cast(gc, cd), cast(gy1, y1d), cast(gy2, y2d), None, None


def time_step(b, c, y1, y2, dt, h):
    return TimeStep.apply(b, c, y1, y2, dt, h)
