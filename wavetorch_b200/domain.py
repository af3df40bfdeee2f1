"""Row-slab domain decomposition of ONE wave simulation across the GPUs of a box (BASELINE config 5).

The reference has nothing like this; it is the multi-GPU mode for grids that do not fit one GPU.
Rank g owns rows [r0, r1) of the [Nx, Ny] grid plus `halo` ghost rows on each interior side.  Time advances in blocks of
`halo` steps during which every rank integrates its extended slab as if it were a whole domain -- the error made at the
artificial slab edges travels inwards one row per step, so after `halo` steps the owned rows are still exact -- then
neighbours refresh each other's ghost rows ("halo depth = temporal block").

    exchange    inside the C library, in the stream, without the host: `wt_slab_forward` / `wt_slab_backward` launch one
                kernel per exchange that stores the boundary rows of both time levels straight into the neighbour's
                ghost rows over NVLink (peer-mapped pointers from torch symmetric memory) and synchronises with the
                neighbours through flag words (csrc/wt_slab.cu).  No NCCL on the data path.
    memory      every rank's state (u1, u2, adj1, adj2) lives in ONE symmetric allocation made once per shape; checkpoints
                (u_t, u_{t-1}) are taken every `checkpoint_every` steps (any multiple of `halo`), the tape exists for one
                checkpoint segment at a time, and `batch_chunk` bounds all of it further (see `memory_model`).
    probes      read by the owning rank, summed over ranks ([B,T,P] all-reduce: a few KB)
    gradient    every rank produces dLoss/dc for ITS rows only; the slabs are all-gathered (cells are disjoint -- no
                reduction), after which the geometry chain runs replicated.

`virtual_ranks=n` runs n slabs inside one process on one GPU, each on its own CUDA stream, through the SAME exchange
kernel and flag protocol (the peers are plain device pointers): the single-GPU tests of the decomposition.

Saturable damping / Kerr terms: forward only.  Their adjoint coefficients depend on the local field, which is inexact in
the ghost rows during a block, so the decomposed adjoint would not be exact; it is refused.
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib
from .distributed import shard_bounds
from .functional import LoopSpec, _dev_index, _f32, _require_cuda


class WtSlab(ctypes.Structure):
    """include/wavetorch_b200.h: wt_slab"""
    _fields_ = [("halo", ctypes.c_int32), ("up", ctypes.c_int32), ("dn", ctypes.c_int32),
                ("up_Nx", ctypes.c_int32), ("dn_Nx", ctypes.c_int32), ("reserved", ctypes.c_int32 * 3),
                ("up_f1", ctypes.c_uint64), ("up_f2", ctypes.c_uint64), ("dn_f1", ctypes.c_uint64),
                ("dn_f2", ctypes.c_uint64), ("up_flags", ctypes.c_uint64), ("dn_flags", ctypes.c_uint64),
                ("flags", ctypes.c_void_p), ("state", ctypes.c_void_p)]


def slab_rows(Nx, world, rank, halo):
    """(r0, r1, e0, e1): owned rows [r0, r1) and extended rows [e0, e1) of `rank`."""
    r0, r1 = shard_bounds(Nx, world, rank)
    return r0, r1, max(0, r0 - halo), min(Nx, r1 + halo)


def memory_model(Nx, Ny, B, T, world, halo, checkpoint_every, batch_chunk=0, rank=None):
    """Bytes of device memory the decomposed forward+backward needs on the fattest rank (what bench.py compares with
    torch.cuda.max_memory_allocated): state, checkpoints, the tape of one checkpoint segment and the library workspace."""
    Bc = batch_chunk if batch_chunk and batch_chunk < B else B
    S = min(checkpoint_every if checkpoint_every else T, T)
    n_seg = (T + S - 1) // S
    rows = max((slab_rows(Nx, world, r, halo)[3] - slab_rows(Nx, world, r, halo)[2]) for r in
               (range(world) if rank is None else [rank]))
    field = Bc * rows * Ny * 4
    state = 4 * field                                   # u1, u2, adj1, adj2 (symmetric allocation)
    ckpt = 2 * field * (n_seg - 1)                      # (u_t, u_{t-1}) at the start of every segment but the first, for ONE
                                                        # chunk: with several chunks the backward re-runs a chunk's forward
    tape = field * S                                    # L(u_{t-1}) per step of the segment being differentiated
    work = 3 * field + 9 * rows * Ny * 4                # ping-pong partners of the blocked kernels + coefficient planes
    return dict(state=state, checkpoints=ckpt, tape=tape, workspace=work, total=state + ckpt + tape + work,
                rows=rows, batch_chunk=Bc, segment=S, segments=n_seg)


def gather_row_slabs(own, Nx, group=None):
    """Assemble a [Nx, Ny] field from every rank's owned rows (`own`: [r1-r0, Ny] of this rank).  The slabs are disjoint, so
    this is an all-gather, not a reduction; ragged slab heights are padded to the tallest one for the collective."""
    world, Ny = dist.get_world_size(group), own.shape[1]
    heights = [shard_bounds(Nx, world, r)[1] - shard_bounds(Nx, world, r)[0] for r in range(world)]
    mine = own.new_zeros((max(heights), Ny))   # noqa: E501
    mine[:own.shape[0]] = own
    hmax = max(heights)
    parts = own.new_empty((world * hmax, Ny))
    dist.all_gather_into_tensor(parts, mine, group=group)
    return torch.cat([parts[r * hmax:r * hmax + heights[r]] for r in range(world)], dim=0)


class _SlabRank:
    """One rank's slab: geometry, its share of the sources / probes, state tensors and exchange descriptors."""

    def __init__(self, rank, world, Nx, Ny, halo, src_ij, prb_ij, prb_sq, dev):
        self.rank, self.world, self.halo, self.dev = rank, world, halo, dev
        self.r0, self.r1, self.e0, self.e1 = slab_rows(Nx, world, rank, halo)
        self.up, self.dn = self.r0 - self.e0, self.e1 - self.r1      # ghost rows above / below
        self.rows = self.e1 - self.e0
        src, prb = src_ij.cpu(), prb_ij.cpu()
        in_ext = lambda t: (t[:, 0] >= self.e0) & (t[:, 0] < self.e1)
        in_own = lambda t: (t[:, 0] >= self.r0) & (t[:, 0] < self.r1)
        off = torch.tensor([[self.e0, 0]], dtype=torch.int32)
        self.src_ext = (src[in_ext(src)] - off).contiguous().to(dev)          # injected in forward (ghosts included)
        self.src_own = (src[in_own(src)] - off).contiguous().to(dev)          # gathered for dLoss/dx
        ids = torch.nonzero(in_ext(prb)).flatten()                            # probes inside the extended slab
        self.prb_ext = (prb[ids] - off).contiguous().to(dev)
        self.prb_owned = in_own(prb[ids]).to(dev)                             # which of them this rank reports
        self.prb_ids = ids.to(dev)
        self.sq = prb_sq[self.prb_ids].contiguous()
        self.n_p = int(ids.numel())
        self.stream = None
        self.ck = {}

    def bind(self, fields, flags, state, peers):
        """fields: my [u1,u2,a1,a2] tensors [Bc,rows,Ny]; peers: {-1/+1: (rows, [ptr u1,u2,a1,a2], flags_ptr)}."""
        self.u1, self.u2, self.a1, self.a2 = fields
        self.flags, self.state = flags, state

        def desc(i1, i2):
            s = WtSlab()
            s.halo, s.up, s.dn = self.halo, self.up, self.dn
            if self.up:
                rows, ptrs, fl = peers[-1]
                s.up_Nx, s.up_f1, s.up_f2, s.up_flags = rows, ptrs[i1], ptrs[i2], fl
            if self.dn:
                rows, ptrs, fl = peers[1]
                s.dn_Nx, s.dn_f1, s.dn_f2, s.dn_flags = rows, ptrs[i1], ptrs[i2], fl
            s.flags, s.state = flags.data_ptr(), state.data_ptr()
            return s

        self.desc_u, self.desc_a = desc(0, 1), desc(2, 3)


class SlabContext:
    """Peer-mapped state of the decomposition for one problem shape (allocated once, reused by every forward/backward)."""

    def __init__(self, Nx, Ny, Bc, halo, spec, dev, group, virtual):
        self.key = (Nx, Ny, Bc, halo, str(dev), virtual, spec.src_ij.data_ptr(), spec.prb_ij.data_ptr())
        self.virtual = virtual
        world = virtual if virtual else dist.get_world_size(group)
        ranks = list(range(world)) if virtual else [dist.get_rank(group)]
        self.world = world
        if halo < 8 or halo % 8:
            raise ValueError("domain decomposition: halo must be a positive multiple of 8 (got %d)" % halo)
        if min(shard_bounds(Nx, world, r)[1] - shard_bounds(Nx, world, r)[0] for r in range(world)) < halo:
            raise ValueError("domain decomposition: every rank needs at least `halo` (= %d) rows" % halo)
        self.slabs = [_SlabRank(r, world, Nx, Ny, halo, spec.src_ij, spec.prb_ij, spec.prb_sq, dev) for r in ranks]
        rows_of = lambda r: slab_rows(Nx, world, r, halo)[3] - slab_rows(Nx, world, r, halo)[2]
        fmax = Bc * max(rows_of(r) for r in range(world)) * Ny            # floats per field slot, same on every rank
        fmax = (fmax + 63) // 64 * 64
        total = 4 * fmax + 64                                             # + flag words (uint32[4]) + epoch state ([2])
        if virtual:
            bufs = [torch.zeros(total, dtype=torch.float32, device=dev) for _ in ranks]
            base = [b.data_ptr() for b in bufs]
            self._keep = bufs
        else:
            import torch.distributed._symmetric_memory as symm_mem
            buf = symm_mem.empty(total, dtype=torch.float32, device=dev)
            buf.zero_()
            handle = symm_mem.rendezvous(buf, group if group is not None else dist.group.WORLD)
            base = [int(p) for p in handle.buffer_ptrs]
            if len(base) != world or any(p == 0 for p in base):
                raise RuntimeError("wavetorch_b200: symmetric memory rendezvous did not map every peer")
            bufs = [buf]
            self._keep = (buf, handle)
            torch.cuda.synchronize(dev)
            dist.barrier(group)            # every rank's flags are zero before anybody's first store can land
        for s, buf in zip(self.slabs, bufs):
            n = Bc * s.rows * Ny
            fields = [buf[i * fmax:i * fmax + n].view(Bc, s.rows, Ny) for i in range(4)]
            words = buf[4 * fmax:4 * fmax + 64].view(torch.int32)
            peers = {}
            for d in (-1, 1):
                q = s.rank + d
                if 0 <= q < world:
                    peers[d] = (rows_of(q), [base[q] + 4 * i * fmax for i in range(4)], base[q] + 4 * 4 * fmax)
            s.bind(fields, words[:4], words[8:10], peers)
            if virtual:
                s.stream = torch.cuda.Stream(device=dev)
        self.exchange_bytes = 2 * halo * Ny * Bc * 4          # per neighbour and direction, per exchange


class _DomainLoop(torch.autograd.Function):
    @staticmethod
    def _problem(spec, s, Ny, B, T, dev, zero_init, n_src):
        flags = (spec.flags | _lib.WT_F_FORCE_STREAM) & ~_lib.WT_F_ZERO_INIT
        if zero_init:
            flags |= _lib.WT_F_ZERO_INIT
        return _lib.make_problem(s.rows, Ny, B, T, n_src, s.n_p, spec.dt, spec.h, spec.b0, spec.uth, spec.c_nl,
                                 flags, _dev_index(dev))

    @staticmethod
    def _on(s):
        """Context: the stream this slab's work is enqueued on (its own stream for virtual ranks)."""
        return torch.cuda.stream(s.stream) if s.stream is not None else _Null()

    @staticmethod
    def _fork(ctxs):
        cur = torch.cuda.current_stream()
        for s in ctxs.slabs:
            if s.stream is not None:
                s.stream.wait_stream(cur)

    @staticmethod
    def _join(ctxs):
        cur = torch.cuda.current_stream()
        for s in ctxs.slabs:
            if s.stream is not None:
                cur.wait_stream(s.stream)

    @staticmethod
    def _forward_segment(lib, spec, cx, L, Ny, nb, xsrc, T, zero_init, hist, dev):
        """wt_slab_forward on every local slab; returns [(probe_out, probe_raw)] per slab.  xsrc: view of x for the segment."""
        res = []
        for s, loc in zip(cx.slabs, L):
            with _DomainLoop._on(s):
                xs = xsrc.contiguous()      # on this slab's stream
                prob = _DomainLoop._problem(spec, s, Ny, nb, T, dev, zero_init, s.src_ext.shape[0])
                plan = _lib.query_plan(prob)
                need = int(plan.workspace_fwd_bytes)
                if loc.get("ws") is None or loc["ws"].numel() < need:
                    loc["ws"] = torch.empty(max(need, 16), device=dev, dtype=torch.uint8)
                h = None
                if hist:
                    hb = max(int(plan.history_bytes), 16)
                    if loc.get("hist") is None or loc["hist"].numel() < hb:
                        loc["hist"] = None
                        loc["hist"] = torch.empty(hb, device=dev, dtype=torch.uint8)
                    h = loc["hist"]
                po = torch.empty((nb, T, max(s.n_p, 1)), device=dev, dtype=torch.float32)
                pr = torch.empty_like(po)
                with torch.cuda.device(dev):
                    st = lib.wt_slab_forward(ctypes.byref(prob), ctypes.byref(s.desc_u), _lib.ptr(loc["c"]),
                                             _lib.ptr(loc["b"]), _lib.ptr(loc["rho"]), _lib.ptr(xs), _lib.ptr(s.src_ext),
                                             _lib.ptr(s.prb_ext), _lib.ptr(s.sq), _lib.ptr(s.u1), _lib.ptr(s.u2),
                                             _lib.ptr(po) if s.n_p else None, _lib.ptr(pr) if s.n_p else None,
                                             _lib.ptr(h), h.numel() if h is not None else 0, _lib.ptr(loc["ws"]),
                                             loc["ws"].numel(), _lib.stream_ptr(dev))
                _lib.check(st, "wt_slab_forward")
                _lib.count_launches(plan.launches_fwd + (T + s.halo - 1) // s.halo)
                res.append((po, pr))
        return res

    @staticmethod
    def _forward_chunk(lib, spec, cx, L, Ny, x32, segs, ci, b0, b1, keep_ck, out, raw, dev):
        """All segments of one batch chunk without a tape; optionally checkpoints, optionally the probe series."""
        nb = b1 - b0
        for k, (s0, s1) in enumerate(segs):
            if keep_ck and k > 0:
                for s in cx.slabs:
                    with _DomainLoop._on(s):
                        s.ck[(ci, k)] = (s.u1[:nb].clone(), s.u2[:nb].clone())
            if k == len(segs) - 1 and out is None:
                break                     # checkpoints only: the last segment's end state is not needed
            res = _DomainLoop._forward_segment(lib, spec, cx, L, Ny, nb, x32[b0:b1, s0:s1], s1 - s0, k == 0, False, dev)
            if out is not None:
                for s, (po, pr) in zip(cx.slabs, res):
                    if s.n_p:
                        with _DomainLoop._on(s):
                            ids = s.prb_ids[s.prb_owned]
                            out[b0:b1, s0:s1, ids] = po[:, :, s.prb_owned]
                            raw[b0:b1, s0:s1, ids] = pr[:, :, s.prb_owned]

    @staticmethod
    def forward(ctx, x, c, b, rho, spec, halo, S, bc, group, virtual, holder):
        lib = _lib.load()
        _require_cuda(x, "the input waveform x")
        _require_cuda(c, "the wave speed c")
        dev = x.device
        x32, c32, b32, rho32 = _f32(x, "x"), _f32(c, "c"), _f32(b, "b"), _f32(rho, "rho")
        B, T = x32.shape
        Nx, Ny = c32.shape
        P = spec.prb_ij.shape[0]
        need = ctx.needs_input_grad
        nonlinear = spec.b0 > 0 or spec.c_nl != 0
        want_grad = spec.track_grad and T > 0 and (need[0] or need[1] or need[2] or (need[3] and nonlinear))
        if want_grad and (nonlinear or need[2]):
            raise NotImplementedError(
                "wavetorch_b200: the domain-decomposed adjoint supports the linear cell only (saturable damping / Kerr "
                "coefficients depend on the field, which is inexact in the ghost rows); run the forward under "
                "torch.no_grad() or use batch sharding")
        Bc = bc if bc and bc < B else B
        cx = holder.context(Nx, Ny, Bc, halo, spec, dev, group, virtual)
        S = max(halo, (int(S) + halo - 1) // halo * halo) if S else T
        segs = [(s0, min(s0 + S, T)) for s0 in range(0, T, S)]
        chunks = [(b0, min(b0 + Bc, B)) for b0 in range(0, B, Bc)]
        L = holder.local(cx, c32, b32, rho32)
        out = torch.zeros((B, T, P), device=dev, dtype=torch.float32)
        raw = torch.zeros((B, T, P), device=dev, dtype=torch.float32)
        for s in cx.slabs:
            s.ck = {}
        # One chunk: its checkpoints are taken now.  Several chunks: keeping every chunk's checkpoints until the backward
        # would defeat the chunking, so the backward re-runs a chunk's forward to take them (one more forward sweep).
        keep_ck = want_grad and len(chunks) == 1
        _DomainLoop._fork(cx)
        for ci, (b0, b1) in enumerate(chunks):
            _DomainLoop._forward_chunk(lib, spec, cx, L, Ny, x32, segs, ci, b0, b1, keep_ck, out, raw, dev)
        _DomainLoop._join(cx)
        if not virtual:
            dist.all_reduce(out, group=group)
            if want_grad:
                dist.all_reduce(raw, group=group)
        ctx.no_tape = not want_grad
        ctx.shape = (Nx, Ny)
        if want_grad:
            ctx.save_for_backward(x32, c32, b32, rho32)
            ctx.meta = (spec, group, virtual, cx, L, segs, chunks, raw, Nx,
                        (x.dtype, c.dtype, b.dtype, rho.dtype if rho is not None else None))
        return out.to(x.dtype)

    @staticmethod
    def backward(ctx, grad_out):
        if ctx.no_tape:
            need = ctx.needs_input_grad
            zero = torch.zeros(ctx.shape, device=grad_out.device, dtype=grad_out.dtype) if need[3] else None
            return (None, None, None, zero) + (None,) * 7
        lib = _lib.load()
        spec, group, virtual, cx, L, segs, chunks, raw, NX, dtypes = ctx.meta
        x32, c32, b32, rho32 = ctx.saved_tensors
        dev = x32.device
        B, T = x32.shape
        Ny = c32.shape[1]
        need = ctx.needs_input_grad
        g = grad_out.detach().to(torch.float32)
        gxs = [torch.zeros((B, T), device=dev, dtype=torch.float32) if need[0] else None for _ in cx.slabs]
        gown = [torch.zeros((s.r1 - s.r0, Ny), device=dev, dtype=torch.float32) for s in cx.slabs]
        tc = [torch.empty((s.rows, Ny), device=dev, dtype=torch.float32) for s in cx.slabs]
        _DomainLoop._fork(cx)
        for ci, (b0, b1) in enumerate(chunks):
            nb = b1 - b0
            if len(chunks) > 1:
                _DomainLoop._forward_chunk(lib, spec, cx, L, Ny, x32, segs, ci, b0, b1, True, None, None, dev)
            for s in cx.slabs:
                with _DomainLoop._on(s):
                    s.a1[:nb].zero_()
                    s.a2[:nb].zero_()
            for k in range(len(segs) - 1, -1, -1):
                s0, s1 = segs[k]
                xs = x32[b0:b1, s0:s1]
                if k > 0:
                    for s in cx.slabs:
                        if s.ck.get((ci, k)) is None:
                            raise RuntimeError("wavetorch_b200: the checkpoints of this graph were consumed by an earlier "
                                               "backward(); run the forward again")
                        with _DomainLoop._on(s):
                            u1, u2 = s.ck.pop((ci, k))
                            s.u1[:nb].copy_(u1)
                            s.u2[:nb].copy_(u2)
                            del u1, u2
                # forward of the segment with a tape (ghost sources included; the same exchanges as in the forward pass)
                _DomainLoop._forward_segment(lib, spec, cx, L, Ny, nb, xs, s1 - s0, k == 0, True, dev)
                # adjoint of the segment: seeds for every probe inside the extended slab, with the EXACT raw samples
                for i, (s, loc) in enumerate(zip(cx.slabs, L)):
                    with _DomainLoop._on(s):
                        pb = _DomainLoop._problem(spec, s, Ny, nb, s1 - s0, dev, k == 0, s.src_own.shape[0])
                        plan_b = _lib.query_plan(pb)
                        needb = int(plan_b.workspace_bwd_bytes)
                        if loc["ws"].numel() < needb:
                            loc["ws"] = None
                            loc["ws"] = torch.empty(needb, device=dev, dtype=torch.uint8)
                        gp = g[b0:b1, s0:s1, s.prb_ids].contiguous() if s.n_p else None
                        rp = raw[b0:b1, s0:s1, s.prb_ids].contiguous() if s.n_p else None
                        gx = torch.empty((nb, s1 - s0), device=dev, dtype=torch.float32) if need[0] else None
                        hist = loc["hist"]
                        with torch.cuda.device(dev):
                            st = lib.wt_slab_backward(ctypes.byref(pb), ctypes.byref(s.desc_a), _lib.ptr(loc["c"]),
                                                      _lib.ptr(loc["b"]), _lib.ptr(loc["rho"]), _lib.ptr(s.src_own),
                                                      _lib.ptr(s.prb_ext), _lib.ptr(s.sq), _lib.ptr(gp), _lib.ptr(rp),
                                                      _lib.ptr(hist), hist.numel(), _lib.ptr(s.a1), _lib.ptr(s.a2),
                                                      _lib.ptr(tc[i]), _lib.ptr(gx), _lib.ptr(loc["ws"]),
                                                      loc["ws"].numel(), _lib.stream_ptr(dev))
                        _lib.check(st, "wt_slab_backward")
                        _lib.count_launches(plan_b.launches_bwd + (s1 - s0 + s.halo - 1) // s.halo)
                        gown[i] += tc[i][s.up:s.up + (s.r1 - s.r0)]
                        if gx is not None:
                            gxs[i][b0:b1, s0:s1] = gx
        _DomainLoop._join(cx)
        grad_x = torch.stack(gxs).sum(0) if need[0] else None
        for loc in L:
            loc["hist"] = None       # the tape of the last segment is not kept between iterations
        # assemble dLoss/dc: the slabs are disjoint -> gather, no reduction
        if virtual:
            grad_c = torch.cat(gown, dim=0)
        else:
            grad_c = gather_row_slabs(gown[0], NX, group)
            if grad_x is not None:
                dist.all_reduce(grad_x, group=group)
        ctx.meta = None
        xd, cd, bd, rd = dtypes
        grad_rho = torch.zeros((NX, Ny), device=dev, dtype=torch.float32) if need[3] else None
        return (grad_x.to(xd) if need[0] else None, grad_c.to(cd) if need[1] else None, None,
                grad_rho.to(rd) if need[3] else None) + (None,) * 7


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class DomainDecomposedWaveRNN(torch.nn.Module):
    """Wraps a WaveRNN: every rank passes the SAME waveforms x and receives the same probe outputs; the grid rows are
    split over the ranks of `group` (or over `virtual_ranks` slabs inside this process).

    halo              ghost rows per interior side = steps between exchanges (multiple of 8)
    checkpoint_every  steps between (u_t, u_{t-1}) checkpoints, rounded up to a multiple of halo (None: the wrapped model's
                      `checkpoint_every`, 0: no checkpoints -- one tape for the whole sequence)
    batch_chunk       waveforms processed at a time (None: the wrapped model's `batch_chunk`, 0: all)
    """

    def __init__(self, model, halo=16, group=None, virtual_ranks=0, checkpoint_every=None, batch_chunk=None):
        super().__init__()
        self.model, self.halo, self.group, self.virtual_ranks = model, int(halo), group, int(virtual_ranks)
        self.checkpoint_every, self.batch_chunk = checkpoint_every, batch_chunk
        self._cx = None
        self._loc = None

    # ---- cached per-shape state -------------------------------------------------------------
    def context(self, Nx, Ny, Bc, halo, spec, dev, group, virtual):
        key = (Nx, Ny, Bc, halo, str(dev), virtual, spec.src_ij.data_ptr(), spec.prb_ij.data_ptr())
        if self._cx is None or self._cx.key != key:
            self._cx, self._loc = None, None
            self._cx = SlabContext(Nx, Ny, Bc, halo, spec, dev, group, virtual)
        return self._cx

    def local(self, cx, c32, b32, rho32):
        """Per-slab coefficient rows (re-sliced every call: c changes with every optimiser step) + reusable scratch."""
        if self._loc is None:
            self._loc = [dict(ws=None, hist=None) for _ in cx.slabs]
        for s, loc in zip(cx.slabs, self._loc):
            sl = slice(s.e0, s.e1)
            loc["c"], loc["b"] = c32[sl].contiguous(), b32[sl].contiguous()
            loc["rho"] = rho32[sl].contiguous() if rho32 is not None else None
        return self._loc

    @property
    def exchange_bytes(self):
        """NVLink bytes one exchange sends to ONE neighbour (model: 2 fields x halo rows x Ny x batch x 4 B)."""
        return self._cx.exchange_bytes if self._cx is not None else None

    def forward(self, x):
        m = self.model
        geom = m.cell.geom
        c, b, rho = geom.c, geom.b, geom.rho
        if not c.is_cuda:
            raise RuntimeError("wavetorch_b200: the model is on %s. The wave-RNN hot path has no CPU fallback." % c.device)
        tab = m._pixel_tables(c.device)
        if not tab["scalar_probes"]:
            raise NotImplementedError("domain decomposition supports scalar-coordinate probes only")
        s = m.cell.host_scalars()
        if getattr(geom, "_h_host", None) is None:
            geom._h_host = float(geom.h)
        spec = LoopSpec(src_ij=tab["src_ij"], prb_ij=tab["prb_ij"], prb_sq=tab["prb_sq"], dt=s["dt"], h=geom._h_host,
                        b0=s["b0"], uth=s["uth"], c_nl=s["c_nl"], track_grad=torch.is_grad_enabled())
        S = self.checkpoint_every if self.checkpoint_every is not None else getattr(m, "checkpoint_every", 0)
        bc = self.batch_chunk if self.batch_chunk is not None else getattr(m, "batch_chunk", 0)
        return _DomainLoop.apply(x, c, b, rho, spec, self.halo, int(S or 0), int(bc or 0), self.group,
                                 self.virtual_ranks, self)
