"""Row-slab domain decomposition of ONE wave simulation across the GPUs of a box (BASELINE config 5).

The reference has nothing like this; it is the multi-GPU mode for grids that do not fit one GPU.
Rank g owns rows [r0, r1) of the [Nx, Ny] grid plus `halo` ghost rows on each interior side.  Time is advanced in
segments of `halo` steps: inside a segment every rank runs the ordinary streaming kernels (wt_forward) on its extended
slab as if it were a whole domain -- the error made at the artificial slab edges travels inwards one row per step, so
after `halo` steps the owned rows are still exact -- then neighbours swap `halo` rows of both time levels
("halo depth = temporal block").  The adjoint does the same in reverse with wt_backward's
adj1/adj2 chaining, re-running each segment's forward from its checkpoint to rebuild the tape.

    exchange:   NCCL send/recv of [B, halo, Ny] row blocks between neighbours (torch.distributed P2P)
    probes:     read by the owning rank, summed over ranks ([B,T,P] all-reduce)
    gradient:   every rank produces dLoss/dc for its owned rows; one all-reduce assembles the full field

`virtual_ranks=n` runs all n slabs inside one process on one GPU, exchanging by plain copies: the same code path, used
by the single-GPU tests of the decomposition.
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib
from .distributed import shard_bounds
from .functional import LoopSpec, _call_forward, _dev_index, _f32, _require_cuda


class _Slab:
    """Geometry of one rank's extended slab and its share of the sources / probes."""

    def __init__(self, rank, world, Nx, halo, src_ij, prb_ij, dev):
        self.rank, self.world = rank, world
        self.r0, self.r1 = shard_bounds(Nx, world, rank)
        self.e0, self.e1 = max(0, self.r0 - halo), min(Nx, self.r1 + halo)
        self.up, self.dn = self.r0 - self.e0, self.e1 - self.r1      # ghost rows above / below
        src, prb = src_ij.cpu(), prb_ij.cpu()
        in_ext = lambda t: (t[:, 0] >= self.e0) & (t[:, 0] < self.e1)
        in_own = lambda t: (t[:, 0] >= self.r0) & (t[:, 0] < self.r1)
        off = torch.tensor([[self.e0, 0]], dtype=torch.int32)
        self.src_ext = (src[in_ext(src)] - off).contiguous().to(dev)          # injected in forward (ghosts included)
        self.src_own = (src[in_own(src)] - off).contiguous().to(dev)          # gathered for dLoss/dx
        self.prb_ids = torch.nonzero(in_ext(prb)).flatten()                   # probes inside the extended slab
        self.prb_ext = (prb[self.prb_ids] - off).contiguous().to(dev)
        self.prb_owned = in_own(prb[self.prb_ids]).to(dev)                    # which of them this rank reports
        self.prb_ids = self.prb_ids.to(dev)


class _DomainLoop(torch.autograd.Function):
    @staticmethod
    def _problem(spec, slab, Ny, B, T, dev, zero_init, n_src, n_prb):
        flags = (spec.flags | _lib.WT_F_FORCE_STREAM) & ~_lib.WT_F_ZERO_INIT
        if zero_init:
            flags |= _lib.WT_F_ZERO_INIT
        return _lib.make_problem(slab.e1 - slab.e0, Ny, B, T, n_src, n_prb, spec.dt, spec.h, spec.b0, spec.uth, spec.c_nl,
                                 flags, _dev_index(dev))

    @staticmethod
    def _exchange(slabs, fields, group, virtual):
        """fields[i] = list of [B, rows_i, Ny] tensors of slab i (same length for all).  Fill every ghost block with the
        neighbour's owned rows."""
        if virtual:
            for i, s in enumerate(slabs):
                own = s.r1 - s.r0
                for fi, f in enumerate(fields[i]):
                    if s.up:      # my upper ghost rows <- last rows owned by the slab above
                        a, fa = slabs[i - 1], fields[i - 1][fi]
                        f[:, :s.up] = fa[:, a.up + (a.r1 - a.r0) - s.up: a.up + (a.r1 - a.r0)]
                    if s.dn:      # my lower ghost rows <- first rows owned by the slab below
                        a, fa = slabs[i + 1], fields[i + 1][fi]
                        f[:, s.up + own:] = fa[:, a.up: a.up + s.dn]
            return
        s, fs = slabs[0], fields[0]
        own0, own1 = s.up, s.up + (s.r1 - s.r0)
        ops, recvs = [], []
        for f in fs:
            if s.rank > 0:
                ops.append(dist.P2POp(dist.isend, f[:, own0:own0 + s.up].contiguous(), s.rank - 1, group))
                buf = torch.empty_like(f[:, :s.up])
                buf = buf.contiguous()
                ops.append(dist.P2POp(dist.irecv, buf, s.rank - 1, group))
                recvs.append((f, slice(0, s.up), buf))
            if s.rank + 1 < s.world:
                ops.append(dist.P2POp(dist.isend, f[:, own1 - s.dn:own1].contiguous(), s.rank + 1, group))
                buf = torch.empty_like(f[:, own1:]).contiguous()
                ops.append(dist.P2POp(dist.irecv, buf, s.rank + 1, group))
                recvs.append((f, slice(own1, own1 + s.dn), buf))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            for f, sl, buf in recvs:
                f[:, sl] = buf

    @staticmethod
    def forward(ctx, x, c, b, rho, spec, halo, group, virtual):
        lib = _lib.load()
        _require_cuda(x, "the input waveform x")
        _require_cuda(c, "the wave speed c")
        dev = x.device
        x32, c32, b32, rho32 = _f32(x, "x"), _f32(c, "c"), _f32(b, "b"), _f32(rho, "rho")
        B, T = x32.shape
        Nx, Ny = c32.shape
        P = spec.prb_ij.shape[0]
        world = virtual if virtual else dist.get_world_size(group)
        ranks = list(range(world)) if virtual else [dist.get_rank(group)]
        slabs = [_Slab(r, world, Nx, halo, spec.src_ij, spec.prb_ij, dev) for r in ranks]
        if min(s.r1 - s.r0 for s in slabs) < halo:
            raise ValueError("domain decomposition: every rank needs at least `halo` (= %d) rows" % halo)
        need = ctx.needs_input_grad
        nonlinear = spec.b0 > 0 or spec.c_nl != 0
        want_grad = spec.track_grad and T > 0 and (need[0] or need[1] or need[2] or (need[3] and nonlinear))
        segs = [(s0, min(s0 + halo, T)) for s0 in range(0, T, halo)]
        loc = []   # per slab: local coefficient slices and state
        for s in slabs:
            sl = slice(s.e0, s.e1)
            rows = s.e1 - s.e0
            loc.append(dict(c=c32[sl].contiguous(), b=b32[sl].contiguous(),
                            rho=rho32[sl].contiguous() if rho32 is not None else None,
                            u1=torch.empty((B, rows, Ny), device=dev, dtype=torch.float32),
                            u2=torch.empty((B, rows, Ny), device=dev, dtype=torch.float32),
                            sq=spec.prb_sq[s.prb_ids].contiguous(), ck=[]))
        out = torch.zeros((B, T, P), device=dev, dtype=torch.float32)
        raw = torch.zeros((B, T, P), device=dev, dtype=torch.float32)
        for k, (s0, s1) in enumerate(segs):
            xs = x32[:, s0:s1].contiguous()
            for s, L in zip(slabs, loc):
                if want_grad and k > 0:
                    L["ck"].append((L["u1"].clone(), L["u2"].clone()))
                n_p = s.prb_ext.shape[0]
                prob = _DomainLoop._problem(spec, s, Ny, B, s1 - s0, dev, k == 0, s.src_ext.shape[0], n_p)
                plan = _lib.query_plan(prob)
                ws = torch.empty(max(int(plan.workspace_fwd_bytes), 16), device=dev, dtype=torch.uint8)
                po = torch.empty((B, s1 - s0, max(n_p, 1)), device=dev, dtype=torch.float32)
                pr = torch.empty_like(po)
                sub = LoopSpec(src_ij=s.src_ext, prb_ij=s.prb_ext, prb_sq=L["sq"], dt=spec.dt, h=spec.h)
                _call_forward(lib, prob, dev, L["c"], L["b"], L["rho"], xs, sub, L["u1"], L["u2"], po if n_p else None,
                              pr if n_p else None, None, None, ws)
                _lib.count_launches(plan.launches_fwd)
                if n_p:
                    ids = s.prb_ids[s.prb_owned]
                    out[:, s0:s1, ids] = po[:, :, s.prb_owned]
                    raw[:, s0:s1, ids] = pr[:, :, s.prb_owned]
            _DomainLoop._exchange(slabs, [[L["u1"], L["u2"]] for L in loc], group, virtual)
        if not virtual:
            dist.all_reduce(out, group=group)
            if want_grad:
                dist.all_reduce(raw, group=group)
        ctx.no_tape = not want_grad
        ctx.shape = (Nx, Ny)
        if want_grad:
            ctx.meta = (spec, halo, group, virtual, slabs, loc, segs, raw, x32, Nx,
                        (x.dtype, c.dtype, b.dtype, rho.dtype if rho is not None else None))
        return out.to(x.dtype)

    @staticmethod
    def backward(ctx, grad_out):
        if ctx.no_tape:
            need = ctx.needs_input_grad
            zero = torch.zeros(ctx.shape, device=grad_out.device, dtype=grad_out.dtype) if need[3] else None
            return None, None, None, zero, None, None, None, None
        lib = _lib.load()
        spec, halo, group, virtual, slabs, loc, segs, raw, x32, NX, dtypes = ctx.meta
        dev = x32.device
        B, T = x32.shape
        Ny = loc[0]["c"].shape[1]
        need = ctx.needs_input_grad
        nonlinear = spec.b0 > 0 or spec.c_nl != 0
        g = grad_out.detach().to(torch.float32)
        # full-size accumulators (only the owned rows of each slab are filled, then summed over ranks)
        grad_c = torch.zeros((NX, Ny), device=dev, dtype=torch.float32)
        grad_rho = torch.zeros((NX, Ny), device=dev, dtype=torch.float32) if (need[3] and nonlinear) else None
        grad_x = torch.zeros((B, T), device=dev, dtype=torch.float32) if need[0] else None
        adj = [[torch.zeros_like(L["u1"]), torch.zeros_like(L["u1"])] for L in loc]
        for k in range(len(segs) - 1, -1, -1):
            s0, s1 = segs[k]
            xs = x32[:, s0:s1].contiguous()
            for s, L, A in zip(slabs, loc, adj):
                rows = s.e1 - s.e0
                n_p = s.prb_ext.shape[0]
                if k > 0:
                    u1, u2 = (t.clone() for t in L["ck"][k - 1])
                else:
                    u1 = torch.empty((B, rows, Ny), device=dev, dtype=torch.float32)
                    u2 = torch.empty_like(u1)
                # forward of the segment with a tape (ghost sources included)
                pf = _DomainLoop._problem(spec, s, Ny, B, s1 - s0, dev, k == 0, s.src_ext.shape[0], n_p)
                plan_f = _lib.query_plan(pf)
                pb = _DomainLoop._problem(spec, s, Ny, B, s1 - s0, dev, k == 0, s.src_own.shape[0], n_p)
                plan_b = _lib.query_plan(pb)
                ws = torch.empty(max(int(plan_f.workspace_fwd_bytes), int(plan_b.workspace_bwd_bytes), 16), device=dev,
                                 dtype=torch.uint8)
                hist = torch.empty(max(int(plan_f.history_bytes), 16), device=dev, dtype=torch.uint8)
                po = torch.empty((B, s1 - s0, max(n_p, 1)), device=dev, dtype=torch.float32)
                sub_f = LoopSpec(src_ij=s.src_ext, prb_ij=s.prb_ext, prb_sq=L["sq"], dt=spec.dt, h=spec.h)
                _call_forward(lib, pf, dev, L["c"], L["b"], L["rho"], xs, sub_f, u1, u2, po if n_p else None,
                              po.clone() if n_p else None, None, hist, ws)
                # adjoint of the segment: seeds for every probe inside the extended slab, with the EXACT raw samples
                gp = g[:, s0:s1, s.prb_ids].contiguous() if n_p else None
                rp = raw[:, s0:s1, s.prb_ids].contiguous() if n_p else None
                tc = torch.empty((rows, Ny), device=dev, dtype=torch.float32)
                tr = torch.empty_like(tc) if grad_rho is not None else None
                gx = torch.empty((B, s1 - s0), device=dev, dtype=torch.float32) if need[0] else None
                with torch.cuda.device(dev):
                    st = lib.wt_backward(ctypes.byref(pb), _lib.ptr(L["c"]), _lib.ptr(L["b"]), _lib.ptr(L["rho"]),
                                         _lib.ptr(s.src_own), _lib.ptr(s.prb_ext), _lib.ptr(L["sq"]), _lib.ptr(gp),
                                         _lib.ptr(rp), None, _lib.ptr(hist), hist.numel(), _lib.ptr(A[0]), _lib.ptr(A[1]),
                                         _lib.ptr(tc), None, _lib.ptr(tr), _lib.ptr(gx), _lib.ptr(ws), ws.numel(),
                                         _lib.stream_ptr(dev))
                _lib.check(st, "wt_backward")
                _lib.count_launches(plan_f.launches_fwd + plan_b.launches_bwd)
                own = slice(s.up, s.up + (s.r1 - s.r0))
                grad_c[s.r0:s.r1] += tc[own]
                if grad_rho is not None:
                    grad_rho[s.r0:s.r1] += tr[own]
                if gx is not None:
                    grad_x[:, s0:s1] += gx
                del hist, u1, u2
            _DomainLoop._exchange(slabs, adj, group, virtual)
        if not virtual:
            flat = [grad_c] + ([grad_rho] if grad_rho is not None else []) + ([grad_x] if grad_x is not None else [])
            buf = torch.cat([t.reshape(-1) for t in flat])
            dist.all_reduce(buf, group=group)
            o = 0
            for t in flat:
                t.copy_(buf[o:o + t.numel()].view_as(t))
                o += t.numel()
        ctx.meta = None
        xd, cd, bd, rd = dtypes
        if need[3] and grad_rho is None:
            grad_rho = torch.zeros((NX, Ny), device=dev, dtype=torch.float32)
        return (grad_x.to(xd) if need[0] else None, grad_c.to(cd) if need[1] else None, None,
                grad_rho.to(rd) if need[3] else None, None, None, None, None)


class DomainDecomposedWaveRNN(torch.nn.Module):
    """Wraps a WaveRNN: every rank passes the SAME waveforms x and receives the same probe outputs; the grid rows are
    split over the ranks of `group` (or over `virtual_ranks` slabs inside this process)."""

    def __init__(self, model, halo=16, group=None, virtual_ranks=0):
        super().__init__()
        self.model, self.halo, self.group, self.virtual_ranks = model, int(halo), group, int(virtual_ranks)

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
        return _DomainLoop.apply(x, c, b, rho, spec, self.halo, self.group, self.virtual_ranks)
