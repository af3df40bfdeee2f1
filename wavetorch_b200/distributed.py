"""Batch sharding across the GPUs of one box (one process per GPU, torch.distributed over NCCL/NVLink).

The reference has no distributed code.  Waveforms are independent (rnn.py:36-41: per-sample
state, shared geometry), so the batch is split contiguously over ranks with the geometry replicated and the
only collective is ONE all-reduce of the gradient that leaves the time loop -- dLoss/dc, plus the direct
dLoss/drho of the nonlinear terms -- issued on the adjoint's stream right after wt_backward.  Reducing before
the geometry chain (blur/projection backward) gives the same rho.grad on every rank as reducing after it,
because that chain is linear in the incoming gradient.
"""
import torch
import torch.distributed as dist


def shard_bounds(n, world_size, rank):
    """Contiguous [lo, hi) slice of n samples owned by `rank`; the first n % world_size ranks get one more."""
    base, extra = divmod(int(n), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(x, group=None, dim=0):
    """This rank's contiguous slice of a batch-first tensor."""
    ws, rk = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_bounds(x.shape[dim], ws, rk)
    return x.narrow(dim, lo, hi - lo)


class _SumGradAcrossRanks(torch.autograd.Function):
    """Identity in forward; all-reduce(sum) of the incoming gradients in backward, as one fused buffer."""

    @staticmethod
    def forward(ctx, group, scale, reducer, *tensors):
        ctx.group, ctx.scale, ctx.reducer = group, scale, reducer
        return tuple(t.view_as(t) for t in tensors)

    @staticmethod
    def backward(ctx, *grads):
        live = [g for g in grads if g is not None]
        if live:
            # float32 (the product path) goes through the peer-memory kernel; a float64 model (set_dtype('float64')) keeps its
            # precision and takes the NCCL all-reduce
            dt = torch.float64 if any(g.dtype == torch.float64 for g in live) else torch.float32
            flat = torch.cat([g.reshape(-1).to(dt) for g in live]) if len(live) > 1 else \
                live[0].reshape(-1).to(dt).contiguous()
            if ctx.reducer is not None and flat.is_cuda and dt == torch.float32 and flat.numel() <= ctx.reducer.capacity:
                flat = ctx.reducer.all_reduce(flat, ctx.scale)       # one kernel over NVLink peer memory (csrc/wt_peer.cu)
            else:
                if live[0].reshape(-1).data_ptr() == flat.data_ptr():
                    flat = flat.clone()                              # never reduce into autograd's own buffer
                dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=ctx.group)
                if ctx.scale != 1.0:
                    flat.mul_(ctx.scale)
            out, o = [], 0
            for g in grads:
                if g is None:
                    out.append(None)
                else:
                    n = g.numel()
                    out.append(flat[o:o + n].view_as(g).to(g.dtype))
                    o += n
        else:
            out = list(grads)
        return (None, None, None) + tuple(out)


def sync_grads(*tensors, group=None, average=False, reducer=None):
    """Return views of `tensors` whose gradients are summed (or averaged) over all ranks in backward.

    reducer: a wavetorch_b200.peer.PeerGradReducer to run the reduction as one peer-memory kernel instead of NCCL."""
    scale = 1.0 / dist.get_world_size(group) if average else 1.0
    return _SumGradAcrossRanks.apply(group, scale, reducer, *tensors)


class BatchShardedWaveRNN(torch.nn.Module):
    """Wraps a WaveRNN: forward takes this rank's shard of the waveforms; in backward the loop gradients
    (dLoss/dc and the direct dLoss/drho) are all-reduced once, so `rho.grad` is the gradient of the SUM of the
    per-rank losses (pass average=True for the mean) on every rank."""

    def __init__(self, model, group=None, average=False, peer_reduce="auto"):
        """peer_reduce: "auto" = reduce the gradient with the NVLink peer-memory kernel when the model lives on CUDA and
        the symmetric-memory rendezvous succeeds, else NCCL/gloo all-reduce; True = require it; False = never."""
        super().__init__()
        self.model = model
        self.group = group
        self.average = average
        self.reducer = None
        self.reduce_mode = "all_reduce"
        dev = next(model.parameters()).device
        if peer_reduce and dev.type == "cuda" and dist.get_backend(group) == "nccl":
            try:
                from .peer import PeerGradReducer
                Nx, Ny = model.cell.geom.domain_shape
                self.reducer = PeerGradReducer(2 * int(Nx) * int(Ny), dev, group)
                self.reduce_mode = "peer-kernel"
            except Exception as exc:
                if peer_reduce is True:
                    raise
                import sys
                sys.stderr.write("wavetorch_b200: peer-memory all-reduce unavailable (%r); using NCCL\n" % (exc,))
        # every rank must take the same branch
        if dist.is_initialized() and dist.get_world_size(group) > 1 and dev.type == "cuda" and dist.get_backend(group) == "nccl":
            ok = torch.tensor([1 if self.reducer is not None else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 0:
                self.reducer, self.reduce_mode = None, "all_reduce"

    def forward(self, x_local, output_fields=False):
        m = self.model
        geom = m.cell.geom
        c, rho, b = geom.c, geom.rho, geom.b
        nonlinear = m.cell.host_scalars()["b0"] > 0 or m.cell.host_scalars()["c_nl"] != 0
        if torch.is_grad_enabled() and (c.requires_grad or (nonlinear and rho.requires_grad)):
            if nonlinear:
                c, rho = sync_grads(c, rho, group=self.group, average=self.average, reducer=self.reducer)
            else:
                (c,) = sync_grads(c, group=self.group, average=self.average, reducer=self.reducer)
        return m._run(x_local, c, b, rho, output_fields)

    def gather_outputs(self, y_local):
        """All-gather the per-rank outputs along the batch axis (equal shard sizes required)."""
        ws = dist.get_world_size(self.group)
        parts = [torch.empty_like(y_local) for _ in range(ws)]
        dist.all_gather(parts, y_local.contiguous(), group=self.group)
        return torch.cat(parts, dim=0)
