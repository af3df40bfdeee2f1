"""Fused loss head of the classifier.

Drop-in for the two lines every training/eval pass of the reference runs on the loop output (train.py:61-62, :69-70,
:82, :99-101; utils.py:35-36):

    yb_pred = normalize_power(model(xb).sum(dim=1))
    loss = torch.nn.CrossEntropyLoss()(yb_pred, yb.argmax(dim=1))

`power_cross_entropy(out, labels)` returns `(loss, yb_pred)` from one kernel launch (`wt_loss_forward`); its backward is a
broadcast of a [B,P] table (`wt_loss_backward`) because dLoss/dout[b,t,p] does not depend on t.
"""
import torch

from . import _lib


class _PowerCrossEntropy(torch.autograd.Function):
    @staticmethod
    def forward(ctx, out, labels, batch_total):
        lib = _lib.load()
        if not out.is_cuda:
            raise RuntimeError("wavetorch_b200: the loss head runs on CUDA tensors only (no CPU fallback)")
        if out.dim() != 3:
            raise ValueError("power_cross_entropy expects the loop output [B, T, n_probes], got %s" % (tuple(out.shape),))
        B, T, P = out.shape
        o32 = out.detach().to(torch.float32).contiguous()
        lab = labels.to(device=out.device, dtype=torch.int64).contiguous()
        if lab.shape != (B,):
            raise ValueError("labels must be class indices of shape [B]")
        dev = out.device
        loss = torch.empty((), device=dev, dtype=torch.float32)
        y_pred = torch.empty((B, P), device=dev, dtype=torch.float32)
        dlds = torch.empty((B, P), device=dev, dtype=torch.float32)
        scratch = torch.empty((B,), device=dev, dtype=torch.float32)
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        with torch.cuda.device(dev):
            st = lib.wt_loss_forward(B, T, P, int(batch_total or 0), _lib.ptr(o32), _lib.ptr(lab), _lib.ptr(loss),
                                     _lib.ptr(y_pred), _lib.ptr(dlds), _lib.ptr(scratch), idx, _lib.stream_ptr(dev))
        _lib.check(st, "wt_loss_forward")
        _lib.count_launches(2)
        ctx.save_for_backward(dlds)
        ctx.shape, ctx.dtype = (B, T, P), out.dtype
        ctx.mark_non_differentiable(y_pred)
        return loss.to(out.dtype), y_pred.to(out.dtype)

    @staticmethod
    def backward(ctx, grad_loss, _grad_pred):
        lib = _lib.load()
        (dlds,) = ctx.saved_tensors
        B, T, P = ctx.shape
        dev = dlds.device
        g = grad_loss.detach().to(torch.float32).contiguous()
        grad = torch.empty((B, T, P), device=dev, dtype=torch.float32)
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        with torch.cuda.device(dev):
            st = lib.wt_loss_backward(B, T, P, _lib.ptr(dlds), _lib.ptr(g), _lib.ptr(grad), idx, _lib.stream_ptr(dev))
        _lib.check(st, "wt_loss_backward")
        _lib.count_launches(1)
        return grad.to(ctx.dtype), None, None


def power_cross_entropy(out, labels, batch_total=None):
    """(loss, yb_pred) for the loop output `out` [B,T,P] and class indices `labels` [B].

    batch_total: number of samples the mean runs over when `out` is one shard of a larger batch (default: B)."""
    return _PowerCrossEntropy.apply(out, labels, batch_total)
