"""CUDA-graph capture of a whole training iteration.

The fused kernels make one iteration of wavetorch/train.py:59-72 a few milliseconds long; what is left on the critical
path of a training loop is host launch latency (geometry ops, loss head, optimizer: ~60 small launches).  Capturing
`zero_grad -> model(x) -> loss -> backward -> optimizer.step -> constrain_to_design_region` once and replaying it removes
that latency: CUDA graphs instead of a tracing compiler.  The wt_* C entry points only enqueue work on the current stream
and allocate nothing, so they are capturable as they are.
"""
import torch


class GraphedTrainStep:
    """Capture `loss = loss_fn(model(x), y); loss.backward(); optimizer.step(); constrain` for fixed-shape (x, y).

    optimizer must be created with capturable=True (torch.optim.Adam(..., capturable=True)).
    Call the object with new inputs (host-pinned or device tensors); it returns the device tensor holding the loss
    of the replayed iteration.
    """

    def __init__(self, model, optimizer, loss_fn, x_example, y_example, warmup=3, post_step=None):
        self.model, self.optimizer, self.loss_fn = model, optimizer, loss_fn
        self.post_step = post_step if post_step is not None else self._default_post_step
        dev = next(model.parameters()).device
        self.x = torch.empty_like(x_example, device=dev)
        self.y = torch.empty_like(y_example, device=dev)
        self.x.copy_(x_example)
        self.y.copy_(y_example)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._iteration()
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        self.optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph):
            self.loss = self._iteration()

    def _default_post_step(self):
        inner = getattr(self.model, "model", self.model)      # unwrap BatchShardedWaveRNN
        inner.cell.geom.constrain_to_design_region()

    def _iteration(self):
        self.optimizer.zero_grad(set_to_none=True)
        loss = self.loss_fn(self.model(self.x), self.y)
        loss.backward()
        self.optimizer.step()
        self.post_step()
        return loss

    def __call__(self, x, y):
        self.x.copy_(x, non_blocking=True)
        self.y.copy_(y, non_blocking=True)
        self.graph.replay()
        return self.loss
