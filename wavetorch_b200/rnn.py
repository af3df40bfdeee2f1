"""WaveRNN: the wave equation unrolled in time (API of wavetorch/rnn.py).

The reference's forward is a Python loop over T steps that calls the cell, every source and every probe per step
(rnn.py:50-67).  Here forward() gathers the static description of the problem once and hands the whole loop to
the fused CUDA kernels through functional.wave_rnn -- one launch for all B x T steps on small grids.
"""
import torch

from . import _lib
from .functional import LoopSpec, wave_rnn


class WaveRNN(torch.nn.Module):
    def __init__(self, cell, sources, probes=[]):
        super().__init__()
        self.cell = cell
        self.sources = torch.nn.ModuleList(sources if type(sources) is list else [sources])
        self.probes = torch.nn.ModuleList(probes if type(probes) is list else [probes])
        self._spec_cache = None
        # optional tuning knobs forwarded to the planner (0 = automatic)
        self.cluster = 0
        self.rows_per_thread = 0
        self.plan_flags = 0
        # memory/recompute trade-off for training long sequences or large grids (0 = keep the whole tape)
        self.checkpoint_every = 0
        self.batch_chunk = 0

    # ------------------------------------------------------------------
    def _pixel_tables(self, device):
        """int32 device tables of source / probe pixels; rebuilt only when a coordinate buffer changes."""
        bufs = [t for m in list(self.sources) + list(self.probes) for t in (m.x, m.y)]
        key = (str(device),) + tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in bufs)
        if self._spec_cache is not None and self._spec_cache[0] == key:
            return self._spec_cache[1]
        Nx, Ny = self.cell.geom.domain_shape

        def gather(mods):
            rows = [m.pixels()[0].cpu() for m in mods]
            cols = [m.pixels()[1].cpu() for m in mods]
            if not rows:
                return torch.zeros((0, 2), dtype=torch.int32), []
            ij = torch.stack([torch.cat(rows), torch.cat(cols)], dim=1)
            if ij.numel() and (ij.min() < 0 or (ij[:, 0] >= Nx).any() or (ij[:, 1] >= Ny).any()):
                raise IndexError("source/probe coordinate outside the %dx%d domain" % (Nx, Ny))
            return ij.to(torch.int32), [r.numel() for r in rows]

        src_ij, _ = gather(self.sources)
        prb_ij, prb_counts = gather(self.probes)
        prb_sq = torch.tensor([int(getattr(p, "squared", False)) for p, n in zip(self.probes, prb_counts)
                               for _ in range(n)], dtype=torch.int32)
        # the on-chip kernels add x at most WT_MAX_SRC_LISTINGS times per pixel (rnn.py:56-57 adds it once per listing)
        many = _lib.validate_pixels(Nx, Ny, src_ij, prb_ij) > _lib.WT_MAX_SRC_LISTINGS
        tables = dict(src_ij=src_ij.contiguous().to(device), prb_ij=prb_ij.contiguous().to(device),
                      prb_sq=prb_sq.to(device), prb_counts=prb_counts, force_stream=many,
                      scalar_probes=all(p.x.dim() == 0 for p in self.probes))
        self._spec_cache = (key, tables)
        return tables

    # ------------------------------------------------------------------
    def forward(self, x, output_fields=False, field_every=1):
        """Propagate for the length of the inputs.

        x : [B, T] input sequences (batch first).  Returns the probe time series [B, T, n_probes]
        (squared for WaveIntensityProbe), or all fields [B, T, Nx, Ny] when there are no probes or
        `output_fields` is set (rnn.py:21-72).

        field_every (extension): with output_fields, keep only every field_every-th field --
        [B, T // field_every, Nx, Ny], snapshot k being the field after step (k+1)*field_every - 1, i.e.
        `model(x, output_fields=True)[:, field_every-1::field_every]` without materialising the full history (what the
        reference's plotting code slices afterwards, plot.py:203-243).  Forward-only: use it under torch.no_grad().
        """
        geom = self.cell.geom
        # evaluated once per forward, like rnn.py:46-47
        return self._run(x, geom.c, geom.b, geom.rho, output_fields, field_every)

    def _run(self, x, c, b, rho, output_fields=False, field_every=1):
        if x.dim() != 2:
            raise ValueError("WaveRNN expects x of shape [batch, time], got %s" % (tuple(x.shape),))
        geom = self.cell.geom
        if not c.is_cuda:
            raise RuntimeError("wavetorch_b200: the model is on %s. The wave-RNN hot path has no CPU fallback; call "
                               "model.to('cuda') (the reference's CPU path is the oracle, not the product)." % c.device)
        if x.device != c.device:
            raise RuntimeError("wavetorch_b200: x is on %s but the model is on %s" % (x.device, c.device))
        tab = self._pixel_tables(c.device)
        fields = bool(output_fields) or len(self.probes) == 0
        s = self.cell.host_scalars()
        if getattr(geom, "_h_host", None) is None:
            geom._h_host = float(geom.h)
        flags = self.plan_flags | (_lib.WT_F_FORCE_STREAM if tab["force_stream"] else 0)
        spec = LoopSpec(src_ij=tab["src_ij"], prb_ij=tab["prb_ij"], prb_sq=tab["prb_sq"], dt=s["dt"], h=geom._h_host,
                        b0=s["b0"], uth=s["uth"], c_nl=s["c_nl"], output_fields=fields,
                        field_every=max(int(field_every), 1), flags=flags,
                        cluster=self.cluster, rows_per_thread=self.rows_per_thread,
                        checkpoint_every=self.checkpoint_every, batch_chunk=self.batch_chunk)
        y = wave_rnn(x, c, b, rho, spec)
        if fields or tab["scalar_probes"]:
            return y
        # multi-pixel probes: the reference stacks per-probe [B, n] readouts along a new last axis
        return torch.stack(torch.split(y, tab["prb_counts"], dim=-1), dim=-1)
