"""Checkpoint format of the reference, unchanged (wavetorch/io.py:13-87), so that `.pt` files written by either
implementation load in the other: keys `model_geom_class_str`, `model_state`, `history`, `history_geom_state`, `cfg`.
"""
import copy
import os

import torch

from . import geom
from .cell import WaveCell
from .probe import WaveIntensityProbe
from .rnn import WaveRNN
from .source import WaveSource
from .utils import set_dtype


def save_model(model, name, savedir='./study/', history=None, history_geom_state=None, cfg=None, verbose=True):
    """Save the model state and history to `savedir + name + '.pt'` (io.py:13-41; `savedir` is a prefix, as there)."""
    inner = getattr(model, "model", model)      # unwrap BatchShardedWaveRNN / DomainDecomposedWaveRNN
    str_filename = name + '.pt'
    if not os.path.exists(savedir):
        os.makedirs(savedir)
    str_savepath = savedir + str_filename
    if history_geom_state is None:
        history_geom_state = [inner.cell.geom.state_reconstruction_args()]
    data = {'model_geom_class_str': inner.cell.geom.__class__.__name__,
            'model_state': {k: v.detach().cpu() for k, v in inner.state_dict().items()},
            'history': history,
            'history_geom_state': history_geom_state,
            'cfg': cfg}
    if verbose:
        print("Saving model to %s" % str_savepath)
    torch.save(data, str_savepath)


def new_geometry(class_str, state):
    cls = getattr(geom, class_str)
    return cls(**copy.deepcopy(state))


def load_model(str_filename, which_iteration=-1, verbose=True):
    """Rebuild (model, history, history_geom_state, cfg) from a checkpoint (io.py:50-87).  Like the reference, probes come
    back as WaveIntensityProbe and sources as WaveSource; the model is returned on the CPU in eval mode -- move it with
    `.to('cuda')` before calling it."""
    if verbose:
        print("Loading model from %s" % str_filename)
    data = torch.load(str_filename, map_location="cpu", weights_only=False)
    set_dtype(data['cfg']['dtype'])
    new_geom = new_geometry(data['model_geom_class_str'], data['history_geom_state'][which_iteration])
    model_state = copy.deepcopy(data['model_state'])
    px = [model_state[k].item() for k in model_state if 'probes' in k and 'x' in k]
    py = [model_state[k].item() for k in model_state if 'probes' in k and 'y' in k]
    sx = [model_state[k].item() for k in model_state if 'sources' in k and 'x' in k]
    sy = [model_state[k].item() for k in model_state if 'sources' in k and 'y' in k]
    new_probes = [WaveIntensityProbe(x, y) for (x, y) in zip(px, py)]
    new_sources = [WaveSource(x, y) for (x, y) in zip(sx, sy)]
    new_cell = WaveCell(model_state['cell.dt'].item(), new_geom)
    new_model = WaveRNN(new_cell, new_sources, new_probes)
    new_model.eval()
    return new_model, data['history'], data['history_geom_state'], data['cfg']
