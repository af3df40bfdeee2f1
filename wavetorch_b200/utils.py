"""Small helpers with the names the reference exports from wavetorch/utils.py (utils.py:6-36)."""
import numpy as np
import torch


def to_tensor(x, dtype=None):
    """Convert numbers / sequences / ndarrays to a tensor of `dtype` (default: torch default dtype)."""
    if dtype is None:
        dtype = torch.get_default_dtype()
    if isinstance(x, torch.Tensor):
        return x.detach().clone().to(dtype)
    if isinstance(x, np.ndarray):
        return torch.from_numpy(np.ascontiguousarray(x)).to(dtype)
    return torch.tensor(x, dtype=dtype)


def set_dtype(dtype=None):
    """Select the global default dtype, 'float32' (default) or 'float64' (utils.py:14-20).

    The CUDA time loop always computes in float32 (BASELINE north star); with a float64 default the
    geometry parameterisation runs in float64 and the fields are cast at the kernel boundary.
    """
    table = {None: torch.float32, "float32": torch.float32, "float64": torch.float64}
    if dtype not in table:
        raise ValueError("Unsupported data type: %s; should be either float32 or float64" % dtype)
    torch.set_default_dtype(table[dtype])


def window_data(X, window_length):
    """Centre crop of a 1-D sample to `window_length` points (utils.py:23-26)."""
    mid = len(X) / 2
    return X[int(mid - window_length / 2):int(mid + window_length / 2)]


def accuracy_onehot(y_pred, y_label):
    """Fraction of rows whose arg-max equals the integer label (utils.py:29-32)."""
    return (y_pred.argmax(dim=1) == y_label).float().mean().item()


def normalize_power(X):
    """Divide every row by its sum over the probe axis (utils.py:35-36)."""
    return X / X.sum(dim=1, keepdim=True)
