"""Deterministic synthetic inputs for benchmarks and demos.

The vowel recordings the reference trains on (wavetorch/data/vowels.py, librosa + a download) are not available offline;
these RNG-free stand-ins have the same shape and normalisation: three formant sinusoids per class under a Hann window,
unit energy per sample like data/vowels.py:12-17.
"""
import math

import numpy as np

_FORMANTS = np.array([[730.0, 1090.0, 2440.0], [270.0, 2290.0, 3010.0], [300.0, 870.0, 2240.0]])
_AMPS = np.array([1.0, 0.5, 0.25])


def synthetic_vowels(B, T, sr=10000.0, dtype=np.float32, first=0):
    """[B, T] waveforms; sample b belongs to class (first + b) % 3."""
    n = np.arange(T, dtype=np.float64)
    env = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / (T - 1))
    x = np.zeros((B, T), dtype=np.float64)
    for row in range(B):
        b = first + row
        jitter = 1.0 + 0.02 * (math.modf(0.7548776662466927 * (b + 1))[0] - 0.5)
        for j in range(3):
            phase = 2.0 * np.pi * math.modf(0.6180339887498949 * (3 * b + j + 1))[0]
            x[row] += _AMPS[j] * np.sin(2.0 * np.pi * _FORMANTS[b % 3, j] * jitter * n / sr + phase)
        x[row] *= env
        x[row] /= np.sqrt((x[row] ** 2).sum())
    return x.astype(dtype)
