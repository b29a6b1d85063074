"""CPU oracle for ``dynsight.analysis.self_time_correlation`` -- TEST INFRASTRUCTURE ONLY.

numpy restatement of ``_internal/analysis/time_correlations.py:60-92``: subtract each particle's
mean, ``c_i(t') = dot(x_i[:F-t'], x_i[t':])``, then per lag the mean of ``c_i`` over the particles
divided by ``F - t'`` and ``std(c_i) / ((F - t') sqrt(N))``, both normalised by the lag-0 mean.
``cross_time_correlation`` restates lines 143-178 the same way.
Pinned by the reference goldens ``tests/analysis/tcorr/{t_corr,std_dev,c_corr,ctd_dev}.npy`` (seeded random walk
of tests/analysis/test_time_correlations.py:26-37) and the doctest value
(time_correlations.py:57).
"""

from __future__ import annotations

import numpy as np


def self_time_correlation(data: np.ndarray, max_delay: int | None = None
                          ) -> tuple[np.ndarray, np.ndarray]:  # fmt: skip
    n_part, n_frames = data.shape
    x = data - np.mean(data, axis=1, keepdims=True)
    if max_delay is None:
        max_delay = n_frames
    corr = np.zeros(max_delay)
    err = np.zeros(max_delay)
    for lag in range(max_delay):
        valid = n_frames - lag
        c = np.einsum("it,it->i", x[:, :valid], x[:, lag : lag + valid])
        corr[lag] = np.mean(c) / valid
        err[lag] = np.std(c) / (valid * np.sqrt(n_part))
    norm = corr[0]
    return corr / norm, err / norm


def cross_time_correlation(data: np.ndarray, max_delay: int | None = None
                           ) -> tuple[np.ndarray, np.ndarray]:  # fmt: skip
    """time_correlations.py:143-178: all ordered pairs i != j, no lag-0 normalisation."""
    n_part, n_frames = data.shape
    x = data - np.mean(data, axis=1, keepdims=True)
    if max_delay is None:
        max_delay = n_frames
    corr = np.zeros(max_delay)
    err = np.zeros(max_delay)
    off = ~np.eye(n_part, dtype=bool)
    for lag in range(max_delay):
        valid = n_frames - lag
        c = (x[:, :valid] @ x[:, lag : lag + valid].T)[off]
        corr[lag] = np.mean(c) / valid
        err[lag] = np.std(c) / (valid * np.sqrt(n_part * (n_part - 1)))
    return corr, err


def reference_walk() -> np.ndarray:
    """The input of tests/analysis/test_time_correlations.py:20-37."""
    rng = np.random.default_rng(12345)
    n_part, n_frames = 5, 100
    walk = np.zeros((n_part, n_frames), dtype=np.float64)
    x_pos = 0.0
    for i in range(n_part):
        for j in range(n_frames):
            x_pos += rng.random()
            walk[i, j] = x_pos
    return walk
