"""CPU oracle for ``dynsight.analysis.spatialaverage`` -- TEST INFRASTRUCTURE ONLY.

Restates ``_internal/analysis/spatial_average.py:31-63`` (``processframe``): per frame, all pairs of
selected particles within ``cutoff`` under the periodic minimum image (MDAnalysis ``FastNS``:
coordinates folded into the box, pairs with ``d <= cutoff``), then for every particle
``np.mean`` of the descriptor over its neighbours and itself.

FastNS is a third-party dependency (MDAnalysis, not installable here); the pair search is restated
with ``scipy.spatial.cKDTree(boxsize=...)`` on the folded float64 copies of the float32 positions.
Pinned by the reference's own golden ``tests/analysis/spavg/test_spavg.npy`` (max abs difference
0.0 on its 2048 x 6 values, see ``tests/golden/make_golden.py``); pairs at *exactly* the cutoff and
the summation order inside ``np.mean`` are not pinned by that test (tolerance 1e-12 relative).
"""

from __future__ import annotations

import numpy as np
from scipy.spatial import cKDTree


def fold(pos: np.ndarray, box: np.ndarray) -> np.ndarray:
    """float32 positions folded into [0, box) the way the device kernel does, as float64."""
    p = np.asarray(pos)
    b = np.asarray(box, dtype=p.dtype)
    w = p - np.floor(p / b) * b
    w = np.where(w >= b, w - b, w)
    w = np.where(w < 0, 0, w)
    return w.astype(np.float64)


def spatial_average(pos: np.ndarray, dims: np.ndarray, values: np.ndarray, r_cut: float
                    ) -> np.ndarray:  # fmt: skip
    """pos (F, N, 3), dims (F, >=3), values (N, F) or (N, F, C) -> same shape as values."""
    out = np.zeros_like(values, dtype=np.float64)
    for f in range(pos.shape[0]):
        box = np.asarray(dims[f, :3], dtype=np.float64)
        w = fold(pos[f], dims[f, :3].astype(pos.dtype))
        w = np.minimum(w, np.nextafter(box, 0))  # cKDTree wants coordinates strictly below the box
        tree = cKDTree(w, boxsize=box)
        for i, nb in enumerate(tree.query_ball_point(w, float(r_cut))):
            others = [j for j in nb if j != i]
            out[i, f] = np.mean(values[[*others, i], f], axis=0)
    return out
