"""numpy restatement of the neighbour-list consumers psi / phi (TEST INFRASTRUCTURE ONLY).

Reference: ``/root/reference/src/dynsight/_internal/descriptors/misc.py``
  * ``orientational_order_param`` :15-95   (arithmetic :76-93)
  * ``compute_mean_alignment``    :98-133
  * ``velocity_alignment``        :136-219 (displacement branch :205-219)
They take the CSR neighbour lists of ``neighbor_list_celllist_centers`` (centres == selection ==
all atoms).  PARITY PIN: bit-compared with the reference module itself (imported by path in
tests/golden/make_golden.py) on seeded inputs and pinned by the reference's goldens
``tests/trajectory/files/{psi,phi}.npy`` (rtol 1e-3 / 1e-4 in the reference's own test).
"""

from __future__ import annotations

import numpy as np


def orientational_order_param(
    positions: np.ndarray, csr: list[tuple[np.ndarray, np.ndarray]], order: int = 6
) -> np.ndarray:
    """positions (F, N, 3) float32; csr[f] = (indptr, indices) -> psi (N, F) float64."""
    n_frames, n_atoms, _ = positions.shape
    psi = np.zeros((n_atoms, n_frames))
    for t in range(n_frames):
        frame = positions[t][:, :2].copy()
        indptr, indices = csr[t]
        for i in range(n_atoms):
            neighbors = indices[indptr[i] : indptr[i + 1]]
            if len(neighbors) <= 1:
                continue
            tmp = 0.0
            for j in neighbors:
                if j != i:
                    x, y = frame[j] - frame[i]
                    theta = np.mod(np.arctan2(y, x), 2 * np.pi)
                    tmp += np.exp(1j * order * theta)
            tmp /= len(neighbors)
            psi[i][t] = np.abs(tmp)
    return psi


def _cosine_similarity(a: np.ndarray, b: np.ndarray) -> float:
    """1 - scipy.spatial.distance.cosine(a, b) (the distance is clipped to [0, 2] by scipy)."""
    uv = np.dot(a, b)
    uu = np.dot(a, a)
    vv = np.dot(b, b)
    dist = 1.0 - uv / np.sqrt(uu * vv)
    dist = max(0.0, min(float(dist), 2.0))
    return float(1 - dist)


def velocity_alignment(
    positions: np.ndarray, csr: list[tuple[np.ndarray, np.ndarray]]
) -> np.ndarray:
    """Displacement branch: phi (N, F-1).  NOTE the reference never updates ``r_0`` inside its loop
    (misc.py:207-213), so the "velocity" of step t is the displacement from the FIRST frame,
    ``r_t - r_0``, evaluated with the neighbours of frame t-1; the golden phi.npy encodes that."""
    n_frames, n_atoms, _ = positions.shape
    phi = np.zeros((n_frames - 1, n_atoms))
    for t in range(1, n_frames):
        vectors = positions[t] - positions[0]
        indptr, indices = csr[t - 1]
        for i in range(n_atoms):
            if not np.any(vectors[i]):
                continue
            neighbors = indices[indptr[i] : indptr[i + 1]]
            valid = [j for j in neighbors if j != i and np.any(vectors[j])]
            if not valid:
                continue
            phi[t - 1, i] = np.mean([_cosine_similarity(vectors[i], vectors[j]) for j in valid])
    return phi.T
