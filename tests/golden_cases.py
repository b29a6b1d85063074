"""Seeded input generators shared by tests/golden/make_golden.py and the parity tests.

Only inputs are generated here (pure numpy, deterministic); the matching outputs of the
reference's numba kernels are stored in tests/golden/numba_cases.npz.
"""

from __future__ import annotations

import numpy as np


def random_csr_cases() -> list[dict]:
    """Ragged / edge-case inputs for ``neighbor_list_celllist_centers`` (lens.py:69-147)."""
    rng = np.random.default_rng(20261019)
    cases: list[dict] = []

    def add(n, m, box, r_cut, pbc, spread=(0.0, 1.0), same=False):
        box = np.asarray(box, dtype=np.float32)
        lo, hi = spread
        pe = ((rng.random((n, 3)) * (hi - lo) + lo) * box).astype(np.float32)
        pc = pe if same else ((rng.random((m, 3)) * (hi - lo) + lo) * box).astype(np.float32)
        cases.append({"pos_env": pe, "pos_cent": pc, "box": box, "r_cut": r_cut, "pbc": pbc})

    add(300, 300, (9.0, 9.0, 9.0), 1.5, True, same=True)  # 0 plain fluid
    add(257, 91, (8.0, 11.0, 6.5), 1.7, True)  # 1 centres != env, anisotropic
    add(200, 200, (7.0, 7.0, 7.0), 1.2, True, spread=(-0.3, 1.3), same=True)  # 2 atoms outside box
    add(180, 64, (7.0, 7.0, 7.0), 1.4, False, spread=(-0.2, 1.2))  # 3 pbc off + outside
    add(150, 150, (4.0, 4.0, 4.0), 1.9, True, same=True)  # 4 ncell pinned at 3, box > 2 r_cut
    add(60, 60, (3.0, 3.5, 2.9), 2.5, True, same=True)  # 5 box < 2 r_cut (images collide)
    add(1, 1, (5.0, 5.0, 5.0), 1.0, True, same=True)  # 6 single atom
    add(400, 5, (3.0, 3.0, 3.0), 1.0, True)  # 7 dense: rows longer than 32 / 64
    add(500, 500, (30.0, 30.0, 30.0), 3, True, same=True)  # 8 integer r_cut, sparse, many empty cells
    add(333, 333, (6.0, 6.0, 6.0), 2.0, False, same=True)  # 9 pbc off, exact 3 cells
    add(64, 200, (10.0, 5.0, 20.0), 1.25, True, spread=(-1.2, 2.3))  # 10 far outside (several boxes)
    return cases


def fluid_case(n_side: int = 13, n_frames: int = 5) -> tuple[np.ndarray, np.ndarray, float]:
    """Perturbed-lattice LJ-like fluid at rho*=0.8 with a Gaussian random walk, wrapped.

    Returns positions (F, N, 3) float32, box (3,) float32, r_cut.
    """
    rng = np.random.default_rng(7)
    n = n_side**3  # 2197
    box_l = (n / 0.8) ** (1.0 / 3.0)
    g = (np.arange(n_side) + 0.5) * (box_l / n_side)
    lattice = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 3)
    pos = np.empty((n_frames, n, 3), dtype=np.float32)
    cur = lattice + rng.normal(0.0, 0.15, lattice.shape)
    for t in range(n_frames):
        cur = cur + rng.normal(0.0, 0.08, cur.shape)
        pos[t] = np.mod(cur, box_l).astype(np.float32)
    # shuffle atom ids so that id order is not spatial order
    perm = rng.permutation(n)
    pos = pos[:, perm]
    box = np.full(3, box_l, dtype=np.float32)
    pos = np.minimum(pos, np.nextafter(box[0], np.float32(0)))
    return pos, box, 1.5


def random_soap_spectra() -> np.ndarray:
    """(N=12, F=9, C=50) float64 spectra with zero vectors and repeated frames."""
    rng = np.random.default_rng(99)
    s = rng.random((12, 9, 50)) * 3.0
    s[0, 3] = 0.0  # zero vector inside a series
    s[1, :] = 0.0  # all-zero particle
    s[2, 5] = s[2, 4]  # identical consecutive frames -> distance 0 (or the ~2e-8 noise floor)
    s[3, 6] = 2.5 * s[3, 5]  # parallel vectors
    s[4, 2] = -s[4, 1]  # anti-parallel -> 2
    s[5] *= 1e-150  # tiny magnitudes
    s[6] *= 1e120  # huge magnitudes
    return s
