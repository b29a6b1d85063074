"""CPU restatement of the SOAP power spectrum the reference obtains from DScribe.

TEST INFRASTRUCTURE ONLY (oracle): imported by ``tests/``, ``__graft_entry__.smoke()``,
``tests/golden/make_golden.py`` and ``bench.py``'s CPU-baseline legs -- never by ``dynsight_b200``.

Reference call site: ``src/dynsight/_internal/soapify/saponify.py:98-122`` --
``dscribe.descriptors.SOAP(species, r_cut, n_max, l_max, periodic).create(...)`` with every other
DScribe default (sigma=1, rbf="gto", weighting=None, average="off", float64, dense).
The arithmetic lives in the third-party package **DScribe** (PyPI ``dscribe``; the reference pins no
version, ``pyproject.toml:15``; not vendored, not installed, no network).  What follows restates
DScribe's published SOAP-GTO algorithm (Himanen et al., Comput. Phys. Commun. 247, 106949 (2020),
and ``dscribe/descriptors/soap.py::get_basis_gto`` + ``dscribe/ext/soapGTO.cpp``):

1. sigma = 1, eta = 1/(2 sigma^2); atoms (all periodic images, the centre included) with
   |d| <= R = r_cut + sigma*sqrt(-2 ln 1e-3) contribute with weight 1 (no cutoff function).
2. radial basis per l: a_k = linspace(1, r_cut, n_max), alpha_lk = -ln(1e-3 / a_k^l) / a_k^2,
   overlap S_kk' = 1/2 Gamma(l+3/2) (alpha_lk + alpha_lk')^-(l+3/2), beta_l = sqrtm(inv(S)).
3. G^Z_lkm = pi^(3/2) sum_j (eta/(eta+alpha))^l (eta+alpha)^(-3/2) r^l exp(-alpha eta r^2/(alpha+eta))
   Y_lm(d_hat)  (orthonormal real spherical harmonics), c^Z_nlm = sum_k beta_l[n,k] G^Z_lkm.
4. p^{ZZ'}_{l n n'} = pi sqrt(8/(2l+1)) sum_m c^Z_nlm c^Z'_n'lm.
5. order: species pairs Z<=Z' ascending atomic number; inside a pair l outermost, then n, then
   n' (n'>=n when Z==Z', full n_max^2 otherwise) -- the order that reproduces the reference's
   golden files (tests/test_oracle_golden.py) and that ``fill_soap_vector_from_dscribe``
   (saponify.py:176-202) assumes.

PARITY PIN: validated against every SOAP golden the reference holds
(``tests/soap/test_soap/c0..c5*.npy``, ``tests/trajectory/files/soap.npy``, the doctest sums in
``saponify.py:86-87`` / ``timesoap.py:50-52,115-116,172-173``); those are single-species, so the
multi-species block order (step 5, species pairs) is "parity unpinned".
"""

from __future__ import annotations

import math

import numpy as np

_SIGMA = 1.0
_ETA = 1.0 / (2.0 * _SIGMA * _SIGMA)
_THRESHOLD = 1e-3

# Atomic numbers for the element symbols the path meets ("X" is ASE's dummy element, Z=0).
ATOMIC_NUMBERS = {
    "X": 0, "H": 1, "He": 2, "Li": 3, "Be": 4, "B": 5, "C": 6, "N": 7, "O": 8, "F": 9, "Ne": 10,
    "Na": 11, "Mg": 12, "Al": 13, "Si": 14, "P": 15, "S": 16, "Cl": 17, "Ar": 18, "K": 19,
    "Ca": 20, "Ti": 22, "Cr": 24, "Mn": 25, "Fe": 26, "Co": 27, "Ni": 28, "Cu": 29, "Zn": 30,
    "Br": 35, "Pd": 46, "Ag": 47, "Pt": 78, "Au": 79,
}  # fmt: skip


def search_radius(r_cut: float) -> float:
    return r_cut + _SIGMA * math.sqrt(-2.0 * math.log(_THRESHOLD))


def gto_basis(r_cut: float, n_max: int, l_max: int) -> tuple[np.ndarray, np.ndarray]:
    """alphas (l_max+1, n_max) and betas (l_max+1, n_max, n_max), DScribe ``get_basis_gto``."""
    from scipy.linalg import inv, sqrtm
    from scipy.special import gamma

    a = np.linspace(1.0, r_cut, n_max)
    alphas = np.zeros((l_max + 1, n_max))
    betas = np.zeros((l_max + 1, n_max, n_max))
    for l in range(l_max + 1):
        al = -np.log(_THRESHOLD / np.power(a, l)) / a**2
        m = al[None, :] + al[:, None]
        s = 0.5 * gamma(l + 1.5) * m ** (-l - 1.5)
        betas[l] = np.real(sqrtm(inv(s)))
        alphas[l] = al
    return alphas, betas


def real_sph_harm(l_max: int, d_hat: np.ndarray) -> list[np.ndarray]:
    """Orthonormal real spherical harmonics; returns list over l of arrays (K, 2l+1).

    Built from the associated Legendre recursion in z and the (x + i y)^m polynomials, so no
    trigonometric call and no singularity on the z axis.
    """
    x, y, z = d_hat[:, 0], d_hat[:, 1], d_hat[:, 2]
    k = x.shape[0]
    # A_m + i B_m = (x + i y)^m
    a_m = [np.ones(k)]
    b_m = [np.zeros(k)]
    for _ in range(l_max):
        a_m.append(a_m[-1] * x - b_m[-1] * y)
        b_m.append(a_m[-2] * y + b_m[-1] * x)
    # Q_l^m(z) = P_l^m / sin^m(theta) (no Condon-Shortley phase needed: it cancels in sum_m)
    q = [[None] * (l_max + 1) for _ in range(l_max + 1)]
    for m in range(l_max + 1):
        q[m][m] = np.full(k, float(np.prod(np.arange(1, 2 * m, 2))) if m > 0 else 1.0)
        if m + 1 <= l_max:
            q[m + 1][m] = (2 * m + 1) * z * q[m][m]
        for l in range(m + 2, l_max + 1):
            q[l][m] = ((2 * l - 1) * z * q[l - 1][m] - (l + m - 1) * q[l - 2][m]) / (l - m)
    out = []
    for l in range(l_max + 1):
        y_l = np.zeros((k, 2 * l + 1))
        for m in range(l + 1):
            norm = math.sqrt(
                (2 * l + 1) / (4 * math.pi) * math.factorial(l - m) / math.factorial(l + m)
            )
            if m == 0:
                y_l[:, l] = norm * q[l][0]
            else:
                y_l[:, l + m] = math.sqrt(2.0) * norm * q[l][m] * a_m[m]
                y_l[:, l - m] = math.sqrt(2.0) * norm * q[l][m] * b_m[m]
        out.append(y_l)
    return out


def neighbours_all_images(
    pos: np.ndarray, centre: np.ndarray, cell: np.ndarray | None, radius: float
) -> tuple[np.ndarray, np.ndarray]:
    """Displacements d (K,3) and atom indices (K,) of every image within ``radius``."""
    d0 = pos - centre[None, :]
    if cell is None:
        shifts = np.zeros((1, 3))
    else:
        reps = [int(math.ceil(radius / cell[k])) + 1 if cell[k] > 0 else 0 for k in range(3)]
        gx, gy, gz = np.meshgrid(
            np.arange(-reps[0], reps[0] + 1),
            np.arange(-reps[1], reps[1] + 1),
            np.arange(-reps[2], reps[2] + 1),
            indexing="ij",
        )
        shifts = np.stack([gx.ravel(), gy.ravel(), gz.ravel()], axis=1) * cell[None, :]
    d = d0[None, :, :] + shifts[:, None, :]
    r2 = np.sum(d * d, axis=-1)
    s_idx, a_idx = np.nonzero(r2 <= radius * radius)
    return d[s_idx, a_idx], a_idx


def soap_frame(
    pos: np.ndarray,
    species_z: np.ndarray,
    centre_idx: np.ndarray,
    cell: np.ndarray | None,
    r_cut: float,
    n_max: int,
    l_max: int,
    basis: tuple[np.ndarray, np.ndarray] | None = None,
) -> np.ndarray:
    """SOAP power spectrum of one frame -> (n_centres, C) float64.

    ``pos`` (N,3) any float dtype (computed in float64), ``species_z`` (N,) atomic numbers,
    ``cell`` (3,) orthorhombic lengths or None for a non-periodic frame.
    """
    pos = np.asarray(pos, dtype=np.float64)
    alphas, betas = basis if basis is not None else gto_basis(r_cut, n_max, l_max)
    radius = search_radius(r_cut)
    zs = np.unique(species_z)
    n_sp = len(zs)
    n_feat = (n_sp * n_max) * (n_sp * n_max + 1) // 2 * (l_max + 1)
    out = np.zeros((len(centre_idx), n_feat))
    pref = math.pi**1.5
    for ci, c_atom in enumerate(centre_idx):
        d, a_idx = neighbours_all_images(pos, pos[c_atom], cell, radius)
        r2 = np.sum(d * d, axis=-1)
        r = np.sqrt(r2)
        safe_r = np.where(r > 0, r, 1.0)
        d_hat = d / safe_r[:, None]
        d_hat[r == 0] = (0.0, 0.0, 1.0)
        ylm = real_sph_harm(l_max, d_hat)
        # c[z][l] : (n_max, 2l+1)
        c = [[None] * (l_max + 1) for _ in range(n_sp)]
        for zi, z in enumerate(zs):
            sel = species_z[a_idx] == z
            for l in range(l_max + 1):
                al = alphas[l]  # (n_max,)
                rl = np.where(r[sel] > 0, r[sel] ** l, 1.0 if l == 0 else 0.0)
                radial = (
                    pref
                    * (_ETA / (_ETA + al)) ** l
                    * (_ETA + al) ** -1.5
                    * rl[:, None]
                    * np.exp(-al[None, :] * _ETA / (al[None, :] + _ETA) * r2[sel, None])
                )  # (K, n_max) primitives
                g = radial.T @ ylm[l][sel]  # (n_max, 2l+1)
                c[zi][l] = betas[l] @ g
        col = 0
        for zi in range(n_sp):
            for zj in range(zi, n_sp):
                for l in range(l_max + 1):
                    block = math.pi * math.sqrt(8.0 / (2 * l + 1)) * (c[zi][l] @ c[zj][l].T)
                    vals = block[np.triu_indices(n_max)] if zi == zj else block.ravel()
                    out[ci, col : col + vals.size] = vals
                    col += vals.size
    return out


def soap_trajectory(
    positions: np.ndarray,
    species_z: np.ndarray,
    centre_idx: np.ndarray,
    cells: np.ndarray | None,
    r_cut: float,
    n_max: int,
    l_max: int,
) -> np.ndarray:
    """(F,N,3) -> (n_centres, F, C), the layout ``saponify_trajectory`` returns (saponify.py:124)."""
    basis = gto_basis(r_cut, n_max, l_max)
    frames = [
        soap_frame(
            positions[f],
            species_z,
            centre_idx,
            None if cells is None else np.asarray(cells[f], dtype=np.float64),
            r_cut,
            n_max,
            l_max,
            basis,
        )
        for f in range(positions.shape[0])
    ]
    return np.transpose(np.stack(frames, axis=0), (1, 0, 2))


def fill_soap_vector(soap: np.ndarray, lmax: int = 8, nmax: int = 8) -> np.ndarray:
    """Restatement of ``fill_soap_vector_from_dscribe`` (saponify.py:176-202), single species."""
    shape = soap.shape
    if soap.ndim == 1:
        soap = soap[None, None, :]
    elif soap.ndim == 2:  # noqa: PLR2004
        soap = soap[:, None, :]
    n_p, n_f, _ = soap.shape
    iu = np.triu_indices(nmax)
    tri = soap.reshape(n_p, n_f, lmax + 1, -1)
    full = np.zeros((n_p, n_f, lmax + 1, nmax, nmax))
    full[:, :, :, iu[0], iu[1]] = tri
    full[:, :, :, iu[1], iu[0]] = tri
    out = full.reshape(n_p, n_f, -1)
    if len(shape) == 1:
        return out.squeeze()
    if len(shape) == 2:  # noqa: PLR2004
        return out.squeeze(1)
    return out


# ---------------------------------------------------------------------------------------------
# C / OpenMP version (oracle/soap_oracle.c): same algorithm, used as the timed CPU baseline
# ---------------------------------------------------------------------------------------------
_CLIB = None


def _clib():
    global _CLIB  # noqa: PLW0603
    if _CLIB is None:
        import ctypes
        import subprocess
        from pathlib import Path

        here = Path(__file__).resolve().parent
        so_path, src = here / "libsoap_oracle.so", here / "soap_oracle.c"
        if not so_path.exists() or so_path.stat().st_mtime < src.stat().st_mtime:
            subprocess.run(
                ["/usr/bin/gcc" if __import__("os").path.exists("/usr/bin/gcc") else "gcc", "-O2", "-fPIC", "-fopenmp", "-std=c11", "-shared", "-o", str(so_path),
                 str(src), "-lm"],
                check=True,
            )  # fmt: skip
        _CLIB = ctypes.CDLL(str(so_path))
        _CLIB.dsgo_soap_frame.restype = ctypes.c_int
        _CLIB.dsgo_soap_frame.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
            ctypes.c_int, ctypes.c_void_p, ctypes.c_double, ctypes.c_int, ctypes.c_int,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
        ]  # fmt: skip
        _CLIB.dsgo_soap_set_num_threads.argtypes = [ctypes.c_int]
    return _CLIB


def soap_frame_c(
    pos: np.ndarray,
    species_z: np.ndarray,
    centre_idx: np.ndarray,
    cell: np.ndarray | None,
    r_cut: float,
    n_max: int,
    l_max: int,
    basis: tuple[np.ndarray, np.ndarray] | None = None,
    n_threads: int | None = None,
) -> np.ndarray:
    """Same contract as :func:`soap_frame`, evaluated by the C/OpenMP restatement."""
    lib = _clib()
    if n_threads is not None:
        lib.dsgo_soap_set_num_threads(int(n_threads))
    alphas, betas = basis if basis is not None else gto_basis(r_cut, n_max, l_max)
    alphas = np.ascontiguousarray(alphas)
    betas = np.ascontiguousarray(betas)
    p = np.ascontiguousarray(pos, dtype=np.float64)
    zs, sp = np.unique(species_z, return_inverse=True)
    sp = np.ascontiguousarray(sp, dtype=np.int32)
    cen = np.ascontiguousarray(centre_idx, dtype=np.int32)
    n_sp = len(zs)
    n_feat = (n_sp * n_max) * (n_sp * n_max + 1) // 2 * (l_max + 1)
    out = np.zeros((len(cen), n_feat))
    cell_arr = None if cell is None else np.ascontiguousarray(cell, dtype=np.float64)
    rc = lib.dsgo_soap_frame(
        p.ctypes.data, sp.ctypes.data, p.shape[0], cen.ctypes.data, len(cen), n_sp,
        None if cell_arr is None else cell_arr.ctypes.data, float(r_cut), n_max, l_max,
        alphas.ctypes.data, betas.ctypes.data, out.ctypes.data,
    )  # fmt: skip
    if rc != 0:
        msg = "dsgo_soap_frame: unsupported arguments"
        raise ValueError(msg)
    return out
