"""Host-side GTO radial basis of the SOAP descriptor (float64, numpy only).

DScribe's ``get_basis_gto`` (called inside ``SOAP(...)`` at soapify/saponify.py:98-104):
``a_k = linspace(1, r_cut, n_max)``, ``alpha_lk = -ln(1e-3 / a_k^l) / a_k^2``, overlap
``S_kk' = 1/2 Gamma(l+3/2) (alpha_lk + alpha_lk')^-(l+3/2)`` and the Loewdin matrix
``beta_l = S^(-1/2) = sqrtm(inv(S))``.
"""

from __future__ import annotations

import math
from functools import lru_cache

import numpy as np

THRESHOLD = 1e-3
SIGMA = 1.0


def search_radius(r_cut: float) -> float:
    """r_cut + sigma*sqrt(-2 ln 1e-3): every atom image inside it contributes."""
    return r_cut + SIGMA * math.sqrt(-2.0 * math.log(THRESHOLD))


@lru_cache(maxsize=32)
def gto_basis(r_cut: float, n_max: int, l_max: int) -> tuple[np.ndarray, np.ndarray]:
    """Return ``alphas (l_max+1, n_max)`` and ``betas (l_max+1, n_max, n_max)`` (C-contiguous)."""
    if n_max < 1 or l_max < 0:
        msg = "n_max must be >= 1 and l_max >= 0"
        raise ValueError(msg)
    if not r_cut > 1.0:
        msg = "r_cut must be greater than 1 angstrom for the GTO radial basis"
        raise ValueError(msg)
    a = np.linspace(1.0, float(r_cut), n_max)
    alphas = np.empty((l_max + 1, n_max))
    betas = np.empty((l_max + 1, n_max, n_max))
    for l in range(l_max + 1):
        # same floating-point expressions as DScribe, so that S agrees to the last bit
        al = -np.log(THRESHOLD / np.power(a, l)) / a**2
        m = np.zeros((n_max, n_max))
        m[:, :] = al
        m = m + m.transpose()
        s = 0.5 * _gamma(l + 3.0 / 2.0) * m ** (-l - 3.0 / 2.0)
        alphas[l] = al
        betas[l] = _inv_sqrt_spd(s)
    return np.ascontiguousarray(alphas), np.ascontiguousarray(betas)


def _gamma(x: float) -> float:
    try:
        from scipy.special import gamma
    except ImportError:  # pragma: no cover
        return math.gamma(x)
    return float(gamma(x))


def _inv_sqrt_spd(s: np.ndarray) -> np.ndarray:
    """S^(-1/2).  The overlap matrices are badly conditioned (|beta| reaches 1e3..1e6), so the
    result depends on the algorithm at the 1e-7 level; DScribe evaluates ``sqrtm(inv(S))`` with
    scipy, and the same two calls are used here when scipy is importable so that the spectra agree
    with DScribe's to ~1e-11 even for ill-conditioned bases.  Fallback: symmetric eigendecomposition.
    """
    try:
        from scipy.linalg import inv, sqrtm
    except ImportError:  # pragma: no cover - scipy is part of the image
        w, v = np.linalg.eigh(s)
        if np.any(w <= 0):
            msg = "GTO overlap matrix is numerically singular (n_max too large for this r_cut)"
            raise ValueError(msg) from None
        return (v / np.sqrt(w)[None, :]) @ v.T
    return np.real(sqrtm(inv(s)))


def n_features(n_species: int, n_max: int, l_max: int) -> int:
    sn = n_species * n_max
    return sn * (sn + 1) // 2 * (l_max + 1)


def scratch_doubles(n_species: int, n_max: int, l_max: int) -> int:
    """Doubles of device scratch per centre-frame: with several species or n_max > 8 the SOAP
    kernel keeps the primitive sums ``(S, (l+1)^2, n)`` in global memory between its two passes."""
    if n_species * ((n_max + 7) // 8) <= 1:
        return 0
    return n_species * (l_max + 1) ** 2 * n_max
