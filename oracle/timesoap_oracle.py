"""numpy restatement of the reference's timeSOAP routines (TEST INFRASTRUCTURE ONLY).

Reference: ``/root/reference/src/dynsight/_internal/timesoap/timesoap.py``
  * ``normalize_soap`` :13-58 (arithmetic :54-58)
  * ``soap_distance``  :61-127 (arithmetic :118-127)
  * ``timesoap``       :130-179 (bounds check :175-177)
PARITY PIN: bit-compared against the reference module itself (oracle/ref_loader.py) in
tests/golden/make_golden.py and pinned by tests/tsoap/test_tsoap/c0,c2 + trajectory/files/tsoap.npy.
"""

from __future__ import annotations

import numpy as np


def normalize_soap(soap: np.ndarray) -> np.ndarray:
    norms = np.linalg.norm(soap, axis=-1)
    mask = norms > 0.0
    out = np.zeros(soap.shape)
    out[mask] = soap[mask] / norms[..., np.newaxis][mask]
    return out


def soap_distance(v_1: np.ndarray, v_2: np.ndarray) -> np.ndarray:
    n_1 = normalize_soap(v_1)
    n_2 = normalize_soap(v_2)
    cos_theta = np.clip(np.sum(n_1 * n_2, axis=-1), -1.0, 1.0)
    return np.sqrt(np.maximum(0, 2 - 2 * cos_theta))


def timesoap(soap: np.ndarray, delay: int = 1) -> np.ndarray:
    if delay < 1 or delay >= soap.shape[1]:
        msg = "delay value outside bounds"
        raise ValueError(msg)
    return soap_distance(soap[:, :-delay, :], soap[:, delay:, :])
