"""GPU parity: fused timeSOAP kernel through the C ABI vs the reference's numpy routine.

Tolerance (stated in BASELINE.md section 5): |d| <= 1e-5 * t + 1e-7 -- the reference's own
sqrt(2 - 2 cos) has a ~2e-8 absolute noise floor near t = 0.
"""

from __future__ import annotations

import numpy as np
import pytest
import torch
from golden_cases import random_soap_spectra

from oracle import timesoap_oracle as to

pytestmark = pytest.mark.gpu


def _close(got: np.ndarray, want: np.ndarray) -> bool:
    return bool(np.all(np.abs(got - want) <= 1e-5 * np.abs(want) + 1e-7))


@pytest.mark.parametrize("delay", [1, 3])
def test_vs_reference_module_outputs(numba_cases, delay):
    from dynsight_b200 import engine

    s = random_soap_spectra()
    got = engine.timesoap_device(torch.as_tensor(s).cuda(), delay).cpu().numpy()
    want = numba_cases[f"tsoap_d{delay}"]
    assert got.shape == want.shape
    assert _close(got, want)
    # exact edge values: zero vectors -> sqrt(2); anti-parallel -> 2
    assert np.allclose(got[1], np.sqrt(2.0), atol=1e-12)
    if delay == 1:
        assert abs(got[4, 1] - 2.0) < 1e-12
        assert got[2, 4] <= 1e-7  # identical consecutive frames
        assert got[3, 5] <= 1e-7  # parallel vectors


@pytest.mark.parametrize(("n_comp", "delay"), [(50, 1), (324, 1), (324, 4), (700, 1), (3300, 2)])
def test_random_sizes_and_strides(n_comp, delay):
    from dynsight_b200 import engine

    rng = np.random.default_rng(n_comp + delay)
    s = rng.random((37, 23, n_comp))
    want = to.timesoap(s, delay)
    t = torch.as_tensor(s).cuda()
    assert _close(engine.timesoap_device(t, delay).cpu().numpy(), want)
    # the transposed view saponify_trajectory returns (frames outermost in memory)
    tv = t.permute(1, 0, 2).contiguous().permute(1, 0, 2)
    assert not tv.is_contiguous()
    assert _close(engine.timesoap_device(tv, delay).cpu().numpy(), want)


def test_bounds_error():
    from dynsight_b200 import engine

    t = torch.zeros((3, 5, 8), dtype=torch.float64, device="cuda")
    for bad in (0, 5, 6):
        with pytest.raises(ValueError, match="delay value outside bounds"):
            engine.timesoap_device(t, bad)


def test_tsoap_golden_from_golden_soap_subsample(ref_soap):
    """The stored SOAP golden holds frames 0::10, so delay-1 timeSOAP on it equals the reference's
    timesoap(soap[:, ::10]); compare the kernel with the oracle on the reference's own spectra."""
    from dynsight_b200 import engine

    soap = ref_soap["trj_soap"]  # (7, 21, 324) DScribe output
    want = to.timesoap(soap, 1)
    got = engine.timesoap_device(torch.as_tensor(soap).cuda().contiguous(), 1).cpu().numpy()
    assert _close(got, want)
