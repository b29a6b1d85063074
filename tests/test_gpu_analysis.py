"""GPU tests of the spatial average through the drop-in API, written
like the reference's tests/analysis/test_spatialaverage.py against its golden
(tests/analysis/spavg/test_spavg.npy, stored in tests/golden/coex.npz) and against the oracle."""

from __future__ import annotations

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REL = 1e-12  # summation order inside the mean is the only difference from the oracle


@pytest.fixture(scope="module")
def coex():
    from conftest import GOLDEN

    return dict(np.load(GOLDEN / "coex.npz"))


@pytest.fixture(scope="module")
def coex_trj(coex):
    from dynsight_b200.trajectory import Trj
    from dynsight_b200.universe import ArrayUniverse

    n = coex["o_positions"].shape[1]
    return Trj(ArrayUniverse(coex["o_positions"], coex["dims"], ["OW"] * n, ["O"] * n))


def test_spavg_reference_golden(coex, coex_trj):
    """tests/analysis/test_spatialaverage.py:53-64."""
    from dynsight_b200.trajectory import Insight

    coords = coex_trj.get_coordinates("type O")[:, :, 0].T.astype(np.float64)
    out = Insight(coords).spatial_average(coex_trj, r_cut=5.0, selection="type O", n_jobs=1)
    assert out.meta == {"sp_av_r_cut": 5.0, "selection": "type O"}
    assert out.dataset.shape == coex["spavg"].shape
    assert np.allclose(out.dataset, coex["spavg"])
    assert np.abs(out.dataset - coex["spavg"]).max() <= REL * np.abs(coex["spavg"]).max()


@pytest.mark.parametrize("n_comp", [1, 7, 324])
def test_spavg_vs_oracle_random_fluid(n_comp):
    from oracle import spatial_oracle

    from dynsight_b200 import analysis
    from dynsight_b200.universe import ArrayUniverse

    rng = np.random.default_rng(n_comp)
    n, nf, box_l = 700, 3, 14.0
    box = np.array([box_l, box_l * 1.2, box_l * 0.9], dtype=np.float32)
    pos = ((rng.random((nf, n, 3)) * 1.4 - 0.2) * box).astype(np.float32)  # some atoms outside
    dims = np.tile(np.concatenate([box, [90, 90, 90]]).astype(np.float32), (nf, 1))
    dims[1, :3] *= 1.05  # per-frame box
    values = rng.normal(size=(n, nf) if n_comp == 1 else (n, nf, n_comp))
    uni = ArrayUniverse(pos, dims, ["A"] * n, ["A"] * n)
    got = analysis.spatialaverage(uni, values, "all", 2.5)
    want = spatial_oracle.spatial_average(pos, dims, values, 2.5)
    assert got.shape == values.shape and got.dtype == np.float64
    assert np.abs(got - want).max() <= REL * max(1.0, np.abs(want).max())


def test_spavg_selection_and_trajslice():
    from oracle import spatial_oracle

    from dynsight_b200 import analysis
    from dynsight_b200.universe import ArrayUniverse

    rng = np.random.default_rng(5)
    n, nf = 400, 6
    box = np.array([12.0, 12.0, 12.0], dtype=np.float32)
    pos = (rng.random((nf, n, 3)) * box).astype(np.float32)
    dims = np.tile(np.concatenate([box, [90, 90, 90]]).astype(np.float32), (nf, 1))
    names = np.where(np.arange(n) % 3 == 0, "B", "A")
    uni = ArrayUniverse(pos, dims, names, names)
    sel = np.where(names == "A")[0]
    sl = slice(1, 6, 2)
    values = rng.normal(size=(sel.size, 3))
    got = analysis.spatialaverage(uni, values, "name A", 3.0, trajslice=sl)
    want = spatial_oracle.spatial_average(pos[sl][:, sel], dims[sl], values, 3.0)
    assert np.abs(got - want).max() <= REL * max(1.0, np.abs(want).max())


def test_spavg_errors():
    from dynsight_b200 import analysis
    from dynsight_b200.universe import ArrayUniverse

    pos = np.zeros((2, 5, 3), dtype=np.float32)
    dims = np.tile(np.array([5, 5, 5, 90, 90, 90], dtype=np.float32), (2, 1))
    uni = ArrayUniverse(pos, dims, ["A"] * 5, ["A"] * 5)
    with pytest.raises(ValueError, match="ndim == 2 or ndim == 3"):
        analysis.spatialaverage(uni, np.zeros(5), "all", 1.0)
    with pytest.raises(ValueError, match="simulation box"):
        analysis.spatialaverage(ArrayUniverse(pos, None, ["A"] * 5, ["A"] * 5), np.zeros((5, 2)),
                                "all", 1.0)  # fmt: skip


def test_spavg_large_fluid_properties():
    """1M particles (BASELINE cfg3 density): a constant field is a fixed point, and
    sum_i (n_i + 1) out_i = sum_j (n_j + 1) v_j because the neighbour relation is symmetric."""
    import torch

    from dynsight_b200 import engine

    n, nf = 1_000_000, 2
    box_l = (n / 0.8) ** (1 / 3)
    g = torch.Generator(device="cuda").manual_seed(3)
    pos = (torch.rand((nf, n, 3), device="cuda", generator=g) * box_l).float()
    box = np.full((nf, 3), box_l)
    values = torch.rand((n, nf), device="cuda", generator=g, dtype=torch.float64)
    wrapped = engine.wrap_positions(pos, box)
    nr = engine.neighbor_rows(wrapped, box, 1.5, inclusive_cutoff=True)
    assert not nr.overflowed()
    out = engine.spatial_average_rows(nr, values, torch.empty_like(values))
    ones = engine.spatial_average_rows(nr, torch.ones_like(values), torch.empty_like(values))
    assert torch.equal(ones, torch.ones_like(ones))
    w = (nr.counts.T + 1).double()
    lhs, rhs = (w * out).sum(0), (w * values).sum(0)
    assert torch.allclose(lhs, rhs, rtol=1e-10)
    assert 10.0 < float(nr.counts.double().mean()) < 12.5


# ---- self time-correlation function --------------------------------
def test_tcf_reference_goldens():
    """tests/analysis/test_time_correlations.py:55-63 and the doctest of time_correlations.py:57."""
    from conftest import GOLDEN
    from oracle import tcf_oracle

    from dynsight_b200.trajectory import Insight

    fx = np.load(GOLDEN / "tcorr.npz")
    corr, std = Insight(tcf_oracle.reference_walk()).get_time_correlation()
    assert np.allclose(corr, fx["t_corr"])
    assert np.allclose(std, fx["std_dev"])
    assert np.allclose(corr, fx["t_corr"], rtol=1e-10, atol=1e-13)
    np.random.seed(1234)  # noqa: NPY002
    from dynsight_b200 import analysis

    time_corr, _ = analysis.self_time_correlation(np.random.rand(100, 100))  # noqa: NPY002
    assert np.isclose(time_corr[1], float(fx["kat_rand_lag1"]))


@pytest.mark.parametrize(("n_part", "n_frames", "max_delay"),
                         [(37, 1000, None), (300, 257, 100), (3, 30000, 64), (50, 9, 9)])  # fmt: skip
def test_tcf_vs_oracle(n_part, n_frames, max_delay):
    from oracle import tcf_oracle

    from dynsight_b200 import analysis

    rng = np.random.default_rng(n_frames)
    data = np.cumsum(rng.normal(size=(n_part, n_frames)), axis=1) + rng.normal(size=(n_part, 1))
    got_c, got_e = analysis.self_time_correlation(data, max_delay)
    want_c, want_e = tcf_oracle.self_time_correlation(data, max_delay)
    assert got_c.shape == want_c.shape
    assert np.abs(got_c - want_c).max() <= 1e-10
    assert np.abs(got_e - want_e).max() <= 1e-10 * max(1.0, np.abs(want_e).max())


def test_tcf_errors():
    from dynsight_b200 import analysis

    with pytest.raises(ValueError, match="max_delay"):
        analysis.self_time_correlation(np.zeros((3, 10)), 11)


def test_cross_tcf_reference_goldens_and_oracle():
    """tests/analysis/test_time_correlations.py:66-73 and the doctest of time_correlations.py:141,
    plus random series against the oracle (all ordered pairs i != j)."""
    from conftest import GOLDEN
    from oracle import tcf_oracle

    from dynsight_b200 import analysis

    fx = np.load(GOLDEN / "tcorr.npz")
    corr, std = analysis.cross_time_correlation(tcf_oracle.reference_walk())
    assert np.allclose(corr, fx["c_corr"])
    assert np.allclose(std, fx["ctd_dev"])
    np.random.seed(1234)  # noqa: NPY002
    time_corr, _ = analysis.cross_time_correlation(np.random.rand(100, 100))  # noqa: NPY002
    assert np.isclose(time_corr[0], float(fx["kat_cross_lag0"]))
    rng = np.random.default_rng(8)
    for n_part, n_frames, max_delay in ((2, 50, None), (17, 300, 40), (6, 30000, 16)):
        data = np.cumsum(rng.normal(size=(n_part, n_frames)), axis=1)
        got_c, got_e = analysis.cross_time_correlation(data, max_delay)
        want_c, want_e = tcf_oracle.cross_time_correlation(data, max_delay)
        assert np.allclose(got_c, want_c, rtol=1e-9, atol=1e-9 * np.abs(want_c).max())
        assert np.allclose(got_e, want_e, rtol=1e-8, atol=1e-9 * np.abs(want_e).max())
    with pytest.raises(ValueError, match="two particles"):
        analysis.cross_time_correlation(np.zeros((1, 10)))
