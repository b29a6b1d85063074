"""GPU tests through the drop-in Python API, written like the reference's own tests
(tests/lens/test_lens.py, tests/neighbors/test_nn.py, tests/soap/test_soap.py,
tests/tsoap/test_tsoap.py, tests/trajectory/test_trj.py) against the reference's golden files.
"""

from __future__ import annotations

import numpy as np
import pytest
from test_oracle_golden import SOAP_CASES

pytestmark = pytest.mark.gpu

# (r_cut, delay, centers, selection, n_jobs) -- tests/lens/conftest.py:6-74
LENS_API_CASES = {
    "c0": (3, 1, "all", "all", 1),
    "c1": (4, 1, "all", "all", 1),
    "c2": (4, 1, "id 1", "all", 1),
    "c3": (4, 1, "all", "id 1", 1),
    "c4": (4, 1, "all", "all", 2),
    "c5": (4, 2, "all", "all", 1),
}
SOAP_API_CENTERS = {
    "soap_c0": "all", "soap_c1": "all", "soap_c2": "all", "soap_c3": "not name C5",
    "soap_c4": "name C3 or name C6", "soap_c5": "all", "trj_soap": "all",
}  # fmt: skip


@pytest.fixture(scope="module")
def trj(balls7):
    from dynsight_b200.trajectory import Trj
    from dynsight_b200.universe import ArrayUniverse

    return Trj(
        ArrayUniverse(
            balls7["positions"], balls7["dims"], balls7["names"], balls7["types"], balls7["ids"]
        )
    )


@pytest.mark.parametrize("case", sorted(LENS_API_CASES))
def test_lens(trj, ref_lens, case):
    r_cut, delay, centers, selection, n_jobs = LENS_API_CASES[case]
    got = trj.get_lens(r_cut=r_cut, delay=delay, centers=centers, selection=selection,
                       n_jobs=n_jobs)  # fmt: skip
    assert got.meta == {"name": "lens", "r_cut": r_cut, "delay": delay, "centers": centers,
                        "selection": selection}  # fmt: skip
    assert got.dataset.dtype == np.float64
    assert np.allclose(ref_lens["lens_" + case], got.dataset, atol=1e-6)
    assert np.array_equal(ref_lens["lens_" + case], got.dataset)


@pytest.mark.parametrize("case", ["c0", "c1", "c2", "c3", "c4"])
def test_nn(trj, ref_lens, case):
    r_cut, _, centers, selection, n_jobs = LENS_API_CASES[case]
    neigh, got = trj.get_coord_number(r_cut=r_cut, centers=centers, selection=selection,
                                      n_jobs=n_jobs)  # fmt: skip
    assert np.array_equal(ref_lens["nn_" + case], got.dataset)
    assert len(neigh) == 201 and len(neigh[0]) == got.dataset.shape[0]


def test_lens_2d(ref_lens, tmp_path):
    """tests/lens/test_lens.py:66-92, including the box assignment loop over an XYZ universe."""
    from dynsight_b200.trajectory import Trj

    rng = np.random.default_rng(42)
    coords = rng.random((100, 100, 3))
    coords[..., 2] = 0.0
    path = tmp_path / "random_2d.xyz"
    with path.open("w") as fh:
        for frame in coords:
            print(coords.shape[1], file=fh)
            print("# comment line", file=fh)
            for atom in frame:
                print("C", atom[0], atom[1], atom[2], file=fh)
    trj_2d = Trj.init_from_xyz(path, dt=1.0)
    for ts in trj_2d.universe.trajectory:
        pos = trj_2d.universe.atoms.positions
        lengths = pos.max(axis=0) - pos.min(axis=0)
        lengths[2] = 0.5
        ts.dimensions = np.concatenate([lengths, np.array([90, 90, 90])])
    lens_pbc = trj_2d.get_lens(r_cut=0.1, respect_pbc=True)
    _ = trj_2d.get_coord_number(r_cut=0.1, respect_pbc=True)
    assert np.array_equal(ref_lens["lens_2d_pbc"], lens_pbc.dataset)
    lens_fbc = trj_2d.get_lens(r_cut=0.1, respect_pbc=False)
    _ = trj_2d.get_coord_number(r_cut=0.1, respect_pbc=False)
    assert np.array_equal(ref_lens["lens_2d_fbc"], lens_fbc.dataset)


@pytest.mark.parametrize("case", sorted(SOAP_CASES))
def test_soap(trj, ref_soap, case):
    r_cut, n_max, l_max, pbc, _ = SOAP_CASES[case]
    got = trj.get_soap(r_cut=r_cut, n_max=n_max, l_max=l_max, respect_pbc=pbc,
                       centers=SOAP_API_CENTERS[case])  # fmt: skip
    assert got.meta["name"] == "soap" and got.meta["centers"] == SOAP_API_CENTERS[case]
    assert np.allclose(ref_soap[case], got.dataset[:, ::10])  # test_soap.py:37 (rtol 1e-5)


def test_tsoap(trj, ref_soap):
    soap = trj.get_soap(r_cut=8, n_max=4, l_max=4)
    for delay, key in ((1, "tsoap_c0"), (4, "tsoap_c2")):
        _, ts = trj.get_timesoap(soap_insight=soap, delay=delay)
        assert ts.meta["name"] == "timesoap" and ts.meta["delay"] == delay
        assert np.allclose(ref_soap[key], ts.dataset, atol=1e-6)
    soap4 = trj.get_soap(r_cut=4, n_max=4, l_max=4)
    _, ts = trj.get_timesoap(soap_insight=soap4, delay=1)
    assert np.allclose(ref_soap["tsoap_c1"], ts.dataset, atol=1e-6)


def test_get_descriptors_like_test_trj(trj, ref_lens, ref_soap):
    """tests/trajectory/test_trj.py:75-106 (hot-path descriptors only)."""
    r_cut = 15.0
    neigcounts, n_c = trj.get_coord_number(r_cut=r_cut)
    lens = trj.get_lens(r_cut=r_cut)
    soap, tsoap = trj.get_timesoap(r_cut=10.0, n_max=8, l_max=8)
    np.testing.assert_allclose(ref_lens["trj_n_c"], n_c.dataset, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(ref_lens["trj_lens"], lens.dataset, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(ref_soap["trj_soap"], soap.dataset[:, ::10], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(ref_soap["trj_tsoap"], tsoap.dataset, rtol=1e-3, atol=1e-6)
    # neighbour AtomGroups: same members as the CSR rows
    first = neigcounts[0][0]
    assert len(first) == int(n_c.dataset[0, 0])
    # second call re-using the neighbour list
    _, n_c2 = trj.get_coord_number(r_cut=r_cut, neigcounts=neigcounts)
    assert np.array_equal(n_c2.dataset, n_c.dataset)
    _, n_c3 = trj.get_coord_number(r_cut=r_cut, neigcounts=[list(fr) for fr in neigcounts[:3]])
    assert np.array_equal(n_c3.dataset, n_c.dataset[:, :3])


def test_trajslice_and_batching_match_oracle(balls7):
    """Sliced trajectories and small frame batches (halo of `delay` frames between batches)."""
    from dynsight_b200 import lens as dlens
    from dynsight_b200 import soap as dsoap
    from dynsight_b200.universe import ArrayUniverse
    from golden_cases import fluid_case

    from oracle import lens_oracle as lo
    from oracle import soap_oracle as so
    from oracle import timesoap_oracle as to

    pos, box, r_cut = fluid_case(n_side=9, n_frames=12)
    dims = np.tile(np.concatenate([box, [90, 90, 90]]).astype(np.float32), (12, 1))
    u = ArrayUniverse(pos, dims, names=["O"] * pos.shape[1])
    sl = slice(1, 12, 2)
    frames = list(range(12))[sl]
    for delay in (1, 2):
        want = lo.compute_lens_arrays(pos, dims, r_cut, delay=delay, frames=frames)
        for bf in (None, 3, 4):
            got = dlens.compute_lens_device(u, r_cut, delay, trajslice=sl, batch_frames=bf)
            assert np.array_equal(got.cpu().numpy(), want), (delay, bf)
    with pytest.raises(RuntimeError, match="No valid pairs found."):
        dlens.compute_lens(u, r_cut, delay=6, trajslice=sl)
    # SOAP -> timeSOAP without materialising SOAP, in overlapping batches
    cent = "index 0:40"
    soap = so.soap_trajectory(pos[frames] * 3.0, np.full(pos.shape[1], 8), np.arange(41),
                              dims[frames, :3] * 3.0, 4.0, 4, 3)  # fmt: skip
    u3 = ArrayUniverse(pos * 3.0, dims * np.array([3, 3, 3, 1, 1, 1], dtype=np.float32),
                       names=["O"] * pos.shape[1])  # fmt: skip
    for bf in (None, 3):
        got = dsoap.timesoap_from_trajectory(u3, 4.0, 4, 3, delay=2, centers=cent, trajslice=sl,
                                             batch_frames=bf).cpu().numpy()  # fmt: skip
        want = to.timesoap(soap, 2)
        assert np.all(np.abs(got - want) <= 1e-5 * np.abs(want) + 1e-7)


def test_soap_small_api_functions():
    from dynsight_b200 import soap as dsoap
    from golden_cases import random_soap_spectra

    from oracle import timesoap_oracle as to

    s = random_soap_spectra()
    assert np.allclose(dsoap.normalize_soap(s), to.normalize_soap(s), rtol=1e-12, atol=0)
    d = dsoap.soap_distance(s[:, 0], s[:, 1])
    assert np.allclose(d, to.soap_distance(s[:, 0], s[:, 1]), atol=1e-7)
    assert np.isclose(dsoap.soap_distance(s[7, 0], s[7, 1]), to.soap_distance(s[7, 0], s[7, 1]))
    assert np.allclose(dsoap.timesoap(s, 2), to.timesoap(s, 2), atol=1e-7)
    with pytest.raises(ValueError, match="delay value outside bounds"):
        dsoap.timesoap(s, 9)


def test_multi_species_via_types():
    """Species come from atom types; order = ascending atomic number (DScribe), unpinned by the
    reference, checked against the oracle's convention."""
    from dynsight_b200 import soap as dsoap
    from dynsight_b200.universe import ArrayUniverse

    from oracle import soap_oracle as so

    rng = np.random.default_rng(8)
    n = 120
    box = np.array([26.5, 27.0, 28.0], dtype=np.float32)
    pos = (rng.random((2, n, 3)) * box).astype(np.float32)
    types = np.array(["O", "H", "Cu"])[rng.integers(0, 3, n)]
    types[:3] = ["Cu", "H", "O"]
    dims = np.tile(np.concatenate([box, [90, 90, 90]]).astype(np.float32), (2, 1))
    u = ArrayUniverse(pos, dims, names=types, types=types)
    got = dsoap.saponify_trajectory(u, 4.5, 4, 4, centers="type O or type Cu")
    z = np.array([{"H": 1, "O": 8, "Cu": 29}[t] for t in types])
    cent = np.nonzero((types == "O") | (types == "Cu"))[0]
    want = so.soap_trajectory(pos, z, cent, dims[:, :3].astype(float), 4.5, 4, 4)
    assert got.shape == want.shape == (len(cent), 2, 3 * 4 * (3 * 4 + 1) // 2 * 5)
    scale = np.abs(want).max(axis=-1, keepdims=True)
    assert (np.abs(got - want) / scale).max() < 1e-8
    with pytest.raises(KeyError, match="not a chemical symbol"):
        dsoap.saponify_trajectory(ArrayUniverse(pos, dims, names=["Qq"] * n, types=["Qq"] * n), 4.5)


def test_psi_phi_like_test_trj(trj, ref_lens):
    """tests/trajectory/test_trj.py:80,92,97-99 through Trj."""
    neigcounts, phi = trj.get_velocity_alignment(r_cut=15.0)
    assert phi.meta == {"name": "velocity_alignement", "r_cut": 15.0, "centers": "all",
                        "selection": "all"}  # fmt: skip
    np.testing.assert_allclose(ref_lens["trj_phi"], phi.dataset, rtol=1e-4, atol=1e-6)
    _, psi = trj.get_orientational_op(r_cut=15.0, neigcounts=neigcounts)
    assert psi.meta["name"] == "orientational_op" and psi.meta["order"] == 6
    np.testing.assert_allclose(ref_lens["trj_psi"], psi.dataset, rtol=1e-3, atol=1e-6)
    # a plain list[list[AtomGroup]] works too
    plain = [list(fr) for fr in neigcounts]
    from dynsight_b200 import descriptors

    psi2 = descriptors.orientational_order_param(trj.universe, plain, order=6)
    assert np.array_equal(psi2, psi.dataset)


def test_psi_phi_doctest_and_random_vs_oracle(xyz60, ref_lens):
    from dynsight_b200 import descriptors
    from dynsight_b200.trajectory import Trj
    from dynsight_b200.universe import ArrayUniverse
    from golden_cases import fluid_case

    from oracle import descriptors_oracle as do
    from oracle import lens_oracle as lo

    t60 = Trj(ArrayUniverse(xyz60["positions"], None, names=xyz60["types"]))
    neig, _ = t60.get_coord_number(r_cut=3.0)
    psi = descriptors.orientational_order_param(t60.universe, neig)
    phi = descriptors.velocity_alignment(t60.universe, neig)
    assert np.isclose(psi[0][0], float(ref_lens["kat_psi00"]))
    assert np.isclose(phi[0][0], float(ref_lens["kat_phi00"]))
    # random fluid, several orders, against the restated reference loops
    pos, box, r_cut = fluid_case(n_side=7, n_frames=4)
    pos[2, 5] = pos[0, 5]  # a zero displacement vector
    dims = np.tile(np.concatenate([box, [90, 90, 90]]).astype(np.float32), (4, 1))
    u = ArrayUniverse(pos, dims)
    tr = Trj(u)
    neig, _ = tr.get_coord_number(r_cut=r_cut)
    csr = [lo.neighbors(pos[f], pos[f], r_cut, box, True) for f in range(4)]
    for order in (1, 4, 6):
        got = descriptors.orientational_order_param(u, neig, order=order)
        np.testing.assert_allclose(got, do.orientational_order_param(pos, csr, order), rtol=1e-5,
                                   atol=1e-6)  # fmt: skip
    np.testing.assert_allclose(descriptors.velocity_alignment(u, neig),
                               do.velocity_alignment(pos, csr), rtol=1e-5, atol=1e-6)  # fmt: skip


def test_coord_number_does_not_materialise_the_lists(trj, ref_lens):
    """Trj.get_coord_number needs the counts only: the canonical CSR (two traversals, row sorts,
    device -> host copy of all indices) is built on the first access to an element."""
    neigh, got = trj.get_coord_number(r_cut=4, centers="all", selection="all")
    assert neigh._batch is None and neigh._csr is None  # noqa: SLF001
    assert np.array_equal(ref_lens["nn_c1"], got.dataset)
    first = neigh[0][0]
    assert neigh._batch is not None and neigh._csr is not None  # noqa: SLF001
    assert len(first) == int(got.dataset[0, 0])
    assert [len(ag) for ag in neigh[5]] == [int(v) for v in got.dataset[:, 5]]


def test_empty_selections_behave_like_the_reference(trj):
    """No centres -> empty arrays; no environment atoms -> every neighbour row is empty, LENS 0.0
    (lens.py:171-173) and coordination number 0."""
    from dynsight_b200 import lens as dlens

    uni = trj.universe
    got = dlens.compute_lens(uni, 4.0, centers="name ZZ")
    assert got.shape == (0, 200) and got.dtype == np.float64
    got = dlens.compute_lens(uni, 4.0, selection="name ZZ")
    assert got.shape == (7, 200) and not got.any()
    neigh, nc = trj.get_coord_number(r_cut=4, selection="name ZZ")
    assert nc.dataset.shape == (7, 201) and not nc.dataset.any()
    assert len(neigh) == 201 and len(neigh[3]) == 7 and len(neigh[3][2]) == 0
    soap = trj.get_soap(r_cut=4.0, n_max=4, l_max=4, centers="name ZZ")
    assert soap.dataset.shape[0] == 0 and soap.dataset.shape[1] == 201


def test_load_or_compute_soap_roundtrip(trj, tmp_path):
    """utilities.load_or_compute_soap (utilities.py:67-103): computes and saves, then loads."""
    from dynsight_b200 import utilities

    path = tmp_path / "soap.json"
    first = utilities.load_or_compute_soap(trj.with_slice(slice(0, 5)), 4.0, 4, 4, soap_path=path)
    assert path.exists() and first.meta["name"] == "soap"
    again = utilities.load_or_compute_soap(trj, 99.0, 1, 1, soap_path=path)  # loaded, not computed
    assert np.array_equal(first.dataset, again.dataset) and again.meta == first.meta


def test_cfg1_full_size_lens_and_coordination_bit_exact():
    """BASELINE cfg1 at full size (2 197 particles x 200 frames, r_cut = 1.5 sigma): LENS and the
    coordination numbers through the Trj API equal the oracle bit for bit."""
    from golden_cases import fluid_case
    from oracle import lens_oracle as lo

    from dynsight_b200.trajectory import Trj
    from dynsight_b200.universe import ArrayUniverse

    pos, box, r_cut = fluid_case(n_side=13, n_frames=200)
    dims = np.tile(np.concatenate([box, [90, 90, 90]]).astype(np.float32), (200, 1))
    t = Trj(ArrayUniverse(pos, dims, ["A"] * pos.shape[1], ["A"] * pos.shape[1]))
    lens = t.get_lens(r_cut=r_cut)
    _, n_c = t.get_coord_number(r_cut=r_cut)
    assert np.array_equal(lens.dataset, lo.compute_lens_arrays(pos, dims, r_cut))
    assert np.array_equal(n_c.dataset, lo.coord_number_arrays(pos, dims, r_cut))


@pytest.mark.parametrize("scale", [1.0, 1.35, 1.8, 2.4])
def test_psi_phi_lane_groups_vs_oracle(scale):
    """The psi / phi kernels use 4, 8, 16 or 32 lanes per atom depending on the mean row length:
    every variant against the restated reference loops (rows of ~11 ... ~150 entries)."""
    from golden_cases import fluid_case
    from oracle import descriptors_oracle as do
    from oracle import lens_oracle as lo

    from dynsight_b200 import descriptors
    from dynsight_b200.trajectory import Trj
    from dynsight_b200.universe import ArrayUniverse

    pos, box, r_cut = fluid_case(n_side=8, n_frames=3)
    r_cut *= scale
    dims = np.tile(np.concatenate([box, [90, 90, 90]]).astype(np.float32), (3, 1))
    u = ArrayUniverse(pos, dims)
    neig, n_c = Trj(u).get_coord_number(r_cut=r_cut)
    csr = [lo.neighbors(pos[f], pos[f], r_cut, box, True) for f in range(3)]
    assert abs(n_c.dataset.mean() - np.mean([np.diff(c[0]).mean() for c in csr])) < 1e-9
    np.testing.assert_allclose(descriptors.orientational_order_param(u, neig, order=6),
                               do.orientational_order_param(pos, csr, 6), rtol=1e-5, atol=1e-6)  # fmt: skip
    np.testing.assert_allclose(descriptors.velocity_alignment(u, neig),
                               do.velocity_alignment(pos, csr), rtol=1e-5, atol=1e-6)  # fmt: skip


def test_psi_phi_with_a_subset_selection_match_the_plain_list_path():
    """The lazy NeighbourList stores positions inside the environment selection; the psi / phi
    kernels index the universe's arrays, so rows are translated to universe indices like the
    reference's ``neigh[t][i].indices`` (misc.py:79-88, 190-205)."""
    from dynsight_b200 import descriptors
    from dynsight_b200.trajectory import Trj
    from dynsight_b200.universe import ArrayUniverse
    from golden_cases import fluid_case

    pos, box, r_cut = fluid_case(n_side=7, n_frames=4)
    n = pos.shape[1]
    dims = np.tile(np.concatenate([box, [90, 90, 90]]).astype(np.float32), (4, 1))
    types = np.where(np.arange(n) % 3 == 0, "C", "O")
    u = ArrayUniverse(pos, dims, names=types, types=types)
    tr = Trj(u)
    neig, _ = tr.get_coord_number(r_cut=r_cut, centers="all", selection="type O")
    plain = [list(fr) for fr in neig]   # list[list[AtomGroup]]: .indices are universe indices
    assert not np.array_equal(neig.env_indices, np.arange(n))
    for order in (4, 6):
        lazy = descriptors.orientational_order_param(u, neig, order=order)
        assert np.array_equal(lazy, descriptors.orientational_order_param(u, plain, order=order))
    assert np.array_equal(descriptors.velocity_alignment(u, neig),
                          descriptors.velocity_alignment(u, plain))  # fmt: skip


def test_pipeline_one_pass_equals_the_separate_calls(balls7):
    """pipeline.descriptors_from_trajectory feeds every device batch to the LENS and the timeSOAP
    job: same values as compute_lens / Trj.get_coord_number / timesoap_from_trajectory, for the
    pair path, the rows path and explicit centres, with small batches that overlap."""
    from dynsight_b200 import lens as dlens
    from dynsight_b200 import pipeline
    from dynsight_b200 import soap as dsoap
    from dynsight_b200.trajectory import Trj
    from dynsight_b200.universe import ArrayUniverse

    rng = np.random.default_rng(4)
    n, nf = 700, 9
    box = np.array([28.0, 29.0, 27.5], dtype=np.float32)
    pos = (rng.random((1, n, 3)) * box + rng.normal(0, 0.2, (nf, n, 3)).cumsum(0)).astype(np.float32)
    dims = np.tile(np.concatenate([box, [90, 90, 90]]).astype(np.float32), (nf, 1))
    types = np.where(np.arange(n) % 4 == 0, "C", "O")
    uni = ArrayUniverse(pos, dims, names=types, types=["O"] * n)
    for kw, r_lens in ((dict(), 4.0), (dict(), 9.5), (dict(centers="name C"), 4.0)):
        for delay in (1, 2):
            res = pipeline.descriptors_from_trajectory(
                uni, lens_r_cut=r_lens, soap_r_cut=4.5, soapnmax=4, soaplmax=4, delay=delay,
                coord_number=True, batch_frames=4, **kw,
            )
            assert np.array_equal(res["lens"], dlens.compute_lens(uni, r_lens, delay=delay, **kw))
            _, nc = Trj(uni).get_coord_number(r_cut=r_lens, **kw)
            assert np.array_equal(res["coord_number"], nc.dataset)
            want_ts = dsoap.timesoap_from_trajectory(uni, 4.5, 4, 4, delay=delay, as_numpy=True,
                                                     **kw)  # fmt: skip
            # SOAP sums run in cell order, and the order of atoms inside a cell (atomic scatter)
            # differs from launch to launch: equal to rounding, not bit for bit
            assert np.abs(res["timesoap"] - want_ts).max() < 1e-12
    # default batch size: a short trajectory is uploaded before the jobs are set up and the LENS
    # kernels are enqueued before the SOAP job exists (the early path of the pipeline); also with a
    # sub-selection, a frame slice, LENS only and timeSOAP only
    for kw in (dict(), dict(selection="name O"), dict(trajslice=slice(1, 8)),
               dict(centers="name C", trajslice=slice(0, 9, 2))):
        res = pipeline.descriptors_from_trajectory(uni, lens_r_cut=4.0, soap_r_cut=4.5, soapnmax=4,
                                                   soaplmax=4, coord_number=True, **kw)  # fmt: skip
        assert np.array_equal(res["lens"], dlens.compute_lens(uni, 4.0, **kw))
        trj_kw = {k: v for k, v in kw.items() if k != "trajslice"}
        _, nc = Trj(uni).with_slice(kw.get("trajslice")).get_coord_number(r_cut=4.0, **trj_kw)
        assert np.array_equal(res["coord_number"], nc.dataset)
        want_ts = dsoap.timesoap_from_trajectory(uni, 4.5, 4, 4, as_numpy=True, **kw)
        assert np.abs(res["timesoap"] - want_ts).max() < 1e-12
    only_lens = pipeline.descriptors_from_trajectory(uni, lens_r_cut=4.0)
    assert set(only_lens) == {"lens"}
    assert np.array_equal(only_lens["lens"], dlens.compute_lens(uni, 4.0))
    only_ts = pipeline.descriptors_from_trajectory(uni, soap_r_cut=4.5, soapnmax=4, soaplmax=4)
    assert set(only_ts) == {"timesoap"}
    assert np.abs(only_ts["timesoap"] - dsoap.timesoap_from_trajectory(uni, 4.5, 4, 4, as_numpy=True)).max() < 1e-12
    with pytest.raises(ValueError, match="nothing to compute"):
        pipeline.descriptors_from_trajectory(uni)
    with pytest.raises(ValueError, match="delay value outside bounds"):
        pipeline.descriptors_from_trajectory(uni, soap_r_cut=4.5, delay=9)
