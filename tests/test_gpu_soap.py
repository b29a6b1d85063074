"""GPU parity: SOAP power spectrum through the C ABI vs the reference's DScribe goldens and the
CPU oracle.

Tolerance (BASELINE.json north_star: 1e-5 relative): |d| <= 1e-5 * max|p| per vector; the kernel
is float64 so the observed error is ~1e-10, and the tests assert a 1e-8 bar to catch regressions.
"""

from __future__ import annotations

import numpy as np
import pytest
import torch
from test_oracle_golden import SOAP_CASES

from oracle import soap_oracle as so
from oracle import timesoap_oracle as to

pytestmark = pytest.mark.gpu

REL = 1e-8


def _rel_err(got: np.ndarray, want: np.ndarray) -> float:
    scale = np.abs(want).max(axis=-1, keepdims=True)
    return float((np.abs(got - want) / scale).max())


def _soap_gpu(pos, species_idx, centres, box, r_cut, n_max, l_max, dtype=torch.float32):
    from dynsight_b200 import engine

    return (
        engine.soap_power_spectrum(
            torch.as_tensor(np.ascontiguousarray(pos)).to(dtype).cuda(),
            torch.as_tensor(np.asarray(species_idx, dtype=np.int32)).cuda(),
            torch.as_tensor(np.asarray(centres, dtype=np.int32)).cuda(),
            box, r_cut, n_max, l_max,
        )
        .cpu()
        .numpy()
    )  # fmt: skip


@pytest.mark.parametrize("case", sorted(SOAP_CASES))
def test_dscribe_goldens(balls7, ref_soap, case):
    """tests/soap/conftest.py:6-74 + tests/trajectory/test_trj.py:82 (golden frames 0::10)."""
    r_cut, n_max, l_max, pbc, cent = SOAP_CASES[case]
    pos = balls7["positions"]
    box = balls7["dims"][:, :3] if pbc else None
    got = _soap_gpu(pos, np.zeros(7), cent, box, r_cut, n_max, l_max)
    gold = ref_soap[case]
    assert got.shape == (len(cent), 201, gold.shape[2])
    assert _rel_err(got[:, ::10], gold) < REL


def test_tsoap_goldens_end_to_end(balls7, ref_soap):
    """tests/tsoap/conftest.py:6-35 and test_trj.py:104-106: SOAP kernel -> timeSOAP kernel."""
    from dynsight_b200 import engine

    pos = torch.as_tensor(balls7["positions"]).cuda()
    sp = torch.zeros(7, dtype=torch.int32, device="cuda")
    cen = torch.arange(7, dtype=torch.int32, device="cuda")
    box = balls7["dims"][:, :3]
    soap44 = engine.soap_power_spectrum(pos, sp, cen, box, 8.0, 4, 4)
    t1 = engine.timesoap_device(soap44, 1).cpu().numpy()
    t4 = engine.timesoap_device(soap44, 4).cpu().numpy()
    assert np.allclose(t1, ref_soap["tsoap_c0"], atol=1e-6, rtol=0)
    assert np.allclose(t4, ref_soap["tsoap_c2"], atol=1e-6, rtol=0)
    soap88 = engine.soap_power_spectrum(pos, sp, cen, box, 10.0, 8, 8)
    t = engine.timesoap_device(soap88, 1).cpu().numpy()
    assert np.allclose(t, ref_soap["trj_tsoap"], atol=1e-6, rtol=1e-3)
    # rc=4: no neighbours inside the cutoff region -> identical spectra -> timeSOAP ~ 0 (golden c1)
    soap_rc4 = engine.soap_power_spectrum(pos, sp, cen, box, 4.0, 4, 4)
    t = engine.timesoap_device(soap_rc4, 1).cpu().numpy()
    assert np.allclose(t, ref_soap["tsoap_c1"], atol=1e-6, rtol=0)


def test_doctest_known_answers(xyz60):
    """saponify.py:86-87 / :174 and timesoap.py:50-52, :115-116, :172-173 (rtol 1e-3), no box."""
    from dynsight_b200 import engine

    soap_t = engine.soap_power_spectrum(
        torch.as_tensor(xyz60["positions"]).cuda(),
        torch.zeros(60, dtype=torch.int32, device="cuda"),
        torch.arange(60, dtype=torch.int32, device="cuda"),
        None, 2.0, 8, 8,
    )  # fmt: skip
    soap = soap_t.cpu().numpy()
    assert soap.shape == (60, 20, 324)
    assert np.isclose(soap[0].sum(), float(xyz60["kat_soap0_sum"]), atol=1e-6, rtol=1e-3)
    tsoap = engine.timesoap_device(soap_t, 1).cpu().numpy()
    assert np.isclose(tsoap.sum(), float(xyz60["kat_tsoap_sum"]), atol=1e-6, rtol=1e-3)
    assert np.isclose(tsoap[0, 0], float(xyz60["kat_dist01"]), atol=1e-6, rtol=1e-3)
    want = so.soap_trajectory(xyz60["positions"][:2], np.zeros(60, int), np.arange(60), None, 2.0, 8, 8)
    assert _rel_err(soap[:, :2], want) < REL


@pytest.mark.parametrize(
    ("n_atoms", "box_l", "r_cut", "n_max", "l_max", "n_species", "dtype"),
    [
        (400, 31.0, 5.0, 8, 8, 1, torch.float32),   # wide cells on all axes (31 > 3 * 8.72)
        (300, 19.0, 5.0, 8, 8, 1, torch.float64),   # narrow: explicit image loop, float64 input
        (60, 7.5, 4.0, 4, 5, 1, torch.float32),     # box < search radius: several images per atom
        (250, 27.0, 5.0, 8, 10, 3, torch.float32),  # 3 species, l_max = 10 (cfg4 parameters)
        (200, 28.0, 4.5, 6, 3, 2, torch.float32),   # 2 species, odd l_max
        (150, 30.0, 5.5, 12, 2, 1, torch.float32),  # n_max > 8: two radial blocks
        (220, 33.0, 6.0, 3, 12, 2, torch.float32),  # l_max = 12
        (180, 29.0, 5.0, 9, 4, 2, torch.float32),   # odd n_max > 8 with 2 species (split mode)
        (120, 12.0, 4.0, 5, 6, 3, torch.float64),   # narrow box, 3 species
        # list mode (one species, >= 3 cells per axis) for every l_max instantiation
        (300, 30.0, 5.0, 7, 4, 1, torch.float32),
        (300, 30.0, 5.0, 8, 6, 1, torch.float32),
        (300, 30.0, 5.0, 8, 7, 1, torch.float64),
        (200, 30.0, 5.0, 6, 10, 1, torch.float32),
        (200, 30.0, 5.0, 4, 12, 1, torch.float32),
    ],
)
def test_random_boxes_vs_oracle(n_atoms, box_l, r_cut, n_max, l_max, n_species, dtype):
    rng = np.random.default_rng(n_atoms + n_max * 100 + l_max)
    box = np.array([box_l, box_l * 1.1, box_l * 0.93])
    pos = (rng.random((2, n_atoms, 3)) * 1.3 - 0.15) * box  # some atoms outside the box
    if dtype == torch.float32:
        pos = pos.astype(np.float32)
    sp_idx = rng.integers(0, n_species, n_atoms)
    sp_idx[:n_species] = np.arange(n_species)
    zs = np.array([1, 8, 29])[:n_species]
    centres = np.sort(rng.choice(n_atoms, size=min(n_atoms, 23), replace=False))
    got = _soap_gpu(pos, sp_idx, centres, box, r_cut, n_max, l_max, dtype)
    want = so.soap_trajectory(
        pos, zs[sp_idx], centres, np.tile(box, (2, 1)), r_cut, n_max, l_max
    )
    assert got.shape == want.shape
    assert _rel_err(got, want) < REL


def test_list_rows_overflow_falls_back_to_the_in_kernel_search():
    """Neighbour rows are sized from the mean density; a dense droplet in a dilute box overflows
    them and the launch must fall back to the in-kernel search with identical results."""
    rng = np.random.default_rng(21)
    box = np.array([40.0, 41.0, 39.0])
    pos = rng.random((2, 400, 3)) * box
    pos[:, :170] = box / 2 + rng.normal(size=(2, 170, 3)) * 2.0   # 170 atoms within a few A
    pos = pos.astype(np.float32)
    centres = np.arange(0, 400, 9)
    got = _soap_gpu(pos, np.zeros(400), centres, box, 5.0, 8, 8)
    want = so.soap_trajectory(pos, np.zeros(400, int), centres, np.tile(box, (2, 1)), 5.0, 8, 8)
    assert _rel_err(got, want) < REL
    # and a frame batch in which only one frame overflows
    pos[1] = (rng.random((400, 3)) * box).astype(np.float32)
    got = _soap_gpu(pos, np.zeros(400), centres, box, 5.0, 8, 8)
    want = so.soap_trajectory(pos, np.zeros(400, int), centres, np.tile(box, (2, 1)), 5.0, 8, 8)
    assert _rel_err(got, want) < REL
    # species-major rows (split mode): two species, the droplet still overflows
    sp_idx = rng.integers(0, 2, 400)
    got = _soap_gpu(pos, sp_idx, centres, box, 5.0, 6, 5)
    want = so.soap_trajectory(
        pos, np.array([1, 8])[sp_idx], centres, np.tile(box, (2, 1)), 5.0, 6, 5
    )
    assert _rel_err(got, want) < REL


def test_non_periodic_two_species_vs_oracle():
    rng = np.random.default_rng(9)
    pos = (rng.normal(size=(2, 400, 3)) * 8.0).astype(np.float32)
    sp_idx = rng.integers(0, 2, 400)
    centres = np.arange(0, 400, 13)
    got = _soap_gpu(pos, sp_idx, centres, None, 5.0, 6, 7)
    want = so.soap_trajectory(pos, np.array([6, 79])[sp_idx], centres, None, 5.0, 6, 7)
    assert _rel_err(got, want) < REL


def test_non_periodic_cluster_vs_oracle():
    rng = np.random.default_rng(4)
    pos = (rng.normal(size=(3, 500, 3)) * 9.0).astype(np.float32)
    centres = np.arange(0, 500, 17)
    got = _soap_gpu(pos, np.zeros(500), centres, None, 5.0, 8, 8)
    want = so.soap_trajectory(pos, np.zeros(500, int), centres, None, 5.0, 8, 8)
    assert _rel_err(got, want) < REL


def test_invariances_at_bench_scale():
    """cfg2-like frame (10k atoms): rotation + translation + permutation invariance of the power
    spectrum, and agreement with the oracle on a few centres."""
    from dynsight_b200 import engine

    rng = np.random.default_rng(0)
    n = 10_000
    box_l = (n / 0.0334) ** (1 / 3)
    pos = (rng.random((1, n, 3)) * box_l).astype(np.float64)
    box = np.full(3, box_l)
    sp = torch.zeros(n, dtype=torch.int32, device="cuda")
    cen = torch.arange(n, dtype=torch.int32, device="cuda")
    base = engine.soap_power_spectrum(torch.as_tensor(pos).cuda(), sp, cen, box, 5.0, 8, 8)
    # cubic-group rotation (x,y,z)->(y,z,x) keeps the box, plus a translation and a permutation
    perm = rng.permutation(n)
    pos2 = (pos[:, :, [1, 2, 0]] + np.array([3.3, -7.1, 11.9]))[:, perm]
    other = engine.soap_power_spectrum(torch.as_tensor(pos2).cuda(), sp, cen, box, 5.0, 8, 8)
    a = base.cpu().numpy()[perm, 0]
    b = other.cpu().numpy()[:, 0]
    assert _rel_err(b, a) < 1e-9
    pick = np.arange(0, n, n // 12)[:12]
    want = so.soap_trajectory(pos, np.zeros(n, int), pick, np.tile(box, (1, 1)), 5.0, 8, 8)
    assert _rel_err(base.cpu().numpy()[pick], want) < REL
    # timeSOAP of identical frames is (numerically) zero
    two = torch.cat([base, base], dim=1)
    assert float(engine.timesoap_device(two, 1).abs().max()) <= 1e-7
    _ = to


def test_multi_species_invariances_at_cfg4_parameters():
    """BASELINE cfg4 parameters (3 species by index hash, l_max = 10, metal density) on 8000 atoms:
    split mode (accumulate kernel + spectrum kernel).  Permutation + translation + axis rotation
    leave the spectra unchanged and a few centres agree with the oracle."""
    from dynsight_b200 import engine

    rng = np.random.default_rng(4)
    n = 8_000
    box_l = (n / 0.0857) ** (1 / 3)
    pos = (rng.random((1, n, 3)) * box_l).astype(np.float64)
    box = np.full(3, box_l)
    sp_idx = (np.arange(n) * 7919) % 3
    cen = torch.arange(n, dtype=torch.int32, device="cuda")
    base = engine.soap_power_spectrum(
        torch.as_tensor(pos).cuda(), torch.as_tensor(sp_idx, dtype=torch.int32).cuda(), cen, box,
        5.0, 8, 10, n_species=3,
    )  # fmt: skip
    assert base.shape == (n, 1, 3300)
    perm = rng.permutation(n)
    pos2 = (pos[:, :, [2, 0, 1]] + np.array([-4.2, 9.9, 1.7]))[:, perm]
    other = engine.soap_power_spectrum(
        torch.as_tensor(pos2).cuda(), torch.as_tensor(sp_idx[perm], dtype=torch.int32).cuda(), cen,
        box, 5.0, 8, 10, n_species=3,
    )  # fmt: skip
    assert _rel_err(other.cpu().numpy()[:, 0], base.cpu().numpy()[perm, 0]) < 1e-9
    pick = np.arange(0, n, n // 6)[:6]
    zs = np.array([13, 28, 29])
    want = so.soap_trajectory(pos, zs[sp_idx], pick, np.tile(box, (1, 1)), 5.0, 8, 10)
    assert _rel_err(base.cpu().numpy()[pick], want) < REL


def test_cfg2_full_size_materialised_equals_streamed():
    """BASELINE cfg2 at full size (10 000 atoms x 1 000 frames, SOAP(5 A, 8, 8): a 25.9 GB tensor):
    SOAP materialised on the device + timeSOAP equals the streamed ``timesoap_from_trajectory``
    (frame batches overlapping by ``delay``), and every spectrum is finite with unit-safe norms."""
    from dynsight_b200 import engine, soap
    from dynsight_b200.universe import ArrayUniverse

    n, nf = 10_000, 1_000
    box_l = (n / 0.0334) ** (1 / 3)
    g = torch.Generator(device="cuda").manual_seed(2)
    cur = torch.rand((n, 3), device="cuda", generator=g) * box_l
    steps = 0.1 * torch.randn((nf, n, 3), device="cuda", generator=g)
    pos = torch.remainder(cur[None] + steps.cumsum(0), box_l).float()
    pos.clamp_(max=float(np.nextafter(np.float32(box_l), np.float32(0))))
    dims = np.tile(np.array([box_l] * 3 + [90] * 3, dtype=np.float32), (nf, 1))
    uni = ArrayUniverse(pos.cpu().numpy(), dims, ["O"] * n, ["O"] * n)
    del steps
    spectra = soap.saponify_trajectory_device(uni, 5.0, 8, 8)
    assert spectra.shape == (n, nf, 324)
    assert bool(torch.isfinite(spectra[:, ::97]).all())
    full = engine.timesoap_device(spectra, 2)
    del spectra
    streamed = soap.timesoap_from_trajectory(uni, 5.0, 8, 8, delay=2, batch_frames=128)
    assert full.shape == streamed.shape == (n, nf - 2)
    assert float((full - streamed).abs().max()) < 1e-9
    assert 0.0 < float(full.mean()) < 1.5


# ---------------------------------------------------------------------------------------------
# multi-species layout: pins DERIVED from single-species runs (no reference golden has S > 1;
# DScribe is called at saponify.py:96-122 with species = set(types))
# ---------------------------------------------------------------------------------------------
def _blocks(spec, n_species, n_max, l_max):
    """Split a multi-species spectrum (..., C) into {(zi, zj): (..., l+1, n, n)} full matrices
    (diagonal blocks symmetrised from their upper triangle)."""
    out, pos = {}, 0
    iu = np.triu_indices(n_max)
    for zi in range(n_species):
        for zj in range(zi, n_species):
            if zi == zj:
                size = (l_max + 1) * iu[0].size
                tri = spec[..., pos : pos + size].reshape(*spec.shape[:-1], l_max + 1, -1)
                full = np.zeros((*spec.shape[:-1], l_max + 1, n_max, n_max))
                full[..., iu[0], iu[1]] = tri
                full[..., iu[1], iu[0]] = tri
            else:
                size = (l_max + 1) * n_max * n_max
                full = spec[..., pos : pos + size].reshape(*spec.shape[:-1], l_max + 1, n_max, n_max)
            out[(zi, zj)] = full
            pos += size
    assert pos == spec.shape[-1]
    return out


def test_multi_species_blocks_sum_to_the_single_species_spectrum():
    """Relabelling the atoms of a one-species system into three species must not change
    c_nlm = sum_Z c^Z_nlm, hence p_lnn' = sum_Z p^ZZ_lnn' + sum_{Z<Z'} (p^ZZ'_lnn' + p^ZZ'_ln'n):
    pins the content and the (l, n, n') layout of every block against the single-species kernel
    (itself pinned by the seven DScribe goldens)."""
    from dynsight_b200 import soap as dsoap
    from dynsight_b200.universe import ArrayUniverse

    rng = np.random.default_rng(21)
    n, n_max, l_max = 300, 5, 4
    box = np.array([27.0, 28.5, 26.2], dtype=np.float32)
    pos = (rng.random((2, n, 3)) * box).astype(np.float32)
    dims = np.tile(np.concatenate([box, [90, 90, 90]]).astype(np.float32), (2, 1))
    one = dsoap.saponify_trajectory(ArrayUniverse(pos, dims, names=["O"] * n, types=["O"] * n),
                                    4.5, n_max, l_max)  # fmt: skip
    types = np.array(["Cu", "H", "O"])[rng.integers(0, 3, n)]
    three = dsoap.saponify_trajectory(ArrayUniverse(pos, dims, names=types, types=types), 4.5,
                                      n_max, l_max)  # fmt: skip
    blk = _blocks(three, 3, n_max, l_max)
    total = sum(blk[(z, z)] for z in range(3))
    for zi in range(3):
        for zj in range(zi + 1, 3):
            total = total + blk[(zi, zj)] + np.swapaxes(blk[(zi, zj)], -1, -2)
    iu = np.triu_indices(n_max)
    got = total[..., iu[0], iu[1]].reshape(n, 2, -1)
    scale = np.abs(one).max(axis=-1, keepdims=True)
    assert (np.abs(got - one) / scale).max() < 1e-10


def test_multi_species_block_order_and_orientation_derived():
    """Which block is (Z, Z') and which index of a cross block belongs to which species, derived
    from single-species runs.  All neighbours sit on the z axis through the centre, so only m = 0
    survives and p^ZZ'_lnn' = u_n v_n' factorises for l >= 1; u and v follow (up to a sign) from the
    single-species spectra of {centre + near atoms} and {centre + far atoms}.  Near and far atoms
    excite different radial functions, so u v^T != v u^T.  Species order = ascending atomic number:
    swapping the labels swaps the diagonal blocks and transposes the cross block."""
    from dynsight_b200 import soap as dsoap
    from dynsight_b200.universe import ArrayUniverse

    n_max, l_max, r_cut = 4, 2, 5.0
    box = np.array([60.0, 60.0, 60.0], dtype=np.float32)
    c0 = np.array([30.0, 30.0, 30.0])
    near = c0 + np.array([[0, 0, 1.0], [0, 0, -1.3]])
    far = c0 + np.array([[0, 0, 4.2], [0, 0, -3.9]])
    dims = np.concatenate([box, [90, 90, 90]]).astype(np.float32)[None]

    def run(points, types):
        pos = np.asarray(points, dtype=np.float32)[None]
        u = ArrayUniverse(pos, dims, names=list(types), types=list(types))
        return dsoap.saponify_trajectory(u, r_cut, n_max, l_max)[0, 0]   # centre = atom 0

    iu = np.triu_indices(n_max)

    def factor(spec, l):
        tri = spec.reshape(l_max + 1, -1)[l]
        m = np.zeros((n_max, n_max))
        m[iu] = tri
        m = m + m.T - np.diag(np.diag(m))
        k = int(np.argmax(np.diag(m)))
        return m[k] / np.sqrt(m[k, k])

    p_near = run(np.vstack([c0, near]), ["O"] * 3)
    p_far = run(np.vstack([c0, far]), ["O"] * 3)
    pts = np.vstack([c0, near, far])
    # centre + near atoms are hydrogen (Z = 1), far atoms copper (Z = 29): H is species 0
    lo = _blocks(run(pts, ["H", "H", "H", "Cu", "Cu"]), 2, n_max, l_max)
    # the other way round: centre + near atoms copper, far atoms hydrogen: H (far) is species 0
    hi = _blocks(run(pts, ["Cu", "Cu", "Cu", "H", "H"]), 2, n_max, l_max)
    for l in (1, 2):
        u, v = factor(p_near, l), factor(p_far, l)
        uv = np.outer(u, v)
        scale = np.abs(uv).max()
        assert np.abs(uv - uv.T).max() > 0.05 * scale          # the configuration is polarised
        for blocks, a, b in ((lo, u, v), (hi, v, u)):         # species 0 = the hydrogen atoms
            cross = blocks[(0, 1)][l]
            want = np.outer(a, b)
            err = min(np.abs(cross - want).max(), np.abs(cross + want).max())
            assert err < 1e-9 * scale, (l, err)
            assert np.abs(blocks[(0, 0)][l] - np.outer(a, a)).max() < 1e-9 * scale
            assert np.abs(blocks[(1, 1)][l] - np.outer(b, b)).max() < 1e-9 * scale


# ---------------------------------------------------------------------------------------------
# fused SOAP -> timeSOAP (dsg_soap_timesoap): frame-sequential warps, ring of `delay` vectors
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize(
    ("n_species", "n_max", "l_max", "periodic"),
    [(1, 8, 8, True), (1, 4, 4, True), (1, 6, 6, False), (3, 8, 10, True), (2, 4, 3, True),
     (1, 12, 5, True), (4, 5, 4, True), (2, 8, 12, False)],
)
def test_fused_timesoap_equals_two_kernel_path(n_species, n_max, l_max, periodic):
    """engine.soap_timesoap == soap_power_spectrum + timesoap_device (to rounding) for the fused
    single-species kernels (list mode and in-kernel search), split mode (several species,
    n_max > 8), delays 1..3, runs that split the frame axis, a subset of centres, and with the
    spectra requested as well."""
    from dynsight_b200 import engine

    rng = np.random.default_rng(17 + n_species + n_max)
    n, n_frames = 260, 23
    box = np.array([27.0, 28.5, 26.5], dtype=np.float32)
    pos = (rng.random((1, n, 3)) * box + rng.normal(0, 0.15, (n_frames, n, 3)).cumsum(0)).astype(
        np.float32)
    dpos = torch.as_tensor(pos).cuda()
    species = torch.as_tensor(rng.integers(0, n_species, n).astype(np.int32)).cuda()
    centres = torch.as_tensor(np.sort(rng.choice(n, 37, replace=False)).astype(np.int32)).cuda()
    bx = box if periodic else None
    ref = engine.soap_power_spectrum(dpos, species, centres, bx, 4.5, n_max, l_max,
                                     n_species=n_species)
    for delay in (1, 2, 3):
        want = engine.timesoap_device(ref, delay)
        got = engine.soap_timesoap(dpos, species, centres, bx, 4.5, n_max, l_max, delay=delay,
                                   n_species=n_species)
        assert got.shape == want.shape == (37, n_frames - delay)
        # two launches sum the neighbours in different orders (atomic scatter inside a cell), and
        # sqrt(2 - 2 cos) amplifies 1e-13 on the spectra to ~1e-10 on a distance of 0.05
        assert float((got - want).abs().max()) < 2e-9
    spectra = torch.full_like(ref, float("nan"))
    got = engine.soap_timesoap(dpos, species, centres, bx, 4.5, n_max, l_max, delay=2,
                               n_species=n_species, soap_out=spectra)
    scale = ref.abs().amax(dim=-1, keepdim=True)
    assert float(((spectra - ref).abs() / scale).max()) < 1e-9   # summation order, see above
    assert float((got - engine.timesoap_device(ref, 2)).abs().max()) < 2e-9


def test_fused_timesoap_vs_oracle_and_api():
    """The fused path against the oracle chain (SOAP restatement -> timesoap.py restatement) and
    through timesoap_from_trajectory with small overlapping batches."""
    from dynsight_b200 import soap as dsoap
    from dynsight_b200.universe import ArrayUniverse

    rng = np.random.default_rng(5)
    n, n_frames = 150, 9
    box = np.array([26.5, 27.0, 28.0], dtype=np.float32)
    pos = (rng.random((1, n, 3)) * box + rng.normal(0, 0.2, (n_frames, n, 3)).cumsum(0)).astype(
        np.float32)
    dims = np.tile(np.concatenate([box, [90, 90, 90]]).astype(np.float32), (n_frames, 1))
    uni = ArrayUniverse(pos, dims, names=["O"] * n, types=["O"] * n)
    cent = np.arange(0, n, 5)
    want_soap = so.soap_trajectory(pos, np.full(n, 8), cent, dims[:, :3].astype(float), 5.0, 8, 8)
    for delay, batch in ((1, 4), (2, 5), (1, None)):
        want = to.timesoap(want_soap, delay)
        got = dsoap.timesoap_from_trajectory(uni, 5.0, 8, 8, delay=delay, batch_frames=batch,
                                             as_numpy=True)[cent]
        assert np.all(np.abs(got - want) <= 1e-5 * np.abs(want) + 1e-7)


# ---------------------------------------------------------------------------------------------
# the BASELINE configurations at their stated sizes (sampled centres vs the oracle)
# ---------------------------------------------------------------------------------------------
def _random_walk(n: int, n_frames: int, box_l: float, step: float, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    top = np.nextafter(np.float32(box_l), np.float32(0))
    cur = rng.random((n, 3)) * box_l
    out = np.empty((n_frames, n, 3), dtype=np.float32)
    for f in range(n_frames):
        cur = np.mod(cur + rng.normal(0, step, cur.shape), box_l)
        out[f] = np.minimum(cur.astype(np.float32), top)
    return out


def test_cfg5_size_soap_and_fused_timesoap_vs_oracle():
    """BASELINE.json configs[4] shape (what bench.py times): 100,000 atoms at water-oxygen density,
    SOAP(5 A, 8, 8) in list mode on the tensor path, three frames.  The spectra of 128 sampled centres
    and the fused SOAP -> timeSOAP values of the same centres against the oracle chain."""
    import os

    from dynsight_b200 import engine

    n, n_frames = 100_000, 3
    box_l = (n / 0.0334) ** (1 / 3)
    pos = _random_walk(n, n_frames, box_l, 0.1, 21)
    box = np.full(3, np.float32(box_l), dtype=np.float32)
    dpos = torch.as_tensor(pos).cuda()
    sp = torch.zeros(n, dtype=torch.int32, device="cuda")
    sample = np.linspace(0, n - 1, 128).astype(np.int32)
    basis = so.gto_basis(5.0, 8, 8)
    want = np.stack(
        [so.soap_frame_c(pos[f], np.full(n, 8), sample, box.astype(np.float64), 5.0, 8, 8, basis,
                         n_threads=os.cpu_count() or 1) for f in range(n_frames)], axis=1,
    )  # fmt: skip
    got = engine.soap_power_spectrum(dpos, sp, torch.as_tensor(sample).cuda(), box, 5.0, 8, 8)
    assert _rel_err(got.cpu().numpy(), want) < REL
    # the fused kernel on ALL centres (frame-sequential warps, one ring slot per centre)
    ts = engine.soap_timesoap(dpos, sp, torch.arange(n, dtype=torch.int32, device="cuda"), box,
                              5.0, 8, 8, delay=1).cpu().numpy()  # fmt: skip
    want_t = to.timesoap(want, 1)
    assert ts.shape == (n, n_frames - 1)
    assert np.all(np.abs(ts[sample] - want_t) <= 1e-5 * np.abs(want_t) + 1e-7)
    assert np.isfinite(ts).all() and float(ts.min()) >= 0.0


def test_cfg4_size_three_species_l10_vs_oracle():
    """BASELINE.json configs[3] at its stated size: 250,000 atoms, 3 species, l_max = 10, metal
    density (K ~ 239), two frames through the split-mode kernels (accumulate on the tensor path +
    soap_spectrum_mma_kernel): spectra of 36 sampled centres and their timeSOAP vs the oracle."""
    import os

    from dynsight_b200 import engine

    n, n_frames, n_sp = 250_000, 2, 3
    box_l = (n / 0.0857) ** (1 / 3)
    pos = _random_walk(n, n_frames, box_l, 0.05, 22)
    box = np.full(3, np.float32(box_l), dtype=np.float32)
    sp_idx = (np.arange(n) * 7919) % n_sp
    z = np.asarray([26, 28, 29])[sp_idx]   # Fe, Ni, Cu: ascending atomic numbers
    dpos = torch.as_tensor(pos).cuda()
    sp = torch.as_tensor(sp_idx.astype(np.int32)).cuda()
    sample = np.linspace(0, n - 1, 36).astype(np.int32)
    basis = so.gto_basis(5.0, 8, 10)
    want = np.stack(
        [so.soap_frame_c(pos[f], z, sample, box.astype(np.float64), 5.0, 8, 10, basis,
                         n_threads=os.cpu_count() or 1) for f in range(n_frames)], axis=1,
    )  # fmt: skip
    got = engine.soap_power_spectrum(dpos, sp, torch.as_tensor(sample).cuda(), box, 5.0, 8, 10,
                                     n_species=n_sp)  # fmt: skip
    assert got.shape == (36, n_frames, 3300)
    assert _rel_err(got.cpu().numpy(), want) < REL
    ts = engine.soap_timesoap(dpos, sp, torch.arange(n, dtype=torch.int32, device="cuda"), box,
                              5.0, 8, 10, delay=1, n_species=n_sp).cpu().numpy()  # fmt: skip
    want_t = to.timesoap(want, 1)
    assert np.all(np.abs(ts[sample] - want_t) <= 1e-5 * np.abs(want_t) + 1e-7)
    assert np.isfinite(ts).all()
