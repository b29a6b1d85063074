"""GPU parity: neighbour CSR / LENS / coordination number through the C ABI vs the CPU oracle.

Bar: bit-exact (integer index work + one IEEE float64 division).
"""

from __future__ import annotations

import numpy as np
import pytest
import torch
from golden_cases import fluid_case, random_csr_cases
from test_oracle_golden import LENS_CASES, NN_CASES, lens_2d_inputs

from oracle import lens_oracle as lo

pytestmark = pytest.mark.gpu


def _dev(x, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda().contiguous()


def _check_batch_vs_oracle(batch, pos_env, pos_cent, box, r_cut, pbc):
    """Compare every frame of a NeighborBatch with the oracle CSR."""
    counts = batch.counts.cpu().numpy()
    indptr = batch.indptr.cpu().numpy()
    frame_off = batch.frame_off.cpu().numpy()
    indices = batch.indices.cpu().numpy()
    for f in range(pos_env.shape[0]):
        bx = box[f] if np.ndim(box) == 2 else box  # noqa: PLR2004
        ip, ix = lo.neighbors(pos_env[f], pos_env[f] if pos_cent is None else pos_cent[f], r_cut, bx,
                              pbc)  # fmt: skip
        assert np.array_equal(indptr[f], ip), f"indptr differs in frame {f}"
        assert np.array_equal(counts[f], np.diff(ip)), f"counts differ in frame {f}"
        got = indices[frame_off[f] : frame_off[f + 1]]
        assert np.array_equal(got, ix), f"indices differ in frame {f}"


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_random_csr_cases_vs_numba_golden(numba_cases, dtype):
    from dynsight_b200 import engine

    for k, case in enumerate(random_csr_cases()):
        pe = _dev(case["pos_env"][None], dtype)
        same = case["pos_cent"] is case["pos_env"]
        pc = None if same else _dev(case["pos_cent"][None], dtype)
        batch = engine.neighbor_lists(pe, case["box"], case["r_cut"], pc, case["pbc"])
        ip, ix = batch.frame_csr(0)
        assert np.array_equal(ip.cpu().numpy(), numba_cases[f"csr{k}_indptr"]), k
        assert np.array_equal(ix.cpu().numpy(), numba_cases[f"csr{k}_indices"]), k


def test_random_cases_batched_with_explicit_centres():
    """Several frames per launch, centres passed explicitly even when equal to env."""
    from dynsight_b200 import engine

    rng = np.random.default_rng(5)
    n_frames, n, m = 7, 211, 97
    box = (rng.random((n_frames, 3)) * 6 + 4).astype(np.float32)  # per-frame box (NPT-like)
    pe = ((rng.random((n_frames, n, 3)) * 1.2 - 0.1) * box[:, None, :]).astype(np.float32)
    pc = ((rng.random((n_frames, m, 3)) * 1.2 - 0.1) * box[:, None, :]).astype(np.float32)
    for pbc in (True, False):
        batch = engine.neighbor_lists(_dev(pe), box, 1.6, _dev(pc), pbc)
        _check_batch_vs_oracle(batch, pe, pc, box, 1.6, pbc)


def test_float64_positions_not_representable_in_float32():
    from dynsight_b200 import engine

    rng = np.random.default_rng(11)
    box = np.array([7.3, 6.1, 8.9])
    pe = rng.random((3, 500, 3)) * box
    batch = engine.neighbor_lists(_dev(pe), box, 1.45, None, True)
    _check_batch_vs_oracle(batch, pe, None, box, 1.45, True)


def test_guard_band_pairs_exactly_at_cutoff():
    """Pairs at |d| == r_cut - 1e-6 +- ulps must follow the float64 rule (strict <)."""
    from dynsight_b200 import engine

    r_cut = 1.5
    box = np.array([12.0, 12.0, 12.0])
    base = np.array([3.0, 3.0, 3.0])
    rc = r_cut - 1e-6
    offs = [rc, np.nextafter(rc, 0), np.nextafter(rc, 9), rc * (1 - 1e-7), rc * (1 + 1e-7), r_cut]
    pts = [base]
    for k, o in enumerate(offs):
        v = np.zeros(3)
        v[k % 3] = o
        pts.append(base + v)
        pts.append(base - v * 0.999999)
    pe = np.asarray(pts)[None]
    for arr in (pe, pe.astype(np.float32)):
        batch = engine.neighbor_lists(_dev(arr), box, r_cut, None, True)
        _check_batch_vs_oracle(batch, arr, None, box, r_cut, True)


def test_fluid_lens_and_counts_vs_numba_golden(numba_cases):
    from dynsight_b200 import engine

    pos, box, r_cut = fluid_case()
    batch = engine.neighbor_lists(_dev(pos), box, r_cut, None, True)
    lens = engine.lens_pairs(batch, 1).cpu().numpy()
    assert np.array_equal(lens, numba_cases["fluid_lens"])
    nc = engine.coordination_numbers(batch.counts).cpu().numpy()
    assert np.array_equal(nc, numba_cases["fluid_counts"])
    ip, ix = batch.frame_csr(pos.shape[0] - 1)
    assert np.array_equal(ip.cpu().numpy(), numba_cases["fluid_last_indptr"])
    assert np.array_equal(ix.cpu().numpy(), numba_cases["fluid_last_indices"])
    # delay 2 against the oracle
    lens2 = engine.lens_pairs(batch, 2).cpu().numpy()
    dims = np.tile(np.concatenate([box, [90, 90, 90]]).astype(np.float32), (pos.shape[0], 1))
    assert np.array_equal(lens2, lo.compute_lens_arrays(pos, dims, r_cut, delay=2))


def test_lens_from_two_csr_signature():
    from dynsight_b200 import engine

    pos, box, r_cut = fluid_case(n_side=8, n_frames=2)
    a = lo.neighbors(pos[0], pos[0], r_cut, box, True)
    b = lo.neighbors(pos[1], pos[1], r_cut, box, True)
    got = engine.lens_from_two_csr(_dev(a[0]), _dev(a[1]), _dev(b[0]), _dev(b[1])).cpu().numpy()
    assert np.array_equal(got, lo.lens_from_two_csr(a[0], a[1], b[0], b[1]))


def test_long_rows_sorted_paths():
    """Rows longer than 32 (shared-memory bitonic) and longer than 512 (global fallback)."""
    from dynsight_b200 import engine

    rng = np.random.default_rng(3)
    box = np.array([6.0, 6.0, 6.0], dtype=np.float32)
    for n, r_cut in ((300, 1.9), (4000, 1.99)):
        pe = (rng.random((2, n, 3)) * box).astype(np.float32)
        batch = engine.neighbor_lists(_dev(pe), box, r_cut, None, True)
        assert int(batch.counts.max()) > (32 if n == 300 else 512)  # noqa: PLR2004
        _check_batch_vs_oracle(batch, pe, None, box, r_cut, True)
        lens = engine.lens_pairs(batch, 1).cpu().numpy()
        dims = np.tile(np.concatenate([box, [90, 90, 90]]).astype(np.float32), (2, 1))
        assert np.array_equal(lens, lo.compute_lens_arrays(pe, dims, r_cut))


def test_balls7_reference_goldens(balls7, ref_lens):
    """tests/trajectory/test_trj.py:79-94 numbers (r_cut=15) straight through the engine."""
    from dynsight_b200 import engine

    pos, dims = balls7["positions"], balls7["dims"]
    batch = engine.neighbor_lists(_dev(pos), dims[:, :3], 15.0, None, True)
    assert np.array_equal(engine.lens_pairs(batch, 1).cpu().numpy(), ref_lens["trj_lens"])
    assert np.array_equal(
        engine.coordination_numbers(batch.counts).cpu().numpy(), ref_lens["trj_n_c"]
    )


@pytest.mark.parametrize("pbc", [True, False])
def test_lens_2d_goldens(ref_lens, pbc):
    from dynsight_b200 import engine

    pos, dims = lens_2d_inputs()
    batch = engine.neighbor_lists(_dev(pos), dims[:, :3], 0.1, None, pbc)
    got = engine.lens_pairs(batch, 1).cpu().numpy()
    assert np.array_equal(got, ref_lens["lens_2d_pbc" if pbc else "lens_2d_fbc"])


@pytest.mark.parametrize("case", sorted(LENS_CASES))
def test_lens_cases_goldens(balls7, ref_lens, case):
    from dynsight_b200 import engine

    r_cut, delay, cent, env = LENS_CASES[case]
    pos, dims = balls7["positions"], balls7["dims"]
    pe = pos if env is None else pos[:, env]
    pc = None if cent is None and env is None else (pos if cent is None else pos[:, cent])
    batch = engine.neighbor_lists(
        _dev(pe), dims[:, :3], r_cut, None if pc is None else _dev(pc), True
    )
    assert np.array_equal(
        engine.lens_pairs(batch, delay).cpu().numpy(), ref_lens["lens_" + case]
    )
    if case in NN_CASES:
        assert np.array_equal(
            engine.coordination_numbers(batch.counts).cpu().numpy(), ref_lens["nn_" + case]
        )


def test_error_paths():
    from dynsight_b200 import _lib, engine

    pos = torch.zeros((1, 4, 3), dtype=torch.float32, device="cuda")
    with pytest.raises(_lib.DsgError):
        engine.neighbor_lists(pos, np.array([0.0, 1.0, 1.0]), 1.0)  # non-positive box
    with pytest.raises(_lib.DsgError):
        engine.neighbor_lists(pos, np.array([1.0, 1.0, 1.0]), -1.0)
    with pytest.raises(RuntimeError):
        engine.neighbor_lists(pos.cpu(), np.array([1.0, 1.0, 1.0]), 1.0)
    batch = engine.neighbor_lists(pos, np.array([5.0, 5.0, 5.0]), 1.0)
    with pytest.raises(RuntimeError, match="No valid pairs found."):
        engine.lens_pairs(batch, 1)


def test_large_fluid_properties():
    """cfg3-sized frame (1e6 particles): properties that need no oracle at this size --
    rows strictly ascending, neighbour relation symmetric (sum of ids trick), counts == row
    lengths, LENS of a frame with itself == 0 -- plus an oracle spot check on a sub-block."""
    from dynsight_b200 import engine

    n_side = 100
    n = n_side**3
    box_l = (n / 0.8) ** (1 / 3)
    g = torch.Generator(device="cuda").manual_seed(0)
    grid = (torch.arange(n_side, device="cuda", dtype=torch.float32) + 0.5) * (box_l / n_side)
    lat = torch.stack(torch.meshgrid(grid, grid, grid, indexing="ij"), dim=-1).reshape(-1, 3)
    pos0 = torch.remainder(lat + 0.25 * torch.randn(lat.shape, device="cuda", generator=g), box_l)
    pos0 = pos0.clamp_(max=float(np.nextafter(np.float32(box_l), np.float32(0))))
    pos = torch.stack([pos0, pos0]).contiguous()
    box = np.full(3, np.float32(box_l), dtype=np.float32)
    batch = engine.neighbor_lists(pos, box, 1.5, None, True)
    indptr = batch.indptr[0].long()
    ix = batch.indices[: int(batch.frame_off[1])].long()
    assert torch.equal(batch.counts[0].long(), indptr[1:] - indptr[:-1])
    # strictly ascending inside rows: positions where the next entry is not larger must be row ends
    row_of = torch.repeat_interleave(torch.arange(n, device="cuda"), batch.counts[0].long())
    same_row = row_of[1:] == row_of[:-1]
    assert bool(torch.all(ix[1:][same_row] > ix[:-1][same_row]))
    # symmetry: j in N(i) <=> i in N(j)  => multiset of (i,j) equals multiset of (j,i)
    key_fwd = row_of * n + ix
    key_bwd = ix * n + row_of
    assert torch.equal(torch.sort(key_fwd).values, torch.sort(key_bwd).values)
    lens = engine.lens_pairs(batch, 1)
    assert float(lens.abs().max()) == 0.0
    mean_nc = float(batch.counts[0].float().mean())
    assert 9.0 < mean_nc < 14.0  # rho*=0.8, r_cut=1.5 -> ~11.3
    # oracle spot check of 2000 centres against the full frame
    sel = torch.arange(0, n, n // 2000, device="cuda")[:2000]
    p_host = pos0.cpu().numpy()
    ip, ixo = lo.neighbors(p_host, p_host[sel.cpu().numpy()], 1.5, box, True)
    # the oracle's "j != i" compares env index with *centre-local* index, so emulate by checking
    # rows through set equality after removing/adding the self / local-index quirk
    indptr_h = indptr.cpu().numpy()
    ix_h = ix.cpu().numpy()
    for k, c in enumerate(sel.cpu().numpy()):
        want = set(ixo[ip[k] : ip[k + 1]].tolist())
        want.discard(int(c))  # self is a neighbour in the oracle call (centre-local index k != c)
        got = set(ix_h[indptr_h[c] : indptr_h[c + 1]].tolist())
        # env atom with index == k is skipped by the oracle call; tolerate exactly that one id
        assert got - {k} == want - {k}, (k, c)


# ---------------------------------------------------------------------------------------------
# rows path (single traversal, unsorted rows, atomically reserved offsets) -- the LENS fast path
# ---------------------------------------------------------------------------------------------
def _check_rows_vs_oracle(nr, pos_env, pos_cent, box, r_cut, pbc):
    assert not nr.overflowed()
    counts = nr.counts.cpu().numpy()
    starts = nr.row_start.cpu().numpy()
    rows = nr.rows.cpu().numpy()
    for f in range(pos_env.shape[0]):
        bx = box[f] if np.ndim(box) == 2 else box  # noqa: PLR2004
        ip, ix = lo.neighbors(pos_env[f], pos_env[f] if pos_cent is None else pos_cent[f], r_cut, bx,
                              pbc)  # fmt: skip
        assert np.array_equal(counts[f], np.diff(ip)), f"counts differ in frame {f}"
        for i in range(len(ip) - 1):
            got = np.sort(rows[starts[f, i] : starts[f, i] + counts[f, i]])
            assert np.array_equal(got, ix[ip[i] : ip[i + 1]]), (f, i)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_rows_path_random_cases(numba_cases, dtype):
    from dynsight_b200 import engine

    for k, case in enumerate(random_csr_cases()):
        pe = _dev(case["pos_env"][None], dtype)
        same = case["pos_cent"] is case["pos_env"]
        pc = None if same else _dev(case["pos_cent"][None], dtype)
        nr = engine.neighbor_rows(pe, case["box"], case["r_cut"], pc, case["pbc"])
        assert not nr.overflowed()
        counts = nr.counts[0].cpu().numpy()
        starts = nr.row_start[0].cpu().numpy()
        rows = nr.rows.cpu().numpy()
        ip, ix = numba_cases[f"csr{k}_indptr"], numba_cases[f"csr{k}_indices"]
        assert np.array_equal(counts, np.diff(ip)), k
        for i in range(len(ip) - 1):
            assert np.array_equal(np.sort(rows[starts[i] : starts[i] + counts[i]]),
                                  ix[ip[i] : ip[i + 1]]), (k, i)  # fmt: skip


def test_rows_path_shift_mode_with_atoms_outside_box():
    """>= 6 cells per axis (image implied by the cell offset) with atoms several boxes away,
    per-frame boxes, explicit centres, pbc on and off."""
    from dynsight_b200 import engine

    rng = np.random.default_rng(17)
    n_frames, n, m = 4, 900, 300
    box = (rng.random((n_frames, 3)) * 4 + 12).astype(np.float32)  # 12..16 with r_cut 1.5 -> 8+ cells
    pe = ((rng.random((n_frames, n, 3)) * 3.4 - 1.2) * box[:, None, :]).astype(np.float32)
    pc = ((rng.random((n_frames, m, 3)) * 3.4 - 1.2) * box[:, None, :]).astype(np.float32)
    for pbc in (True, False):
        nr = engine.neighbor_rows(_dev(pe), box, 1.5, _dev(pc), pbc)
        _check_rows_vs_oracle(nr, pe, pc, box, 1.5, pbc)
        nr = engine.neighbor_rows(_dev(pe), box, 1.5, None, pbc)
        _check_rows_vs_oracle(nr, pe, None, box, 1.5, pbc)
    pe64 = (rng.random((2, n, 3)) * 1.6 - 0.3) * box[:2, None, :].astype(np.float64)
    nr = engine.neighbor_rows(_dev(pe64), box[:2], 1.5, None, True)
    _check_rows_vs_oracle(nr, pe64, None, box[:2], 1.5, True)


def test_rows_path_dense_cells_and_overflow_retry():
    """More than 128 candidates per cell (chunked traversal) and a deliberately small capacity."""
    from dynsight_b200 import engine

    rng = np.random.default_rng(23)
    box = np.array([6.0, 6.0, 6.0], dtype=np.float32)
    pe = (rng.random((3, 1200, 3)) * box).astype(np.float32)
    nr = engine.neighbor_rows(_dev(pe), box, 1.9, None, True)
    _check_rows_vs_oracle(nr, pe, None, box, 1.9, True)
    small = engine.neighbor_rows(_dev(pe), box, 1.9, None, True, capacity=256 * 8)
    assert small.overflowed() and small.needed_capacity() > 256 * 8
    dims = np.tile(np.concatenate([box, [90, 90, 90]]).astype(np.float32), (3, 1))
    want = lo.compute_lens_arrays(pe, dims, 1.9)
    lens, nr2 = engine.lens_from_positions(_dev(pe), box, 1.9, 1, capacity=256 * 8)
    assert not nr2.overflowed()
    assert np.array_equal(lens.cpu().numpy(), want)


def test_rows_path_thread_and_warp_kernels_agree():
    """Sparse fluid: the thread-per-centre kernel (auto) and the warp-per-cell kernel (forced) give
    the same rows; in a dense spot the thread kernel writes the long rows in a second walk."""
    from dynsight_b200 import engine

    pos, box, r_cut = fluid_case(n_side=16, n_frames=3)
    dims = np.tile(np.concatenate([box, [90, 90, 90]]).astype(np.float32), (3, 1))
    for pbc in (True, False):
        auto = engine.neighbor_rows(_dev(pos), box, r_cut, None, pbc)
        warp = engine.neighbor_rows(_dev(pos), box, r_cut, None, pbc, force_warp_kernel=True)
        _check_rows_vs_oracle(auto, pos, None, box, r_cut, pbc)
        _check_rows_vs_oracle(warp, pos, None, box, r_cut, pbc)
    # 60 atoms packed into one corner -> more than 32 neighbours for those centres
    dense = pos.copy()
    dense[:, :60] = (dense[:, :60] % 1.0) * 0.9 + 0.05
    nr = engine.neighbor_rows(_dev(dense), box, r_cut, None, True)
    assert not nr.overflowed()  # rows longer than the tile are written by a second walk
    assert int(nr.counts.max()) > 32
    _check_rows_vs_oracle(nr, dense, None, box, r_cut, True)
    lens, nr2 = engine.lens_from_positions(_dev(dense), box, r_cut, 1)
    assert not nr2.overflowed()
    assert np.array_equal(lens.cpu().numpy(), lo.compute_lens_arrays(dense, dims, r_cut))


def test_rows_path_lens_goldens(balls7, ref_lens, numba_cases):
    from dynsight_b200 import engine

    pos, dims = balls7["positions"], balls7["dims"]
    lens, nr = engine.lens_from_positions(_dev(pos), dims[:, :3], 15.0, 1)
    assert np.array_equal(lens.cpu().numpy(), ref_lens["trj_lens"])
    assert np.array_equal(engine.coordination_numbers(nr.counts).cpu().numpy(), ref_lens["trj_n_c"])
    posf, box, r_cut = fluid_case()
    lens, nr = engine.lens_from_positions(_dev(posf), box, r_cut, 1)
    assert np.array_equal(lens.cpu().numpy(), numba_cases["fluid_lens"])
    assert np.array_equal(
        engine.coordination_numbers(nr.counts).cpu().numpy(), numba_cases["fluid_counts"]
    )
    for pbc in (True, False):
        p2, d2 = lens_2d_inputs()
        lens, _ = engine.lens_from_positions(_dev(p2), d2[:, :3], 0.1, 1, respect_pbc=pbc)
        assert np.array_equal(lens.cpu().numpy(), ref_lens["lens_2d_pbc" if pbc else "lens_2d_fbc"])


@pytest.mark.parametrize("force_warp", [False, True])
def test_rows_path_near_threshold_pairs_in_a_large_box(force_warp):
    """Adversarial for the float32 guard band of the shift-mode rows kernels: thousands of pairs
    whose distance is within a few float32 ulps of the cutoff, at coordinates ~100 (float32
    spacing 8e-6), many of them across the periodic boundary.  Membership must follow the
    reference's float64 rule bit for bit."""
    from dynsight_b200 import engine

    rng = np.random.default_rng(99)
    r_cut = 1.5
    rc = r_cut - 1e-6
    box = np.array([107.7, 96.3, 121.1], dtype=np.float32)
    n_pairs = 6000
    base = rng.random((n_pairs, 3)) * box
    # a third of the pairs hug a face of the box so that the partner lands in the periodic image
    face = rng.integers(0, 3, n_pairs)
    hug = rng.random(n_pairs) < 0.33
    base[hug, face[hug]] = rng.random(hug.sum()) * 0.4
    d = rng.normal(size=(n_pairs, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    r = rc * (1.0 + rng.uniform(-4e-6, 4e-6, n_pairs))
    partner = base - d * r[:, None]
    pts = np.concatenate([base, partner]).astype(np.float32)
    pts = np.mod(pts, box).astype(np.float32)
    frames = np.stack([pts, np.roll(pts, 1, axis=0)])
    nr = engine.neighbor_rows(_dev(frames), box, r_cut, None, True, force_warp_kernel=force_warp)
    assert not nr.overflowed()
    _check_rows_vs_oracle(nr, frames, None, box, r_cut, True)
    counts = nr.counts.cpu().numpy()
    assert 0.2 < (counts[0, :n_pairs] > 0).mean() < 0.8  # both outcomes occur


def test_lens_rows_thread_and_warp_kernels_agree_on_mixed_rows():
    """delay 1: thread-per-centre LENS kernel (short rows), its in-kernel switch for warps that
    hold a row longer than 32, and the warp-per-centre kernel (bit 30) give identical results."""
    from dynsight_b200 import _lib, engine

    pos, box, r_cut = fluid_case(n_side=16, n_frames=5)
    dense = pos.copy()
    dense[:, :80] = (dense[:, :80] % 1.0) * 0.9 + 0.05   # rows of > 32 entries for some centres
    dims = np.tile(np.concatenate([box, [90, 90, 90]]).astype(np.float32), (5, 1))
    want = lo.compute_lens_arrays(dense, dims, r_cut)
    nr = engine.neighbor_rows(_dev(dense), box, r_cut, None, True)
    assert not nr.overflowed() and int(nr.counts.max()) > 32
    got_thread = engine.lens_pairs_rows(nr, 1)
    nr.k_est = 1e9  # force the dense-rows flag
    got_warp = engine.lens_pairs_rows(nr, 1)
    assert np.array_equal(got_thread.cpu().numpy(), want)
    assert np.array_equal(got_warp.cpu().numpy(), want)
    del _lib


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_canonical_csr_one_traversal_equals_two_pass(numba_cases, dtype):
    """engine.neighbor_lists builds the canonical CSR from the single-traversal rows
    (dsg_rows_csr_offsets + dsg_rows_to_csr); the count / fill protocol (two traversals) must give
    the same arrays, and both equal the numba goldens."""
    from dynsight_b200 import engine

    for k, case in enumerate(random_csr_cases()):
        pe = _dev(case["pos_env"].astype(dtype)[None])
        pc = _dev(case["pos_cent"].astype(dtype)[None])
        one = engine.neighbor_lists(pe, case["box"], case["r_cut"], pc, case["pbc"])
        two = engine.neighbor_lists(pe, case["box"], case["r_cut"], pc, case["pbc"], two_pass=True)
        for a, b in ((one.indptr, two.indptr), (one.frame_off, two.frame_off),
                     (one.indices, two.indices), (one.counts, two.counts)):  # fmt: skip
            assert torch.equal(a, b)
        if dtype == np.float32:
            continue
        assert np.array_equal(one.indptr[0].cpu().numpy(), numba_cases[f"csr{k}_indptr"])
        assert np.array_equal(one.indices.cpu().numpy(), numba_cases[f"csr{k}_indices"])


def test_canonical_csr_thread_sort_with_a_dense_spot():
    """Sparse fluid (expected row length ~11 -> one thread sorts one row) with 80 atoms packed
    into a corner: the warps that hold a row longer than 32 switch to the warp-cooperative
    copy-and-sort.  The CSR must equal the oracle's."""
    from dynsight_b200 import engine

    pos, box, r_cut = fluid_case(n_side=16, n_frames=3)
    dense = pos.copy()
    dense[:, :80] = (dense[:, :80] % 1.0) * 0.9 + 0.05
    batch = engine.neighbor_lists(_dev(dense), box, r_cut, None, True)
    assert int(batch.counts.max()) > 32
    _check_batch_vs_oracle(batch, dense, None, box, r_cut, True)


def test_clustered_atoms_overflow_retry_terminates():
    """Regression (found by scripts/fuzz_neighbors.py): atoms clustered in a large, mostly empty box
    with one short axis (warp-per-cell kernel, rint image on that axis).  The density-based
    capacity is far too small, cell 0 is empty, and the retry must size itself from the number of
    reservation regions reported in the status block -- it used to loop forever."""
    from dynsight_b200 import engine

    rng = np.random.default_rng(169)
    nf, n, r_cut = 3, 1500, 2.5
    box = np.array([120.5, 9.26, 114.3], dtype=np.float32)
    centres = rng.random((nf, n // 40, 3)) * box
    pos = (centres[:, rng.integers(0, centres.shape[1], n)] + rng.normal(0, 0.3, (nf, n, 3))).astype(
        np.float32)  # fmt: skip
    first = engine.neighbor_rows(_dev(pos), box, r_cut, None, True, capacity=256 * 64)
    assert first.overflowed() and int(first.status[1]) > 0   # regions reported although cell 0 is empty
    assert first.needed_capacity() >= int(first.counts.sum()) // 2
    nr_ok = engine.neighbor_rows_checked(_dev(pos), box, r_cut, None, True, capacity=256 * 64)
    _check_rows_vs_oracle(nr_ok, pos, None, box, r_cut, True)
    batch = engine.neighbor_lists(_dev(pos), box, r_cut, None, True)
    _check_batch_vs_oracle(batch, pos, None, box, r_cut, True)
    dims = np.tile(np.concatenate([box, [90, 90, 90]]).astype(np.float32), (nf, 1))
    lens, nr = engine.lens_from_positions(_dev(pos), box, r_cut, 1, capacity=256 * 64)
    assert not nr.overflowed()
    assert np.array_equal(lens.cpu().numpy(), lo.compute_lens_arrays(pos, dims, r_cut))


# ---------------------------------------------------------------------------------------------
# pair path (dsg_lens_direct): half-shell pair search, partner-frame lookup, no lists
# ---------------------------------------------------------------------------------------------
def _dims(box, n_frames):
    box = np.asarray(box, dtype=np.float32)
    if box.ndim == 1:
        box = np.tile(box, (n_frames, 1))
    return np.concatenate([box, np.full((n_frames, 3), 90, np.float32)], axis=1)


def _check_direct_vs_oracle(pos, box, r_cut, delay, pbc, dtype=None):
    from dynsight_b200 import engine

    n_frames = pos.shape[0]
    assert engine.lens_direct_supported(n_frames, pos.shape[1], box, r_cut, pbc)
    lens, counts = engine.lens_direct(_dev(pos, dtype), box, r_cut, delay, pbc)
    dims = _dims(box, n_frames)
    want = lo.compute_lens_arrays(pos, dims, r_cut, delay=delay, respect_pbc=pbc)
    want_c = lo.coord_number_arrays(pos, dims, r_cut, respect_pbc=pbc)
    assert np.array_equal(counts.cpu().numpy().T.astype(np.float64), want_c), "counts differ"
    assert np.array_equal(lens.cpu().numpy(), want), "LENS differs"
    # counts-only mode (Trj.get_coord_number)
    _, counts0 = engine.lens_direct(_dev(pos, dtype), box, r_cut, 0, pbc, want_lens=False)
    assert torch.equal(counts0, counts)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_pair_path_random_cases(dtype):
    """Random boxes from 3 (pbc off) / 6 (pbc on) to 12 cells per axis, atoms inside and far outside
    the box (the reference bins those by truncation + modulo), several delays, per-frame boxes."""
    rng = np.random.default_rng(2024)
    for case in range(36):
        n = int(rng.integers(200, 1500))
        n_frames = int(rng.integers(2, 7))
        delay = int(rng.integers(1, n_frames))
        pbc = bool(case % 2)
        r_cut = float(rng.uniform(0.8, 1.7))
        ncx = rng.integers(6 if pbc else 3, 13, size=3)
        box = (ncx * r_cut * rng.uniform(1.0, 1.15, size=3)).astype(np.float32)
        spread = (1.0, 1.0, 1.3, 3.0)[case % 4]
        pos = ((rng.random((n_frames, n, 3)) - 0.5) * spread + 0.5) * box
        pos[1:] = pos[0] + rng.normal(0, 0.2 * r_cut, (n_frames - 1, n, 3)).cumsum(axis=0)
        pos = pos.astype(np.float32)
        if spread == 1.0:
            pos = np.clip(pos, 0, np.nextafter(box, np.float32(0))).astype(np.float32)
        boxes = box
        if case % 3 == 0:  # NPT-like: every frame has its own box (same cell counts or not)
            boxes = (box[None, :] * rng.uniform(0.97, 1.03, (n_frames, 3))).astype(np.float32)
        p = pos if dtype == torch.float32 else pos.astype(np.float64) + rng.normal(0, 1e-9, pos.shape)
        from dynsight_b200 import engine

        if not engine.lens_direct_supported(n_frames, n, boxes, r_cut, pbc):
            continue  # too dense for one thread per atom: covered by the rows path tests
        _check_direct_vs_oracle(p, boxes, r_cut, delay, pbc, dtype)


def test_pair_path_near_threshold_pairs_in_both_frames():
    """Adversarial for BOTH float32 guard bands of the pair path: thousands of pairs within a few
    float32 ulps of the cutoff at coordinates ~100, a third of them across the periodic boundary;
    frame 1 holds the same points moved rigidly by a random lattice of tiny shifts, so the
    partner-frame test (plain float32 minimum image) sees near-threshold pairs too."""
    rng = np.random.default_rng(77)
    r_cut = 1.5
    rc = r_cut - 1e-6
    box = np.array([107.7, 96.3, 121.1], dtype=np.float32)
    n_pairs = 6000
    base = rng.random((n_pairs, 3)) * box
    face = rng.integers(0, 3, n_pairs)
    hug = rng.random(n_pairs) < 0.33
    base[hug, face[hug]] = rng.random(hug.sum()) * 0.4
    d = rng.normal(size=(n_pairs, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    r = rc * (1.0 + rng.uniform(-4e-6, 4e-6, n_pairs))
    partner = base - d * r[:, None]
    pts = np.mod(np.concatenate([base, partner]), box).astype(np.float32)
    # frame 1: every pair re-drawn at a fresh near-threshold distance around the same base point
    r1 = rc * (1.0 + rng.uniform(-4e-6, 4e-6, n_pairs))
    pts1 = np.mod(np.concatenate([base, base - d * r1[:, None]]), box).astype(np.float32)
    top = np.nextafter(box, np.float32(0))
    frames = np.minimum(np.stack([pts, pts1, pts]), top).astype(np.float32)
    _check_direct_vs_oracle(frames, box, r_cut, 1, True)
    _check_direct_vs_oracle(frames, box, r_cut, 2, True)


def test_pair_path_dense_spot_settles_inline():
    """A cluster whose half rows exceed the 32-entry staging tile in an otherwise dilute box: the
    thread walks its cells a second time and settles every hit at once."""
    rng = np.random.default_rng(8)
    box = np.array([30.0, 31.0, 29.0], dtype=np.float32)
    n = 4000
    pos = (rng.random((3, n, 3)) * box).astype(np.float32)
    pos[:, :90] = (np.array([14.2, 15.1, 13.9]) + rng.normal(0, 0.25, (3, 90, 3))).astype(np.float32)
    _check_direct_vs_oracle(pos, box, 1.5, 1, True)
    _check_direct_vs_oracle(pos, box, 1.5, 1, False)


def test_pair_path_unsupported_and_fallback(balls7):
    """Dense cells / few cells per axis are refused through the C ABI; the API falls back to the
    rows path there (r_cut = 15 A in a 30 A box: the reference golden)."""
    from dynsight_b200 import _lib, engine

    pos = balls7["positions"][:3]
    box = balls7["dims"][0, :3]
    assert not engine.lens_direct_supported(3, pos.shape[1], box, 15.0, True)
    with pytest.raises(_lib.DsgError, match="pair path"):
        engine.lens_direct(_dev(pos), box, 15.0, 1, True)


def test_cfg3_size_two_different_frames_vs_oracle():
    """BASELINE.json configs[2] at its stated size: 1,000,000 particles, rho* = 0.8, r_cut = 1.5.
    Three DIFFERENT frames through the public driver (pair path) and through the rows path
    (neigh_rows_thread_kernel + lens_rows_thread_kernel); LENS and coordination numbers must match
    the oracle bit for bit (about 10 s of CPU work)."""
    from dynsight_b200 import engine
    from dynsight_b200 import lens as dlens
    from dynsight_b200.universe import ArrayUniverse

    n_side = 100
    n = n_side**3
    box_l = np.float32((n / 0.8) ** (1 / 3))
    rng = np.random.default_rng(3)
    grid = (np.arange(n_side, dtype=np.float64) + 0.5) * (float(box_l) / n_side)
    lat = np.stack(np.meshgrid(grid, grid, grid, indexing="ij"), axis=-1).reshape(-1, 3)
    lat = lat[rng.permutation(n)]
    cur = lat + rng.normal(0, 0.2, lat.shape)
    frames = []
    for _ in range(3):
        cur = np.mod(cur + rng.normal(0, 0.07, lat.shape), float(box_l))
        frames.append(np.minimum(cur.astype(np.float32), np.nextafter(box_l, np.float32(0))))
    pos = np.stack(frames)
    box = np.full(3, box_l, dtype=np.float32)
    dims = _dims(box, 3)
    want = lo.compute_lens_arrays(pos, dims, 1.5, delay=1)
    want_c = lo.coord_number_arrays(pos, dims, 1.5)
    assert 0.05 < want.mean() < 0.9  # the frames really differ
    uni = ArrayUniverse(pos, dims)
    got = dlens.compute_lens_device(uni, 1.5).cpu().numpy()
    assert np.array_equal(got, want), "pair path LENS differs at 1M particles"
    dpos = _dev(pos)
    _, counts = engine.lens_direct(dpos, box, 1.5, 0, True, want_lens=False)
    assert np.array_equal(counts.cpu().numpy().T.astype(np.float64), want_c)
    res, nr = engine.lens_from_positions(dpos, box, 1.5, 1)
    assert np.array_equal(res.cpu().numpy(), want), "rows path LENS differs at 1M particles"
    assert np.array_equal(nr.counts.cpu().numpy().T.astype(np.float64), want_c)


def test_cfg5_size_lens_vs_oracle():
    """BASELINE.json configs[4] shape for LENS: 100,000 atoms at water-oxygen density, r_cut = 5 A
    (k ~ 17.5), three frames of a random walk, pair path and rows path vs the oracle."""
    from dynsight_b200 import engine

    n = 100_000
    rng = np.random.default_rng(12)
    box_l = np.float32((n / 0.0334) ** (1 / 3))
    cur = rng.random((n, 3)) * float(box_l)
    frames = []
    for _ in range(3):
        cur = np.mod(cur + rng.normal(0, 0.1, cur.shape), float(box_l))
        frames.append(np.minimum(cur.astype(np.float32), np.nextafter(box_l, np.float32(0))))
    pos = np.stack(frames)
    box = np.full(3, box_l, dtype=np.float32)
    _check_direct_vs_oracle(pos, box, 5.0, 1, True)
    want = lo.compute_lens_arrays(pos, _dims(box, 3), 5.0, delay=1)
    res, _ = engine.lens_from_positions(_dev(pos), box, 5.0, 1)
    assert np.array_equal(res.cpu().numpy(), want)


def test_compute_lens_delay_zero_returns_zeros(balls7):
    """delay = 0 pairs every frame with itself in the reference (lens.py:255-257): zeros."""
    from dynsight_b200 import lens as dlens
    from dynsight_b200.universe import ArrayUniverse

    uni = ArrayUniverse(balls7["positions"][:5], balls7["dims"][:5])
    got = dlens.compute_lens(uni, 10.0, delay=0)
    assert got.shape == (7, 5)
    assert not got.any()


def test_pair_path_dense_sampling_across_the_guard_band():
    """300,000 pairs whose squared distance is spread evenly over r2 +- 1.5e-3, i.e. across the whole
    float32 guard band INCLUDING its edges (a few pairs per float32 ulp): catches a pre-test of the
    band that is a rounding error too narrow.  Both frames, pair path and rows path."""
    from dynsight_b200 import engine

    rng = np.random.default_rng(31)
    r_cut = 1.5
    r2 = (r_cut - 1e-6) ** 2
    box = np.array([107.7, 96.3, 121.1], dtype=np.float32)
    n_pairs = 300_000
    frames = []
    base = rng.random((n_pairs, 3)) * box
    for _ in range(2):
        d = rng.normal(size=(n_pairs, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        r = np.sqrt(r2 + rng.uniform(-1.5e-3, 1.5e-3, n_pairs))
        pts = np.mod(np.concatenate([base, base - d * r[:, None]]), box).astype(np.float32)
        frames.append(np.minimum(pts, np.nextafter(box, np.float32(0))))
    pos = np.stack(frames).astype(np.float32)
    dims = _dims(box, 2)
    want = lo.compute_lens_arrays(pos, dims, r_cut, delay=1)
    want_c = lo.coord_number_arrays(pos, dims, r_cut)
    lens, counts = engine.lens_direct(_dev(pos), box, r_cut, 1, True)
    assert np.array_equal(counts.cpu().numpy().T.astype(np.float64), want_c)
    assert np.array_equal(lens.cpu().numpy(), want)
    res, nr = engine.lens_from_positions(_dev(pos), box, r_cut, 1)
    assert np.array_equal(res.cpu().numpy(), want)
