"""CPU-only tests of the host logic: Universe shim, selections, readers, frame batching plans,
the C-ABI surface (symbols vs header) and the "no oracle / no fallback in the product" rule."""

from __future__ import annotations

import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _universe(balls7):
    from dynsight_b200.universe import ArrayUniverse

    return ArrayUniverse(
        balls7["positions"], balls7["dims"], balls7["names"], balls7["types"], balls7["ids"]
    )


def test_selection_language(balls7):
    u = _universe(balls7)
    assert u.select_atoms("all").n_atoms == 7
    assert u.select_atoms("id 1").indices.tolist() == [0]
    assert u.select_atoms("name C3 or name C6").indices.tolist() == [2, 5]
    assert u.select_atoms("not name C5").indices.tolist() == [0, 1, 2, 3, 5, 6]
    assert u.select_atoms("type C").n_atoms == 7
    assert u.select_atoms("type O").n_atoms == 0
    assert u.select_atoms("id 2:4 and not name C3").indices.tolist() == [1, 3]
    assert u.select_atoms("(name C1 C2) or index 6").indices.tolist() == [0, 1, 6]
    sub = u.select_atoms("not name C5")
    assert sub.select_atoms("name C6 C5").indices.tolist() == [5]
    with pytest.raises(ValueError):
        u.select_atoms("frobnicate 3")
    with pytest.raises(ValueError):
        u.select_atoms("name")


def test_universe_surface(balls7):
    u = _universe(balls7)
    assert u.trajectory.n_frames == 201 and len(u.trajectory) == 201
    ag = u.select_atoms("name C3 or name C6")
    u.trajectory[5]
    assert np.array_equal(ag.positions, balls7["positions"][5][[2, 5]])
    assert np.array_equal(u.trajectory.ts.dimensions, balls7["dims"][5])
    assert np.array_equal(u.atoms.positions, balls7["positions"][5])
    seen = [ts.frame for ts in u.trajectory[10:20:3]]
    assert seen == [10, 13, 16, 19]
    assert ag[np.array([1])].indices.tolist() == [5]
    assert len(ag) == 2 and ag.atoms.types.tolist() == ["C", "C"]


def test_xyz_box_assignment_broadcasts_like_mdanalysis():
    """with a box-less reader the last assigned box wins."""
    from dynsight_b200.universe import ArrayUniverse

    u = ArrayUniverse(np.zeros((4, 3, 3), dtype=np.float32))
    assert u.trajectory.ts.dimensions is None
    for k, ts in enumerate(u.trajectory):
        ts.dimensions = np.array([1.0 + k, 2, 3, 90, 90, 90])
    u.trajectory[0]
    assert u.trajectory.ts.dimensions[0] == 4.0


def test_readers_roundtrip(tmp_path, balls7):
    from dynsight_b200 import io as dio

    # xyz
    p = tmp_path / "t.xyz"
    pos = balls7["positions"][:3]
    with p.open("w") as fh:
        for fr in pos:
            fh.write(f"{fr.shape[0]}\ncomment\n")
            for a in fr:
                fh.write(f"C {a[0]!r} {a[1]!r} {a[2]!r}\n".replace("np.float32(", "").replace(")", ""))
    topo, got = dio.read_xyz(p)
    assert np.array_equal(got, pos) and topo.types.tolist() == ["C"] * 7
    # uncompressed xtc written by hand
    import struct

    raw = b""
    for k, fr in enumerate(pos):
        raw += struct.pack(">iiif", 1995, 7, k, float(k))
        raw += struct.pack(">9f", 3, 0, 0, 0, 3, 0, 0, 0, 3)
        raw += struct.pack(">i", 7)
        raw += struct.pack(">21f", *(fr / np.float32(10)).ravel().tolist())
    q = tmp_path / "t.xtc"
    q.write_bytes(raw)
    got, dims = dio.read_xtc_uncompressed(q)
    assert got.shape == (3, 7, 3) and np.allclose(got, pos, atol=1e-4)
    assert np.array_equal(dims[0], np.array([30, 30, 30, 90, 90, 90], dtype=np.float32))
    (tmp_path / "bad.xtc").write_bytes(struct.pack(">ii", 1995, 100) + b"\0" * 100)
    with pytest.raises(NotImplementedError):
        dio.read_xtc_uncompressed(tmp_path / "bad.xtc")


def test_frame_sequence_and_batch_plan():
    from dynsight_b200 import frames

    assert frames.frame_sequence(10, None) == list(range(10))
    assert frames.frame_sequence(10, slice(2, 9, 3)) == [2, 5, 8]
    assert frames.pick_batch_frames(1000, 100.0, 3, budget_bytes=1e6) == 10
    assert frames.pick_batch_frames(10**6, 100.0, 3, budget_bytes=1e6) == 3
    assert frames.pick_batch_frames(10, 1.0, 1) <= 4096


def test_basis_matches_dscribe_recipe():
    from dynsight_b200 import soap_basis

    alphas, betas = soap_basis.gto_basis(8.0, 4, 4)
    assert alphas.shape == (5, 4) and betas.shape == (5, 4, 4)
    # alpha_{l,0} = ln(1000) for every l because a_0 = 1
    assert np.allclose(alphas[:, 0], np.log(1000.0))
    # betas orthonormalise the overlap: beta S beta^T = 1
    from math import gamma

    for l in range(5):
        m = alphas[l][:, None] + alphas[l][None, :]
        s = 0.5 * gamma(l + 1.5) * m ** (-(l + 1.5))
        assert np.allclose(betas[l] @ s @ betas[l].T, np.eye(4), atol=1e-8)
    assert soap_basis.n_features(1, 8, 8) == 324
    assert soap_basis.n_features(3, 8, 10) == 3300
    with pytest.raises(ValueError):
        soap_basis.gto_basis(0.5, 4, 4)


def test_fill_soap_vector_shapes():
    from dynsight_b200 import soap

    x = np.arange(2 * 3 * 324, dtype=np.float64).reshape(2, 3, 324)
    full = soap.fill_soap_vector_from_dscribe(x)
    assert full.shape == (2, 3, 576)  # doctest saponify.py:174
    blk = full[0, 0].reshape(9, 8, 8)
    assert np.array_equal(blk, np.transpose(blk, (0, 2, 1)))
    assert soap.fill_soap_vector_from_dscribe(x[0]).shape == (3, 576)
    assert soap.fill_soap_vector_from_dscribe(x[0, 0]).shape == (576,)
    from oracle import soap_oracle

    assert np.array_equal(full, soap_oracle.fill_soap_vector(x))


def test_fill_soap_vector_multi_species_extension():
    """Extension of saponify.py:127-202 to DScribe's multi-species layout: per l the symmetric
    (S n) x (S n) matrix; the power spectrum p^{ZZ'}_{nn'} = p^{Z'Z}_{n'n} fixes the transpose."""
    from dynsight_b200 import soap

    rng = np.random.default_rng(0)
    s_, n, lm = 3, 4, 2
    c = rng.normal(size=(lm + 1, s_ * n, 5))          # fake coefficients c[l][Zn][m]
    p = np.einsum("lam,lbm->lab", c, c)                # symmetric (S n) x (S n) per l
    flat = []
    for zi in range(s_):
        for zj in range(zi, s_):
            for l in range(lm + 1):
                blk = p[l, zi * n:(zi + 1) * n, zj * n:(zj + 1) * n]
                flat.append(blk[np.triu_indices(n)] if zi == zj else blk.ravel())
    vec = np.concatenate(flat)
    full = soap.fill_soap_vector_from_dscribe(vec, lmax=lm, nmax=n, n_species=s_)
    assert full.shape == ((lm + 1) * (s_ * n) ** 2,)
    assert np.allclose(full.reshape(lm + 1, s_ * n, s_ * n), p)
    assert soap.fill_soap_vector_from_dscribe(np.tile(vec, (2, 3, 1)), lm, n, s_).shape == (
        2, 3, full.size)  # fmt: skip
    with pytest.raises(ValueError, match="expected"):
        soap.fill_soap_vector_from_dscribe(vec[:-1], lm, n, s_)


def test_api_errors_without_gpu(balls7):
    """Errors the reference raises before any compute must not need a GPU either."""
    from dynsight_b200.trajectory import Insight, Trj

    trj = Trj(_universe(balls7))
    assert (trj.n_atoms, trj.n_frames) == (7, 201)
    assert trj.with_slice(slice(0, 50, 5)).n_frames == 10
    with pytest.raises(ValueError, match="r_cut, n_max e l_max cannot be None"):
        trj.get_timesoap()
    bad = Insight(np.zeros((2, 3, 4)), {"name": "lens"})
    with pytest.raises(ValueError, match="must be 'soap'"):
        trj.get_timesoap(soap_insight=bad)
    with pytest.raises(ValueError, match="dataset.ndim != 3."):
        Insight(np.zeros((2, 3)), {}).get_angular_velocity()
    with pytest.raises(ValueError, match="delay value outside bounds"):
        Insight(np.zeros((2, 3, 4)), {"name": "soap"}).get_angular_velocity(delay=3)
    coords = trj.with_slice(slice(0, 4)).get_coordinates("name C1")
    assert coords.shape == (4, 1, 3)


def test_insight_json_roundtrip(tmp_path):
    from dynsight_b200.trajectory import Insight

    ins = Insight(np.arange(6, dtype=np.float64).reshape(2, 3), {"name": "lens", "r_cut": 3.0})
    ins.dump_to_json(tmp_path / "a.json")
    back = Insight.load_from_json(tmp_path / "a.json")
    assert np.array_equal(back.dataset, ins.dataset) and back.meta == ins.meta
    (tmp_path / "e.json").write_text("{}")
    with pytest.raises(ValueError, match="'dataset_file' key not found"):
        Insight.load_from_json(tmp_path / "e.json")


def test_c_abi_exports_every_declared_symbol():
    """The shared library loads and exports every function include/dynsight_b200.h declares
    (no compute calls here: there is no GPU in the CPU test environment)."""
    from dynsight_b200 import _lib
    from dynsight_b200.build import build_library

    build_library()
    header = (ROOT / "include" / "dynsight_b200.h").read_text()
    declared = set(re.findall(r"\b(dsg_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SIGNATURES), "ctypes table and header disagree"
    loaded = _lib.load()
    assert b"sm_100a" in loaded.dsg_version()
    assert loaded.dsg_soap_n_features(1, 8, 8) == 324
    # argument validation happens before any CUDA call
    n = ctypes.c_size_t(0)
    box = np.array([10.0, 10.0, 10.0])
    assert loaded.dsg_neigh_workspace_bytes(1, 10, 10, 1, box.ctypes.data, 2.0, ctypes.byref(n)) == 0
    assert n.value > 0
    assert loaded.dsg_neigh_workspace_bytes(1, 10, 10, 1, box.ctypes.data, -1.0, ctypes.byref(n)) < 0
    assert b"r_cut" in loaded.dsg_last_error()


def test_soap_workspace_plans_neighbour_rows_only_where_list_mode_applies():
    """dsg_soap_workspace_bytes is host code: periodic frames with >= 3 search-radius cells per
    axis get per-centre neighbour rows (4 B x (1.5 x mean neighbours + 64), rounded to 32 entries),
    narrow boxes and non-periodic frames do not."""
    from dynsight_b200 import _lib

    lib = _lib.load()

    def ws(n_frames, n_atoms, box, n_species=1, n_max=8, l_max=8, r_cut=5.0, n_cent=None):
        n = ctypes.c_size_t(0)
        ptr = None
        if box is not None:
            arr = np.ascontiguousarray(np.tile(np.asarray(box, dtype=np.float64), (n_frames, 1)))
            ptr = arr.ctypes.data
        rc = lib.dsg_soap_workspace_bytes(n_frames, n_atoms, n_cent or n_atoms, n_species, n_max,
                                          l_max, ptr, r_cut, ctypes.byref(n))  # fmt: skip
        assert rc == 0, lib.dsg_last_error()
        return n.value

    n_atoms, n_frames = 600, 4      # about 93 neighbours inside the search radius
    radius = 5.0 + np.sqrt(-2.0 * np.log(1e-3))
    wide = 3.0 * radius * 1.001       # three cells per axis
    narrow = 3.0 * radius * 0.999     # one cell + explicit images on every axis
    k_mean = n_atoms * 4.0 / 3.0 * np.pi * radius**3 / wide**3
    cap = (int(1.5 * k_mean) + 64 + 31) & ~31
    rows = 4 * n_frames * n_atoms * cap
    extra = ws(n_frames, n_atoms, [wide] * 3) - ws(n_frames, n_atoms, [narrow] * 3)
    # the two layouts differ by the rows, the per-row counters and the cell arrays (27 vs 1 cell)
    assert rows <= extra < rows + 4 * n_frames * n_atoms * 2 + 64 * 1024
    # one narrow axis is enough to switch the rows off
    mixed = ws(n_frames, n_atoms, [wide, wide, narrow])
    assert mixed < ws(n_frames, n_atoms, [wide] * 3) - rows // 2
    # too dense for rows (mean above 600 neighbours): the in-kernel search is kept
    dense = ws(n_frames, 20 * n_atoms, [wide] * 3) - ws(n_frames, 20 * n_atoms, [narrow] * 3)
    assert dense < 64 * 1024
    # a handful of centres in a large system: rows for every atom would be wasted memory
    big_atoms, big_box = 1_000_000, [(1_000_000 / 0.0334) ** (1 / 3)] * 3
    assert ws(8, big_atoms, big_box, n_cent=10) < ws(8, big_atoms, big_box) // 4
    # several species: species-major rows, one offset per species and centre more
    extra3 = ws(n_frames, n_atoms, [wide] * 3, n_species=3) - ws(n_frames, n_atoms, [narrow] * 3,
                                                                  n_species=3)  # fmt: skip
    assert rows <= extra3 < rows + 4 * n_frames * n_atoms * 4 + 256 * 1024


def test_product_never_touches_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use oracle/."""
    for path in (ROOT / "dynsight_b200").rglob("*"):
        if path.suffix in {".py", ".cu", ".cuh", ".h"}:
            text = path.read_text()
            assert "oracle" not in text.lower(), f"{path} mentions the oracle"
            assert "/root/reference" not in text, f"{path} reads the reference tree"


def test_no_cuda_means_loud_failure(balls7):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from dynsight_b200.trajectory import Trj

    trj = Trj(_universe(balls7))
    with pytest.raises(RuntimeError, match="CUDA device"):
        trj.get_lens(r_cut=15.0)
    with pytest.raises(RuntimeError, match="CUDA device"):
        trj.get_soap(r_cut=8.0, n_max=4, l_max=4)


def test_trj_get_slice_keeps_coordinates_only(balls7):
    """trajectory.py:115-137: a MemoryReader copy of the chosen frames, without a box."""
    from dynsight_b200.trajectory import Trj
    from dynsight_b200.universe import ArrayUniverse

    trj = Trj(ArrayUniverse(balls7["positions"], balls7["dims"], balls7["names"], balls7["types"],
                            balls7["ids"]))  # fmt: skip
    sub = trj.get_slice(3, 40, 7)
    assert sub.universe.trajectory.n_frames == len(range(3, 40, 7))
    assert np.array_equal(sub.get_coordinates("all"), balls7["positions"][3:40:7])
    assert sub.universe.trajectory.ts.dimensions is None
    assert sub.universe.select_atoms("name C3").n_atoms == 1


def test_c_abi_limits_are_refused_before_any_cuda_call():
    """The limits the header states are checked on the host: n_frames * n_atoms
    below 2^31 per batch, at most 65 535 frames per batch, SOAP search radius below 37 A, basis
    sizes, species count, positive boxes, delay inside the batch -- each with its error code and a
    message, and the fused call's workspace has no global ring where the spectrum kernel keeps it
    in shared memory."""
    from dynsight_b200 import _lib

    lib = _lib.load()
    n = ctypes.c_size_t(0)
    box1 = np.array([60.0, 60.0, 60.0])
    ok = lib.dsg_soap_workspace_bytes(1, 1000, 1000, 1, 8, 8, box1.ctypes.data, 5.0, ctypes.byref(n))
    assert ok == 0 and n.value > 0
    err_invalid, err_unsupported, err_overflow = -1, -4, -5
    # n_frames * n_atoms >= 2^31
    box_many = np.tile(box1, (4, 1))
    assert lib.dsg_soap_workspace_bytes(4, 600_000_000, 10, 1, 8, 8, box_many.ctypes.data, 5.0,
                                        ctypes.byref(n)) == err_overflow
    assert b"2^31" in lib.dsg_last_error()
    # more than 65 535 frames in one batch
    assert lib.dsg_soap_workspace_bytes(70_000, 10, 10, 1, 8, 8, None, 5.0,
                                        ctypes.byref(n)) == err_invalid
    # search radius above 37 A (gamma r^2 would leave the range of the branch-free exponential)
    assert lib.dsg_soap_workspace_bytes(1, 10, 10, 1, 8, 8, None, 40.0,
                                        ctypes.byref(n)) == err_unsupported
    assert b"37" in lib.dsg_last_error()
    # basis and species limits, DScribe's r_cut > 1
    for args in ((1, 33, 8), (1, 8, 13), (17, 8, 8)):
        assert lib.dsg_soap_workspace_bytes(1, 10, 10, args[0], args[1], args[2], None, 5.0,
                                            ctypes.byref(n)) == err_unsupported
    assert lib.dsg_soap_workspace_bytes(1, 10, 10, 1, 8, 8, None, 0.9,
                                        ctypes.byref(n)) == err_invalid
    # a periodic frame needs a positive box
    bad_box = np.array([60.0, 0.0, 60.0])
    assert lib.dsg_soap_workspace_bytes(1, 10, 10, 1, 8, 8, bad_box.ctypes.data, 5.0,
                                        ctypes.byref(n)) == err_invalid
    # fused call: delay inside the batch
    box3 = np.tile(box1, (3, 1))
    assert lib.dsg_soap_timesoap_workspace_bytes(3, 100, 100, 1, 8, 8, box3.ctypes.data, 5.0, 3,
                                                 ctypes.byref(n)) == err_invalid
    assert lib.dsg_soap_timesoap_workspace_bytes(3, 100, 100, 1, 8, 8, box3.ctypes.data, 5.0, 0,
                                                 ctypes.byref(n)) == err_invalid
    # three species, delay 1: the previous spectra live in shared memory (no ring in the workspace);
    # delay 2: a global ring of 2 x (n_feat + 1) doubles per centre
    m1, m2 = ctypes.c_size_t(0), ctypes.c_size_t(0)
    box9 = np.tile(box1, (9, 1))
    assert lib.dsg_soap_timesoap_workspace_bytes(9, 2000, 2000, 3, 8, 10, box9.ctypes.data, 5.0, 1,
                                                 ctypes.byref(m1)) == 0
    assert lib.dsg_soap_timesoap_workspace_bytes(9, 2000, 2000, 3, 8, 10, box9.ctypes.data, 5.0, 2,
                                                 ctypes.byref(m2)) == 0
    ring2 = 8 * 2000 * 2 * (3300 + 1)
    assert m2.value - m1.value >= ring2 - 4096
    # LENS pair path: same overflow rule
    assert lib.dsg_lens_direct_workspace_bytes(4, 600_000_000, box_many.ctypes.data, 5.0,
                                               ctypes.byref(n)) < 0
