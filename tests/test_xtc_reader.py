"""Native XTC reader (dynsight_b200/csrc/xtc.cpp): host code, runs
without a GPU.  Checked against (a) frames written by the independent test writer in
``xtc_writer.py``, (b) the numpy reader of the uncompressed layout, (c) the decode of the
reference's ``tests/systems/coex/test_coex.xtc`` stored in ``tests/golden/coex.npz`` (the O-atom
coordinates and a SHA-256 of all coordinates; the file itself is read only where the reference tree
is mounted)."""

from __future__ import annotations

import hashlib
import struct
from pathlib import Path

import numpy as np
import pytest
from xtc_writer import MAGICINTS, write_frame

from dynsight_b200 import io as dsg_io

GOLDEN = Path(__file__).resolve().parent / "golden"
REF_COEX = Path("/root/reference/tests/systems/coex/test_coex.xtc")


def _synthetic(rng: np.random.Generator, n_groups: int, spread: int, smallidx: int):
    """Integer coordinates + layout: waters with adapting small tables and single atoms."""
    coords, layout = [], []
    idx = smallidx
    for g in range(n_groups):
        if g % 5 == 4:  # noqa: PLR2004
            coords.append(rng.integers(-spread, spread, size=3))
            layout.append(("single", 0))
            continue
        adapt = int(rng.integers(-1, 2))
        if idx + adapt < 9 or idx + adapt > 40:  # noqa: PLR2004
            adapt = 0
        half = MAGICINTS[idx] // 2
        p0 = rng.integers(-spread, spread, size=3)
        p1 = p0 + rng.integers(-half + 1, half - 1, size=3) // 2
        p2 = p0 + rng.integers(-half + 1, half - 1, size=3) // 2
        coords += [p0, p1, p2]
        layout.append(("water", adapt))
        idx += adapt
    return np.asarray(coords, dtype=np.int64), layout


@pytest.mark.parametrize(("spread", "smallidx", "precision"),
                         [(3000, 21, 1000.0), (200, 12, 100.0), (2_000_000, 30, 1000.0),
                          (40_000_000, 33, 1000.0)])  # fmt: skip  (last: sizes above 2^24)
def test_reader_decodes_written_frames(tmp_path, spread, smallidx, precision):
    rng = np.random.default_rng(spread + smallidx)
    frames, expect = [], []
    for f in range(5):
        coords, layout = _synthetic(rng, 40 + f, spread, smallidx)
        if f > 0:  # same atom count in every frame
            coords, layout = _synthetic(np.random.default_rng(f), 40, spread, smallidx)
        box = np.diag([3.0 + f, 4.0, 5.0]).astype(np.float32)
        frames.append(write_frame(coords, box, f, 0.5 * f, precision, smallidx, layout))
        inv = np.float32(1.0 / np.float32(precision))
        expect.append(coords.astype(np.float32) * inv * np.float32(10.0))
    n0 = expect[0].shape[0]
    keep = [k for k, e in enumerate(expect) if e.shape[0] == n0]
    path = tmp_path / "synthetic.xtc"
    path.write_bytes(b"".join(frames[k] for k in keep))
    pos, dims = dsg_io.read_xtc(path, n_threads=3)
    assert pos.shape == (len(keep), n0, 3)
    for row, k in enumerate(keep):
        assert np.array_equal(pos[row], expect[k])
        assert np.allclose(dims[row], [10.0 * (3.0 + k), 40.0, 50.0, 90, 90, 90])


def test_uncompressed_layout_matches_numpy_reader(tmp_path):
    rng = np.random.default_rng(0)
    n, nf = 7, 4
    xyz = rng.random((nf, n, 3)).astype(np.float32) * 3
    raw = b""
    for f in range(nf):
        raw += struct.pack(">iiif", 1995, n, f, 0.1 * f)
        raw += struct.pack(">9f", 3, 0, 0, 0, 3, 0, 0, 0, 3)
        raw += struct.pack(">i", n) + xyz[f].astype(">f4").tobytes()
    path = tmp_path / "small.xtc"
    path.write_bytes(raw)
    p1, d1 = dsg_io.read_xtc(path)
    p2, d2 = dsg_io.read_xtc_uncompressed(path)
    assert np.array_equal(p1, p2)
    assert np.array_equal(d1, d2)


def test_truncated_file_fails_loudly(tmp_path):
    rng = np.random.default_rng(1)
    coords, layout = _synthetic(rng, 30, 1000, 20)
    blob = write_frame(coords, np.eye(3, dtype=np.float32), 0, 0.0, 1000.0, 20, layout)
    path = tmp_path / "cut.xtc"
    path.write_bytes(blob[: len(blob) - 8])
    with pytest.raises(Exception, match="XTC"):
        dsg_io.read_xtc(path)


@pytest.mark.skipif(not REF_COEX.exists(), reason="reference tree not mounted")
def test_reference_coex_trajectory_decodes_to_the_stored_fixture():
    fx = np.load(GOLDEN / "coex.npz")
    pos, dims = dsg_io.read_xtc(REF_COEX)
    assert hashlib.sha256(np.ascontiguousarray(pos).tobytes()).hexdigest() == str(fx["sha256_all"])
    assert np.array_equal(pos[:, fx["o_index"]], fx["o_positions"])
    assert np.array_equal(dims, fx["dims"])


def test_all_zero_box_means_no_box_like_mdanalysis(tmp_path):
    """MDAnalysis hands out ``ts.dimensions = None`` for a frame whose box is all zero (then LENS
    uses the bounding-box rule of lens.py:271-274): ``from_gro_xtc`` turns such a file into a
    box-less universe; a file that mixes boxed and box-less frames is refused."""
    from dynsight_b200.universe import ArrayUniverse

    rng = np.random.default_rng(2)
    n, nf = 5, 3
    xyz = rng.random((nf, n, 3)).astype(np.float32) * 3

    def write(boxes, name):
        raw = b""
        for f in range(nf):
            raw += struct.pack(">iiif", 1995, n, f, 0.1 * f)
            b = boxes[f]
            raw += struct.pack(">9f", b, 0, 0, 0, b, 0, 0, 0, b)
            raw += struct.pack(">i", n) + xyz[f].astype(">f4").tobytes()
        path = tmp_path / name
        path.write_bytes(raw)
        return path

    gro = tmp_path / "t.gro"
    lines = ["t", f"{n:5d}"]
    for i in range(n):
        lines.append(f"{1:5d}{'RES':<5s}{'O':>5s}{i + 1:5d}{xyz[0, i, 0]:8.3f}{xyz[0, i, 1]:8.3f}{xyz[0, i, 2]:8.3f}")
    lines.append("   3.00000   3.00000   3.00000")
    gro.write_text("\n".join(lines) + "\n")
    uni = ArrayUniverse.from_gro_xtc(gro, write([0.0, 0.0, 0.0], "nobox.xtc"))
    assert uni.trajectory[0].dimensions is None
    boxed = ArrayUniverse.from_gro_xtc(gro, write([3.0, 3.0, 3.0], "box.xtc"))
    assert np.allclose(boxed.trajectory[1].dimensions, [30.0, 30.0, 30.0, 90, 90, 90])
    with pytest.raises(ValueError, match="all-zero"):
        ArrayUniverse.from_gro_xtc(gro, write([3.0, 0.0, 3.0], "mixed.xtc"))
