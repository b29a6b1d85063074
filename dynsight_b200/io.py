"""Minimal trajectory readers that need no MDAnalysis.

They produce exactly what the hot path reads from an ``MDAnalysis.Universe`` (SURVEY.md section
8b): float32 positions in angstrom ``(F, N, 3)``, per-frame ``dimensions`` ``[lx, ly, lz, 90, 90,
90]`` float32 (or ``None``), atom names / types / 1-based ids.

Formats: GROMACS ``.gro`` topology, ``.xtc`` (compressed frames are decoded by the native
multi-threaded reader in ``csrc/xtc.cpp`` through the C ABI; ``read_xtc_uncompressed`` is a pure
numpy reader for the ``natoms <= 9`` layout of the reference's ``tests/systems/balls_7_nvt.xtc``),
and ``.xyz``.
"""

from __future__ import annotations

import re
import struct
from dataclasses import dataclass
from pathlib import Path

import numpy as np

_XTC_MAGIC = 1995


@dataclass
class Topology:
    names: np.ndarray  # (N,) str
    types: np.ndarray  # (N,) str  (MDAnalysis guesses the element from the name)
    ids: np.ndarray  # (N,) int, 1-based serials
    resnames: np.ndarray  # (N,) str


def guess_type_from_name(name: str) -> str:
    """Element guess the way MDAnalysis does for simple names (``C1`` -> ``C``)."""
    stripped = re.sub(r"[^A-Za-z]", "", name)
    if not stripped:
        return name
    if stripped.capitalize() in _TWO_LETTER:
        return stripped.capitalize()
    return stripped[0].upper()


_TWO_LETTER = {"Cl", "Br", "Na", "Mg", "Al", "Si", "Fe", "Cu", "Zn", "Ag", "Au", "Ni", "Pt"}


def read_gro(path: str | Path) -> tuple[Topology, np.ndarray, np.ndarray]:
    """Return ``(topology, positions (N,3) float32 angstrom, dimensions (6,) float32)``."""
    lines = Path(path).read_text().splitlines()
    n_atoms = int(lines[1])
    names, resnames, ids, pos = [], [], [], []
    for line in lines[2 : 2 + n_atoms]:
        resnames.append(line[5:10].strip())
        names.append(line[10:15].strip())
        ids.append(int(line[15:20]))
        pos.append([float(line[20:28]), float(line[28:36]), float(line[36:44])])
    box = [float(x) for x in lines[2 + n_atoms].split()[:3]]
    positions = np.asarray(pos, dtype=np.float32) * np.float32(10.0)
    dims = np.array([*box, 0, 0, 0], dtype=np.float32) * np.float32(10.0)
    dims[3:] = 90.0
    names_a = np.asarray(names)
    topo = Topology(
        names=names_a,
        types=np.asarray([guess_type_from_name(n) for n in names]),
        ids=np.asarray(ids, dtype=np.int64),
        resnames=np.asarray(resnames),
    )
    return topo, positions, dims


def read_xtc_uncompressed(path: str | Path) -> tuple[np.ndarray, np.ndarray]:
    """Read an XTC whose frames are stored uncompressed (natoms <= 9).

    Frame layout (big endian): magic 1995, natoms, step, time f32, 9 x f32 box (nm),
    natoms, 3*natoms x f32 coordinates (nm).  nm -> angstrom is a float32 multiply by 10, as
    MDAnalysis does.

    Returns ``(positions (F,N,3) float32, dimensions (F,6) float32)``.
    """
    raw = Path(path).read_bytes()
    magic, n_atoms = struct.unpack(">ii", raw[:8])
    if magic != _XTC_MAGIC:
        msg = f"{path}: not an XTC file (magic {magic})"
        raise ValueError(msg)
    if n_atoms > 9:  # noqa: PLR2004
        msg = (
            f"{path}: {n_atoms} atoms -> compressed XTC frames; only uncompressed XTC "
            "(natoms <= 9) is supported by dynsight_b200.io"
        )
        raise NotImplementedError(msg)
    frame_bytes = 16 + 36 + 4 + 12 * n_atoms
    if len(raw) % frame_bytes:
        msg = f"{path}: size {len(raw)} is not a multiple of the frame size {frame_bytes}"
        raise ValueError(msg)
    n_frames = len(raw) // frame_bytes
    buf = np.frombuffer(raw, dtype=">f4").reshape(n_frames, frame_bytes // 4)
    box = buf[:, 4:13].astype(np.float32).reshape(n_frames, 3, 3)
    pos = buf[:, 14:].astype(np.float32).reshape(n_frames, n_atoms, 3) * np.float32(10.0)
    dims = np.empty((n_frames, 6), dtype=np.float32)
    dims[:, 0] = box[:, 0, 0] * np.float32(10.0)
    dims[:, 1] = box[:, 1, 1] * np.float32(10.0)
    dims[:, 2] = box[:, 2, 2] * np.float32(10.0)
    dims[:, 3:] = 90.0
    return np.ascontiguousarray(pos), dims


def triclinic_dimensions(box: np.ndarray) -> np.ndarray:
    """Box vectors (F, 3, 3) float32 -> ``[lx, ly, lz, alpha, beta, gamma]`` (F, 6) float32, the
    way MDAnalysis' ``triclinic_box`` reports them (lengths of the vectors, angles in degrees;
    an all-zero box stays all-zero)."""
    box = np.asarray(box, dtype=np.float32)
    x, y, z = box[:, 0], box[:, 1], box[:, 2]
    lx = np.linalg.norm(x, axis=1)
    ly = np.linalg.norm(y, axis=1)
    lz = np.linalg.norm(z, axis=1)
    dims = np.zeros((box.shape[0], 6), dtype=np.float32)
    ok = (lx > 0) & (ly > 0) & (lz > 0)
    with np.errstate(invalid="ignore", divide="ignore"):
        alpha = np.rad2deg(np.arccos(np.einsum("ij,ij->i", y, z) / (ly * lz)))
        beta = np.rad2deg(np.arccos(np.einsum("ij,ij->i", x, z) / (lx * lz)))
        gamma = np.rad2deg(np.arccos(np.einsum("ij,ij->i", x, y) / (lx * ly)))
    dims[ok, 0], dims[ok, 1], dims[ok, 2] = lx[ok], ly[ok], lz[ok]
    dims[ok, 3], dims[ok, 4], dims[ok, 5] = alpha[ok], beta[ok], gamma[ok]
    return dims


def read_xtc(path: str | Path, n_threads: int = 0) -> tuple[np.ndarray, np.ndarray]:
    """Read any XTC trajectory (compressed or not) with the native reader (``dsg_xtc_read``).

    Returns ``(positions (F,N,3) float32 angstrom, dimensions (F,6) float32)`` -- the arrays an
    ``MDAnalysis.Universe`` built from the same file holds (``trajectory.py:68-79``).
    """
    import ctypes

    from dynsight_b200 import _lib

    lib = _lib.load()
    cpath = str(path).encode()
    n_frames = ctypes.c_int64(0)
    n_atoms = ctypes.c_int32(0)
    _lib.check(lib.dsg_xtc_index(cpath, ctypes.byref(n_frames), ctypes.byref(n_atoms), None, 0))
    offsets = np.empty(n_frames.value, dtype=np.int64)
    _lib.check(lib.dsg_xtc_index(cpath, ctypes.byref(n_frames), ctypes.byref(n_atoms),
                                 offsets.ctypes.data, offsets.size))  # fmt: skip
    pos = np.empty((n_frames.value, n_atoms.value, 3), dtype=np.float32)
    box = np.empty((n_frames.value, 9), dtype=np.float32)
    _lib.check(lib.dsg_xtc_read(cpath, offsets.ctypes.data, n_frames.value, n_atoms.value,
                                pos.ctypes.data, box.ctypes.data, None, None, int(n_threads)))  # fmt: skip
    return pos, triclinic_dimensions(box.reshape(-1, 3, 3))


def read_xyz(path: str | Path) -> tuple[Topology, np.ndarray]:
    """Read a multi-frame XYZ file -> ``(topology, positions (F,N,3) float32)``."""
    lines = Path(path).read_text().splitlines()
    n_atoms = int(lines[0].split()[0])
    stride = n_atoms + 2
    n_frames = len([ln for ln in lines if ln.strip()]) // stride
    names: list[str] = []
    pos = np.empty((n_frames, n_atoms, 3), dtype=np.float32)
    for f in range(n_frames):
        block = lines[f * stride + 2 : f * stride + 2 + n_atoms]
        for a, line in enumerate(block):
            tok = line.split()
            if f == 0:
                names.append(tok[0])
            pos[f, a] = (np.float32(tok[1]), np.float32(tok[2]), np.float32(tok[3]))
    names_a = np.asarray(names)
    topo = Topology(
        names=names_a,
        types=names_a.copy(),
        ids=np.arange(1, n_atoms + 1, dtype=np.int64),
        resnames=np.asarray([""] * n_atoms),
    )
    return topo, pos
