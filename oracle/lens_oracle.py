"""ctypes front-end of ``oracle/lens_oracle.c`` + restated trajectory drivers (TEST ONLY).

Reference lines followed (``/root/reference/src/dynsight/_internal/lens/lens.py``):
  * ``neighbors``                   <- ``neighbor_list_celllist_centers`` lens.py:69-147
  * ``lens_from_two_csr``           <- lens.py:150-189
  * ``compute_lens_arrays``         <- ``compute_lens`` driver loop lens.py:246-310
  * ``coord_number_arrays``         <- ``list_neighbours_along_trajectory`` lens.py:360-388 +
                                       ``Trj.get_coord_number`` trajectory.py:168-186
PARITY PIN: see ``lens_oracle.c`` header.
"""

from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB: ctypes.CDLL | None = None


def build(force: bool = False) -> Path:
    so = _HERE / "liblens_oracle.so"
    src = _HERE / "lens_oracle.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(
            ["/usr/bin/gcc" if __import__("os").path.exists("/usr/bin/gcc") else "gcc", "-O2", "-fPIC", "-fopenmp", "-std=c11", "-shared", "-o", str(so), str(src),
             "-lm"],
            check=True,
        )  # fmt: skip
    return so


def lib() -> ctypes.CDLL:
    global _LIB  # noqa: PLW0603
    if _LIB is None:
        _LIB = ctypes.CDLL(str(build()))
        _LIB.dsgo_neighbors.restype = ctypes.c_longlong
        _LIB.dsgo_neighbors.argtypes = [
            ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_double,
            ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p),
        ]  # fmt: skip
        _LIB.dsgo_free.argtypes = [ctypes.c_void_p]
        _LIB.dsgo_lens_from_two_csr.argtypes = [*([ctypes.c_void_p] * 4), ctypes.c_int,
                                                ctypes.c_void_p]  # fmt: skip
        _LIB.dsgo_set_num_threads.argtypes = [ctypes.c_int]
        _LIB.dsgo_max_threads.restype = ctypes.c_int
    return _LIB


def set_num_threads(n: int) -> None:
    lib().dsgo_set_num_threads(int(n))


def max_threads() -> int:
    return int(lib().dsgo_max_threads())


def neighbors(
    positions_env: np.ndarray,
    positions_cent: np.ndarray,
    r_cut: float,
    box: np.ndarray,
    respect_pbc: bool,
) -> tuple[np.ndarray, np.ndarray]:
    """CSR neighbour list of the centres (indptr int32 (M+1,), indices int32 (nnz,))."""
    pe = np.ascontiguousarray(positions_env, dtype=np.float64)
    pc = np.ascontiguousarray(positions_cent, dtype=np.float64)
    bx = np.ascontiguousarray(box, dtype=np.float64)
    m = pc.shape[0]
    indptr = np.empty(m + 1, dtype=np.int32)
    ptr = ctypes.c_void_p()
    nnz = lib().dsgo_neighbors(
        pe.ctypes.data, pe.shape[0], pc.ctypes.data, m, float(r_cut), bx.ctypes.data,
        int(bool(respect_pbc)), indptr.ctypes.data, ctypes.byref(ptr),
    )  # fmt: skip
    if nnz < 0:
        raise MemoryError
    buf = (ctypes.c_int32 * max(int(nnz), 1)).from_address(ptr.value)
    indices = np.frombuffer(buf, dtype=np.int32, count=int(nnz)).copy()
    lib().dsgo_free(ptr)
    return indptr, indices


def lens_from_two_csr(
    indptr1: np.ndarray, indices1: np.ndarray, indptr2: np.ndarray, indices2: np.ndarray
) -> np.ndarray:
    n = len(indptr1) - 1
    out = np.zeros(n, dtype=np.float64)
    a = [np.ascontiguousarray(x, dtype=np.int32) for x in (indptr1, indices1, indptr2, indices2)]
    a = [x if x.size else np.zeros(1, np.int32) for x in a]
    lib().dsgo_lens_from_two_csr(*(x.ctypes.data for x in a), n, out.ctypes.data)
    return out


def frame_box(
    positions_all_f32: np.ndarray, dims: np.ndarray | None, scale: float | None
) -> np.ndarray:
    """Box of one frame: ``ts.dimensions[:3]`` or the bounding box of *all* atoms.

    ``scale=1.01`` restates lens.py:271-274 (compute_lens), ``scale=None`` lens.py:368-371
    (list_neighbours_along_trajectory: no 1.01 factor).  float32 arithmetic, as numpy does.
    """
    if dims is not None:
        return np.asarray(dims[:3])
    mins = positions_all_f32.min(axis=0)
    maxs = positions_all_f32.max(axis=0)
    box = maxs - mins
    if scale is not None:
        box = box * scale
    return box


def compute_lens_arrays(
    positions: np.ndarray,
    dims: np.ndarray | None,
    r_cut: float,
    delay: int = 1,
    cent_idx: np.ndarray | None = None,
    env_idx: np.ndarray | None = None,
    frames: list[int] | None = None,
    respect_pbc: bool = True,
) -> np.ndarray:
    """``compute_lens`` on arrays: positions (F,N,3) float32, dims (F,6) float32 or None."""
    n_frames, n_atoms, _ = positions.shape
    fr = list(range(n_frames)) if frames is None else list(frames)
    cent_idx = np.arange(n_atoms) if cent_idx is None else np.asarray(cent_idx)
    env_idx = np.arange(n_atoms) if env_idx is None else np.asarray(env_idx)
    pairs = [(fr[i], fr[i + delay]) for i in range(len(fr) - delay)]
    if not pairs:
        msg = "No valid pairs found."
        raise RuntimeError(msg)
    out = np.zeros((len(cent_idx), len(pairs)), dtype=np.float64)
    for k, (t1, t2) in enumerate(pairs):
        csr = []
        for t in (t1, t2):
            box = frame_box(positions[t], None if dims is None else dims[t], 1.01)
            csr.append(
                neighbors(positions[t][env_idx], positions[t][cent_idx], r_cut, box, respect_pbc)
            )
        out[:, k] = lens_from_two_csr(csr[0][0], csr[0][1], csr[1][0], csr[1][1])
    return out


def coord_number_arrays(
    positions: np.ndarray,
    dims: np.ndarray | None,
    r_cut: float,
    cent_idx: np.ndarray | None = None,
    env_idx: np.ndarray | None = None,
    frames: list[int] | None = None,
    respect_pbc: bool = True,
) -> np.ndarray:
    """Coordination numbers (n_centres, n_frames) float64 as ``Trj.get_coord_number`` returns."""
    n_frames, n_atoms, _ = positions.shape
    fr = list(range(n_frames)) if frames is None else list(frames)
    cent_idx = np.arange(n_atoms) if cent_idx is None else np.asarray(cent_idx)
    env_idx = np.arange(n_atoms) if env_idx is None else np.asarray(env_idx)
    out = np.zeros((len(cent_idx), len(fr)), dtype=np.float64)
    for k, t in enumerate(fr):
        box = frame_box(positions[t], None if dims is None else dims[t], None)
        indptr, _ = neighbors(
            positions[t][env_idx], positions[t][cent_idx], r_cut, box, respect_pbc
        )
        out[:, k] = np.diff(indptr)
    return out
