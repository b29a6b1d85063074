"""Device-level operators: torch CUDA tensors in, torch CUDA tensors out, through the C ABI.

These mirror the reference's L2 kernels (SURVEY.md section 8b "lower-level kernel boundary"),
batched over frames:

* :func:`neighbor_lists`      <- ``neighbor_list_celllist_centers``  (lens/lens.py:69-147)
* :func:`lens_pairs` / :func:`lens_from_two_csr` <- ``lens_from_two_csr`` (lens/lens.py:150-189)
* :func:`coordination_numbers` <- ``Trj.get_coord_number`` counting loop (trajectory.py:168-184)
* :func:`timesoap_device`     <- ``timesoap`` (timesoap/timesoap.py:130-179)
* :func:`soap_power_spectrum` <- DScribe ``SOAP.create`` as called at soapify/saponify.py:98-122

PyTorch is plumbing only (device memory, streams).  Every operator calls hand-written sm_100a
kernels; there is no CPU or eager-torch fallback.
"""

from __future__ import annotations

import ctypes
from dataclasses import dataclass

import numpy as np
import torch

from dynsight_b200 import _lib


def launch_count() -> int:
    """Kernel launches issued by libdynsight_b200 in this process (counted inside the library)."""
    return int(_lib.load().dsg_launch_count())


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return _lib.DSG_F32
    if t.dtype == torch.float64:
        return _lib.DSG_F64
    msg = f"positions must be float32 or float64, got {t.dtype}"
    raise TypeError(msg)


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        msg = f"{name} must be a CUDA tensor (dynsight_b200 has no CPU path)"
        raise RuntimeError(msg)
    if not t.is_contiguous():
        msg = f"{name} must be contiguous"
        raise ValueError(msg)


def _host_box(box: np.ndarray | torch.Tensor, n_frames: int) -> np.ndarray:
    if isinstance(box, torch.Tensor):
        box = box.detach().cpu().numpy()
    b = np.ascontiguousarray(np.asarray(box, dtype=np.float64))
    if b.ndim == 1:
        b = np.ascontiguousarray(np.broadcast_to(b[None, :3], (n_frames, 3)))
    b = np.ascontiguousarray(b[:, :3])
    if b.shape != (n_frames, 3):
        msg = f"box must have shape ({n_frames}, 3), got {b.shape}"
        raise ValueError(msg)
    return b


@dataclass
class NeighborBatch:
    """Sorted CSR neighbour lists of a batch of frames (device resident).

    Row ``i`` of frame ``f`` is ``indices[frame_off[f] + indptr[f, i] : frame_off[f] + indptr[f, i+1]]``.
    """

    counts: torch.Tensor  # (F, M) int32
    indptr: torch.Tensor  # (F, M+1) int32, frame-local
    frame_off: torch.Tensor  # (F+1,) int64
    indices: torch.Tensor  # (nnz,) int32
    n_frames: int
    n_cent: int

    def frame_csr(self, f: int) -> tuple[torch.Tensor, torch.Tensor]:
        """(indptr int32 (M+1,), indices int32 (nnz_f,)) of one frame -- the numba return value."""
        off = self.frame_off[f : f + 2].tolist()
        return self.indptr[f], self.indices[off[0] : off[1]]


def neighbor_lists(
    pos_env: torch.Tensor,
    box: np.ndarray | torch.Tensor,
    r_cut: float,
    pos_cent: torch.Tensor | None = None,
    respect_pbc: bool = True,
    counts_only: bool = False,
    two_pass: bool = False,
) -> NeighborBatch:
    """Sorted CSR neighbour lists for every frame of a batch.

    ``pos_env`` (F, N, 3) and optional ``pos_cent`` (F, M, 3): CUDA float32/float64.
    ``box`` (F, 3) or (3,) box lengths (host).  ``pos_cent=None`` means centres == env atoms.

    Default: the single-traversal rows path followed by a copy-and-sort of every row into its
    canonical CSR place (``dsg_rows_csr_offsets`` + ``dsg_rows_to_csr``).  ``two_pass=True`` uses
    the count / fill protocol (``dsg_neigh_count`` + ``dsg_neigh_fill``, two traversals); both give
    identical arrays.
    """
    lib = _lib.load()
    _require_cuda(pos_env, "pos_env")
    if pos_env.ndim != 3 or pos_env.shape[2] != 3:  # noqa: PLR2004
        msg = f"pos_env must have shape (F, N, 3), got {tuple(pos_env.shape)}"
        raise ValueError(msg)
    n_frames, n_env, _ = pos_env.shape
    code = _dtype_code(pos_env)
    if pos_cent is not None:
        _require_cuda(pos_cent, "pos_cent")
        if pos_cent.dtype != pos_env.dtype or pos_cent.shape[0] != n_frames:
            msg = "pos_cent must match pos_env in dtype and frame count"
            raise ValueError(msg)
        n_cent = pos_cent.shape[1]
    else:
        n_cent = n_env
    h_box = _host_box(box, n_frames)
    if not two_pass and not counts_only:
        return _neighbor_lists_from_rows(pos_env, h_box, r_cut, pos_cent, respect_pbc)
    dev = pos_env.device
    ws_bytes = ctypes.c_size_t(0)
    _lib.check(
        lib.dsg_neigh_workspace_bytes(
            n_frames, n_env, n_cent, int(pos_cent is None), h_box.ctypes.data, float(r_cut),
            ctypes.byref(ws_bytes),
        )
    )  # fmt: skip
    workspace = torch.empty(ws_bytes.value, dtype=torch.uint8, device=dev)
    counts = torch.empty((n_frames, n_cent), dtype=torch.int32, device=dev)
    indptr = torch.empty((n_frames, n_cent + 1), dtype=torch.int32, device=dev)
    frame_off = torch.empty(n_frames + 1, dtype=torch.int64, device=dev)
    cent_ptr = pos_cent.data_ptr() if pos_cent is not None else None
    common = (
        pos_env.data_ptr(), cent_ptr, code, n_frames, n_env, n_cent, h_box.ctypes.data,
        float(r_cut), int(bool(respect_pbc)), workspace.data_ptr(), ws_bytes.value,
    )  # fmt: skip
    _lib.check(
        lib.dsg_neigh_count(
            *common, counts.data_ptr(), indptr.data_ptr(), frame_off.data_ptr(), _stream_ptr()
        )
    )
    if counts_only:
        indices = torch.empty(0, dtype=torch.int32, device=dev)
        return NeighborBatch(counts, indptr, frame_off, indices, n_frames, n_cent)
    nnz = int(frame_off[-1].item())  # the one host sync of the two-phase protocol
    indices = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
    _lib.check(
        lib.dsg_neigh_fill(
            *common, indptr.data_ptr(), frame_off.data_ptr(), indices.data_ptr(), _stream_ptr()
        )
    )
    return NeighborBatch(counts, indptr, frame_off, indices[:nnz], n_frames, n_cent)


_MAX_ROW_RETRIES = 8


def _next_capacity(nr: "NeighborRows", previous: int | None) -> int:
    """Rows capacity for the retry after a reservation overflow: what the cursors say was needed
    (+10 %), and never less than twice the previous attempt, so the retry loop always ends."""
    need = int(nr.needed_capacity() * 1.1) + _ROW_PARTS * 2048
    prev = int(previous) if previous is not None else int(nr.rows.numel())
    return max(need, 2 * prev)


def neighbor_rows_checked(
    pos_env: torch.Tensor,
    box: np.ndarray | torch.Tensor,
    r_cut: float,
    pos_cent: torch.Tensor | None = None,
    respect_pbc: bool = True,
    capacity: int | None = None,
    inclusive_cutoff: bool = False,
) -> "NeighborRows":
    """:func:`neighbor_rows` + the overflow check (one host sync) and the retries."""
    force_warp = False
    for _ in range(_MAX_ROW_RETRIES):
        nr = neighbor_rows(pos_env, box, r_cut, pos_cent, respect_pbc, capacity, force_warp,
                           inclusive_cutoff)  # fmt: skip
        if not nr.overflowed():
            return nr
        if nr.needs_warp_kernel():
            force_warp = True
        else:
            capacity = _next_capacity(nr, capacity)
    msg = f"neighbour rows still overflow after {_MAX_ROW_RETRIES} attempts (capacity {capacity})"
    raise RuntimeError(msg)


def _neighbor_lists_from_rows(
    pos_env: torch.Tensor,
    box: np.ndarray | torch.Tensor,
    r_cut: float,
    pos_cent: torch.Tensor | None,
    respect_pbc: bool,
) -> NeighborBatch:
    lib = _lib.load()
    nr = neighbor_rows_checked(pos_env, box, r_cut, pos_cent, respect_pbc)
    dev = pos_env.device
    n_frames, n_cent = nr.n_frames, nr.n_cent
    ws_bytes = ctypes.c_size_t(0)
    _lib.check(lib.dsg_rows_csr_workspace_bytes(n_frames, n_cent, ctypes.byref(ws_bytes)))
    workspace = torch.empty(ws_bytes.value, dtype=torch.uint8, device=dev)
    indptr = torch.empty((n_frames, n_cent + 1), dtype=torch.int32, device=dev)
    frame_off = torch.empty(n_frames + 1, dtype=torch.int64, device=dev)
    _lib.check(
        lib.dsg_rows_csr_offsets(
            nr.counts.data_ptr(), n_frames, n_cent, workspace.data_ptr(), ws_bytes.value,
            indptr.data_ptr(), frame_off.data_ptr(), _stream_ptr(),
        )
    )
    nnz = int(frame_off[-1].item())
    indices = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
    _lib.check(
        lib.dsg_rows_to_csr(
            nr.counts.data_ptr(), nr.row_start.data_ptr(), nr.rows.data_ptr(), indptr.data_ptr(),
            frame_off.data_ptr(), n_frames, n_cent, indices.data_ptr(),
            int(np.ceil(nr.k_est)) if nr.k_est > 0 else 0, _stream_ptr(),
        )
    )
    return NeighborBatch(nr.counts, indptr, frame_off, indices[:nnz], n_frames, n_cent)


@dataclass
class NeighborRows:
    """Unsorted neighbour rows of a batch of frames at atomically reserved offsets (LENS fast path).

    Row ``i`` of frame ``f`` is ``rows[row_start[f, i] : row_start[f, i] + counts[f, i]]``.
    """

    counts: torch.Tensor  # (F, M) int32
    row_start: torch.Tensor  # (F, M) int64
    rows: torch.Tensor  # (capacity,) int32
    status: torch.Tensor  # (2 + 256,) int64: [0] overflow flag, [2:] region cursors
    n_frames: int
    n_cent: int
    k_est: float = 0.0  # expected row length from the density (host estimate)

    def overflowed(self) -> bool:
        return bool(self.status[0].item() != 0)

    def needs_warp_kernel(self) -> bool:
        """The thread-per-centre kernel met a row longer than its staging tile (status 2)."""
        return bool(self.status[0].item() == 2)

    def needed_capacity(self) -> int:
        return int(self.status[2:].max().item()) * max(1, int(self.status[1].item()))


_ROW_PARTS = 256
_LENS_DENSE_ROWS = 1 << 30  # bit 30 of dsg_lens_pairs_rows' delay: warp-per-centre kernel
_LENS_THREAD_MAX_K = 24.0  # above this expected row length one warp per centre is faster


def neighbor_rows(
    pos_env: torch.Tensor,
    box: np.ndarray | torch.Tensor,
    r_cut: float,
    pos_cent: torch.Tensor | None = None,
    respect_pbc: bool = True,
    capacity: int | None = None,
    force_warp_kernel: bool = False,
    inclusive_cutoff: bool = False,
) -> NeighborRows:
    """Single-traversal neighbour rows (same membership as :func:`neighbor_lists`, rows unsorted).

    ``inclusive_cutoff`` switches the membership rule from the numba kernels'
    ``dr2 < (r_cut - 1e-6)**2`` to FastNS's ``dr2 <= r_cut**2`` (spatial average).

    The library picks one thread per centre for sparse cells and one warp per cell otherwise;
    ``force_warp_kernel`` disables the former.

    The caller must check ``overflowed()`` once the work is done (one host sync) and retry with
    ``needed_capacity()``; :func:`lens_from_positions` does that.
    """
    lib = _lib.load()
    _require_cuda(pos_env, "pos_env")
    n_frames, n_env, _ = pos_env.shape
    code = _dtype_code(pos_env)
    if pos_cent is not None:
        _require_cuda(pos_cent, "pos_cent")
        n_cent = pos_cent.shape[1]
    else:
        n_cent = n_env
    h_box = _host_box(box, n_frames)
    dev = pos_env.device
    vol = np.prod(np.maximum(h_box, 1e-30), axis=1)
    k_est = np.minimum(n_env, n_env / vol * 4.18879 * float(r_cut) ** 3)
    if capacity is None:
        capacity = int(1.3 * float(np.sum(k_est)) * n_cent) + _ROW_PARTS * 2048
    capacity = max(int(capacity), _ROW_PARTS) // _ROW_PARTS * _ROW_PARTS
    ws_bytes = ctypes.c_size_t(0)
    _lib.check(
        lib.dsg_neigh_workspace_bytes(
            n_frames, n_env, n_cent, int(pos_cent is None), h_box.ctypes.data, float(r_cut),
            ctypes.byref(ws_bytes),
        )
    )  # fmt: skip
    workspace = torch.empty(ws_bytes.value, dtype=torch.uint8, device=dev)
    counts = torch.empty((n_frames, n_cent), dtype=torch.int32, device=dev)
    row_start = torch.empty((n_frames, n_cent), dtype=torch.int64, device=dev)
    rows = torch.empty(capacity, dtype=torch.int32, device=dev)
    status = torch.empty(2 + _ROW_PARTS, dtype=torch.int64, device=dev)
    _lib.check(
        lib.dsg_neigh_rows(
            pos_env.data_ptr(), pos_cent.data_ptr() if pos_cent is not None else None, code,
            n_frames, n_env, n_cent, h_box.ctypes.data, float(r_cut),
            int(bool(respect_pbc)) | (2 if force_warp_kernel else 0)
            | (4 if inclusive_cutoff else 0),
            workspace.data_ptr(), ws_bytes.value, counts.data_ptr(), row_start.data_ptr(),
            rows.data_ptr(), capacity, status.data_ptr(), _stream_ptr(),
        )
    )  # fmt: skip
    return NeighborRows(counts, row_start, rows, status, n_frames, n_cent, float(np.max(k_est)))


def lens_pairs_rows(
    nr: NeighborRows, delay: int = 1, out: torch.Tensor | None = None, col0: int = 0
) -> torch.Tensor:
    """LENS of frames (a, a+delay) inside one rows batch -> (M, n_pairs) float64."""
    lib = _lib.load()
    n_pairs = nr.n_frames - delay
    if n_pairs < 1:
        msg = "No valid pairs found."
        raise RuntimeError(msg)
    if out is None:
        out = torch.empty((nr.n_cent, n_pairs), dtype=torch.float64, device=nr.counts.device)
        col0 = 0
    _lib.check(
        lib.dsg_lens_pairs_rows(
            nr.counts.data_ptr(), nr.row_start.data_ptr(), nr.rows.data_ptr(), nr.n_frames,
            nr.n_cent, int(delay) | (_LENS_DENSE_ROWS if nr.k_est > _LENS_THREAD_MAX_K else 0),
            out.data_ptr(), out.stride(0), int(col0), _stream_ptr(),
        )
    )  # fmt: skip
    return out


def lens_from_positions(
    pos_env: torch.Tensor,
    box: np.ndarray | torch.Tensor,
    r_cut: float,
    delay: int = 1,
    pos_cent: torch.Tensor | None = None,
    respect_pbc: bool = True,
    out: torch.Tensor | None = None,
    col0: int = 0,
    capacity: int | None = None,
    force_warp: bool = False,
) -> tuple[torch.Tensor, NeighborRows]:
    """LENS for all frame pairs of a batch straight from positions (rows path + retry on the rare
    reservation overflow).  Returns the (M, n_pairs) result and the rows batch (for the counts).
    ``force_warp`` starts with the warp-per-cell kernel (a caller that already saw status 2)."""
    for _ in range(_MAX_ROW_RETRIES):
        nr = neighbor_rows(pos_env, box, r_cut, pos_cent, respect_pbc, capacity, force_warp)
        res = lens_pairs_rows(nr, delay, out=out, col0=col0)
        if not nr.overflowed():
            return res, nr
        if nr.needs_warp_kernel():
            force_warp = True  # rows longer than the thread kernel's tile: one warp per cell
        else:
            capacity = _next_capacity(nr, capacity)
    msg = f"neighbour rows still overflow after {_MAX_ROW_RETRIES} attempts (capacity {capacity})"
    raise RuntimeError(msg)


def lens_direct_supported(
    n_frames: int, n_atoms: int, box: np.ndarray | torch.Tensor, r_cut: float,
    respect_pbc: bool = True,
) -> bool:  # fmt: skip
    """Whether :func:`lens_direct` (pair path, centres == environment) applies to such a batch."""
    if n_frames < 1 or n_atoms < 1:
        return False
    h_box = _host_box(box, n_frames)
    if not bool(np.all(np.isfinite(h_box) & (h_box > 0.0))):
        return False
    rc = _lib.load().dsg_lens_direct_supported(
        n_frames, n_atoms, h_box.ctypes.data, float(r_cut), int(bool(respect_pbc))
    )
    return rc == 1


def lens_direct(
    pos: torch.Tensor,
    box: np.ndarray | torch.Tensor,
    r_cut: float,
    delay: int = 1,
    respect_pbc: bool = True,
    out: torch.Tensor | None = None,
    col0: int = 0,
    want_lens: bool = True,
    want_counts: bool = True,
) -> tuple[torch.Tensor | None, torch.Tensor | None]:
    """LENS and coordination numbers of a batch straight from positions, centres == environment,
    without neighbour lists (``dsg_lens_direct``: half-shell pair search, every pair found in frame
    t is looked up in frame t + delay).  Returns ``(lens (N, F - delay) float64 | None,
    counts (F, N) int32 | None)``; raises ``DsgError`` (unsupported) when
    :func:`lens_direct_supported` is false."""
    lib = _lib.load()
    _require_cuda(pos, "pos")
    n_frames, n_atoms, _ = pos.shape
    h_box = _host_box(box, n_frames)
    dev = pos.device
    ws_bytes = ctypes.c_size_t(0)
    _lib.check(
        lib.dsg_lens_direct_workspace_bytes(
            n_frames, n_atoms, h_box.ctypes.data, float(r_cut), ctypes.byref(ws_bytes)
        )
    )
    workspace = torch.empty(ws_bytes.value, dtype=torch.uint8, device=dev)
    counts = torch.empty((n_frames, n_atoms), dtype=torch.int32, device=dev) if want_counts else None
    if want_lens:
        if n_frames - delay < 1 or delay < 1:
            msg = "No valid pairs found."
            raise RuntimeError(msg)
        if out is None:
            out = torch.empty((n_atoms, n_frames - delay), dtype=torch.float64, device=dev)
            col0 = 0
    else:
        out = None
    _lib.check(
        lib.dsg_lens_direct(
            pos.data_ptr(), _dtype_code(pos), n_frames, n_atoms, h_box.ctypes.data, float(r_cut),
            int(bool(respect_pbc)), int(delay) if want_lens else 0, workspace.data_ptr(),
            ws_bytes.value, counts.data_ptr() if counts is not None else None,
            out.data_ptr() if out is not None else None,
            out.stride(0) if out is not None else 0, int(col0), _stream_ptr(),
        )
    )  # fmt: skip
    return out, counts


def wrap_positions(pos: torch.Tensor, box: np.ndarray | torch.Tensor) -> torch.Tensor:
    """Positions (F, N, 3) folded into ``[0, box)`` per frame, in the input precision."""
    lib = _lib.load()
    _require_cuda(pos, "pos")
    n_frames, n_atoms, _ = pos.shape
    h_box = _host_box(box, n_frames)
    d_box = torch.as_tensor(h_box, device=pos.device)
    out = torch.empty_like(pos)
    _lib.check(
        lib.dsg_wrap_positions(
            pos.data_ptr(), _dtype_code(pos), n_frames, n_atoms, d_box.data_ptr(), out.data_ptr(),
            _stream_ptr(),
        )
    )
    return out


def spatial_average_rows(
    nr: NeighborRows, values: torch.Tensor, out: torch.Tensor, frame0: int = 0
) -> torch.Tensor:
    """Mean of ``values`` over every particle's neighbours and itself for the frames of ``nr``.

    ``values`` / ``out``: (N, F_total) or (N, F_total, C) float64 CUDA tensors (components
    contiguous); the batch covers the columns ``frame0 .. frame0 + nr.n_frames``.
    """
    lib = _lib.load()
    _require_cuda(values, "values")
    if values.dtype != torch.float64 or out.dtype != torch.float64:
        msg = "values and out must be float64"
        raise TypeError(msg)
    n_comp = 1 if values.ndim == 2 else values.shape[2]  # noqa: PLR2004
    if values.ndim == 3 and (values.stride(2) != 1 or out.stride(2) != 1):  # noqa: PLR2004
        msg = "components must be contiguous"
        raise ValueError(msg)
    v = values[:, frame0:]
    o = out[:, frame0:]
    _lib.check(
        lib.dsg_spatial_average_rows(
            nr.counts.data_ptr(), nr.row_start.data_ptr(), nr.rows.data_ptr(), nr.n_frames,
            nr.n_cent, int(n_comp), v.data_ptr(), v.stride(0), v.stride(1), o.data_ptr(),
            o.stride(0), o.stride(1), int(np.ceil(nr.k_est)) if nr.k_est > 0 else 0,
            _stream_ptr(),
        )
    )  # fmt: skip
    return out


def self_time_correlation_device(
    data: torch.Tensor, max_delay: int, cross: bool = False
) -> tuple[torch.Tensor, torch.Tensor]:
    """Per-lag mean and standard deviation over the particles of the un-normalised self
    correlation ``c_i(t') = sum_t x_i(t) x_i(t+t')`` of the mean-subtracted rows of ``data``
    (``cross=True``: over all ordered pairs ``i != j`` of ``sum_t x_i(t) x_j(t+t')``)."""
    lib = _lib.load()
    _require_cuda(data, "data")
    if data.dtype != torch.float64 or data.ndim != 2 or data.stride(1) != 1:  # noqa: PLR2004
        msg = "data must be a (n_particles, n_frames) float64 tensor with contiguous rows"
        raise ValueError(msg)
    n_part, n_frames = data.shape
    ws_bytes = ctypes.c_size_t(0)
    ws_fn = lib.dsg_cross_tcf_workspace_bytes if cross else lib.dsg_tcf_workspace_bytes
    run_fn = lib.dsg_cross_time_correlation if cross else lib.dsg_self_time_correlation
    _lib.check(ws_fn(n_part, n_frames, int(max_delay), ctypes.byref(ws_bytes)))
    workspace = torch.empty(ws_bytes.value, dtype=torch.uint8, device=data.device)
    mean = torch.empty(int(max_delay), dtype=torch.float64, device=data.device)
    std = torch.empty_like(mean)
    _lib.check(
        run_fn(
            data.data_ptr(), data.stride(0), n_part, n_frames, int(max_delay),
            workspace.data_ptr(), ws_bytes.value, mean.data_ptr(), std.data_ptr(), _stream_ptr(),
        )
    )  # fmt: skip
    return mean, std


def lens_pairs(
    batch: NeighborBatch, delay: int = 1, out: torch.Tensor | None = None, col0: int = 0
) -> torch.Tensor:
    """LENS of frames (a, a+delay) for all a inside one batch -> (M, n_pairs) float64."""
    lib = _lib.load()
    n_pairs = batch.n_frames - delay
    if n_pairs < 1:
        msg = "No valid pairs found."
        raise RuntimeError(msg)
    if out is None:
        out = torch.empty((batch.n_cent, n_pairs), dtype=torch.float64, device=batch.indptr.device)
        col0 = 0
    _lib.check(
        lib.dsg_lens_pairs(
            batch.indptr.data_ptr(), batch.frame_off.data_ptr(), batch.indices.data_ptr(),
            batch.n_frames, batch.n_cent, int(delay), out.data_ptr(), out.stride(0), int(col0),
            _stream_ptr(),
        )
    )  # fmt: skip
    return out


def lens_from_two_csr(
    indptr1: torch.Tensor, indices1: torch.Tensor, indptr2: torch.Tensor, indices2: torch.Tensor
) -> torch.Tensor:
    """Same signature as the reference's ``lens_from_two_csr`` (lens.py:150-189), on device."""
    lib = _lib.load()
    for name, t in (("indptr1", indptr1), ("indices1", indices1), ("indptr2", indptr2),
                    ("indices2", indices2)):  # fmt: skip
        _require_cuda(t, name)
        if t.dtype != torch.int32:
            msg = f"{name} must be int32"
            raise TypeError(msg)
    n_cent = indptr1.numel() - 1
    out = torch.empty(n_cent, dtype=torch.float64, device=indptr1.device)
    _lib.check(
        lib.dsg_lens_from_two_csr(
            indptr1.data_ptr(), indices1.data_ptr(), indptr2.data_ptr(), indices2.data_ptr(),
            n_cent, out.data_ptr(), _stream_ptr(),
        )
    )  # fmt: skip
    return out


def coordination_numbers(
    counts: torch.Tensor, out: torch.Tensor | None = None, col0: int = 0
) -> torch.Tensor:
    """counts (F, M) int32 -> (M, F) float64, the layout of the coord_number Insight."""
    lib = _lib.load()
    _require_cuda(counts, "counts")
    n_frames, n_cent = counts.shape
    if out is None:
        out = torch.empty((n_cent, n_frames), dtype=torch.float64, device=counts.device)
        col0 = 0
    _lib.check(
        lib.dsg_counts_to_f64(
            counts.data_ptr(), n_frames, n_cent, out.data_ptr(), out.stride(0), int(col0),
            _stream_ptr(),
        )
    )  # fmt: skip
    return out


def timesoap_device(
    soap: torch.Tensor, delay: int = 1, out: torch.Tensor | None = None, col0: int = 0
) -> torch.Tensor:
    """timeSOAP of a (n_particles, n_frames, n_comp) float64 CUDA tensor (any frame/particle strides,
    unit stride along components) -> (n_particles, n_frames - delay) float64."""
    lib = _lib.load()
    if not soap.is_cuda:
        msg = "soap must be a CUDA tensor (dynsight_b200 has no CPU path)"
        raise RuntimeError(msg)
    if soap.dtype != torch.float64 or soap.ndim != 3:  # noqa: PLR2004
        msg = "soap must be float64 with shape (n_particles, n_frames, n_comp)"
        raise ValueError(msg)
    n_p, n_f, n_c = soap.shape
    if delay < 1 or delay >= n_f:
        msg = "delay value outside bounds"
        raise ValueError(msg)
    if soap.stride(2) != 1:
        soap = soap.contiguous()
    if out is None:
        out = torch.empty((n_p, n_f - delay), dtype=torch.float64, device=soap.device)
        col0 = 0
    _lib.check(
        lib.dsg_timesoap(
            soap.data_ptr(), n_p, n_f, n_c, soap.stride(0), soap.stride(1), int(delay),
            out.data_ptr(), out.stride(0), int(col0), _stream_ptr(),
        )
    )  # fmt: skip
    return out


def normalize_soap_device(soap: torch.Tensor) -> torch.Tensor:
    """Unit-normalise the last axis of a float64 CUDA tensor (zero vectors stay zero)."""
    lib = _lib.load()
    if not soap.is_cuda or soap.dtype != torch.float64:
        msg = "soap must be a float64 CUDA tensor"
        raise RuntimeError(msg)
    src = soap.contiguous()
    out = torch.empty_like(src)
    n_comp = src.shape[-1]
    _lib.check(
        lib.dsg_normalize_soap(
            src.data_ptr(), out.data_ptr(), src.numel() // max(n_comp, 1), n_comp, _stream_ptr()
        )
    )
    return out


def soap_power_spectrum(
    pos: torch.Tensor,
    species: torch.Tensor,
    centres: torch.Tensor,
    box: np.ndarray | torch.Tensor | None,
    r_cut: float,
    n_max: int,
    l_max: int,
    n_species: int | None = None,
    out: torch.Tensor | None = None,
    frame0: int = 0,
) -> torch.Tensor:
    """SOAP power spectrum of a batch of frames -> (n_centres, n_frames, n_feat) float64.

    ``pos`` (F, N, 3) CUDA float32/float64; ``species`` (N,) int32 in [0, n_species) (index into the
    ascending-atomic-number species list); ``centres`` (M,) int32 selection-local indices;
    ``box`` (F, 3) / (3,) orthorhombic lengths, or ``None`` for a non-periodic calculation.
    If ``out`` (M, F_total, C) is given, frames are written at ``out[:, frame0:frame0+F]``.
    """
    from dynsight_b200 import soap_basis

    lib = _lib.load()
    _require_cuda(pos, "pos")
    _require_cuda(species, "species")
    _require_cuda(centres, "centres")
    if species.dtype != torch.int32 or centres.dtype != torch.int32:
        msg = "species and centres must be int32"
        raise TypeError(msg)
    n_frames, n_atoms, _ = pos.shape
    n_cent = centres.numel()
    if n_species is None:
        n_species = int(species.max().item()) + 1
    alphas, betas = soap_basis.gto_basis(float(r_cut), int(n_max), int(l_max))
    n_feat = soap_basis.n_features(n_species, n_max, l_max)
    h_box = None if box is None else _host_box(box, n_frames)
    box_ptr = None if h_box is None else h_box.ctypes.data
    ws_bytes = ctypes.c_size_t(0)
    _lib.check(
        lib.dsg_soap_workspace_bytes(
            n_frames, n_atoms, n_cent, n_species, n_max, l_max, box_ptr, float(r_cut),
            ctypes.byref(ws_bytes),
        )
    )  # fmt: skip
    workspace = torch.empty(ws_bytes.value, dtype=torch.uint8, device=pos.device)
    if out is None:
        out = torch.empty((n_cent, n_frames, n_feat), dtype=torch.float64, device=pos.device)
        frame0 = 0
    view = out[:, frame0 : frame0 + n_frames]
    if view.shape != (n_cent, n_frames, n_feat) or view.stride(2) != 1:
        msg = f"out has shape {tuple(out.shape)}; expected (n_centres, >= frame0+F, {n_feat})"
        raise ValueError(msg)
    _lib.check(
        lib.dsg_soap(
            pos.data_ptr(), _dtype_code(pos), species.data_ptr(), centres.data_ptr(), n_frames,
            n_atoms, n_cent, n_species, box_ptr, float(r_cut), int(n_max), int(l_max),
            alphas.ctypes.data, betas.ctypes.data, workspace.data_ptr(), ws_bytes.value,
            view.data_ptr(), view.stride(0), view.stride(1), _stream_ptr(),
        )
    )  # fmt: skip
    return out


def soap_timesoap(
    pos: torch.Tensor,
    species: torch.Tensor,
    centres: torch.Tensor,
    box: np.ndarray | torch.Tensor | None,
    r_cut: float,
    n_max: int,
    l_max: int,
    delay: int = 1,
    n_species: int | None = None,
    out: torch.Tensor | None = None,
    col0: int = 0,
    soap_out: torch.Tensor | None = None,
    frame0: int = 0,
) -> torch.Tensor:
    """timeSOAP of a batch of frames straight from positions (``dsg_soap_timesoap``): the SOAP
    spectrum of every frame is folded into the distance to the spectrum ``delay`` frames earlier
    inside the spectrum epilogue; the spectra are not written unless ``soap_out`` is given.

    Returns / fills ``out[:, col0 : col0 + F - delay]`` (n_centres, >= col0 + F - delay) float64.
    Arguments as :func:`soap_power_spectrum`."""
    from dynsight_b200 import soap_basis

    lib = _lib.load()
    _require_cuda(pos, "pos")
    _require_cuda(species, "species")
    _require_cuda(centres, "centres")
    if species.dtype != torch.int32 or centres.dtype != torch.int32:
        msg = "species and centres must be int32"
        raise TypeError(msg)
    n_frames, n_atoms, _ = pos.shape
    n_cent = centres.numel()
    if delay < 1 or delay >= n_frames:
        msg = "delay value outside bounds"
        raise ValueError(msg)
    if n_species is None:
        n_species = int(species.max().item()) + 1
    alphas, betas = soap_basis.gto_basis(float(r_cut), int(n_max), int(l_max))
    n_feat = soap_basis.n_features(n_species, n_max, l_max)
    h_box = None if box is None else _host_box(box, n_frames)
    box_ptr = None if h_box is None else h_box.ctypes.data
    ws_bytes = ctypes.c_size_t(0)
    _lib.check(
        lib.dsg_soap_timesoap_workspace_bytes(
            n_frames, n_atoms, n_cent, n_species, n_max, l_max, box_ptr, float(r_cut), int(delay),
            ctypes.byref(ws_bytes),
        )
    )  # fmt: skip
    workspace = torch.empty(ws_bytes.value, dtype=torch.uint8, device=pos.device)
    if out is None:
        out = torch.empty((n_cent, n_frames - delay), dtype=torch.float64, device=pos.device)
        col0 = 0
    soap_ptr, s_i, s_f = None, 0, 0
    if soap_out is not None:
        view = soap_out[:, frame0 : frame0 + n_frames]
        if view.shape != (n_cent, n_frames, n_feat) or view.stride(2) != 1:
            msg = f"soap_out has shape {tuple(soap_out.shape)}; expected (n_centres, >= F, {n_feat})"
            raise ValueError(msg)
        soap_ptr, s_i, s_f = view.data_ptr(), view.stride(0), view.stride(1)
    _lib.check(
        lib.dsg_soap_timesoap(
            pos.data_ptr(), _dtype_code(pos), species.data_ptr(), centres.data_ptr(), n_frames,
            n_atoms, n_cent, n_species, box_ptr, float(r_cut), int(n_max), int(l_max),
            alphas.ctypes.data, betas.ctypes.data, int(delay), workspace.data_ptr(),
            ws_bytes.value, out.data_ptr(), out.stride(0), int(col0), soap_ptr, s_i, s_f,
            _stream_ptr(),
        )
    )  # fmt: skip
    return out


def _mean_row_len(batch: NeighborBatch) -> int:
    """Mean neighbours per (atom, frame) from sizes already known on the host (0 = unknown)."""
    cells = batch.n_frames * batch.n_cent
    return int(np.ceil(batch.indices.numel() / cells)) if cells > 0 else 0


def orientational_op_device(
    pos: torch.Tensor, batch: NeighborBatch, order: int = 6, out: torch.Tensor | None = None,
    col0: int = 0,
) -> torch.Tensor:  # fmt: skip
    """|psi_order| for every atom and frame of a batch -> (N, F) float64 (misc.py:15-95).
    ``batch`` must hold the lists of all atoms (centres == selection == all)."""
    lib = _lib.load()
    _require_cuda(pos, "pos")
    n_frames, n_atoms, _ = pos.shape
    if batch.n_frames != n_frames or batch.n_cent != n_atoms:
        msg = "neighbour lists and positions disagree in shape"
        raise ValueError(msg)
    if out is None:
        out = torch.empty((n_atoms, n_frames), dtype=torch.float64, device=pos.device)
        col0 = 0
    _lib.check(
        lib.dsg_orientational_op(
            pos.data_ptr(), _dtype_code(pos), batch.indptr.data_ptr(), batch.frame_off.data_ptr(),
            batch.indices.data_ptr(), n_frames, n_atoms, int(order), out.data_ptr(), out.stride(0),
            int(col0), _mean_row_len(batch), _stream_ptr(),
        )
    )  # fmt: skip
    return out


def velocity_alignment_device(
    vec: torch.Tensor, batch: NeighborBatch, displacement: bool, out: torch.Tensor | None = None,
    col0: int = 0,
) -> torch.Tensor:  # fmt: skip
    """phi (misc.py:98-219).  ``displacement=True``: ``vec`` = positions (F, N, 3), output (N, F-1)
    uses pos[f+1]-pos[0] with the rows of frame f; else ``vec`` = velocities (F, N, 3) -> (N, F)."""
    lib = _lib.load()
    _require_cuda(vec, "vec")
    n_f, n_atoms, _ = vec.shape
    n_out = n_f - 1 if displacement else n_f
    if n_out < 1 or batch.n_cent != n_atoms or batch.n_frames < n_out:
        msg = "neighbour lists and vectors disagree in shape"
        raise ValueError(msg)
    if out is None:
        out = torch.empty((n_atoms, n_out), dtype=torch.float64, device=vec.device)
        col0 = 0
    _lib.check(
        lib.dsg_velocity_alignment(
            vec.data_ptr(), _dtype_code(vec), int(bool(displacement)), batch.indptr.data_ptr(),
            batch.frame_off.data_ptr(), batch.indices.data_ptr(), n_out, n_atoms, out.data_ptr(),
            out.stride(0), int(col0), _mean_row_len(batch), _stream_ptr(),
        )
    )  # fmt: skip
    return out
