"""Trajectory ingestion: frames of a (duck-typed) Universe -> CUDA tensors, in batches.

This is the host-side plumbing between the reference's per-frame driver loops
(``universe.trajectory[t]; ag.positions.astype(np.float64)``, lens/lens.py:263-300 and
soapify/saponify.py:106-117) and the batched device operators in :mod:`dynsight_b200.engine`.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Iterator

import numpy as np
import torch

from dynsight_b200.universe import ArrayUniverse


@dataclass
class FrameBatch:
    """Positions of *all* atoms for a run of frames, resident on the device."""

    seq0: int  # position of the first frame inside the requested frame sequence
    frames: list[int]  # trajectory frame numbers
    pos_all: torch.Tensor  # (F, N_all, 3) float32 / float64, CUDA
    dims: np.ndarray | None  # (F, 6) float32 box [lx, ly, lz, alpha, beta, gamma] or None

    def select(self, ix: np.ndarray | Selection, n_all: int) -> torch.Tensor:
        """Positions of the atoms ``ix`` (universe indices) -> (F, n, 3) contiguous."""
        sel = ix if isinstance(ix, Selection) else Selection(ix, n_all)
        if sel.identity:
            return self.pos_all
        return self.pos_all.index_select(1, sel.device_index(self.pos_all.device)).contiguous()


class Selection:
    """Atom indices of a selection, analysed once per API call (not once per batch): whether it
    is the whole universe in order (then batches are used as they are) and, if not, the index
    tensor on the device."""

    def __init__(self, ix: np.ndarray, n_all: int) -> None:
        self.ix = np.asarray(ix)
        self.identity = bool(
            self.ix.size == n_all
            and (n_all == 0 or (self.ix[0] == 0 and self.ix[-1] == n_all - 1
                                and np.array_equal(self.ix, np.arange(n_all))))
        )  # fmt: skip
        self._dev_index: torch.Tensor | None = None

    def device_index(self, dev: torch.device) -> torch.Tensor:
        if self._dev_index is None or self._dev_index.device != dev:
            self._dev_index = torch.as_tensor(self.ix, dtype=torch.long, device=dev)
        return self._dev_index


def to_device_async(arr: np.ndarray) -> torch.Tensor:
    """Small host array -> device without blocking the host: a pageable ``torch.as_tensor(...,
    device=...)`` is a synchronous copy on the current stream, i.e. it waits for every kernel
    already enqueued there (the LENS kernels of the same call, for the SOAP job's index arrays).
    The page-locked staging buffer comes from torch's caching host allocator, which keeps it alive
    until the copy has run."""
    host = torch.from_numpy(np.ascontiguousarray(arr))
    try:
        host = host.pin_memory()
    except RuntimeError:
        return host.to(device())
    return host.to(device(), non_blocking=True)


def device() -> torch.device:
    if not torch.cuda.is_available():
        msg = (
            "dynsight_b200 needs a CUDA device (B200, sm_100a): there is no CPU fallback for the "
            "descriptor engine"
        )
        raise RuntimeError(msg)
    return torch.device("cuda", torch.cuda.current_device())


def frame_sequence(n_frames: int, trajslice: slice | None) -> list[int]:
    """``list(range(n_frames))[trajslice]`` (lens/lens.py:251-254)."""
    idx = list(range(n_frames))
    return idx if trajslice is None else idx[trajslice]


_COPY_STREAMS: dict[int, torch.cuda.Stream] = {}


def _copy_stream(dev: torch.device) -> torch.cuda.Stream:
    """One side stream per device for host -> device copies."""
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    if key not in _COPY_STREAMS:
        _COPY_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _COPY_STREAMS[key]


def iter_batches(universe: Any, frames: list[int], batch_frames: int, overlap: int = 0
                 ) -> Iterator[FrameBatch]:  # fmt: skip
    """Yield device batches covering ``frames``; consecutive batches share ``overlap`` frames.

    Double-buffered: two device buffers are reused alternately, and the host -> device copy of a
    batch runs on a side stream.  The copy is issued when the consumer asks for the next batch,
    i.e. after the kernels of the previous batch were enqueued on the compute stream, so it
    overlaps that compute instead of queueing behind it; the compute stream waits on the copy's
    event before the batch is handed out, and a buffer is overwritten only after the kernels
    that read it (recorded when the consumer came back for more) have finished.  A consumer may
    therefore keep using a batch until it has asked for the batch after the next one.
    """
    dev = device()
    n = len(frames)
    batch_frames = max(int(batch_frames), overlap + 1)
    side = _copy_stream(dev)
    compute = torch.cuda.current_stream(dev)
    bufs: list[torch.Tensor | None] = [None, None]
    released: list[torch.cuda.Event | None] = [None, None]
    start = 0
    k = 0
    while True:
        stop = min(n, start + batch_frames)
        chunk = frames[start:stop]
        host, dims = _load_host(universe, chunk)
        slot = k & 1
        if bufs[slot] is None or bufs[slot].dtype != host.dtype or bufs[slot].shape[0] < len(chunk):
            bufs[slot] = torch.empty((min(batch_frames, n), *host.shape[1:]), dtype=host.dtype,
                                     device=dev)  # fmt: skip
            released[slot] = torch.cuda.Event()
            released[slot].record(compute)  # the allocator may hand out memory still in use there
        dst = bufs[slot][: len(chunk)]
        with torch.cuda.stream(side):
            side.wait_event(released[slot])
            dst.copy_(host, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(side)
        compute.wait_event(ready)
        yield FrameBatch(start, chunk, dst, dims)
        released[slot] = torch.cuda.Event()
        released[slot].record(compute)
        if stop >= n:
            return
        start = stop - overlap
        k += 1


def _load_host(universe: Any, chunk: list[int]) -> tuple[torch.Tensor, np.ndarray | None]:
    """Positions of all atoms for the frames ``chunk`` as one host tensor (+ per-frame boxes)."""
    if isinstance(universe, ArrayUniverse):
        reader = universe.trajectory
        coords = reader.coordinates
        contiguous_run = len(chunk) > 0 and chunk == list(range(chunk[0], chunk[0] + len(chunk)))
        host = coords[chunk[0] : chunk[0] + len(chunk)] if contiguous_run else coords[chunk]
        dims = None if reader.dimensions_array is None else reader.dimensions_array[chunk]
        return torch.from_numpy(np.ascontiguousarray(host)), dims
    # generic duck-typed Universe (e.g. MDAnalysis): seek + copy frame by frame into pinned memory
    n_all = len(universe.atoms)
    first = np.asarray(universe.atoms.positions)
    dtype = torch.float64 if first.dtype == np.float64 else torch.float32
    stage = torch.empty((len(chunk), n_all, 3), dtype=dtype, pin_memory=True)
    stage_np = stage.numpy()
    dims_list: list[np.ndarray | None] = []
    for k, fr in enumerate(chunk):
        universe.trajectory[fr]
        stage_np[k] = universe.atoms.positions
        d = universe.trajectory.ts.dimensions
        dims_list.append(None if d is None else np.asarray(d, dtype=np.float32))
    if all(d is None for d in dims_list):
        dims = None
    elif any(d is None for d in dims_list):
        msg = "trajectory mixes frames with and without a box"
        raise ValueError(msg)
    else:
        dims = np.stack(dims_list)  # type: ignore[arg-type]
    return stage, dims


def lens_boxes(batch: FrameBatch, scale: float | None) -> np.ndarray:
    """Per-frame box (F, 3) float64 of the neighbour-list path.

    ``ts.dimensions[:3]`` when present, else the bounding box of *all* atoms, multiplied by 1.01
    in ``compute_lens`` (lens/lens.py:268-274) but not in ``list_neighbours_along_trajectory``
    (lens/lens.py:365-371).  float32 arithmetic like numpy's, then widened exactly to float64.
    """
    if batch.dims is not None:
        return np.ascontiguousarray(batch.dims[:, :3].astype(np.float64))
    pos = batch.pos_all
    box = pos.amax(dim=1) - pos.amin(dim=1)
    if scale is not None:
        box = box * scale
    return np.ascontiguousarray(box.cpu().numpy().astype(np.float64))


def soap_boxes(batch: FrameBatch, periodic: bool) -> np.ndarray | None:
    """Orthorhombic cell lengths (F, 3) for SOAP or ``None`` when the calculation is not periodic."""
    if not periodic:
        return None
    if batch.dims is None:
        msg = "soap_respectpbc=True needs a simulation box (ts.dimensions is None)"
        raise ValueError(msg)
    if not np.allclose(batch.dims[:, 3:], 90.0, atol=1e-4):
        msg = "dynsight_b200 SOAP supports orthorhombic boxes only"
        raise NotImplementedError(msg)
    return np.ascontiguousarray(batch.dims[:, :3].astype(np.float64))


def pick_batch_frames(n_atoms: int, bytes_per_atom_frame: float, minimum: int,
                      budget_bytes: float = 4e9, maximum: int = 4096) -> int:  # fmt: skip
    b = int(budget_bytes / max(1.0, n_atoms * bytes_per_atom_frame))
    b = min(b, maximum, max(1, (2**31 - 1) // max(1, n_atoms) - 1))
    return max(b, minimum)


_PINNED_RESULT_LIMIT = 2 << 30  # bytes; larger results go through a pageable copy


def to_host(t: torch.Tensor) -> np.ndarray:
    """Device result -> numpy array.  Results up to 2 GiB land in page-locked memory from torch's
    caching host allocator (DMA at link speed, no first-touch page faults on a fresh numpy buffer);
    the array keeps the buffer alive and returns it to the cache when it is dropped."""
    if t.device.type != "cuda":
        return t.numpy()
    if 0 < t.numel() * t.element_size() <= _PINNED_RESULT_LIMIT:
        try:
            host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        except RuntimeError:
            return t.cpu().numpy()
        host.copy_(t, non_blocking=True)
        torch.cuda.current_stream(t.device).synchronize()
        return host.numpy()
    return t.cpu().numpy()


class PendingHostCopy:
    """A device -> host copy in flight on the side stream (:func:`to_host_async`)."""

    def __init__(self, host: torch.Tensor | None, done: torch.cuda.Event | None,
                 fallback: torch.Tensor | None = None) -> None:  # fmt: skip
        self._host, self._done, self._fallback = host, done, fallback

    def result(self) -> np.ndarray:
        if self._host is None:
            return to_host(self._fallback)
        self._done.synchronize()
        return self._host.numpy()


def to_host_async(t: torch.Tensor) -> PendingHostCopy:
    """Start copying a device result to page-locked host memory on the copy stream, behind the
    work already enqueued on the current stream; ``result()`` waits for the copy only."""
    n_bytes = t.numel() * t.element_size()
    if t.device.type != "cuda" or not 0 < n_bytes <= _PINNED_RESULT_LIMIT:
        return PendingHostCopy(None, None, t)
    try:
        host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    except RuntimeError:
        return PendingHostCopy(None, None, t)
    side = _copy_stream(t.device)
    ready = torch.cuda.Event()
    ready.record(torch.cuda.current_stream(t.device))
    with torch.cuda.stream(side):
        side.wait_event(ready)
        host.copy_(t, non_blocking=True)
        done = torch.cuda.Event()
        done.record(side)
    t.record_stream(side)
    return PendingHostCopy(host, done)
