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

    def select(self, ix: np.ndarray, n_all: int) -> torch.Tensor:
        """Positions of the atoms ``ix`` (universe indices) -> (F, n, 3) contiguous."""
        if ix.size == n_all and np.array_equal(ix, np.arange(n_all)):
            return self.pos_all
        idx = torch.as_tensor(ix, dtype=torch.long, device=self.pos_all.device)
        return self.pos_all.index_select(1, idx).contiguous()


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


def iter_batches(universe: Any, frames: list[int], batch_frames: int, overlap: int = 0
                 ) -> Iterator[FrameBatch]:  # fmt: skip
    """Yield device batches covering ``frames``; consecutive batches share ``overlap`` frames."""
    dev = device()
    n = len(frames)
    batch_frames = max(int(batch_frames), overlap + 1)
    start = 0
    while True:
        stop = min(n, start + batch_frames)
        chunk = frames[start:stop]
        yield _load(universe, chunk, start, dev)
        if stop >= n:
            return
        start = stop - overlap


def _load(universe: Any, chunk: list[int], seq0: int, dev: torch.device) -> FrameBatch:
    if isinstance(universe, ArrayUniverse):
        reader = universe.trajectory
        coords = reader.coordinates
        contiguous_run = len(chunk) > 0 and chunk == list(range(chunk[0], chunk[0] + len(chunk)))
        host = coords[chunk[0] : chunk[0] + len(chunk)] if contiguous_run else coords[chunk]
        dims = None if reader.dimensions_array is None else reader.dimensions_array[chunk]
        t = torch.from_numpy(np.ascontiguousarray(host))
        return FrameBatch(seq0, chunk, t.to(dev, non_blocking=True), dims)
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
    return FrameBatch(seq0, chunk, stage.to(dev, non_blocking=True), dims)


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
