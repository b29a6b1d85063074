"""Drop-in for the pieces of ``dynsight.analysis`` that sit right after the descriptor hot path.

* :func:`spatialaverage` -- ``dynsight.analysis.spatialaverage``
  (``_internal/analysis/spatial_average.py:66-211``): neighbour search + mean over each particle's
  neighbourhood, on the device.
* :func:`self_time_correlation` -- ``dynsight.analysis.self_time_correlation``
  (``_internal/analysis/time_correlations.py:13-92``), SURVEY.md section 8f row 4.
"""

from __future__ import annotations

from typing import TYPE_CHECKING, Any

import numpy as np
import torch

from dynsight_b200 import engine, frames

if TYPE_CHECKING:
    from numpy.typing import NDArray

_SCALAR_NDIM = 2
_VECTOR_NDIM = 3


def spatialaverage_device(
    universe: Any,
    descriptor_array: NDArray[np.float64] | torch.Tensor,
    selection: str,
    r_cut: float,
    trajslice: slice | None = None,
    batch_frames: int | None = None,
) -> torch.Tensor:
    """:func:`spatialaverage` returning a CUDA tensor (no device -> host copy)."""
    sel = universe.select_atoms(selection)
    sel_ix = np.asarray(sel.indices)
    if descriptor_array.ndim not in (_SCALAR_NDIM, _VECTOR_NDIM):
        msg = "descriptor_array must have ndim == 2 or ndim == 3."
        raise ValueError(msg)
    fr_idx = frames.frame_sequence(universe.trajectory.n_frames, trajslice)
    dev = frames.device()
    values = torch.as_tensor(descriptor_array).to(device=dev, dtype=torch.float64).contiguous()
    out = torch.zeros_like(values)  # frames beyond the slice stay zero (np.zeros in the reference)
    n_all = len(universe.atoms)
    if values.shape[0] != sel_ix.size or values.shape[1] < len(fr_idx):
        msg = (
            f"descriptor_array has shape {tuple(values.shape)}; the selection has {sel_ix.size} "
            f"atoms and the slice {len(fr_idx)} frames"
        )
        raise IndexError(msg)
    if batch_frames is None:
        batch_frames = frames.pick_batch_frames(sel_ix.size, 400.0, 1)
    sel_all = frames.Selection(sel_ix, n_all)
    for fb in frames.iter_batches(universe, fr_idx, batch_frames):
        if fb.dims is None:
            msg = "spatialaverage needs a simulation box (ts.dimensions is None)"
            raise ValueError(msg)
        if not np.allclose(fb.dims[:, 3:], 90.0, atol=1e-4):
            msg = "dynsight_b200 spatialaverage supports orthorhombic boxes only"
            raise NotImplementedError(msg)
        box = np.ascontiguousarray(fb.dims[:, :3].astype(np.float64))
        pos = engine.wrap_positions(fb.select(sel_all, n_all), box)  # FastNS bins wrapped positions
        capacity = None
        force_warp = False
        for attempt in range(engine._MAX_ROW_RETRIES):  # noqa: SLF001
            nr = engine.neighbor_rows(pos, box, float(r_cut), None, True, capacity, force_warp,
                                      inclusive_cutoff=True)  # fmt: skip
            engine.spatial_average_rows(nr, values, out, frame0=fb.seq0)
            if not nr.overflowed():
                break
            if nr.needs_warp_kernel():
                force_warp = True
            else:
                capacity = engine._next_capacity(nr, capacity)  # noqa: SLF001
        else:
            msg = "neighbour rows still overflow after several attempts"
            raise RuntimeError(msg)
        del attempt
    return out


def spatialaverage(
    universe: Any,
    descriptor_array: NDArray[np.float64],
    selection: str,
    r_cut: float,
    trajslice: slice | None = None,
    n_jobs: int = 1,  # noqa: ARG001  (kept for signature parity; the GPU path ignores it)
) -> NDArray[np.float64]:
    """Compute spatially averaged descriptor values over neighboring particles.

    For every particle of ``selection`` and every frame of ``trajslice`` the descriptor is averaged
    over the particle itself and the particles within ``r_cut`` (periodic minimum image, pairs at
    exactly ``r_cut`` included like MDAnalysis' FastNS).  ``descriptor_array`` has shape
    ``(n_atoms, n_frames)`` or ``(n_atoms, n_frames, n_features)``; the result has the same shape.

    Raises:
        ValueError: if ``descriptor_array`` has neither 2 nor 3 dimensions.
    """
    return frames.to_host(
        spatialaverage_device(universe, descriptor_array, selection, r_cut, trajslice)
    )


def self_time_correlation(
    data: NDArray[np.float64] | torch.Tensor,
    max_delay: int | None = None,
) -> tuple[NDArray[np.float64], NDArray[np.float64]]:
    """Computes the mean self time correlation function for time-series.

    ``data`` has shape (n_particles, n_frames).  Returns the self-TCF averaged over the particles
    for the delays ``0 .. max_delay - 1`` (all delays if ``None``), normalised to 1 at delay 0,
    and its standard error (time_correlations.py:13-92).

    Raises:
        ValueError: if ``max_delay`` is outside ``1 .. n_frames`` (the reference silently divides
            by zero there).
    """
    dev = frames.device()
    values = torch.as_tensor(data).to(device=dev, dtype=torch.float64).contiguous()
    n_part, n_frames = values.shape
    if max_delay is None:
        max_delay = n_frames
    if max_delay < 1 or max_delay > n_frames:
        msg = f"max_delay={max_delay} must be in 1..{n_frames}"
        raise ValueError(msg)
    mean, std = engine.self_time_correlation_device(values, int(max_delay))
    valid = n_frames - np.arange(max_delay, dtype=np.float64)
    correlation = frames.to_host(mean) / valid
    correlation_error = frames.to_host(std) / (valid * np.sqrt(n_part))
    norm_fact = correlation[0]
    return correlation / norm_fact, correlation_error / norm_fact


def cross_time_correlation(
    data: NDArray[np.float64] | torch.Tensor,
    max_delay: int | None = None,
) -> tuple[NDArray[np.float64], NDArray[np.float64]]:
    """Computes the mean cross time correlation function for time-series.

    For each ordered pair of particles ``i != j`` the cross-TCF of the mean-subtracted series is
    computed; returns its average over the pairs and the standard error for the delays
    ``0 .. max_delay - 1`` (time_correlations.py:95-178; not normalised at delay 0).

    Raises:
        ValueError: if ``max_delay`` is outside ``1 .. n_frames`` or there are fewer than two
            particles (the reference returns NaN there).
    """
    dev = frames.device()
    values = torch.as_tensor(data).to(device=dev, dtype=torch.float64).contiguous()
    n_part, n_frames = values.shape
    if max_delay is None:
        max_delay = n_frames
    if max_delay < 1 or max_delay > n_frames:
        msg = f"max_delay={max_delay} must be in 1..{n_frames}"
        raise ValueError(msg)
    if n_part < 2:  # noqa: PLR2004
        msg = "cross_time_correlation needs at least two particles"
        raise ValueError(msg)
    mean, std = engine.self_time_correlation_device(values, int(max_delay), cross=True)
    valid = n_frames - np.arange(max_delay, dtype=np.float64)
    cross_correlation = frames.to_host(mean) / valid
    correlation_error = frames.to_host(std) / (valid * np.sqrt(n_part * (n_part - 1)))
    return cross_correlation, correlation_error
