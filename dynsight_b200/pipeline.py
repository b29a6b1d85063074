"""One pass over a trajectory for several descriptors of the hot path.

The reference API computes every descriptor in its own pass (``Trj.get_lens``,
``Trj.get_coord_number``, ``Trj.get_timesoap`` each walk the trajectory: trajectory.py:139-320).
On the GPU the pass itself -- decoding / copying the frames to the device -- is a visible share of
the end-to-end time, so this module feeds every device batch to all requested jobs: positions
cross the host -> device link once.  The results are exactly those of the separate calls
(``lens.compute_lens``, ``Trj.get_coord_number``, ``soap.timesoap_from_trajectory``): the jobs
are the same objects those functions run.
"""

from __future__ import annotations

import itertools
from typing import Any

import numpy as np
import torch

from dynsight_b200 import frames
from dynsight_b200.lens import LensJob
from dynsight_b200.soap import TimesoapJob

__all__ = ["descriptors_from_trajectory"]

_EARLY_UPLOAD_FRAMES = 64


def descriptors_from_trajectory(
    universe: Any,
    lens_r_cut: float | None = None,
    soap_r_cut: float | None = None,
    soapnmax: int = 8,
    soaplmax: int = 8,
    delay: int = 1,
    centers: str = "all",
    selection: str = "all",
    trajslice: slice | None = None,
    respect_pbc: bool = True,
    coord_number: bool = False,
    batch_frames: int | None = None,
    as_numpy: bool = True,
) -> dict[str, np.ndarray | torch.Tensor]:
    """LENS, coordination numbers and timeSOAP of a trajectory in ONE pass over its frames.

    ``lens_r_cut`` switches on ``"lens"`` (n_centers, n_frames - delay) and, with
    ``coord_number=True``, ``"coord_number"`` (n_centers, n_frames) -- the values of
    ``compute_lens`` / ``Trj.get_coord_number`` except that the box-less fallback of the
    coordination numbers uses the LENS box rule (the 1.01 factor of lens.py:271-274);
    ``soap_r_cut`` switches on ``"timesoap"`` (n_centers, n_frames - delay), the values of
    ``Trj.get_timesoap`` computed without materialising the SOAP tensor.
    """
    # A short trajectory is one device batch: its host -> device copy is issued before the jobs are
    # set up (selections, species, output buffers: ~1.5 ms of host work at 1e5 atoms, during which
    # the GPU would otherwise idle); a long one is amortised over its batches anyway.
    early = None
    if batch_frames is None and (lens_r_cut is not None or soap_r_cut is not None):
        seq = frames.frame_sequence(universe.trajectory.n_frames, trajslice)
        if delay < len(seq) <= _EARLY_UPLOAD_FRAMES:
            early_it = frames.iter_batches(universe, seq, len(seq), overlap=delay)
            early = (len(seq), next(early_it), early_it)
    jobs: dict[str, Any] = {}
    early_out: dict[str, frames.PendingHostCopy] = {}
    lens_done = False
    if lens_r_cut is not None:
        jobs["lens"] = job = LensJob(universe, lens_r_cut, delay, centers, selection, trajslice,
                                     respect_pbc, want_counts=coord_number)  # fmt: skip
        if early is not None and soap_r_cut is not None and job.default_batch_frames() >= early[0]:
            # the whole trajectory is on its way: enqueue the LENS kernels now, they run (and their
            # results travel to the host) while the SOAP job is still being set up
            job.process(early[1])
            lens_done = True
            if as_numpy:
                early_out["lens"] = frames.to_host_async(job.finish())
                if coord_number:
                    early_out["coord_number"] = frames.to_host_async(job.counts_out)
    if soap_r_cut is not None:
        jobs["timesoap"] = TimesoapJob(universe, soap_r_cut, soapnmax, soaplmax, delay, selection,
                                       respect_pbc, centers, trajslice)  # fmt: skip
    if not jobs:
        msg = "nothing to compute: give lens_r_cut and/or soap_r_cut"
        raise ValueError(msg)
    fr_idx = next(iter(jobs.values())).fr_idx
    if batch_frames is None:
        batch_frames = min(job.default_batch_frames() for job in jobs.values())
    batch_frames = max(int(batch_frames), delay + 1)
    if early is not None and batch_frames >= early[0] and len(fr_idx) == early[0]:
        batches = itertools.chain([early[1]], early[2])   # the copy is already in flight
    else:
        batches = frames.iter_batches(universe, fr_idx, batch_frames, overlap=delay)
    for fb in batches:
        last = fb.seq0 + len(fb.frames) >= len(fr_idx)
        for name, job in jobs.items():
            if name == "lens" and lens_done:
                continue
            job.process(fb)
            if last and as_numpy and name == "lens" and "timesoap" in jobs:
                # the LENS kernels of the last batch are enqueued: their results travel to the
                # host on the copy stream while the SOAP kernels of that batch still run
                early_out["lens"] = frames.to_host_async(job.finish())
                if coord_number:
                    early_out["coord_number"] = frames.to_host_async(job.counts_out)
    res: dict[str, torch.Tensor] = {name: job.finish() for name, job in jobs.items()}
    if coord_number and "lens" in jobs:
        res["coord_number"] = jobs["lens"].counts_out
    if not as_numpy:
        return res
    return {name: early_out[name].result() if name in early_out else frames.to_host(t)
            for name, t in res.items()}  # fmt: skip
