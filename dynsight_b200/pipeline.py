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

from typing import Any

import numpy as np
import torch

from dynsight_b200 import frames
from dynsight_b200.lens import LensJob
from dynsight_b200.soap import TimesoapJob

__all__ = ["descriptors_from_trajectory"]


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
    jobs: dict[str, Any] = {}
    if lens_r_cut is not None:
        jobs["lens"] = LensJob(universe, lens_r_cut, delay, centers, selection, trajslice,
                               respect_pbc, want_counts=coord_number)  # fmt: skip
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
    early: dict[str, frames.PendingHostCopy] = {}
    for fb in frames.iter_batches(universe, fr_idx, batch_frames, overlap=delay):
        last = fb.seq0 + len(fb.frames) >= len(fr_idx)
        for name, job in jobs.items():
            job.process(fb)
            if last and as_numpy and name == "lens" and "timesoap" in jobs:
                # the LENS kernels of the last batch are enqueued: their results travel to the
                # host on the copy stream while the SOAP kernels of that batch still run
                early["lens"] = frames.to_host_async(job.finish())
                if coord_number:
                    early["coord_number"] = frames.to_host_async(job.counts_out)
    res: dict[str, torch.Tensor] = {name: job.finish() for name, job in jobs.items()}
    if coord_number and "lens" in jobs:
        res["coord_number"] = jobs["lens"].counts_out
    if not as_numpy:
        return res
    return {name: early[name].result() if name in early else frames.to_host(t)
            for name, t in res.items()}  # fmt: skip
