"""LENS package -- drop-in for ``dynsight.lens`` (reference ``src/dynsight/lens.py:3-11``).

Same names, arguments, defaults, return shapes and error messages as
``dynsight._internal.lens.lens.compute_lens`` (lens.py:192-310) and
``list_neighbours_along_trajectory`` (lens.py:313-390); the work runs on the GPU through
:mod:`dynsight_b200.engine` (batched cell list -> sorted CSR -> warp merge).  ``n_jobs`` is
accepted and ignored (the reference uses it for ``numba.set_num_threads``).
"""

from __future__ import annotations

from typing import TYPE_CHECKING, Any, Iterator, Sequence

import numpy as np
import torch

from dynsight_b200 import engine, frames

if TYPE_CHECKING:
    from numpy.typing import NDArray

__all__ = ["LensJob", "compute_lens", "compute_lens_device", "list_neighbours_along_trajectory"]


def _mean_neighbours_estimate(n_env: int, box: np.ndarray, r_cut: float) -> float:
    vol = float(np.prod(np.maximum(box[0], 1e-12)))
    return min(float(n_env), n_env / vol * 4.18879 * float(r_cut) ** 3)


class LensJob:
    """LENS (and, on request, coordination numbers) of a trajectory, fed one device batch at a
    time: ``job.process(fb)`` for every batch of ``frames.iter_batches(..., overlap=delay)``, then
    ``job.finish()``.  ``compute_lens_device`` runs it alone; ``pipeline.descriptors_from_trajectory``
    feeds the same batches to the SOAP / timeSOAP job as well (one ingest for all descriptors)."""

    def __init__(self, universe: Any, r_cut: float, delay: int = 1, centers: str = "all",
                 selection: str = "all", trajslice: slice | None = None, respect_pbc: bool = True,
                 want_counts: bool = False) -> None:  # fmt: skip
        ag_env = universe.select_atoms(selection)
        ag_cent = universe.select_atoms(centers)
        self.universe = universe
        self.r_cut, self.delay, self.respect_pbc = r_cut, delay, respect_pbc
        self.fr_idx = frames.frame_sequence(universe.trajectory.n_frames, trajslice)
        n_pairs = len(self.fr_idx) - delay
        if delay < 0 or n_pairs < 1:  # "pairs" would be empty (lens.py:255-260)
            msg = "No valid pairs found."
            raise RuntimeError(msg)
        self.n_all = len(universe.atoms)
        self.ix_env = np.asarray(ag_env.indices)
        self.ix_cent = np.asarray(ag_cent.indices)
        self.same = self.ix_env.shape == self.ix_cent.shape and bool(
            np.array_equal(self.ix_env, self.ix_cent)
        )
        dev = frames.device()
        self.out = torch.empty((self.ix_cent.size, n_pairs), dtype=torch.float64, device=dev)
        self.counts_out = (
            torch.zeros((self.ix_cent.size, len(self.fr_idx)), dtype=torch.float64, device=dev)
            if want_counts else None
        )  # fmt: skip
        # no centres: empty result; no environment atoms: every row is empty and
        # lens_from_two_csr returns 0.0 for a + b == 0 (lens.py:171-173); delay = 0 pairs every
        # frame with itself (lens.py:255-257): identical lists, LENS = 0.0 everywhere
        self.trivial = self.ix_cent.size == 0 or self.ix_env.size == 0 or delay == 0
        if self.trivial:
            self.out.zero_()
        self.sel_env = frames.Selection(self.ix_env, self.n_all)
        self.sel_cent = frames.Selection(self.ix_cent, self.n_all)
        self._pending: tuple | None = None

    @property
    def needs_batches(self) -> bool:
        return not self.trivial or (self.counts_out is not None and self.ix_cent.size > 0
                                    and self.ix_env.size > 0)  # fmt: skip

    def default_batch_frames(self) -> int:
        """Frames per batch from a 4 GB budget.  Rows path: 1.3 k int32 row entries + counts / row
        starts / cell-sorted copies per atom-frame (k from the density of the first frame); the
        pair path needs 44 B."""
        host, dims = frames._load_host(self.universe, self.fr_idx[:1])  # noqa: SLF001
        box0 = (dims[:, :3].astype(np.float64) if dims is not None else
                (host.amax(dim=1) - host.amin(dim=1)).numpy().astype(np.float64) * 1.01)  # fmt: skip
        k_est = _mean_neighbours_estimate(max(1, self.ix_env.size), box0, self.r_cut)
        return frames.pick_batch_frames(max(self.ix_env.size, self.ix_cent.size, 1),
                                        60.0 + 5.2 * k_est, self.delay + 1)  # fmt: skip

    def _settle(self) -> None:
        """The one host sync per rows-path batch: a reservation overflow (rare) reruns it."""
        job, self._pending = self._pending, None
        if job is None:
            return
        fb, box, pos_env, pos_cent, nr = job
        if nr.overflowed():
            warp = nr.needs_warp_kernel()
            _, nr2 = engine.lens_from_positions(
                pos_env, box, self.r_cut, max(self.delay, 1), pos_cent, self.respect_pbc,
                out=self.out if not self.trivial else None, col0=fb.seq0,
                capacity=None if warp else engine._next_capacity(nr, None),  # noqa: SLF001
                force_warp=warp,
            )
            if self.counts_out is not None:
                engine.coordination_numbers(nr2.counts, out=self.counts_out, col0=fb.seq0)

    def process(self, fb: frames.FrameBatch) -> None:
        """Enqueue the kernels of one batch.  The status of a rows-path batch is read only after
        the kernels of the next one are enqueued (software pipeline, one sync per batch)."""
        if not self.needs_batches:
            return
        n_fb = len(fb.frames)
        with_lens = not self.trivial and n_fb > self.delay
        box = frames.lens_boxes(fb, 1.01)
        pos_env = fb.select(self.sel_env, self.n_all)
        if self.same and engine.lens_direct_supported(n_fb, pos_env.shape[1], box, self.r_cut,
                                                      self.respect_pbc):  # fmt: skip
            # centres == environment, sparse cells: pair path, no lists and nothing to retry
            _, counts = engine.lens_direct(
                pos_env, box, self.r_cut, self.delay, self.respect_pbc, out=self.out,
                col0=fb.seq0, want_lens=with_lens, want_counts=self.counts_out is not None,
            )
            if counts is not None:
                engine.coordination_numbers(counts, out=self.counts_out, col0=fb.seq0)
            self._settle()
            return
        pos_cent = None if self.same else fb.select(self.sel_cent, self.n_all)
        nr = engine.neighbor_rows(pos_env, box, self.r_cut, pos_cent, self.respect_pbc)
        if with_lens:
            engine.lens_pairs_rows(nr, self.delay, out=self.out, col0=fb.seq0)
        if self.counts_out is not None:
            engine.coordination_numbers(nr.counts, out=self.counts_out, col0=fb.seq0)
        self._settle()
        self._pending = (fb, box, pos_env, pos_cent, nr)

    def finish(self) -> torch.Tensor:
        self._settle()
        return self.out


def compute_lens_device(
    universe: Any,
    r_cut: float,
    delay: int = 1,
    centers: str = "all",
    selection: str = "all",
    trajslice: slice | None = None,
    respect_pbc: bool = True,
    batch_frames: int | None = None,
) -> torch.Tensor:
    """:func:`compute_lens` that leaves the (n_centers, n_pairs) float64 result on the device."""
    job = LensJob(universe, r_cut, delay, centers, selection, trajslice, respect_pbc)
    if not job.needs_batches:
        return job.out
    if batch_frames is None:
        batch_frames = job.default_batch_frames()
    batch_frames = max(batch_frames, delay + 1)
    for fb in frames.iter_batches(universe, job.fr_idx, batch_frames, overlap=delay):
        job.process(fb)
    return job.finish()


def compute_lens(
    universe: Any,
    r_cut: float,
    delay: int = 1,
    centers: str = "all",
    selection: str = "all",
    trajslice: slice | None = None,
    respect_pbc: bool = True,
    n_jobs: int = 1,  # noqa: ARG001  (numba thread count in the reference)
) -> NDArray[np.float64]:
    """Compute the LENS descriptor for all frames along a trajectory.

    Returns LENS values for each centre and each pair of frames, shape (n_centers, n_pairs),
    float64 -- identical (bit for bit) to the reference's result.

    ``delay=0`` pairs every frame with itself like the reference does and returns zeros of shape
    (n_centers, n_frames).

    Raises:
        RuntimeError: "No valid pairs found." when the frame selection is shorter than ``delay``.
    """
    return frames.to_host(
        compute_lens_device(universe, r_cut, delay, centers, selection, trajslice, respect_pbc)
    )


class FrameNeighbours(Sequence):
    """Neighbours of every centre in one frame; item ``i`` is ``ag_env[indices[s:e]]``."""

    def __init__(self, ag_env: Any, indptr: np.ndarray, indices: np.ndarray) -> None:
        self._ag_env = ag_env
        self.indptr = indptr
        self.indices = indices

    def __len__(self) -> int:
        return len(self.indptr) - 1

    def __getitem__(self, i: int) -> Any:  # type: ignore[override]
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        return self._ag_env[self.indices[self.indptr[i] : self.indptr[i + 1]]]

    def __iter__(self) -> Iterator[Any]:
        for i in range(len(self)):
            yield self[i]


class NeighbourList(Sequence):
    """``list[list[AtomGroup]]`` of the reference, materialised lazily -- twice over.

    The reference builds one AtomGroup per centre and frame in a Python loop (lens.py:383-388),
    which dominates its run time.  Here only the coordination numbers ``counts_device``
    (n_frames, n_centers) are computed up front (single-traversal rows path, nothing sorted, nothing
    copied to the host): that is all ``Trj.get_coord_number`` needs.  The canonical CSR (rows
    ascending, numba-identical ``indptr``) is built on the first access to ``indptr`` /
    ``frame_off`` / ``indices`` or to any element, and an AtomGroup is only created when an element
    is accessed.
    """

    def __init__(self, ag_env: Any, counts: torch.Tensor, build_csr: Any,
                 env_indices: np.ndarray | None = None, centre_indices: np.ndarray | None = None,
                 ) -> None:  # fmt: skip
        self._ag_env = ag_env
        # universe indices of the environment / centre atoms: the CSR ``indices`` are positions
        # inside ``ag_env`` (what the numba kernel returns), consumers that index the universe's
        # arrays (psi / phi kernels) translate through these
        self.env_indices = env_indices
        self.centre_indices = centre_indices
        self.counts_device = counts  # (F, M) int32 CUDA
        self._build_csr = build_csr  # () -> engine.NeighborBatch covering all frames (device)
        self._batch: engine.NeighborBatch | None = None
        self._csr: tuple[np.ndarray, np.ndarray, np.ndarray] | None = None

    def device_batch(self) -> engine.NeighborBatch:
        """Canonical CSR of all frames on the device (built once; what the psi / phi kernels read)."""
        if self._batch is None:
            self._batch = self._build_csr()
            self._build_csr = None
        return self._batch

    def _ensure(self) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
        if self._csr is None:
            nb = self.device_batch()
            self._csr = (nb.indptr.cpu().numpy(), nb.frame_off.cpu().numpy(),
                         nb.indices.cpu().numpy())  # fmt: skip
        return self._csr

    @property
    def indptr(self) -> np.ndarray:
        return self._ensure()[0]

    @property
    def frame_off(self) -> np.ndarray:
        return self._ensure()[1]

    @property
    def indices(self) -> np.ndarray:
        return self._ensure()[2]

    def __len__(self) -> int:
        return int(self.counts_device.shape[0])

    def __getitem__(self, f: int) -> FrameNeighbours:  # type: ignore[override]
        if isinstance(f, slice):
            return [self[k] for k in range(*f.indices(len(self)))]  # type: ignore[return-value]
        if f < 0:
            f += len(self)
        if not 0 <= f < len(self):
            raise IndexError(f)
        return FrameNeighbours(
            self._ag_env, self.indptr[f], self.indices[self.frame_off[f] : self.frame_off[f + 1]]
        )

    def __iter__(self) -> Iterator[FrameNeighbours]:
        for f in range(len(self)):
            yield self[f]


def list_neighbours_along_trajectory(
    universe: Any,
    r_cut: float,
    centers: str = "all",
    selection: str = "all",
    trajslice: slice | None = None,
    respect_pbc: bool = True,
    n_jobs: int = 1,  # noqa: ARG001
) -> NeighbourList:
    """Produce a per-frame list of neighbours: ``result[frame][centre]`` is an AtomGroup of the
    environment atoms within ``r_cut`` (same membership and ascending order as the reference)."""
    if trajslice is None:
        trajslice = slice(None)
    ag_centers = universe.select_atoms(centers)
    ag_env = universe.select_atoms(selection)
    fr_idx = list(range(*trajslice.indices(universe.trajectory.n_frames)))
    n_all = len(universe.atoms)
    ix_env = np.asarray(ag_env.indices)
    ix_cent = np.asarray(ag_centers.indices)
    same = ix_env.shape == ix_cent.shape and bool(np.array_equal(ix_env, ix_cent))
    sel_env, sel_cent = frames.Selection(ix_env, n_all), frames.Selection(ix_cent, n_all)
    batch_frames = frames.pick_batch_frames(max(ix_env.size, ix_cent.size, 1), 160.0, 1)

    def batches() -> Iterator[tuple]:
        for fb in frames.iter_batches(universe, fr_idx, batch_frames):
            box = frames.lens_boxes(fb, None)  # no 1.01 factor here (lens.py:368-371)
            pos_env = fb.select(sel_env, n_all)
            pos_cent = None if same else fb.select(sel_cent, n_all)
            yield pos_env, pos_cent, box

    def build_csr() -> engine.NeighborBatch:
        dev = frames.device()
        parts = [engine.neighbor_lists(pos_env, box, r_cut, pos_cent, respect_pbc)
                 for pos_env, pos_cent, box in batches()]  # fmt: skip
        if not parts:
            return engine.NeighborBatch(
                torch.empty((0, ix_cent.size), dtype=torch.int32, device=dev),
                torch.zeros((0, ix_cent.size + 1), dtype=torch.int32, device=dev),
                torch.zeros(1, dtype=torch.int64, device=dev),
                torch.zeros(0, dtype=torch.int32, device=dev), 0, ix_cent.size,
            )  # fmt: skip
        if len(parts) == 1:
            return parts[0]
        offs = [torch.zeros(1, dtype=torch.int64, device=dev)]
        base = 0
        for nb in parts:
            offs.append(nb.frame_off[1:] + base)
            base += int(nb.frame_off[-1].item())
        return engine.NeighborBatch(
            torch.cat([nb.counts for nb in parts]), torch.cat([nb.indptr for nb in parts]),
            torch.cat(offs), torch.cat([nb.indices for nb in parts]),
            sum(nb.n_frames for nb in parts), ix_cent.size,
        )  # fmt: skip

    counts = []
    if ix_cent.size == 0 or ix_env.size == 0:  # nothing to search: all rows are empty
        dev = frames.device()
        zero_counts = torch.zeros((len(fr_idx), ix_cent.size), dtype=torch.int32, device=dev)

        def empty_csr() -> engine.NeighborBatch:
            return engine.NeighborBatch(
                zero_counts, torch.zeros((len(fr_idx), ix_cent.size + 1), dtype=torch.int32,
                                         device=dev),
                torch.zeros(len(fr_idx) + 1, dtype=torch.int64, device=dev),
                torch.zeros(0, dtype=torch.int32, device=dev), len(fr_idx), ix_cent.size,
            )  # fmt: skip

        return NeighbourList(ag_env, zero_counts, empty_csr, ix_env, ix_cent)
    if fr_idx:
        for pos_env, pos_cent, box in batches():
            if pos_cent is None and engine.lens_direct_supported(
                pos_env.shape[0], pos_env.shape[1], box, r_cut, respect_pbc
            ):  # centres == environment, sparse cells: half-shell pair counts, no lists
                counts.append(engine.lens_direct(pos_env, box, r_cut, 0, respect_pbc,
                                                 want_lens=False)[1])  # fmt: skip
                continue
            # counts only: the rows go to a scratch buffer and are dropped
            nr = engine.neighbor_rows_checked(pos_env, box, r_cut, pos_cent, respect_pbc)
            counts.append(nr.counts)
    dev_counts = (torch.cat(counts, dim=0) if counts else
                  torch.empty((0, ix_cent.size), dtype=torch.int32, device=frames.device()))  # fmt: skip
    return NeighbourList(ag_env, dev_counts, build_csr, ix_env, ix_cent)
