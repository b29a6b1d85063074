"""descriptors package -- drop-in for the neighbour-list consumers of ``dynsight.descriptors``
(reference ``src/dynsight/descriptors.py`` / ``_internal/descriptors/misc.py:15-219``):
``orientational_order_param``, ``compute_mean_alignment``, ``velocity_alignment``.

SURVEY.md section 8f row 2: per-centre reductions over the CSR rows that
``list_neighbours_along_trajectory`` produces.  ``many_body_tica`` (deeptime) is out of scope.
"""

from __future__ import annotations

from typing import TYPE_CHECKING, Any, Callable

import numpy as np
import torch

from dynsight_b200 import engine, frames
from dynsight_b200.lens import NeighbourList

if TYPE_CHECKING:
    from numpy.typing import NDArray

__all__ = ["compute_mean_alignment", "orientational_order_param", "velocity_alignment"]


def _all_positions(universe: Any) -> torch.Tensor:
    """Positions of all atoms for every frame of the trajectory -> (F, N, 3) on the device."""
    n_frames = len(universe.trajectory)
    # iter_batches reuses two device buffers: keep copies, not views
    chunks = [fb.pos_all.clone() for fb in frames.iter_batches(universe, list(range(n_frames)), 4096)]
    return chunks[0] if len(chunks) == 1 else torch.cat(chunks, dim=0)


def _device_csr(neigh_list_per_frame: Any, n_atoms: int) -> engine.NeighborBatch:
    """CSR of a neighbour list on the device: the lazy NeighbourList hands out its device-resident
    canonical CSR; a plain ``list[list[AtomGroup]]`` is flattened through ``.indices``."""
    dev = frames.device()
    if isinstance(neigh_list_per_frame, NeighbourList):
        nb = neigh_list_per_frame.device_batch()  # stays on the device: no host round trip
        cent = neigh_list_per_frame.centre_indices
        if nb.n_cent != n_atoms or (cent is not None and not np.array_equal(cent, np.arange(n_atoms))):
            msg = "the neighbour list must hold one entry per atom of the universe (centers='all')"
            raise ValueError(msg)
        indices = nb.indices
        env = neigh_list_per_frame.env_indices
        if env is not None and indices.numel() and not (
            env.size == n_atoms and np.array_equal(env, np.arange(n_atoms))
        ):
            # rows hold positions inside the environment selection; the reference reads
            # ``neigh[t][i].indices`` = universe indices (misc.py:60-66, 180-190)
            indices = torch.as_tensor(env.astype(np.int32), device=dev)[indices.long()]
        if indices.numel() == 0:
            indices = torch.zeros(1, dtype=torch.int32, device=dev)
        if indices is not nb.indices:
            nb = engine.NeighborBatch(nb.counts, nb.indptr, nb.frame_off, indices, nb.n_frames,
                                      nb.n_cent)  # fmt: skip
        return nb
    rows = [[np.asarray(ag.indices, dtype=np.int32) for ag in fr] for fr in neigh_list_per_frame]
    indptr = np.zeros((len(rows), n_atoms + 1), dtype=np.int32)
    flat, offs = [], [0]
    for f, fr in enumerate(rows):
        indptr[f, 1:] = np.cumsum([len(r) for r in fr])
        flat.extend(fr)
        offs.append(offs[-1] + int(indptr[f, -1]))
    indices = np.concatenate(flat) if flat else np.zeros(0, np.int32)
    frame_off = np.asarray(offs, dtype=np.int64)
    counts = torch.as_tensor(np.diff(indptr, axis=1), device=dev)
    if indptr.shape[1] != n_atoms + 1:
        msg = "the neighbour list must hold one entry per atom of the universe (centers='all')"
        raise ValueError(msg)
    idx = indices if indices.size else np.zeros(1, np.int32)
    return engine.NeighborBatch(
        counts, torch.as_tensor(indptr, device=dev), torch.as_tensor(frame_off, device=dev),
        torch.as_tensor(idx, device=dev), indptr.shape[0], n_atoms,
    )  # fmt: skip


def orientational_order_param(
    universe: Any,
    neigh_list_per_frame: Any,
    order: int = 6,
) -> NDArray[np.float64]:
    """Magnitude of the orientational order parameter |psi_n| (particles in the (x, y) plane).

    Returns an array of shape (n_atoms, n_frames).
    """
    pos = _all_positions(universe)
    batch = _device_csr(neigh_list_per_frame, pos.shape[1])
    return frames.to_host(engine.orientational_op_device(pos, batch, order))


def compute_mean_alignment(
    neigh_list_t: Any,
    vectors: NDArray[np.float64],
    metric: Callable[[NDArray[np.float64], NDArray[np.float64]], float],
) -> NDArray[np.float64]:
    """Average vector alignment for all the atoms in a frame with an arbitrary Python ``metric``.

    Host utility kept for API completeness (a Python callable cannot run on the device);
    ``velocity_alignment`` does not use it -- its cosine metric is evaluated by the CUDA kernel.
    """
    phi_t = np.zeros(len(vectors))
    for i, atom_i in enumerate(vectors):
        if not np.any(atom_i):
            continue
        neighbors = neigh_list_t[i].indices
        valid = [j for j in neighbors if j != i and np.any(vectors[j])]
        if not valid:
            continue
        phi_t[i] = np.mean([metric(np.array(atom_i), vectors[j]) for j in valid])
    return phi_t


def velocity_alignment(
    universe: Any,
    neigh_list_per_frame: Any,
) -> NDArray[np.float64]:
    """Average velocity alignment phi.

    With velocities in the Universe the output has shape (n_atoms, n_frames); otherwise the
    displacements from the first frame are used (as the reference does) and the shape is
    (n_atoms, n_frames - 1).
    """
    n_frames = len(universe.trajectory)
    has_vel = hasattr(universe.atoms, "velocities") and universe.atoms.velocities is not None
    if has_vel:
        vel = np.empty((n_frames, len(universe.atoms), 3), dtype=np.float32)
        for t, _ in enumerate(universe.trajectory):
            vel[t] = universe.atoms.velocities
        vec = torch.as_tensor(vel, device=frames.device())
    else:
        vec = _all_positions(universe)
    batch = _device_csr(neigh_list_per_frame, vec.shape[1])
    return frames.to_host(
        engine.velocity_alignment_device(vec, batch, displacement=not has_vel)
    )
