"""Frame-block sharding of a trajectory across the GPUs of one box (one process per GPU).

Neighbour lists, coordination numbers and SOAP are independent per frame; LENS and timeSOAP
couple frame ``t`` with ``t + delay`` only.  So the trajectory is cut into
contiguous frame blocks, rank ``r`` owning ``[F*r/W, F*(r+1)/W)``, and the only data that crosses
ranks is a halo of ``delay`` frames of positions that each rank receives from its successor
(``torch.distributed`` send/recv: NCCL over NVLink on GPUs, gloo on CPU tensors in the tests).
Outputs stay on the rank that produced them (rank ``r`` holds the pairs that start in its block);
``gather_columns`` concatenates them on demand.  There is no reduction and no all-to-all.
"""

from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist


@dataclass(frozen=True)
class FrameShard:
    rank: int
    world: int
    n_frames: int  # frames of the whole (sliced) trajectory
    delay: int
    start: int  # first owned frame (position in the frame sequence)
    stop: int  # one past the last owned frame

    @property
    def n_owned(self) -> int:
        return self.stop - self.start

    @property
    def halo(self) -> int:
        """Frames needed from the following ranks so that every owned frame ``t`` with
        ``t + delay < n_frames`` finds its partner."""
        return max(0, min(self.n_frames, self.stop + self.delay) - self.stop)

    @property
    def n_pairs_owned(self) -> int:
        """Pairs (t, t+delay) whose first frame is owned by this rank."""
        return max(0, min(self.stop, self.n_frames - self.delay) - self.start)


def shard_frames(n_frames: int, world: int, rank: int, delay: int = 1) -> FrameShard:
    """Contiguous block ``[F*r/W, F*(r+1)/W)`` of rank ``r``."""
    if world < 1 or not 0 <= rank < world:
        msg = f"bad rank {rank} / world {world}"
        raise ValueError(msg)
    start = n_frames * rank // world
    stop = n_frames * (rank + 1) // world
    return FrameShard(rank, world, n_frames, delay, start, stop)


def exchange_halo(
    owned: torch.Tensor, shard: FrameShard, group: dist.ProcessGroup | None = None
) -> torch.Tensor:
    """Return ``owned`` (n_owned, ...) extended by the halo frames received from the next rank.

    Rank ``r > 0`` sends its first ``halo(r-1)`` frames to rank ``r-1``; rank ``r < W-1`` receives
    ``halo(r)`` frames from rank ``r+1``.  Both sides derive the counts from :func:`shard_frames`,
    so one message per neighbour pair suffices.  Blocks must be at least ``delay`` frames long
    (except the last one).  Works on CUDA tensors (NCCL) and CPU tensors (gloo).
    """
    if shard.world == 1:
        return owned
    if owned.shape[0] != shard.n_owned:
        msg = f"rank {shard.rank}: got {owned.shape[0]} frames, owns {shard.n_owned}"
        raise ValueError(msg)
    ops = []
    recv_buf = None
    if shard.rank > 0:
        prev = shard_frames(shard.n_frames, shard.world, shard.rank - 1, shard.delay)
        if prev.halo > shard.n_owned:
            msg = (
                f"frame block of rank {shard.rank} ({shard.n_owned} frames) is shorter than the "
                f"halo of {prev.halo} frames: use fewer ranks or a smaller delay"
            )
            raise ValueError(msg)
        if prev.halo > 0:
            ops.append(dist.P2POp(dist.isend, owned[: prev.halo].contiguous(), shard.rank - 1,
                                  group))  # fmt: skip
    if shard.rank < shard.world - 1 and shard.halo > 0:
        recv_buf = torch.empty((shard.halo, *owned.shape[1:]), dtype=owned.dtype,
                               device=owned.device)  # fmt: skip
        ops.append(dist.P2POp(dist.irecv, recv_buf, shard.rank + 1, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return owned if recv_buf is None else torch.cat([owned, recv_buf], dim=0)


def gather_columns(
    local: torch.Tensor, group: dist.ProcessGroup | None = None
) -> torch.Tensor | None:
    """Concatenate per-rank result slabs (n_centers, n_local_pairs) along the frame axis on
    rank 0 (other ranks return ``None``).  Slab widths may differ between ranks.

    Only the widths are all-gathered (one int64 per rank); the slabs travel point to point to
    rank 0, which receives them one at a time straight behind each other -- no rank ever holds
    more than its own slab, and rank 0 the result plus one slab in flight."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    width = torch.tensor([local.shape[1]], dtype=torch.int64, device=local.device)
    widths = [torch.zeros_like(width) for _ in range(world)]
    dist.all_gather(widths, width, group=group)
    if rank != 0:
        if local.shape[1] > 0:
            dist.send(local.contiguous(), dist.get_global_rank(group, 0) if group else 0,
                      group=group)  # fmt: skip
        return None
    ws = [int(w.item()) for w in widths]
    out = torch.empty((local.shape[0], sum(ws)), dtype=local.dtype, device=local.device)
    out[:, : ws[0]] = local
    col = ws[0]
    for r in range(1, world):
        if ws[r] > 0:
            slab = torch.empty((local.shape[0], ws[r]), dtype=local.dtype, device=local.device)
            dist.recv(slab, dist.get_global_rank(group, r) if group else r, group=group)
            out[:, col : col + ws[r]] = slab
            col += ws[r]
    return out


# ------------------------------------------------------------------------------------------------
# Sharded drivers: every rank reads its own frame block (plus the halo) from the trajectory, so no
# positions cross ranks; results stay local or are gathered on rank 0.
# ------------------------------------------------------------------------------------------------
def block_slice(n_frames: int, trajslice: slice | None, shard_start: int, shard_stop: int) -> slice:
    """The slice of the trajectory that holds positions ``shard_start .. shard_stop - 1`` of the
    frame sequence ``range(n_frames)[trajslice]``."""
    seq = range(n_frames)[trajslice if trajslice is not None else slice(None)]
    sub = seq[shard_start:shard_stop]
    if len(sub) == 0:
        return slice(0, 0)
    stop = sub[-1] + sub.step
    return slice(sub[0], stop if stop >= 0 else None, sub.step)


def compute_lens_sharded(
    universe: object,
    r_cut: float,
    delay: int = 1,
    centers: str = "all",
    selection: str = "all",
    trajslice: slice | None = None,
    respect_pbc: bool = True,
    gather: bool = True,
    group: dist.ProcessGroup | None = None,
) -> torch.Tensor | None:
    """``compute_lens`` with the frame pairs split over the ranks of ``group`` (one GPU each).

    Rank ``r`` loads frames ``[F*r/W, F*(r+1)/W + delay)`` itself and computes the pairs that start
    in its block.  Returns the full (n_centers, n_pairs) tensor on rank 0 (``None`` elsewhere) or,
    with ``gather=False``, every rank's own slab.
    """
    from dynsight_b200 import frames, lens

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n_seq = len(frames.frame_sequence(universe.trajectory.n_frames, trajslice))
    if delay < 1 or n_seq - delay < 1:
        msg = "No valid pairs found."
        raise RuntimeError(msg)
    shard = shard_frames(n_seq, world, rank, delay)
    n_cent = universe.select_atoms(centers).n_atoms
    if shard.n_pairs_owned > 0:
        sl = block_slice(universe.trajectory.n_frames, trajslice, shard.start,
                         shard.start + shard.n_pairs_owned + delay)  # fmt: skip
        local = lens.compute_lens_device(universe, r_cut, delay, centers, selection, sl,
                                         respect_pbc)  # fmt: skip
    else:
        local = torch.empty((n_cent, 0), dtype=torch.float64, device=frames.device())
    return gather_columns(local, group) if gather else local


def timesoap_sharded(
    universe: object,
    soaprcut: float,
    soapnmax: int = 8,
    soaplmax: int = 8,
    delay: int = 1,
    selection: str = "all",
    soap_respectpbc: bool = True,
    centers: str = "all",
    trajslice: slice | None = None,
    gather: bool = True,
    group: dist.ProcessGroup | None = None,
) -> torch.Tensor | None:
    """timeSOAP straight from positions (``soap.timesoap_from_trajectory``) with the frame pairs
    split over the ranks of ``group``; same ownership rule as :func:`compute_lens_sharded`."""
    from dynsight_b200 import frames, soap

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n_seq = len(frames.frame_sequence(universe.trajectory.n_frames, trajslice))
    if delay < 1 or delay >= n_seq:
        msg = "delay value outside bounds"
        raise ValueError(msg)
    shard = shard_frames(n_seq, world, rank, delay)
    if shard.n_pairs_owned > 0:
        sl = block_slice(universe.trajectory.n_frames, trajslice, shard.start,
                         shard.start + shard.n_pairs_owned + delay)  # fmt: skip
        local = soap.timesoap_from_trajectory(universe, soaprcut, soapnmax, soaplmax, delay,
                                              selection, soap_respectpbc, centers, sl)  # fmt: skip
    else:
        n_cent = len(soap._centre_positions(universe.select_atoms(selection), centers))  # noqa: SLF001
        local = torch.empty((n_cent, 0), dtype=torch.float64, device=frames.device())
    return gather_columns(local, group) if gather else local
