"""world_size-2/3 gloo tests (CPU) of the frame-block sharding + halo exchange host logic.

The per-rank compute is the CPU oracle here (allowed in tests); on the GPU box the same driver
code runs with the CUDA engine and NCCL (bench.py --gpus N)."""

from __future__ import annotations

import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, n_frames: int, delay: int, out_dir: str) -> None:
    for p in (str(ROOT), str(ROOT / "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from golden_cases import fluid_case

        from dynsight_b200 import parallel
        from oracle import lens_oracle as lo
        from oracle import timesoap_oracle as to

        pos, box, r_cut = fluid_case(n_side=6, n_frames=n_frames)
        shard = parallel.shard_frames(n_frames, world, rank, delay)
        owned = torch.as_tensor(pos[shard.start : shard.stop])
        ext = parallel.exchange_halo(owned, shard).numpy()
        assert ext.shape[0] == shard.n_owned + shard.halo
        assert np.array_equal(ext, pos[shard.start : shard.stop + shard.halo])
        dims = np.tile(np.concatenate([box, [90, 90, 90]]).astype(np.float32), (ext.shape[0], 1))
        if shard.n_pairs_owned > 0:
            lens = lo.compute_lens_arrays(ext, dims, r_cut, delay=delay)[:, : shard.n_pairs_owned]
            fake_soap = np.cumsum(ext.reshape(ext.shape[0], -1, 3), axis=2).transpose(1, 0, 2) + 1.0
            ts = to.timesoap(fake_soap, delay)[:, : shard.n_pairs_owned]
        else:
            lens = np.zeros((pos.shape[1], 0))
            ts = np.zeros((pos.shape[1], 0))
        full_lens = parallel.gather_columns(torch.as_tensor(lens))
        full_ts = parallel.gather_columns(torch.as_tensor(ts))
        if rank == 0:
            np.save(Path(out_dir) / "lens.npy", full_lens.numpy())
            np.save(Path(out_dir) / "ts.npy", full_ts.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize(("world", "n_frames", "delay"), [(2, 9, 1), (2, 8, 3), (3, 10, 2)])
def test_sharded_lens_equals_single_process(tmp_path, world, n_frames, delay):
    from golden_cases import fluid_case

    from oracle import lens_oracle as lo
    from oracle import timesoap_oracle as to

    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_frames, delay, str(tmp_path)), nprocs=world, join=True)
    pos, box, r_cut = fluid_case(n_side=6, n_frames=n_frames)
    dims = np.tile(np.concatenate([box, [90, 90, 90]]).astype(np.float32), (n_frames, 1))
    want = lo.compute_lens_arrays(pos, dims, r_cut, delay=delay)
    got = np.load(tmp_path / "lens.npy")
    assert got.shape == want.shape == (pos.shape[1], n_frames - delay)
    assert np.array_equal(got, want)
    fake_soap = np.cumsum(pos.reshape(n_frames, -1, 3), axis=2).transpose(1, 0, 2) + 1.0
    assert np.array_equal(np.load(tmp_path / "ts.npy"), to.timesoap(fake_soap, delay))


def test_shard_arithmetic():
    from dynsight_b200 import parallel

    for n_frames in (1, 7, 100, 10_000):
        for world in (1, 2, 4, 8):
            for delay in (1, 3):
                shards = [parallel.shard_frames(n_frames, world, r, delay) for r in range(world)]
                assert shards[0].start == 0 and shards[-1].stop == n_frames
                assert all(a.stop == b.start for a, b in zip(shards, shards[1:]))
                assert sum(s.n_pairs_owned for s in shards) == max(0, n_frames - delay)
                assert shards[-1].halo == 0
    with pytest.raises(ValueError):
        parallel.shard_frames(10, 2, 2)


def _driver_worker(rank: int, world: int, port: int, n_frames: int, delay: int, out_dir: str
                   ) -> None:  # fmt: skip
    """compute_lens_sharded with the device compute replaced by the CPU oracle (test-only stub):
    checks the driver's slicing / ownership / gather logic without a GPU."""
    for p in (str(ROOT), str(ROOT / "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from golden_cases import fluid_case

        from dynsight_b200 import frames, lens, parallel
        from dynsight_b200.universe import ArrayUniverse
        from oracle import lens_oracle as lo

        pos, box, r_cut = fluid_case(n_side=6, n_frames=n_frames)
        dims = np.tile(np.concatenate([box, [90, 90, 90]]).astype(np.float32), (n_frames, 1))
        uni = ArrayUniverse(pos, dims, ["A"] * pos.shape[1], ["A"] * pos.shape[1])

        def cpu_lens(universe, r_cut, delay, centers, selection, trajslice, respect_pbc):
            idx = list(range(universe.trajectory.n_frames))[trajslice]
            return torch.as_tensor(lo.compute_lens_arrays(pos[idx], dims[idx], r_cut, delay=delay))

        lens.compute_lens_device = cpu_lens
        frames.device = lambda: torch.device("cpu")
        sl = slice(1, None)  # composed with the shard's block
        full = parallel.compute_lens_sharded(uni, r_cut, delay, trajslice=sl)
        if rank == 0:
            np.save(Path(out_dir) / "driver_lens.npy", full.numpy())
        else:
            assert full is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize(("world", "n_frames", "delay"), [(2, 9, 1), (3, 11, 2)])
def test_sharded_driver_slices_and_gathers(tmp_path, world, n_frames, delay):
    from golden_cases import fluid_case

    from oracle import lens_oracle as lo

    port = _free_port()
    mp.spawn(_driver_worker, args=(world, port, n_frames, delay, str(tmp_path)), nprocs=world,
             join=True)  # fmt: skip
    pos, box, r_cut = fluid_case(n_side=6, n_frames=n_frames)
    dims = np.tile(np.concatenate([box, [90, 90, 90]]).astype(np.float32), (n_frames, 1))
    want = lo.compute_lens_arrays(pos[1:], dims[1:], r_cut, delay=delay)
    assert np.array_equal(np.load(tmp_path / "driver_lens.npy"), want)
