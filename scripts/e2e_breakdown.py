"""Where the end-to-end time of pipeline.descriptors_from_trajectory goes (run alone or under
torchrun: python -m torch.distributed.run --nproc-per-node N scripts/e2e_breakdown.py)."""
import os, sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dynsight_b200 import pipeline, frames
from dynsight_b200.universe import ArrayUniverse

rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(rank)
n, nf = 100_000, 33
box_l = (n / 0.0334) ** (1 / 3)
rng = np.random.default_rng(rank)
pos = torch.empty((nf, n, 3), dtype=torch.float32, pin_memory=True)
cur = rng.random((n, 3)) * box_l
for f in range(nf):
    cur = np.mod(cur + rng.normal(0, 0.1, cur.shape), box_l)
    pos[f] = torch.from_numpy(np.minimum(cur, np.nextafter(np.float32(box_l), np.float32(0))).astype(np.float32))
dims = np.tile(np.asarray([box_l] * 3 + [90] * 3, dtype=np.float32), (nf, 1))
uni = ArrayUniverse(pos.numpy(), dims, names=["O"] * n, types=["O"] * n)
def call():
    return pipeline.descriptors_from_trajectory(uni, lens_r_cut=5.0, soap_r_cut=5.0, coord_number=True)
for _ in range(2): call()
torch.cuda.synchronize()
for it in range(4):
    t0 = time.perf_counter(); res = call(); t1 = time.perf_counter()
    print(f"rank {rank} pipeline {1e3*(t1-t0):.2f} ms", flush=True)
# pieces
d = torch.empty((nf, n, 3), dtype=torch.float32, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(pos, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"rank {rank} h2d 39.6MB {1e3*(t1-t0):.2f} ms")
r = torch.empty((n, 32), dtype=torch.float64, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter(); h = frames.to_host(r); t1 = time.perf_counter()
print(f"rank {rank} d2h 25.6MB {1e3*(t1-t0):.2f} ms")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); call(); pr.disable()
if rank == 0:
    pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
