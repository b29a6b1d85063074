"""API latency on BASELINE cfg1 (2048 particles x 200 frames, LENS + coordination number)."""
import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
from golden_cases import fluid_case
from dynsight_b200.trajectory import Trj
from dynsight_b200.universe import ArrayUniverse

pos, box, r_cut = fluid_case(n_side=13, n_frames=200)  # 2197 particles
dims = np.tile(np.concatenate([box, [90, 90, 90]]).astype(np.float32), (pos.shape[0], 1))
trj = Trj(ArrayUniverse(pos, dims, ["A"] * pos.shape[1], ["A"] * pos.shape[1]))
print("system", pos.shape, "r_cut", r_cut)
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    lens = trj.get_lens(r_cut=r_cut)
    t1 = time.perf_counter()
    _, nc = trj.get_coord_number(r_cut=r_cut)
    t2 = time.perf_counter()
    print(f"get_lens {1e3*(t1-t0):7.2f} ms   get_coord_number {1e3*(t2-t1):7.2f} ms   "
          f"({pos.shape[1]*(pos.shape[0]-1)/(t1-t0):.3e} pf/s LENS)")
