"""Which pipeline call is slow, and what the allocators did in it."""
import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dynsight_b200 import pipeline
from dynsight_b200.universe import ArrayUniverse
n, nf = 100_000, 33
box_l = (n / 0.0334) ** (1 / 3)
rng = np.random.default_rng(0)
pos = torch.empty((nf, n, 3), dtype=torch.float32, pin_memory=True)
cur = rng.random((n, 3)) * box_l
for f in range(nf):
    cur = np.mod(cur + rng.normal(0, 0.1, cur.shape), box_l)
    pos[f] = torch.from_numpy(np.minimum(cur, np.nextafter(np.float32(box_l), np.float32(0))).astype(np.float32))
dims = np.tile(np.asarray([box_l] * 3 + [90] * 3, dtype=np.float32), (nf, 1))
uni = ArrayUniverse(pos.numpy(), dims, names=["O"] * n, types=["O"] * n)
def stats():
    d = torch.cuda.memory_stats()
    try:
        h = torch.cuda.host_memory_stats()
        hs = f"host allocs {h.get('num_host_alloc', '?')} frees {h.get('num_host_free', '?')} reserved {h.get('reserved_bytes.current', 0) / 1e6:.0f} MB"
    except Exception as e:  # noqa: BLE001
        hs = f"(no host stats: {e})"
    return f"dev segments {d['segment.all.current']} cudaMalloc {d['num_device_alloc']} cudaFree {d['num_device_free']} reserved {d['reserved_bytes.all.current'] / 1e9:.2f} GB | {hs}"
for it in range(10):
    t0 = time.perf_counter()
    res = pipeline.descriptors_from_trajectory(uni, lens_r_cut=5.0, soap_r_cut=5.0, coord_number=True)
    t1 = time.perf_counter()
    del res
    print(f"call {it}: {1e3 * (t1 - t0):7.2f} ms  {stats()}", flush=True)
