"""Time the REFERENCE's own numba LENS kernels (lens.py:69-189, imported by path, unmodified) at
the bench shape (cfg5: 100,000 atoms, r_cut = 5 A) in the build container -- /root/reference does
not exist on the GPU box, so bench.py surfaces this figure as cpu_baseline.reference_numba
(static, labelled) beside the live "port" number.  Driven exactly as compute_lens drives them:
two neighbour lists + one merge per frame pair (lens.py:263-308).

    python scripts/time_reference_numba.py > profiles/r2_reference_numba_lens.json
"""
import json
import os
import platform
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import lens_oracle, ref_loader  # noqa: E402

import numba  # noqa: E402

ref = ref_loader.load_ref_lens()
assert ref is not None, "needs /root/reference"
n, rho, r_cut = 100_000, 0.0334, 5.0
rng = np.random.default_rng(0)
box_l = (n / rho) ** (1 / 3)
box = np.full(3, np.float32(box_l), dtype=np.float32)
p0 = (rng.random((n, 3)) * box_l).astype(np.float32)
p1 = np.mod(p0 + rng.normal(0, 0.1, p0.shape), box_l).astype(np.float32)
frames = [np.minimum(p, np.nextafter(box[0], np.float32(0))) for p in (p0, p1)]


def pair(n_jobs: int) -> np.ndarray:
    numba.set_num_threads(n_jobs)
    csr = []
    for p in frames:
        p64 = p.astype(np.float64)
        csr.append(ref.neighbor_list_celllist_centers(positions_env=p64, positions_cent=p64,
                                                      r_cut=r_cut, box=box, respect_pbc=True))
    return ref.lens_from_two_csr(indptr1=csr[0][0], indices1=csr[0][1], indptr2=csr[1][0],
                                 indices2=csr[1][1])


cores = os.cpu_count() or 1
out = {"host": platform.processor() or platform.machine(), "cores": cores, "atoms": n,
       "r_cut": r_cut, "what": "reference numba kernels lens.py:69-189, one frame pair = two "
                               "neighbour lists + merge (lens.py:263-308)"}
lens = pair(cores)  # JIT warm-up
dims = np.tile(np.concatenate([box, [90, 90, 90]]).astype(np.float32), (2, 1))
out["equals_oracle_port"] = bool(np.array_equal(
    lens, lens_oracle.compute_lens_arrays(np.stack(frames), dims, r_cut)[:, 0]))
for jobs in (1, cores):
    best = 1e30
    for _ in range(3):
        t0 = time.perf_counter()
        pair(jobs)
        best = min(best, time.perf_counter() - t0)
    out[f"lens_pf_per_s_n_jobs_{jobs}"] = n / best
print(json.dumps(out, indent=1))
