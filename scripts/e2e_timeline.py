"""Host timeline of pipeline.descriptors_from_trajectory at the bench shape (where the GPU idles)."""
import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dynsight_b200 import pipeline, frames, lens as dlens, soap as dsoap
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
marks = []
def mark(name):
    marks.append((name, time.perf_counter()))
def wrap(obj, attr, name):
    fn = getattr(obj, attr)
    def inner(*a, **k):
        mark(name + " >")
        r = fn(*a, **k)
        mark(name + " <")
        return r
    setattr(obj, attr, inner)
wrap(dlens.LensJob, "__init__", "LensJob.__init__")
wrap(dsoap.TimesoapJob, "__init__", "TimesoapJob.__init__")
wrap(dlens.LensJob, "process", "LensJob.process")
wrap(dsoap.TimesoapJob, "process", "TimesoapJob.process")
wrap(frames, "to_host_async", "to_host_async")
wrap(frames, "to_host", "to_host")
pipeline.LensJob, pipeline.TimesoapJob = dlens.LensJob, dsoap.TimesoapJob
def call():
    return pipeline.descriptors_from_trajectory(uni, lens_r_cut=5.0, soap_r_cut=5.0, coord_number=True)
for _ in range(3): call()
torch.cuda.synchronize()
for it in range(2):
    marks.clear()
    t0 = time.perf_counter(); call(); t1 = time.perf_counter()
    print(f"--- call {1e3*(t1-t0):.2f} ms")
    for name, t in marks:
        print(f"  {1e3*(t-t0):8.3f} ms  {name}")
