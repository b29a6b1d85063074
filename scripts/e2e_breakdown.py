"""Wall-clock breakdown of the public-API calls the bench's e2e leg makes."""
import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dynsight_b200 import lens as dlens, soap as dsoap
from dynsight_b200.universe import ArrayUniverse

n, bsz = 100_000, 32
box_l = (n / 0.0334) ** (1 / 3)
rng = np.random.default_rng(0)
host = torch.empty((bsz + 1, n, 3), dtype=torch.float32, pin_memory=True)
cur = rng.random((n, 3)) * box_l
for f in range(bsz + 1):
    cur = np.mod(cur + rng.normal(0, 0.1, cur.shape), box_l)
    host[f] = torch.from_numpy(cur.astype(np.float32))
host.clamp_(max=float(np.nextafter(np.float32(box_l), np.float32(0))))
dims = np.tile(np.array([box_l] * 3 + [90] * 3, dtype=np.float32), (bsz + 1, 1))
uni = ArrayUniverse(host.numpy(), dims, names=["O"] * n, types=["O"] * n)
def t(label, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    print(f"{label:40s} {1e3 * (time.perf_counter() - t0):8.2f} ms"); return r
for it in range(3):
    print("--- iteration", it)
    d = t("compute_lens_device", lambda: dlens.compute_lens_device(uni, 5.0, 1))
    t("  lens .cpu().numpy()", lambda: d.cpu().numpy())
    s = t("timesoap_from_trajectory (device)", lambda: dsoap.timesoap_from_trajectory(uni, 5.0, 8, 8, delay=1))
    t("  tsoap .cpu().numpy()", lambda: s.cpu().numpy())
    t("compute_lens (numpy out)", lambda: dlens.compute_lens(uni, 5.0, 1))
