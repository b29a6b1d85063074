"""Multi-batch LENS through the public API on a 1M-particle fluid: shows that the host->device
copies of batch k+1 overlap the kernels of batch k (side stream in frames.iter_batches)."""
import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dynsight_b200 import lens as dlens, engine
from dynsight_b200.universe import ArrayUniverse

n, nf, bsz = 1_000_000, 36, 9
box_l = (n / 0.8) ** (1 / 3)
g = torch.Generator().manual_seed(0)
host = torch.empty((nf, n, 3), dtype=torch.float32, pin_memory=True)
cur = torch.rand((n, 3), generator=g) * box_l
for f in range(nf):
    cur = torch.remainder(cur + 0.05 * torch.randn((n, 3), generator=g), box_l)
    host[f] = cur
host.clamp_(max=float(np.nextafter(np.float32(box_l), np.float32(0))))
dims = np.tile(np.array([box_l] * 3 + [90] * 3, dtype=np.float32), (nf, 1))
uni = ArrayUniverse(host.numpy(), dims, names=["A"] * n, types=["A"] * n)
def timed(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0), r
for it in range(3):
    t_all, _ = timed(lambda: dlens.compute_lens_device(uni, 1.5, 1, batch_frames=bsz))
    t_h2d, dev = timed(lambda: host.cuda(non_blocking=True))
    box = np.full(3, np.float32(box_l))
    t_cmp, _ = timed(lambda: [engine.lens_from_positions(dev[s:s + bsz], box, 1.5, 1) for s in range(0, nf - 1, bsz - 1)])
    big = torch.empty((n, nf - 1), dtype=torch.float64, device="cuda")
    t_big, _ = timed(lambda: [engine.lens_from_positions(dev[s:s + bsz], box, 1.5, 1, out=big, col0=s) for s in range(0, nf - 1, bsz - 1)])
    print(f"kernels alone, results into the (N, F-1) matrix: {t_big:6.2f} ms")
    print(f"api (copies overlapped) {t_all:7.2f} ms | H2D alone {t_h2d:6.2f} ms | kernels alone {t_cmp:6.2f} ms")

import cProfile, pstats
for b in (36, 9):
    t, _ = timed(lambda: dlens.compute_lens_device(uni, 1.5, 1, batch_frames=b))
    print(f"batch_frames={b}: {t:.2f} ms")
pr = cProfile.Profile(); pr.enable()
dlens.compute_lens_device(uni, 1.5, 1, batch_frames=bsz); torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
