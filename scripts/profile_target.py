"""Small single-purpose launcher for ncu: runs one hot-path stage a few times.

usage: python scripts/profile_target.py soap|soap3|neigh|neigh_fluid|pair|pair_fluid|tsoap|tcf|spavg|spavg324 [atoms] [frames]
"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dynsight_b200 import engine  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "soap"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
nf = int(sys.argv[3]) if len(sys.argv) > 3 else 8
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
if what in ("soap3", "fused3"):  # cfg4-like: 3 species, l_max = 10, metal density
    box_l = (n / 0.0857) ** (1 / 3)
    r_cut = 5.0
elif what in ("neigh_fluid", "pair_fluid", "spavg", "spavg324"):
    box_l = (n / 0.8) ** (1 / 3)
    r_cut = 1.5
else:
    box_l = (n / 0.0334) ** (1 / 3)
    r_cut = 5.0
pos = (torch.rand((nf, n, 3), device=dev, generator=g, dtype=torch.float64) * box_l).float()
if what.startswith("fused"):   # the fused kernels walk a trajectory: frames must be continuous
    step = 0.1 * torch.randn((nf, n, 3), device=dev, generator=g, dtype=torch.float64).cumsum(0)
    pos = torch.remainder(pos[:1].double() + step, box_l).float()
pos.clamp_(max=float(np.nextafter(np.float32(box_l), np.float32(0))))
box = np.full(3, np.float32(box_l), dtype=np.float32)
sp = torch.zeros(n, dtype=torch.int32, device=dev)
cen = torch.arange(n, dtype=torch.int32, device=dev)
for it in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if what == "soap":
        out = engine.soap_power_spectrum(pos, sp, cen, box, 5.0, 8, 8)
    elif what == "fused":
        out = engine.soap_timesoap(pos, sp, cen, box, 5.0, 8, 8, delay=1)
    elif what == "fused3":
        sp3 = (torch.arange(n, device=dev, dtype=torch.int32) * 7919 % 3).to(torch.int32)
        out = engine.soap_timesoap(pos, sp3, cen, box, 5.0, 8, 10, delay=1, n_species=3)
    elif what == "soap3":
        sp3 = (torch.arange(n, device=dev, dtype=torch.int32) * 7919 % 3).to(torch.int32)
        out = engine.soap_power_spectrum(pos, sp3, cen, box, 5.0, 8, 10, n_species=3)
    elif what in ("pair", "pair_fluid"):
        out, _ = engine.lens_direct(pos, box, r_cut, 1, True)
    elif what in ("neigh", "neigh_fluid"):
        if len(sys.argv) > 4 and sys.argv[4] == "csr":
            nb = engine.neighbor_lists(pos, box, r_cut)
            out = engine.lens_pairs(nb, 1)
        else:
            out, _ = engine.lens_from_positions(pos, box, r_cut, 1)
    elif what == "tcf":  # n particles x nf frames, all lags
        data = torch.randn((n, nf), device=dev, dtype=torch.float64).cumsum(1)
        out = engine.self_time_correlation_device(data, nf)
    elif what in ("spavg", "spavg324"):
        n_comp = 324 if what == "spavg324" else 1
        vals = torch.rand((n, nf) if n_comp == 1 else (n, nf, n_comp), device=dev, dtype=torch.float64)
        boxes = np.tile(box.astype(np.float64), (nf, 1))
        nr = engine.neighbor_rows(engine.wrap_positions(pos, boxes), boxes, r_cut, inclusive_cutoff=True)
        out = engine.spatial_average_rows(nr, vals, torch.empty_like(vals))
    else:
        soap = torch.rand((n, nf, 324), device=dev, dtype=torch.float64)
        out = engine.timesoap_device(soap, 1)
    e1.record()
    torch.cuda.synchronize()
    print(what, n, nf, f"{e0.elapsed_time(e1):.3f} ms", f"{n * nf / e0.elapsed_time(e1) * 1e3:.3e} pf/s")
