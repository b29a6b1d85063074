import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dynsight_b200 import engine
for (n_species, n_max, l_max) in ((3, 8, 10), (1, 12, 5), (2, 4, 3)):
    rng = np.random.default_rng(17 + n_species + n_max)
    n, n_frames = 260, 23
    box = np.array([27.0, 28.5, 26.5], dtype=np.float32)
    pos = (rng.random((1, n, 3)) * box + rng.normal(0, 0.15, (n_frames, n, 3)).cumsum(0)).astype(np.float32)
    dpos = torch.as_tensor(pos).cuda()
    species = torch.as_tensor(rng.integers(0, n_species, n).astype(np.int32)).cuda()
    centres = torch.as_tensor(np.sort(rng.choice(n, 37, replace=False)).astype(np.int32)).cuda()
    ref = engine.soap_power_spectrum(dpos, species, centres, box, 4.5, n_max, l_max, n_species=n_species)
    for delay in (1, 2, 3):
        want = engine.timesoap_device(ref, delay)
        spectra = torch.full_like(ref, float("nan"))
        got = engine.soap_timesoap(dpos, species, centres, box, 4.5, n_max, l_max, delay=delay, n_species=n_species, soap_out=spectra)
        err = (got - want).abs()
        serr = ((spectra - ref).abs() / ref.abs().amax(-1, keepdim=True)).amax(-1)
        print(n_species, n_max, l_max, "delay", delay, "max ts err", float(err.max()), "bad cols", (err.amax(0) > 1e-8).nonzero().flatten().tolist(),
              "bad centres", int((err.amax(1) > 1e-8).sum()), "spectra err", float(torch.nan_to_num(serr, nan=9.0).max()),
              "bad spectra frames", (torch.nan_to_num(serr, nan=9.0).amax(0) > 1e-8).nonzero().flatten().tolist())
