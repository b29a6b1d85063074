"""Randomised differential test of the fused SOAP -> timeSOAP kernels (dsg_soap_timesoap: the
frame-sequential tensor-path kernel, soap_spectrum_mma_kernel with the ring in shared or global
memory, soap_spectrum_kernel) against the two-kernel path soap_power_spectrum + timesoap_device,
and of the spectra they write on request.

usage: python scripts/fuzz_fused.py [n_cases] [seed]
"""
import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dynsight_b200 import engine

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad = 0
worst = 0.0
for case in range(n_cases):
    nf = int(rng.integers(3, 12))
    n = int(rng.choice([5, 40, 150, 400]))
    r_cut = float(rng.choice([2.5, 4.0, 5.0]))
    n_max = int(rng.choice([1, 3, 4, 8, 8, 9, 12]))
    l_max = int(rng.choice([0, 2, 4, 6, 8, 8, 10, 12]))
    n_species = int(rng.choice([1, 1, 2, 3, 4]))
    delay = int(rng.integers(1, min(4, nf)))
    radius = r_cut + 3.7169
    periodic = rng.random() < 0.8
    edge = radius * rng.uniform(0.6, 4.5, size=3)
    if rng.random() < 0.5:
        edge[:] = radius * rng.uniform(3.05, 4.5)   # list mode
    base = rng.random((1, n, 3)) * edge
    pos = (base + rng.normal(0, 0.12, (nf, n, 3)).cumsum(0)).astype(np.float32)
    sp = np.unique(rng.integers(0, n_species, n), return_inverse=True)[1]
    n_species = int(sp.max()) + 1
    centres = np.sort(rng.choice(n, size=int(rng.integers(1, min(n, 40) + 1)), replace=False))
    box = np.tile(edge, (nf, 1)) if periodic else None
    label = (f"case {case}: n={n} nf={nf} delay={delay} r_cut={r_cut} n_max={n_max} l_max={l_max} "
             f"S={n_species} periodic={periodic} edge/R={np.round(edge / radius, 2)}")
    if len(sys.argv) > 3:
        print(label, flush=True)
    dpos = torch.as_tensor(pos).cuda()
    dsp = torch.as_tensor(sp, dtype=torch.int32).cuda()
    dcen = torch.as_tensor(centres, dtype=torch.int32).cuda()
    ref = engine.soap_power_spectrum(dpos, dsp, dcen, box, r_cut, n_max, l_max, n_species=n_species)
    want = engine.timesoap_device(ref, delay)
    spectra = torch.full_like(ref, float("nan")) if rng.random() < 0.4 else None
    got = engine.soap_timesoap(dpos, dsp, dcen, box, r_cut, n_max, l_max, delay=delay,
                               n_species=n_species, soap_out=spectra)
    err = float((got - want).abs().max())
    ok = got.shape == want.shape and err < 5e-8   # sqrt(2 - 2 cos) amplifies 1e-13 on the spectra
    if spectra is not None:
        scale = ref.abs().amax(dim=-1, keepdim=True).clamp_min(1e-300)
        ok = ok and float(((spectra - ref).abs() / scale).max()) < 1e-8
    worst = max(worst, err)
    if not ok:
        bad += 1
        print("MISMATCH", label, "abs err", err)
print(f"{n_cases} cases, {bad} mismatches, worst absolute timeSOAP difference {worst:.2e}")
sys.exit(1 if bad else 0)
