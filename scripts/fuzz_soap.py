"""Randomised differential test of the SOAP kernels against the CPU oracle (1e-8 relative).

usage: python scripts/fuzz_soap.py [n_cases] [seed]
Draws periodic boxes from below the search radius (several images per atom) to 5 cells per axis,
open (non-periodic) clusters, 1-3 species, n_max 1..12, l_max 0..12, centre subsets, float32 and
float64 inputs, atoms outside the box and coincident atoms.
"""
import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dynsight_b200 import engine
from oracle import soap_oracle as so

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad = 0
worst = 0.0
for case in range(n_cases):
    nf = int(rng.integers(1, 4))
    n = int(rng.choice([1, 3, 20, 120, 400]))
    r_cut = float(rng.choice([2.0, 3.5, 5.0, 6.5]))
    n_max = int(rng.choice([1, 2, 4, 8, 9, 12]))
    l_max = int(rng.choice([0, 1, 3, 6, 8, 10, 12]))
    n_species = int(rng.choice([1, 1, 2, 3]))
    radius = r_cut + 3.7169
    periodic = rng.random() < 0.75
    edge = radius * rng.uniform(0.45, 5.0, size=3)
    if rng.random() < 0.3:
        edge[:] = edge[0]
    box = np.tile(edge, (nf, 1)) * rng.uniform(0.98, 1.02, size=(nf, 1))
    spread = rng.choice([1.0, 1.6])
    pos = (rng.random((nf, n, 3)) * spread - (spread - 1) / 2) * edge
    if n > 2 and rng.random() < 0.2:
        pos[:, 1] = pos[:, 0] + 1e-3   # nearly coincident atoms
    dtype = np.float32 if rng.random() < 0.6 else np.float64
    pos = pos.astype(dtype)
    sp = np.unique(rng.integers(0, n_species, n), return_inverse=True)[1]  # dense 0..S-1
    n_species = int(sp.max()) + 1
    zs = np.array([1, 8, 29])[:n_species]
    centres = np.sort(rng.choice(n, size=int(rng.integers(1, min(n, 12) + 1)), replace=False))
    label = (f"case {case}: n={n} nf={nf} r_cut={r_cut} n_max={n_max} l_max={l_max} S={n_species} "
             f"periodic={periodic} edge/R={np.round(edge / radius, 2)} dtype={dtype.__name__}")
    if len(sys.argv) > 3:
        print(label, flush=True)
    got = engine.soap_power_spectrum(
        torch.as_tensor(pos).cuda(), torch.as_tensor(sp, dtype=torch.int32).cuda(),
        torch.as_tensor(centres, dtype=torch.int32).cuda(), box if periodic else None, r_cut,
        n_max, l_max, n_species=n_species,
    ).cpu().numpy()
    want = so.soap_trajectory(pos, zs[sp], centres, box if periodic else None, r_cut, n_max, l_max)
    scale = np.abs(want).max(axis=-1, keepdims=True)
    err = float((np.abs(got - want) / np.maximum(scale, 1e-300)).max())
    worst = max(worst, err)
    if not (got.shape == want.shape and err < 1e-8):
        bad += 1
        print("MISMATCH", label, "rel err", err)
print(f"{n_cases} cases, {bad} mismatches, worst relative error {worst:.2e}")
sys.exit(1 if bad else 0)
