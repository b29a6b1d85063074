"""Randomised differential test of the neighbour / LENS paths against the CPU oracle (bit-exact).

usage: python scripts/fuzz_neighbors.py [n_cases] [seed]
Draws boxes (cubic / slab / tiny, with 3 .. 40 cells per axis), particle counts 1 .. 4000, cutoffs,
pbc on/off, explicit centres, atoms far outside the box, coincident atoms and lattice positions
(many exact ties in distance), float32 and float64 inputs, and compares the canonical CSR (both
protocols), the unsorted rows and LENS with the oracle.
"""
import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dynsight_b200 import engine
from oracle import lens_oracle as lo

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad = 0
for case in range(n_cases):
    nf = int(rng.integers(2, 5))
    n = int(rng.choice([1, 2, 7, 50, 300, 1500, 4000]))
    r_cut = float(rng.choice([0.7, 1.0, 1.5, 2.5, 6.0]))
    cells = rng.integers(1, 41, size=3)
    box = (cells * r_cut * rng.uniform(1.0, 1.3, size=3)).astype(np.float32)
    if rng.random() < 0.2:
        box[:] = box[0]
    box_f = np.tile(box, (nf, 1)) * rng.uniform(0.97, 1.03, size=(nf, 1)).astype(np.float32)
    kind = rng.choice(["uniform", "outside", "lattice", "clustered"])
    if kind == "lattice":   # exact distance ties, many at the cutoff for r_cut 1.0 / 1.5
        g = np.stack(np.meshgrid(*[np.arange(int(np.ceil(n ** (1 / 3))) + 1)] * 3, indexing="ij"),
                     -1).reshape(-1, 3)[:n]
        pos = np.tile(g[None].astype(np.float64) * 0.5, (nf, 1, 1))
    elif kind == "clustered":
        centres = rng.random((nf, max(1, n // 40), 3)) * box
        pos = centres[:, rng.integers(0, centres.shape[1], n)] + rng.normal(0, 0.3, (nf, n, 3))
    else:
        spread = 1.0 if kind == "uniform" else 4.0
        pos = (rng.random((nf, n, 3)) * spread - (spread - 1) / 2) * box
    if rng.random() < 0.3 and n > 3:   # coincident atoms
        pos[:, 1] = pos[:, 0]
    dtype = np.float32 if rng.random() < 0.7 else np.float64
    pos = pos.astype(dtype)
    pbc = bool(rng.random() < 0.75)
    explicit = rng.random() < 0.3 and n > 2
    pc = np.ascontiguousarray(pos[:, rng.choice(n, size=max(1, n // 3), replace=False)]) if explicit else None
    dev = lambda a: torch.as_tensor(a).cuda()
    if len(sys.argv) > 3:  # verbose: say what is about to run (to find a slow or stuck case)
        print(f"case {case}: n={n} nf={nf} r_cut={r_cut} box={box} kind={kind} "
              f"dtype={dtype.__name__} pbc={pbc} explicit={explicit}", flush=True)
    try:
        one = engine.neighbor_lists(dev(pos), box_f, r_cut, None if pc is None else dev(pc), pbc)
        two = engine.neighbor_lists(dev(pos), box_f, r_cut, None if pc is None else dev(pc), pbc,
                                    two_pass=True)
        for f in range(nf):
            ip, ix = lo.neighbors(pos[f], pos[f] if pc is None else pc[f], r_cut, box_f[f], pbc)
            for nb in (one, two):
                gi, gx = nb.frame_csr(f)
                assert np.array_equal(gi.cpu().numpy(), ip) and np.array_equal(gx.cpu().numpy(), ix)
        lens, nr = engine.lens_from_positions(dev(pos), box_f, r_cut, 1,
                                              None if pc is None else dev(pc), pbc)
        want = np.stack([
            lo.lens_from_two_csr(*lo.neighbors(pos[f], pos[f] if pc is None else pc[f], r_cut,
                                               box_f[f], pbc),
                                 *lo.neighbors(pos[f + 1], pos[f + 1] if pc is None else pc[f + 1],
                                               r_cut, box_f[f + 1], pbc))
            for f in range(nf - 1)], axis=1)
        assert np.array_equal(lens.cpu().numpy(), want)
    except AssertionError:
        bad += 1
        print(f"MISMATCH case {case}: n={n} nf={nf} r_cut={r_cut} box={box} kind={kind} "
              f"dtype={dtype.__name__} pbc={pbc} explicit={explicit}")
print(f"{n_cases} cases, {bad} mismatches")
sys.exit(1 if bad else 0)
