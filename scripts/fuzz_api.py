"""Randomised differential test of the Python API layer (selections, trajslice, delays, frame
batching with overlap, double-buffered ingest, box-less trajectories) against the oracle.

usage: python scripts/fuzz_api.py [n_cases] [seed]
"""
import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dynsight_b200 import lens as dlens, soap as dsoap
from dynsight_b200.universe import ArrayUniverse
from oracle import lens_oracle as lo, soap_oracle as so, timesoap_oracle as to

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad = 0
for case in range(n_cases):
    n, nf = int(rng.choice([5, 40, 300])), int(rng.integers(4, 14))
    r_cut = float(rng.choice([1.2, 2.0]))
    box = (rng.integers(3, 9, size=3) * r_cut * rng.uniform(1.0, 1.3, size=3)).astype(np.float32)
    pos = ((rng.random((nf, n, 3)) * 1.4 - 0.2) * box).astype(np.float32)
    with_box = rng.random() < 0.8
    dims = np.tile(np.concatenate([box, [90, 90, 90]]).astype(np.float32), (nf, 1)) if with_box else None
    names = rng.choice(["A", "B"], size=n)
    names[:2] = ["A", "B"]
    uni = ArrayUniverse(pos, dims, names, names)
    sel = str(rng.choice(["all", "name A", "name B"]))
    cen = str(rng.choice(["all", "name A", sel]))
    env_idx = np.arange(n) if sel == "all" else np.where(names == sel[-1])[0]
    cent_idx = np.arange(n) if cen == "all" else np.where(names == cen[-1])[0]
    start, step = int(rng.integers(0, 3)), int(rng.integers(1, 3))
    sl = slice(start, None, step)
    fr = list(range(nf))[sl]
    delay = int(rng.integers(1, 4))
    pbc = bool(rng.random() < 0.8)
    bf = int(rng.integers(delay + 1, delay + 4))
    if len(fr) - delay < 1:
        continue
    got = dlens.compute_lens_device(uni, r_cut, delay, cen, sel, sl, pbc, batch_frames=bf).cpu().numpy()
    want = lo.compute_lens_arrays(pos, dims, r_cut, delay, cent_idx, env_idx, fr, pbc)
    if not np.array_equal(got, want):
        bad += 1
        print(f"MISMATCH lens case {case}: n={n} nf={nf} sel={sel} cen={cen} slice={sl} delay={delay} "
              f"pbc={pbc} batch={bf} box={with_box}")
    if with_box and case % 4 == 0:   # streamed timeSOAP with small batches vs oracle SOAP + timeSOAP
        sub = pos * 3.0
        uni3 = ArrayUniverse(sub, dims * np.array([3, 3, 3, 1, 1, 1], np.float32), ["O"] * n, ["O"] * n)
        rc = 3.0 * r_cut
        got_ts = dsoap.timesoap_from_trajectory(uni3, rc, 4, 3, delay, trajslice=sl, batch_frames=bf).cpu().numpy()
        spectra = so.soap_trajectory(sub[fr], np.full(n, 8), np.arange(n), 3.0 * dims[fr, :3], rc, 4, 3)
        want_ts = to.timesoap(spectra, delay)
        if not np.all(np.abs(got_ts - want_ts) <= 1e-5 * want_ts + 1e-7):
            bad += 1
            print(f"MISMATCH tsoap case {case}: max diff {np.abs(got_ts - want_ts).max():.2e}")
print(f"{n_cases} cases, {bad} mismatches")
sys.exit(1 if bad else 0)
