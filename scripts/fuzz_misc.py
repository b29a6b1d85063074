"""Randomised differential tests of timeSOAP, the self-TCF and the spatial average vs the oracles.

usage: python scripts/fuzz_misc.py [n_cases] [seed]
"""
import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dynsight_b200 import analysis, engine
from dynsight_b200.universe import ArrayUniverse
from oracle import spatial_oracle, tcf_oracle, timesoap_oracle

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad = 0
for case in range(n_cases):
    # ---- timeSOAP ----
    n, f, c = int(rng.integers(1, 40)), int(rng.integers(2, 30)), int(rng.choice([1, 5, 32, 324, 400, 600]))
    delay = int(rng.integers(1, min(7, f)))
    soap = rng.normal(size=(n, f, c)) * rng.choice([1e-6, 1.0, 1e6])
    soap[rng.random((n, f)) < 0.1] = 0.0
    if rng.random() < 0.3:
        soap[:, 1:] = soap[:, :-1] * (1 + 1e-9)   # nearly identical consecutive vectors
    got = engine.timesoap_device(torch.as_tensor(soap).cuda(), delay).cpu().numpy()
    want = timesoap_oracle.timesoap(soap, delay)
    if not (got.shape == want.shape and np.all(np.abs(got - want) <= 1e-5 * want + 1e-7)):
        bad += 1
        print(f"MISMATCH timesoap case {case}: n={n} f={f} c={c} delay={delay} "
              f"max diff {np.abs(got - want).max():.2e}")
    # ---- self time-correlation ----
    n, f = int(rng.integers(1, 60)), int(rng.integers(2, 400))
    md = None if rng.random() < 0.4 else int(rng.integers(1, f + 1))
    data = np.cumsum(rng.normal(size=(n, f)), axis=1) * rng.choice([1e-3, 1.0, 1e3]) + rng.normal()
    gc, ge = analysis.self_time_correlation(data, md)
    wc, we = tcf_oracle.self_time_correlation(data, md)
    if not (np.allclose(gc, wc, rtol=1e-9, atol=1e-10) and np.allclose(ge, we, rtol=1e-8, atol=1e-10)):
        bad += 1
        print(f"MISMATCH tcf case {case}: n={n} f={f} max_delay={md} "
              f"{np.abs(gc - wc).max():.2e} {np.abs(ge - we).max():.2e}")
    # ---- spatial average ----
    n, f = int(rng.choice([3, 40, 500])), int(rng.integers(1, 4))
    r_cut = float(rng.choice([1.0, 2.5]))
    box = (rng.integers(1, 12, size=3) * r_cut * rng.uniform(1.0, 1.4, size=3)).astype(np.float32)
    pos = ((rng.random((f, n, 3)) * 1.6 - 0.3) * box).astype(np.float32)
    dims = np.tile(np.concatenate([box, [90, 90, 90]]).astype(np.float32), (f, 1))
    comp = int(rng.choice([0, 1, 3, 130]))
    values = rng.normal(size=(n, f) if comp == 0 else (n, f, comp))
    uni = ArrayUniverse(pos, dims, ["A"] * n, ["A"] * n)
    got = analysis.spatialaverage(uni, values, "all", r_cut)
    want = spatial_oracle.spatial_average(pos, dims, values, r_cut)
    if not np.abs(got - want).max() <= 1e-12 * max(1.0, np.abs(want).max()):
        bad += 1
        print(f"MISMATCH spatial case {case}: n={n} f={f} r_cut={r_cut} box={box} comp={comp} "
              f"{np.abs(got - want).max():.2e}")
print(f"{n_cases} cases x 3 operators, {bad} mismatches")
sys.exit(1 if bad else 0)
