"""CPU model of the pair path of ``dsg_lens_direct`` (csrc/neighbors.cu:lens_pair_kernel), checked
against the oracle: half of the 27 reference cells per atom, every hit credited to both atoms, a
hit of frame t looked up directly in frame t + delay (plus the cell-adjacency test when that frame
has atoms outside [0, box)).  Validates the ALGORITHM (symmetry, half-shell enumeration under the
periodic wrap, the adjacency rule); the CUDA code is validated on the GPU by tests/test_gpu_lens.py.

    python scripts/pair_path_model.py [n_cases]
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import lens_oracle  # noqa: E402


def py_floordiv(a, b):
    return np.floor_divide(a, b)


def cells_of(x, box, ncell):
    v = x / box * ncell
    t = np.trunc(v).astype(np.int64)
    return np.mod(t, ncell), (t >= 0) & (t < ncell) & (x >= 0)


def dr2(pi, pj, box, pbc):
    d = pj.astype(np.float64) - pi.astype(np.float64)
    if pbc:
        d = d - np.rint(d / box) * box
    return float(np.dot(d, d))


def pair_model(pos, box, r_cut, delay, pbc):
    n_frames, n, _ = pos.shape
    r2 = (r_cut - 1e-6) ** 2
    box = np.asarray(box, dtype=np.float64)
    ncell = np.maximum(3, py_floordiv(box, r_cut).astype(np.int64))
    cnt = np.zeros((n_frames, n), dtype=np.int64)
    inter = np.zeros((n_frames, n), dtype=np.int64)
    cell = np.zeros((n_frames, n, 3), dtype=np.int64)
    regular = np.ones(n_frames, dtype=bool)
    for f in range(n_frames):
        for k in range(3):
            c, ok = cells_of(pos[f, :, k].astype(np.float64), box[k], ncell[k])
            cell[f, :, k] = c
            regular[f] &= bool(ok.all())
    half = [(0, 0, 1), (0, 1, -1), (0, 1, 0), (0, 1, 1)] + [(1, b, c) for b in (-1, 0, 1) for c in (-1, 0, 1)]
    for f in range(n_frames):
        # cell-sorted slots
        key = (cell[f, :, 0] * ncell[1] + cell[f, :, 1]) * ncell[2] + cell[f, :, 2]
        order = np.argsort(key, kind="stable")
        members = {}
        for s, i in enumerate(order):
            members.setdefault(int(key[i]), []).append((s, int(i)))
        f2 = f + delay
        paired = delay > 0 and f2 < n_frames
        for s, i in enumerate(order):
            ci = cell[f, i]
            cand = [j for (sj, j) in members[int(key[i])] if sj > s]
            for off in half:
                cc = (ci + np.asarray(off)) % ncell
                cand += [j for (_, j) in members.get(int((cc[0] * ncell[1] + cc[1]) * ncell[2] + cc[2]), [])]
            for j in cand:
                if dr2(pos[f, i], pos[f, j], box, pbc) < r2:
                    cnt[f, i] += 1
                    cnt[f, j] += 1
                    if paired:
                        hit2 = dr2(pos[f2, i], pos[f2, j], box, pbc) < r2
                        if hit2 and not regular[f2]:
                            e = np.abs(cell[f2, i] - cell[f2, j])
                            hit2 = bool(np.all((e <= 1) | (e == ncell - 1)))
                        inter[f, i] += hit2
                        inter[f, j] += hit2
    n_pairs = n_frames - delay
    lens = np.zeros((n, n_pairs))
    for t in range(n_pairs):
        a, b = cnt[t], cnt[t + delay]
        den = a + b
        lens[:, t] = np.where(den > 0, (den - 2 * inter[t]) / np.maximum(den, 1), 0.0)
    return lens, cnt.T.astype(np.float64)


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rng = np.random.default_rng(5)
    for case in range(n_cases):
        n = int(rng.integers(30, 160))
        n_frames = int(rng.integers(2, 5))
        delay = int(rng.integers(1, n_frames))
        pbc = bool(rng.integers(0, 2))
        r_cut = float(rng.uniform(0.8, 1.6))
        ncx = rng.integers(3 if not pbc else 6, 9, size=3)
        box = (ncx * r_cut * rng.uniform(1.0, 1.15, size=3)).astype(np.float32)
        spread = rng.choice([1.0, 1.0, 1.4, 3.0])   # > 1: atoms outside the box
        pos = ((rng.random((n_frames, n, 3)) - 0.5) * spread + 0.5) * box
        pos[1:] = pos[0] + rng.normal(0, 0.25 * r_cut, (n_frames - 1, n, 3))
        pos = pos.astype(np.float32)
        if spread == 1.0:
            pos = np.clip(pos, 0, np.nextafter(box, 0)).astype(np.float32)
        dims = np.tile(np.concatenate([box, [90, 90, 90]]).astype(np.float32), (n_frames, 1))
        want = lens_oracle.compute_lens_arrays(pos, dims, r_cut, delay=delay, respect_pbc=pbc)
        want_c = lens_oracle.coord_number_arrays(pos, dims, r_cut, respect_pbc=pbc)
        got, got_c = pair_model(pos, box.astype(np.float64), r_cut, delay, pbc)
        assert np.array_equal(got_c, want_c), f"case {case}: counts differ"
        assert np.array_equal(got, want), f"case {case}: LENS differs"
        print(f"case {case}: n={n} F={n_frames} delay={delay} pbc={pbc} spread={spread} "
              f"ncell={list(map(int, np.maximum(3, box.astype(np.float64) // r_cut)))} ok")
    print("pair path model == oracle on", n_cases, "cases")


if __name__ == "__main__":
    main()
