"""Test utility: a small XTC *writer* for synthetic integer coordinates.

It produces valid compressed XTC frames (the published GROMACS xdr3dfcoord layout) so that the
native reader (``dynsight_b200/csrc/xtc.cpp``) can be checked on machines where the reference's
test files do not exist.  It is not a general encoder: atoms come in groups of three
("water molecules": one full-range triple followed by a run of two small offsets, first two
atoms swapped, as GROMACS does), optionally with single full-range atoms in between, and the
small-offset table index is adapted by the caller-supplied schedule.
"""

from __future__ import annotations

import struct

import numpy as np

MAGICINTS = [
    0, 0, 0, 0, 0, 0, 0, 0, 0, 8, 10, 12, 16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256,
    322, 406, 512, 645, 812, 1024, 1290, 1625, 2048, 2580, 3250, 4096, 5060, 6501, 8192, 10321,
    13003, 16384, 20642, 26007, 32768, 41285, 52015, 65536, 82570, 104031, 131072, 165140, 208063,
    262144, 330280, 416127, 524287, 660561, 832255, 1048576, 1321122, 1664510, 2097152, 2642245,
    3329021, 4194304, 5284491, 6658042, 8388607, 10568983, 13316085, 16777216,
]  # fmt: skip
FIRSTIDX = 9


class BitWriter:
    def __init__(self) -> None:
        self.bits: list[int] = []

    def put(self, value: int, nbits: int) -> None:
        for b in range(nbits - 1, -1, -1):
            self.bits.append((value >> b) & 1)

    def put_triple(self, vals: tuple[int, int, int], sizes: tuple[int, int, int], nbits: int
                   ) -> None:  # fmt: skip
        big = (vals[0] * sizes[1] + vals[1]) * sizes[2] + vals[2]
        while nbits > 8:  # low byte first, the last (partial) byte carries the remaining bits
            self.put(big & 0xFF, 8)
            big >>= 8
            nbits -= 8
        if nbits > 0:
            self.put(big, nbits)

    def tobytes(self) -> bytes:
        bits = self.bits + [0] * (-len(self.bits) % 8)
        return bytes(
            sum(bit << (7 - k) for k, bit in enumerate(bits[i : i + 8]))
            for i in range(0, len(bits), 8)
        )


def bits_for_product(sizes: tuple[int, int, int]) -> int:
    prod = sizes[0] * sizes[1] * sizes[2]
    nbytes = 0
    while prod >> (8 * (nbytes + 1)):
        nbytes += 1
    top = prod >> (8 * nbytes)
    bits, num = 0, 1
    while top >= num:
        bits += 1
        num *= 2
    return bits + 8 * nbytes


def write_frame(coords: np.ndarray, box_nm: np.ndarray, step: int, time: float, precision: float,
                smallidx: int, layout: list[tuple[str, int]]) -> bytes:  # fmt: skip
    """One compressed frame.  ``coords`` (N, 3) int; ``layout`` lists ('single', 0) or
    ('water', adapt) items in atom order, adapt in {-1, 0, 1} changing smallidx after the group."""
    n = coords.shape[0]
    assert n > 9  # noqa: PLR2004
    minint = coords.min(axis=0)
    maxint = coords.max(axis=0)
    sizes = tuple(int(v) for v in (maxint - minint + 1))
    wide = max(sizes) > 0xFFFFFF  # noqa: PLR2004  -> three fixed-width fields per full triple
    bitsize = 0 if wide else bits_for_product(sizes)
    field_bits = [int(v).bit_length() for v in sizes]  # bits for one value in [0, size]
    bw = BitWriter()

    def put_full(vals: tuple[int, int, int]) -> None:
        if wide:
            for v, nb in zip(vals, field_bits):
                bw.put(v, nb)
        else:
            bw.put_triple(vals, sizes, bitsize)

    idx = smallidx
    a = 0
    run = 0  # the decoder keeps the last run length until a flag bit changes it
    for kind, adapt in layout:
        if kind == "single":
            put_full(tuple(int(v) for v in coords[a] - minint))
            if run == 0:
                bw.put(0, 1)  # no run information: the previous run length (0) stays
            else:
                bw.put(1, 1)
                bw.put(0 + 0 + 1, 5)  # run = 0, is_smaller = 0
                run = 0
            a += 1
            continue
        # water: file order is (second atom as the full triple), then offsets of the first atom
        # from the second and of the third atom from the first
        small = MAGICINTS[idx]
        smallnum = small // 2
        p0, p1, p2 = (coords[a + k].astype(np.int64) for k in range(3))
        put_full(tuple(int(v) for v in p1 - minint))
        if run == 6 and adapt == 0:  # noqa: PLR2004
            bw.put(0, 1)  # same run length as the previous group, no adaptation
        else:
            bw.put(1, 1)
            bw.put(6 + adapt + 1, 5)  # run = 6 (two small atoms), is_smaller = adapt
        run = 6
        for delta in (p0 - p1 + smallnum, p2 - p0 + smallnum):
            assert delta.min() >= 0 and delta.max() < small, "offset outside the small table"
            bw.put_triple(tuple(int(v) for v in delta), (small, small, small), idx)
        idx += adapt
        a += 3
    assert a == n
    payload = bw.tobytes()
    head = struct.pack(">iiif", 1995, n, step, time)
    head += struct.pack(">9f", *np.asarray(box_nm, dtype=np.float32).reshape(9))
    head += struct.pack(">i", n)
    head += struct.pack(">f", precision)
    head += struct.pack(">3i", *[int(v) for v in minint])
    head += struct.pack(">3i", *[int(v) for v in maxint])
    head += struct.pack(">i", smallidx)
    head += struct.pack(">i", len(payload))
    return head + payload + b"\0" * (-len(payload) % 4)


def write_synthetic_waters(path, n_mol: int, n_frames: int, seed: int = 0) -> int:
    """A compressed XTC of ``n_mol`` water-like triples x ``n_frames`` frames for throughput
    measurements (bench.py): one frame is encoded and repeated with increasing step numbers (the
    decoder does the same work for every frame).  Returns the file size in bytes."""
    rng = np.random.default_rng(seed)
    smallidx = 21
    half = MAGICINTS[smallidx] // 2
    p0 = rng.integers(0, 40_000, size=(n_mol, 3))
    p1 = p0 + rng.integers(-half + 1, half - 1, size=(n_mol, 3)) // 2
    p2 = p0 + rng.integers(-half + 1, half - 1, size=(n_mol, 3)) // 2
    coords = np.stack([p0, p1, p2], axis=1).reshape(-1, 3).astype(np.int64)
    layout = [("water", 0)] * n_mol
    box = np.diag([40.0, 40.0, 40.0]).astype(np.float32)
    frame = write_frame(coords, box, 0, 0.0, 1000.0, smallidx, layout)
    with open(path, "wb") as fh:
        for f in range(n_frames):
            fh.write(frame[:8] + struct.pack(">if", f, 0.5 * f) + frame[16:])
    return len(frame) * n_frames
