"""A minimal array-backed stand-in for ``MDAnalysis.Universe``.

The hot path reads only a small duck-typed surface from a Universe:
``select_atoms``, ``ag.positions`` / ``n_atoms`` / ``indices`` / ``atoms.types`` / ``ag[idx]`` /
``len(ag)``, ``universe.trajectory.n_frames``, ``universe.trajectory[i]`` (seek),
``universe.trajectory[slice]`` (iteration), ``universe.trajectory.ts.dimensions`` and
``universe.atoms.positions``.  :class:`ArrayUniverse` provides exactly that on top of a
``(F, N, 3)`` float32 array, so the drop-in API works without MDAnalysis installed; a real
``MDAnalysis.Universe`` can be passed to the same functions.

Selections understood: ``all``, ``none``, ``name A B``, ``type A B``, ``id 1 2`` / ``bynum``
(1-based), ``index 0 1`` (0-based), ``resname R``, ranges ``id 1:4`` / ``1-4``, ``not``, ``and``,
``or``, parentheses -- a superset of what the reference's hot-path tests use (``"all"``,
``"id 1"``, ``"name C3 or name C6"``, ``"not name C5"``, ``"type O"``).
"""

from __future__ import annotations

import re
from pathlib import Path
from typing import Iterator

import numpy as np

from dynsight_b200 import io as dsg_io


class Timestep:
    """Current-frame view (``universe.trajectory.ts``)."""

    def __init__(self, reader: ArrayReader, frame: int) -> None:
        self._reader = reader
        self.frame = frame

    @property
    def positions(self) -> np.ndarray:
        return self._reader.coordinates[self.frame]

    @property
    def dimensions(self) -> np.ndarray | None:
        d = self._reader.dimensions_array
        return None if d is None else d[self.frame]

    @dimensions.setter
    def dimensions(self, value: np.ndarray | None) -> None:
        # A reader without a native box (XYZ) does not reset ``ts.dimensions`` when it re-reads a
        # frame, so in MDAnalysis the *last* assigned box is what every later frame access sees
        # (relied upon by tests/lens/test_lens.py:27-33).  Mirror
        # that: assignment broadcasts to all frames; readers with a native per-frame box keep it
        # per frame.
        reader = self._reader
        if value is None:
            reader.dimensions_array = None
            return
        val = np.asarray(value, dtype=np.float32)
        if reader.dimensions_array is None or not reader.has_native_box:
            reader.dimensions_array = np.tile(val[None, :], (reader.n_frames, 1))
        else:
            reader.dimensions_array[self.frame] = val


class ArrayReader:
    """Trajectory reader over an in-memory ``(F, N, 3)`` float32 array."""

    def __init__(self, coordinates: np.ndarray, dimensions: np.ndarray | None, dt: float) -> None:
        self.coordinates = coordinates
        self.dimensions_array = dimensions
        self.has_native_box = dimensions is not None
        self.dt = dt
        self.ts = Timestep(self, 0)

    @property
    def n_frames(self) -> int:
        return int(self.coordinates.shape[0])

    def __len__(self) -> int:
        return self.n_frames

    def __getitem__(self, item: int | slice) -> Timestep | Iterator[Timestep]:
        if isinstance(item, slice):
            return self._iterate(range(*item.indices(self.n_frames)))
        idx = int(item)
        if idx < 0:
            idx += self.n_frames
        if not 0 <= idx < self.n_frames:
            msg = f"frame index {item} out of range for {self.n_frames} frames"
            raise IndexError(msg)
        self.ts = Timestep(self, idx)
        return self.ts

    def _iterate(self, frames: range) -> Iterator[Timestep]:
        for f in frames:
            self.ts = Timestep(self, f)
            yield self.ts

    def __iter__(self) -> Iterator[Timestep]:
        return self._iterate(range(self.n_frames))


class ArrayAtomGroup:
    """Subset of the atoms of an :class:`ArrayUniverse` (``ix`` are universe-wide indices)."""

    def __init__(self, universe: ArrayUniverse, ix: np.ndarray) -> None:
        self.universe = universe
        self.ix = np.asarray(ix, dtype=np.int64)

    # --- what the hot path reads -----------------------------------------------------------
    @property
    def positions(self) -> np.ndarray:
        return self.universe.trajectory.ts.positions[self.ix]

    @property
    def n_atoms(self) -> int:
        return int(self.ix.size)

    @property
    def indices(self) -> np.ndarray:
        return self.ix

    @property
    def atoms(self) -> ArrayAtomGroup:
        return self

    @property
    def types(self) -> np.ndarray:
        return self.universe.topology.types[self.ix]

    @property
    def names(self) -> np.ndarray:
        return self.universe.topology.names[self.ix]

    @property
    def ids(self) -> np.ndarray:
        return self.universe.topology.ids[self.ix]

    def __len__(self) -> int:
        return self.n_atoms

    def __getitem__(self, item: int | slice | np.ndarray) -> ArrayAtomGroup:
        sel = self.ix[item]
        return ArrayAtomGroup(self.universe, np.atleast_1d(sel))

    def __iter__(self) -> Iterator[ArrayAtomGroup]:
        for i in self.ix:
            yield ArrayAtomGroup(self.universe, np.array([i]))

    def select_atoms(self, selection: str) -> ArrayAtomGroup:
        if selection.strip() == "all":  # the common case, no mask arithmetic on 1e6 atoms
            return self
        mask = _Selector(self.universe).evaluate(selection)
        return ArrayAtomGroup(self.universe, self.ix[mask[self.ix]])

    def __repr__(self) -> str:
        return f"<ArrayAtomGroup with {self.n_atoms} atoms>"


class ArrayUniverse:
    """Array-backed Universe: ``positions`` (F, N, 3) float32 angstrom, optional per-frame box."""

    def __init__(
        self,
        positions: np.ndarray,
        dimensions: np.ndarray | None = None,
        names: np.ndarray | list[str] | None = None,
        types: np.ndarray | list[str] | None = None,
        ids: np.ndarray | None = None,
        resnames: np.ndarray | list[str] | None = None,
        dt: float = 1.0,
    ) -> None:
        pos = np.asarray(positions)
        if pos.ndim != 3 or pos.shape[2] != 3:  # noqa: PLR2004
            msg = f"positions must have shape (n_frames, n_atoms, 3), got {pos.shape}"
            raise ValueError(msg)
        if pos.dtype != np.float32:
            pos = pos.astype(np.float32)  # MDAnalysis stores coordinates as float32
        n_frames, n_atoms, _ = pos.shape
        dims = None
        if dimensions is not None:
            dims = np.asarray(dimensions, dtype=np.float32)
            if dims.ndim == 1:
                dims = np.tile(dims[None, :], (n_frames, 1))
            if dims.shape[1] == 3:  # noqa: PLR2004
                dims = np.concatenate(
                    [dims, np.full((n_frames, 3), 90.0, dtype=np.float32)], axis=1
                )
            if dims.shape != (n_frames, 6):
                msg = f"dimensions must have shape ({n_frames}, 6) or (6,), got {dims.shape}"
                raise ValueError(msg)
        names_a = np.asarray(names if names is not None else ["X"] * n_atoms).astype(str)
        types_a = (
            np.asarray(types).astype(str)
            if types is not None
            else np.asarray([dsg_io.guess_type_from_name(n) for n in names_a])
        )
        self.topology = dsg_io.Topology(
            names=names_a,
            types=types_a,
            ids=np.asarray(ids if ids is not None else np.arange(1, n_atoms + 1), dtype=np.int64),
            resnames=np.asarray(resnames if resnames is not None else [""] * n_atoms).astype(str),
        )
        self.trajectory = ArrayReader(np.ascontiguousarray(pos), dims, dt)
        self.atoms = ArrayAtomGroup(self, np.arange(n_atoms))

    def select_atoms(self, selection: str) -> ArrayAtomGroup:
        return self.atoms.select_atoms(selection)

    # --- constructors mirroring MDAnalysis.Universe(topology, trajectory) ---------------------
    @classmethod
    def from_gro_xtc(cls, topo_file: str | Path, traj_file: str | Path) -> ArrayUniverse:
        topo, _, _ = dsg_io.read_gro(topo_file)
        pos, dims = dsg_io.read_xtc(traj_file)  # native reader: compressed frames included
        # MDAnalysis reports ``ts.dimensions = None`` for a frame whose box is all zero, which sends
        # LENS to its bounding-box rule (lens.py:271-274) and SOAP / the spatial average to their
        # box-less branches; an array-backed universe has one answer for the whole file
        boxless = ~np.any(dims[:, :3] > 0.0, axis=1)
        if boxless.all():
            dims = None
        elif boxless.any():
            msg = (f"{traj_file}: {int(boxless.sum())} of {len(boxless)} frames have an all-zero "
                   "box; mixed boxed / box-less trajectories are not supported")
            raise ValueError(msg)
        return cls(pos, dims, topo.names, topo.types, topo.ids, topo.resnames)

    @classmethod
    def from_xyz(cls, traj_file: str | Path, dt: float = 1.0) -> ArrayUniverse:
        topo, pos = dsg_io.read_xyz(traj_file)
        return cls(pos, None, topo.names, topo.types, topo.ids, topo.resnames, dt=dt)


# ---------------------------------------------------------------------------------------------
# selection language
# ---------------------------------------------------------------------------------------------
_TOKEN = re.compile(r"\(|\)|[^\s()]+")
_KEYWORDS = {"name", "type", "id", "bynum", "index", "resname", "all", "none", "not", "and", "or"}


class _Selector:
    def __init__(self, universe: ArrayUniverse) -> None:
        self.topo = universe.topology
        self.n = len(self.topo.names)
        self.tokens: list[str] = []
        self.pos = 0

    def evaluate(self, selection: str) -> np.ndarray:
        self.tokens = _TOKEN.findall(selection)
        self.pos = 0
        if not self.tokens:
            msg = "empty selection string"
            raise ValueError(msg)
        mask = self._or()
        if self.pos != len(self.tokens):
            msg = f"cannot parse selection {selection!r} near {self.tokens[self.pos]!r}"
            raise ValueError(msg)
        return mask

    def _peek(self) -> str | None:
        return self.tokens[self.pos] if self.pos < len(self.tokens) else None

    def _or(self) -> np.ndarray:
        mask = self._and()
        while self._peek() == "or":
            self.pos += 1
            mask = mask | self._and()
        return mask

    def _and(self) -> np.ndarray:
        mask = self._not()
        while self._peek() == "and":
            self.pos += 1
            mask = mask & self._not()
        return mask

    def _not(self) -> np.ndarray:
        if self._peek() == "not":
            self.pos += 1
            return ~self._not()
        return self._atom()

    def _values(self) -> list[str]:
        vals = []
        while (tok := self._peek()) is not None and tok not in _KEYWORDS and tok not in "()":
            vals.append(tok)
            self.pos += 1
        if not vals:
            msg = "selection keyword needs at least one value"
            raise ValueError(msg)
        return vals

    def _int_mask(self, field: np.ndarray, vals: list[str]) -> np.ndarray:
        mask = np.zeros(self.n, dtype=bool)
        for v in vals:
            m = re.fullmatch(r"(-?\d+)[:-](-?\d+)", v)
            if m:
                mask |= (field >= int(m.group(1))) & (field <= int(m.group(2)))
            else:
                mask |= field == int(v)
        return mask

    def _atom(self) -> np.ndarray:
        tok = self._peek()
        if tok is None:
            msg = "unexpected end of selection"
            raise ValueError(msg)
        self.pos += 1
        if tok == "(":
            mask = self._or()
            if self._peek() != ")":
                msg = "missing ')' in selection"
                raise ValueError(msg)
            self.pos += 1
            return mask
        if tok == "all":
            return np.ones(self.n, dtype=bool)
        if tok == "none":
            return np.zeros(self.n, dtype=bool)
        if tok == "name":
            return np.isin(self.topo.names, self._values())
        if tok == "type":
            return np.isin(self.topo.types, self._values())
        if tok == "resname":
            return np.isin(self.topo.resnames, self._values())
        if tok in ("id", "bynum"):
            return self._int_mask(self.topo.ids, self._values())
        if tok == "index":
            return self._int_mask(np.arange(self.n), self._values())
        msg = f"unknown selection keyword {tok!r}"
        raise ValueError(msg)
