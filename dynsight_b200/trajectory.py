"""trajectory package -- drop-in for the hot-path part of ``dynsight.trajectory``.

``Trj`` mirrors ``dynsight._internal.trajectory.trajectory.Trj`` (trajectory.py:32-320: the
constructors, ``get_coordinates``, ``with_slice`` and the four descriptor entry points
``get_coord_number`` / ``get_lens`` / ``get_soap`` / ``get_timesoap``, plus the neighbour-list
consumers ``get_orientational_op`` / ``get_velocity_alignment``, :322-423) and ``Insight`` mirrors
``dynsight._internal.trajectory.insight.Insight`` (insight.py:23-92, 136-155: container,
JSON+NPY dump/load, ``get_angular_velocity``).  Meta dictionaries, return types and error
messages are the reference's.  Everything else in those classes (tICA, onion clustering, RDF,
spatial average, ...) is outside the hot path and not provided.
"""

from __future__ import annotations

import json
from dataclasses import dataclass, field
from pathlib import Path
from typing import Any, Literal

import numpy as np

from dynsight_b200 import descriptors as _descriptors
from dynsight_b200 import lens as _lens
from dynsight_b200 import soap as _soap
from dynsight_b200.logs import logger
from dynsight_b200.universe import ArrayUniverse

UNIVAR_DIM = 2

__all__ = ["Insight", "Trj"]


@dataclass(frozen=True)
class Insight:
    """Contains an analysis performed on a trajectory.

    Attributes:
        dataset: The values of some trajectory descriptor (float64 ndarray).
        meta: A dictionary containing the relevant parameters.
    """

    dataset: np.ndarray
    meta: dict[str, Any] = field(default_factory=dict)

    def __post_init__(self) -> None:
        if logger.auto_recording:
            logger.record_data(self)

    def dump_to_json(self, file_path: Path) -> None:
        """Save the Insight to a JSON file and a .npy file (insight.py:39-53)."""
        npy_path = file_path.with_suffix(".npy")
        np.save(npy_path, self.dataset.astype(np.float64))
        json_data = {"dataset_file": npy_path.name, "meta": self.meta}
        with file_path.open("w") as file:
            json.dump(json_data, file, indent=4)
        logger.log(f"Insight saved to {file_path} and dataset to {npy_path}.")

    @classmethod
    def load_from_json(
        cls,
        file_path: Path,
        mmap_mode: Literal["r", "r+", "w+", "c"] | None = None,
    ) -> Insight:
        """Load the Insight object from a .json file (insight.py:55-92).

        Raises:
            ValueError: if required keys are missing.
        """
        with file_path.open("r") as file:
            data = json.load(file)
        dataset_file = data.get("dataset_file")
        if not dataset_file:
            msg = "'dataset_file' key not found in JSON file."
            logger.log(msg)
            raise ValueError(msg)
        dataset_path = file_path.with_name(dataset_file)
        dataset = np.load(dataset_path, mmap_mode=mmap_mode)
        logger.log(f"Insight loaded from {file_path}, dataset from {dataset_path}.")
        return cls(dataset=dataset, meta=data.get("meta", {}))

    def spatial_average(
        self,
        trj: Trj,
        r_cut: float,
        selection: str = "all",
        n_jobs: int = 1,
    ) -> Insight:
        """Average the descriptor over the neighboring particles (insight.py:94-118).

        The returned Insight contains the following meta: sp_av_r_cut, selection.
        """
        from dynsight_b200 import analysis

        averaged_dataset = analysis.spatialaverage(
            universe=trj.universe,
            descriptor_array=self.dataset,
            selection=selection,
            r_cut=r_cut,
            trajslice=trj.trajslice,
            n_jobs=n_jobs,
        )
        attr_dict = {"sp_av_r_cut": r_cut, "selection": selection}
        logger.log(f"Computed spatial average with args {attr_dict}.")
        return Insight(dataset=averaged_dataset, meta=attr_dict)

    def get_time_correlation(
        self,
        max_delay: int | None = None,
    ) -> tuple[np.ndarray, np.ndarray]:
        """Self time correlation function of the time-series signal (insight.py:120-134)."""
        from dynsight_b200 import analysis

        attr_dict = {"max_delay": max_delay}
        logger.log(f"Computed time corrleation function with args {attr_dict}.")
        return analysis.self_time_correlation(self.dataset, max_delay)

    def get_angular_velocity(self, delay: int = 1) -> Insight:
        """Computes the angular displacement of a vectorial descriptor (insight.py:136-155).

        Raises:
            ValueError if the dataset does not have the right dimensions.
        """
        if self.dataset.ndim != UNIVAR_DIM + 1:
            msg = "dataset.ndim != 3."
            logger.log(msg)
            raise ValueError(msg)
        theta = _soap.timesoap(self.dataset, delay=delay)
        attr_dict = self.meta.copy()
        attr_dict.update({"delay": delay})
        if self.meta.get("name") == "soap":
            attr_dict.update({"name": "timesoap"})
        else:
            attr_dict.update({"name": "angular_velocity"})
        logger.log(f"Computed angular velocity with args {attr_dict}.")
        return Insight(dataset=theta, meta=attr_dict)


@dataclass(frozen=True)
class Trj:
    """Contains a trajectory.

    Attributes:
        universe: an ``MDAnalysis.Universe`` or any object with the same duck-typed surface
            (e.g. :class:`dynsight_b200.universe.ArrayUniverse`).
    """

    universe: Any = field()
    trajslice: slice | None = None
    n_atoms: int = field(init=False)
    n_frames: int = field(init=False)

    def __post_init__(self) -> None:
        n_atoms = len(self.universe.atoms)
        if self.trajslice is None:
            n_frames = len(self.universe.trajectory)
        else:
            n_frames = len(range(*self.trajslice.indices(len(self.universe.trajectory))))
        object.__setattr__(self, "n_atoms", n_atoms)
        object.__setattr__(self, "n_frames", n_frames)

    @classmethod
    def init_from_universe(cls, universe: Any) -> Trj:
        """Initialize Trj object from a Universe."""
        logger.log("Created Trj from MDAnalysis.Universe.")
        return Trj(universe)

    @classmethod
    def init_from_xyz(cls, traj_file: Path, dt: float) -> Trj:
        """Initialize Trj object from an .xyz file (``dt`` is the trajectory's time-step)."""
        logger.log(f"Created Trj from {traj_file} with dt = {dt}.")
        return Trj(ArrayUniverse.from_xyz(traj_file, dt=dt))

    @classmethod
    def init_from_xtc(cls, traj_file: Path, topo_file: Path) -> Trj:
        """Initialize Trj object from .gro and .xtc files (native reader, compressed XTC included)."""
        logger.log(f"Created Trj from {traj_file}, {topo_file}.")
        return Trj(ArrayUniverse.from_gro_xtc(topo_file, traj_file))

    def get_coordinates(self, selection: str) -> np.ndarray:
        """Returns the coordinates as an array of shape (n_frames, n_atoms, n_coordinates)."""
        atoms = self.universe.select_atoms(selection)
        trajslice = slice(None) if self.trajslice is None else self.trajslice
        attr_dict = {"selection": selection}
        logger.log(f"Extracted coordinates array with args {attr_dict}.")
        return np.array([atoms.positions.copy() for _ in self.universe.trajectory[trajslice]])

    def with_slice(self, trajslice: slice | None) -> Trj:
        """Returns a Trj with a different frames' slice."""
        attr_dict = {"trajslice": trajslice}
        logger.log(f"Created a sliced Trj with args {attr_dict}.")
        return Trj(self.universe, trajslice=trajslice)

    def get_slice(self, start: int, stop: int, step: int) -> Trj:
        """Returns a Trj with a subset of frames (trajectory.py:115-137; deprecated upstream in
        favour of ``with_slice``).  Like the reference's ``MemoryReader`` copy, the new trajectory
        keeps the coordinates only: it has no simulation box."""
        coords = np.array(
            [ts.positions.copy() for ts in self.universe.trajectory[start:stop:step]],
            dtype=np.float32,
        ).reshape(-1, self.universe.atoms.n_atoms, 3)
        atoms = self.universe.atoms
        u_new = ArrayUniverse(
            coords, None, names=getattr(atoms, "names", None), types=getattr(atoms, "types", None),
            ids=getattr(atoms, "ids", None), resnames=getattr(atoms, "resnames", None),
        )  # fmt: skip
        attr_dict = {"start": start, "stop": stop, "step": step}
        logger.log(f"Created a sliced Trj with args {attr_dict}.")
        return Trj(u_new)

    def get_coord_number(
        self,
        r_cut: float,
        centers: str = "all",
        selection: str = "all",
        respect_pbc: bool = True,
        neigcounts: Any | None = None,
        n_jobs: int = 1,
    ) -> tuple[Any, Insight]:
        """Compute coordination number on the trajectory.

        Returns:
            tuple:
                * neighcounts: a list[list[AtomGroup]]-like sequence, it can be used to speed up
                    subsequent descriptors' computations.
                * An Insight containing the number of neighbors. It has the following meta:
                    name, r_cut, centers, selection.
        """
        if neigcounts is None:
            neigcounts = _lens.list_neighbours_along_trajectory(
                universe=self.universe,
                r_cut=r_cut,
                centers=centers,
                selection=selection,
                trajslice=self.trajslice,
                respect_pbc=respect_pbc,
                n_jobs=n_jobs,
            )
        if isinstance(neigcounts, _lens.NeighbourList):
            from dynsight_b200 import engine, frames

            counts = frames.to_host(engine.coordination_numbers(neigcounts.counts_device))
        else:  # a plain list[list[AtomGroup]] handed in by the caller (trajectory.py:168-174)
            n_frames = len(neigcounts)
            n_atoms = len(neigcounts[0])
            counts = np.zeros((n_atoms, n_frames), dtype=int)
            for f, frame in enumerate(neigcounts):
                for a, atom_group in enumerate(frame):
                    counts[a, f] = len(atom_group)
        attr_dict = {
            "name": "coord_number",
            "r_cut": r_cut,
            "centers": centers,
            "selection": selection,
        }
        logger.log(f"Computed coord_number using args {attr_dict}.")
        return neigcounts, Insight(dataset=counts.astype(np.float64), meta=attr_dict)

    def get_orientational_op(
        self,
        r_cut: float,
        order: int = 6,
        centers: str = "all",
        selection: str = "all",
        respect_pbc: bool = True,
        neigcounts: Any | None = None,
        n_jobs: int = 1,
    ) -> tuple[Any, Insight]:
        """Compute the magnitude of the orientational order parameter (trajectory.py:322-373).

        Returns the neighbour list and an Insight with meta: name, r_cut, order, centers,
        selection.
        """
        if neigcounts is None:
            neigcounts = _lens.list_neighbours_along_trajectory(
                universe=self.universe, r_cut=r_cut, centers=centers, selection=selection,
                trajslice=self.trajslice, respect_pbc=respect_pbc, n_jobs=n_jobs,
            )  # fmt: skip
        psi = _descriptors.orientational_order_param(
            self.universe, neigh_list_per_frame=neigcounts, order=order
        )
        attr_dict = {
            "name": "orientational_op",
            "r_cut": r_cut,
            "order": order,
            "centers": centers,
            "selection": selection,
        }
        logger.log(f"Computed orientational order parameter using args {attr_dict}.")
        return neigcounts, Insight(dataset=psi, meta=attr_dict)

    def get_velocity_alignment(
        self,
        r_cut: float,
        centers: str = "all",
        selection: str = "all",
        respect_pbc: bool = True,
        neigcounts: Any | None = None,
        n_jobs: int = 1,
    ) -> tuple[Any, Insight]:
        """Compute the average velocity alignment (trajectory.py:375-423).

        Returns the neighbour list and an Insight with meta: name, r_cut, centers, selection.
        """
        if neigcounts is None:
            neigcounts = _lens.list_neighbours_along_trajectory(
                universe=self.universe, r_cut=r_cut, centers=centers, selection=selection,
                trajslice=self.trajslice, respect_pbc=respect_pbc, n_jobs=n_jobs,
            )  # fmt: skip
        phi = _descriptors.velocity_alignment(self.universe, neigh_list_per_frame=neigcounts)
        attr_dict = {
            "name": "velocity_alignement",
            "r_cut": r_cut,
            "centers": centers,
            "selection": selection,
        }
        logger.log(f"Computed average velocity alignment using args {attr_dict}.")
        return neigcounts, Insight(dataset=phi, meta=attr_dict)

    def get_lens(
        self,
        r_cut: float,
        delay: int = 1,
        centers: str = "all",
        selection: str = "all",
        respect_pbc: bool = True,
        n_jobs: int = 1,
    ) -> Insight:
        """Compute LENS on the trajectory.

        Returns an Insight containing LENS with meta: name, r_cut, delay, centers, selection.
        """
        lens = _lens.compute_lens(
            universe=self.universe,
            r_cut=r_cut,
            delay=delay,
            centers=centers,
            selection=selection,
            trajslice=self.trajslice,
            respect_pbc=respect_pbc,
            n_jobs=n_jobs,
        )
        attr_dict = {
            "name": "lens",
            "r_cut": r_cut,
            "delay": delay,
            "centers": centers,
            "selection": selection,
        }
        logger.log(f"Computed LENS using args {attr_dict}.")
        return Insight(dataset=lens, meta=attr_dict)

    def get_soap(
        self,
        r_cut: float,
        n_max: int,
        l_max: int,
        selection: str = "all",
        centers: str = "all",
        respect_pbc: bool = True,
        n_jobs: int = 1,
    ) -> Insight:
        """Compute SOAP on the trajectory.

        The returned Insight contains the following meta: name, r_cut, n_max, l_max, respect_pbc,
        selection, centers.
        """
        soap = _soap.saponify_trajectory(
            self.universe,
            soaprcut=r_cut,
            soapnmax=n_max,
            soaplmax=l_max,
            selection=selection,
            soap_respectpbc=respect_pbc,
            centers=centers,
            n_core=n_jobs,
            trajslice=self.trajslice,
        )
        attr_dict = {
            "name": "soap",
            "r_cut": r_cut,
            "n_max": n_max,
            "l_max": l_max,
            "respect_pbc": respect_pbc,
            "selection": selection,
            "centers": centers,
        }
        logger.log(f"Computed SOAP with args {attr_dict}.")
        return Insight(dataset=soap, meta=attr_dict)

    def get_timesoap(
        self,
        r_cut: float | None = None,
        n_max: int | None = None,
        l_max: int | None = None,
        soap_insight: Insight | None = None,
        selection: str = "all",
        centers: str = "all",
        respect_pbc: bool = True,
        n_jobs: int = 1,
        delay: int = 1,
    ) -> tuple[Insight, Insight]:
        """Compute SOAP and then timeSOAP on the trajectory.

        The returned Insights (soap and timesoap) contain the following meta: name, r_cut, n_max,
        l_max, respect_pbc, selection, centers; the timeSOAP Insight also holds the delay.
        """
        if soap_insight is not None:
            if getattr(soap_insight, "meta", {}).get("name") != "soap":
                msg = (
                    f"soap_insight.meta['name'] must be 'soap', found: "
                    f"{soap_insight.meta.get('name', None)}"
                )
                raise ValueError(msg)
            msg = (
                "Loaded existing soap_insight: parameters r_cut, n_max, l_max,"
                " selection, centers, and respect_pbc will be ignored."
            )
            logger.log(msg)
            soap = soap_insight
        else:
            if r_cut is None or n_max is None or l_max is None:
                msg = "r_cut, n_max e l_max cannot be None if the soap_insight is not provided."
                raise ValueError(msg)
            soap = self.get_soap(
                r_cut=r_cut,
                n_max=n_max,
                l_max=l_max,
                selection=selection,
                centers=centers,
                respect_pbc=respect_pbc,
                n_jobs=n_jobs,
            )
            logger.log(f"Computed SOAP with args {soap.meta}.")
        timesoap = soap.get_angular_velocity(delay=delay)
        logger.log(f"Computed timeSOAP with args {timesoap.meta}.")
        return soap, timesoap
