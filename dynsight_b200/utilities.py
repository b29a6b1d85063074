"""``dynsight.utilities`` entry point that sits on the SOAP hot path
(``_internal/utilities/utilities.py:67-103``)."""

from __future__ import annotations

from pathlib import Path

from dynsight_b200.trajectory import Insight, Trj


def load_or_compute_soap(
    trj: Trj,
    r_cut: float,
    n_max: int,
    l_max: int,
    selection: str = "all",
    centers: str = "all",
    respect_pbc: bool = True,
    n_core: int = 1,
    soap_path: Path | None = None,
) -> Insight:
    """Load or compute SOAP.

    If a valid path to a .json file with a SOAP Insight is provided, that Insight is loaded and
    returned.  Otherwise the Insight is computed from the Trj (and saved if a path was given).
    """
    if soap_path is not None and soap_path.exists():
        return Insight.load_from_json(soap_path)
    soap = trj.get_soap(
        r_cut=r_cut, n_max=n_max, l_max=l_max, selection=selection, centers=centers,
        respect_pbc=respect_pbc, n_jobs=n_core,
    )  # fmt: skip
    if soap_path is not None:
        soap.dump_to_json(soap_path)
    return soap
