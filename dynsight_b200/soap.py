"""SOAP package -- drop-in for ``dynsight.soap`` (reference ``src/dynsight/soap.py:3-19``).

``saponify_trajectory`` / ``fill_soap_vector_from_dscribe`` mirror
``dynsight._internal.soapify.saponify`` (saponify.py:16-202); ``normalize_soap`` / ``soap_distance``
/ ``timesoap`` mirror ``dynsight._internal.timesoap.timesoap`` (timesoap.py:13-179).  Same argument
names, defaults, shapes and error messages; the arithmetic runs on the GPU.
"""

from __future__ import annotations

from typing import TYPE_CHECKING, Any

import numpy as np
import torch

from dynsight_b200 import engine, frames, soap_basis

if TYPE_CHECKING:
    from numpy.typing import NDArray

__all__ = [
    "fill_soap_vector_from_dscribe",
    "normalize_soap",
    "saponify_trajectory",
    "saponify_trajectory_device",
    "TimesoapJob",
    "soap_distance",
    "timesoap",
    "timesoap_from_trajectory",
]

_SYMBOLS = (
    "X H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge As Se "
    "Br Kr Rb Sr Y Zr Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe Cs Ba La Ce Pr Nd Pm Sm Eu Gd Tb Dy "
    "Ho Er Tm Yb Lu Hf Ta W Re Os Ir Pt Au Hg Tl Pb Bi Po At Rn Fr Ra Ac Th Pa U Np Pu Am Cm Bk Cf "
    "Es Fm Md No Lr Rf Db Sg Bh Hs Mt Ds Rg Cn Nh Fl Mc Lv Ts Og"
).split()
ATOMIC_NUMBERS = {sym: z for z, sym in enumerate(_SYMBOLS)}


def _species_indices(types: np.ndarray) -> tuple[np.ndarray, int]:
    """Atom types are used as element symbols (saponify.py:108,111-116); DScribe orders species
    by atomic number.  Returns per-atom species index and the species count."""
    types = np.asarray(types)
    if types.size and bool((types == types[0]).all()):
        # one species (the common case): no sort of 1e5 strings on the critical path of a call
        symbols, inverse = np.asarray([str(types[0])]), np.zeros(types.size, dtype=np.intp)
    else:
        symbols, inverse = np.unique(types.astype(str), return_inverse=True)
    try:
        z_of_symbol = np.asarray([ATOMIC_NUMBERS[str(sym)] for sym in symbols], dtype=np.int64)
    except KeyError as exc:
        msg = f"atom type {exc.args[0]!r} is not a chemical symbol (SOAP uses types as elements)"
        raise KeyError(msg) from None
    uniq, rank = np.unique(z_of_symbol, return_inverse=True)  # several symbols may share a Z
    return rank[inverse].astype(np.int32), int(uniq.size)


def _centre_positions(sel: Any, centers: str) -> np.ndarray:
    """Selection-local indices of the centre atoms, in selection order (saponify.py:92-95)."""
    sel_ix = np.asarray(sel.indices)
    cen_ix = np.asarray(sel.select_atoms(centers).indices)
    if cen_ix.size == sel_ix.size:   # every selected atom is a centre (e.g. centers="all")
        return np.arange(sel_ix.size, dtype=np.int32)
    return np.nonzero(np.isin(sel_ix, cen_ix))[0].astype(np.int32)


def saponify_trajectory_device(
    universe: Any,
    soaprcut: float,
    soapnmax: int = 8,
    soaplmax: int = 8,
    selection: str = "all",
    soap_respectpbc: bool = True,
    centers: str = "all",
    trajslice: slice | None = None,
    batch_frames: int | None = None,
) -> torch.Tensor:
    """:func:`saponify_trajectory` that leaves the (n_centers, n_frames, C) result on the device."""
    if trajslice is None:
        trajslice = slice(None)
    sel = universe.select_atoms(selection)
    sel_ix = np.asarray(sel.indices)
    centers_list = _centre_positions(sel, centers)
    sp_idx, n_species = _species_indices(np.asarray(sel.atoms.types))
    fr_idx = frames.frame_sequence(universe.trajectory.n_frames, trajslice)
    dev = frames.device()
    n_feat = soap_basis.n_features(n_species, soapnmax, soaplmax)
    n_cent = len(centers_list)
    out = torch.empty((n_cent, len(fr_idx), n_feat), dtype=torch.float64, device=dev)
    if not fr_idx or n_cent == 0:
        return out
    species_t = torch.as_tensor(sp_idx, device=dev)
    centres_t = torch.as_tensor(centers_list, device=dev)
    n_all = len(universe.atoms)
    if batch_frames is None:
        # workspace per atom-frame: sorted copies + (several species / n_max > 8) primitive sums,
        # or (one species) the neighbour rows of the list kernel (about 1 KB at liquid densities)
        scratch = soap_basis.scratch_doubles(n_species, soapnmax, soaplmax) * 8.0
        batch_frames = frames.pick_batch_frames(sel_ix.size, 80.0 + max(scratch, 1000.0), 1,
                                                maximum=1024)
    sel_all = frames.Selection(sel_ix, n_all)
    for fb in frames.iter_batches(universe, fr_idx, batch_frames):
        box = frames.soap_boxes(fb, bool(soap_respectpbc))
        pos = fb.select(sel_all, n_all)
        engine.soap_power_spectrum(
            pos, species_t, centres_t, box, float(soaprcut), int(soapnmax), int(soaplmax),
            n_species=n_species, out=out, frame0=fb.seq0,
        )  # fmt: skip
    return out


def saponify_trajectory(
    universe: Any,
    soaprcut: float,
    soapnmax: int = 8,
    soaplmax: int = 8,
    selection: str = "all",
    soap_respectpbc: bool = True,
    n_core: int = 1,  # noqa: ARG001  (joblib process count in the reference)
    centers: str = "all",
    trajslice: slice | None = None,
) -> NDArray[np.float64]:
    """Calculate the SOAP fingerprints for each atom in a universe.

    Returns the SOAP spectra for all the particles and frames, shape
    (n_atoms, n_frames, n_components), float64 (DScribe's ordering: species pairs, l, n, n').
    """
    return frames.to_host(
        saponify_trajectory_device(
            universe, soaprcut, soapnmax, soaplmax, selection, soap_respectpbc, centers, trajslice
        )
    )


def fill_soap_vector_from_dscribe(
    soapfromdscribe: NDArray[np.float64],
    lmax: int = 8,
    nmax: int = 8,
    n_species: int = 1,
) -> NDArray[np.float64]:
    """Return the full SOAP spectrum with the symmetric part explicitly stored.

    Expands the ``(l+1, n(n+1)/2)`` upper triangles to ``(l+1, n, n)`` symmetric blocks; accepts
    1-D, 2-D or 3-D input like the reference (saponify.py:176-202).  Pure data movement on the
    host: it is not on the timed hot path.

    ``n_species > 1`` (extension; the reference handles one species
    only): the input is DScribe's multi-species layout -- species pairs ``Z <= Z'``, then ``l``,
    then ``(n, n')`` (upper triangle for ``Z == Z'``, full ``n x n`` otherwise) -- and the output
    holds, per ``l``, the symmetric ``(S n) x (S n)`` matrix with row index ``Z n_max + n``.
    """
    input_shape = soapfromdscribe.shape
    arr = np.asarray(soapfromdscribe)
    if arr.ndim == 1:
        arr = arr[np.newaxis, np.newaxis, :]
    elif arr.ndim == 2:  # noqa: PLR2004
        arr = arr[:, np.newaxis, :]
    n_particles, n_frames, _ = arr.shape
    upper = np.triu_indices(nmax)
    if n_species == 1:
        tri = arr.reshape(n_particles, n_frames, lmax + 1, -1)
        full = np.zeros((n_particles, n_frames, lmax + 1, nmax, nmax))
        full[:, :, :, upper[0], upper[1]] = tri
        full[:, :, :, upper[1], upper[0]] = tri
    else:
        sn = n_species * nmax
        expected = sn * (sn + 1) // 2 * (lmax + 1)
        if arr.shape[2] != expected:
            msg = f"{arr.shape[2]} components, expected {expected} for {n_species} species"
            raise ValueError(msg)
        full = np.zeros((n_particles, n_frames, lmax + 1, sn, sn))
        pos = 0
        for zi in range(n_species):
            for zj in range(zi, n_species):
                rows = slice(zi * nmax, (zi + 1) * nmax)
                cols = slice(zj * nmax, (zj + 1) * nmax)
                if zi == zj:
                    size = (lmax + 1) * upper[0].size
                    tri = arr[:, :, pos : pos + size].reshape(n_particles, n_frames, lmax + 1, -1)
                    block = np.zeros((n_particles, n_frames, lmax + 1, nmax, nmax))
                    block[:, :, :, upper[0], upper[1]] = tri
                    block[:, :, :, upper[1], upper[0]] = tri
                    full[:, :, :, rows, cols] = block
                else:
                    size = (lmax + 1) * nmax * nmax
                    block = arr[:, :, pos : pos + size].reshape(
                        n_particles, n_frames, lmax + 1, nmax, nmax
                    )
                    full[:, :, :, rows, cols] = block
                    full[:, :, :, cols, rows] = np.swapaxes(block, -1, -2)
                pos += size
    output = full.reshape(n_particles, n_frames, -1)
    if len(input_shape) == 1:
        return output.squeeze()
    if len(input_shape) == 2:  # noqa: PLR2004
        return output.squeeze(1)
    return output


def _to_device(a: Any) -> torch.Tensor:
    if isinstance(a, torch.Tensor):
        return a.to(frames.device(), dtype=torch.float64)
    return torch.as_tensor(np.asarray(a, dtype=np.float64)).to(frames.device())


def normalize_soap(soap: NDArray[np.float64]) -> NDArray[np.float64]:
    """Return the SOAP spectra normalised to unit length (zero vectors stay zero).

    Not on the timed path (``timesoap`` fuses the normalisation); a small warp-per-vector kernel
    keeps the API complete (timesoap.py:54-58).
    """
    arr = np.asarray(soap, dtype=np.float64)
    if arr.size == 0:
        return np.zeros(arr.shape)
    return frames.to_host(engine.normalize_soap_device(_to_device(arr)))


def soap_distance(v_1: NDArray[np.float64], v_2: NDArray[np.float64]) -> NDArray[np.float64]:
    """Kernel SOAP distance between spectra of equal shape ``(..., n_comp)`` (timesoap.py:61-127)."""
    a = np.asarray(v_1, dtype=np.float64)
    b = np.asarray(v_2, dtype=np.float64)
    if a.shape != b.shape:
        a, b = np.broadcast_arrays(a, b)
    lead = a.shape[:-1]
    n_comp = a.shape[-1]
    # one "particle" per vector pair, two "frames": the fused kernel returns d(frame0, frame1)
    stacked = np.stack([a.reshape(-1, n_comp), b.reshape(-1, n_comp)], axis=1)
    if stacked.shape[0] == 0:
        return np.zeros(lead)
    out = engine.timesoap_device(_to_device(stacked), 1).cpu().numpy()[:, 0]
    res = out.reshape(lead)
    return res if res.ndim else np.float64(res)  # type: ignore[return-value]


def timesoap(soap: NDArray[np.float64], delay: int = 1) -> NDArray[np.float64]:
    """Perform the 'timeSOAP' analysis on a (n_particles, n_frames, n_comp) SOAP trajectory.

    Returns (n_particles, n_frames - delay) float64.

    Raises:
        ValueError: "delay value outside bounds" if ``delay < 1 or delay >= n_frames``.
    """
    value_error = "delay value outside bounds"
    if delay < 1 or delay >= soap.shape[1]:
        raise ValueError(value_error)
    return frames.to_host(engine.timesoap_device(_to_device(soap), delay))


class TimesoapJob:
    """timeSOAP straight from positions, fed one device batch at a time (batches overlap by
    ``delay`` frames) through the fused SOAP -> timeSOAP kernels (``engine.soap_timesoap``).  No
    SOAP tensor exists, neither of the trajectory nor of a batch."""

    def __init__(self, universe: Any, soaprcut: float, soapnmax: int = 8, soaplmax: int = 8,
                 delay: int = 1, selection: str = "all", soap_respectpbc: bool = True,
                 centers: str = "all", trajslice: slice | None = None) -> None:  # fmt: skip
        sel = universe.select_atoms(selection)
        self.universe = universe
        self.sel_ix = np.asarray(sel.indices)
        centers_list = _centre_positions(sel, centers)
        sp_idx, self.n_species = _species_indices(np.asarray(sel.atoms.types))
        self.fr_idx = frames.frame_sequence(universe.trajectory.n_frames, trajslice)
        if delay < 1 or delay >= len(self.fr_idx):
            msg = "delay value outside bounds"
            raise ValueError(msg)
        self.r_cut, self.n_max, self.l_max = float(soaprcut), int(soapnmax), int(soaplmax)
        self.delay, self.periodic = int(delay), bool(soap_respectpbc)
        dev = frames.device()
        self.n_cent = len(centers_list)
        self.n_feat = soap_basis.n_features(self.n_species, self.n_max, self.l_max)
        self.out = torch.empty((self.n_cent, len(self.fr_idx) - delay), dtype=torch.float64,
                               device=dev)  # fmt: skip
        self.species_t = frames.to_device_async(sp_idx)
        self.centres_t = frames.to_device_async(centers_list)
        self.n_all = len(universe.atoms)
        self.sel_all = frames.Selection(self.sel_ix, self.n_all)

    def default_batch_frames(self) -> int:
        # per frame: the primitive sums of split mode (several species / n_max > 8) or the
        # neighbour rows of list mode (~1 KB per atom); the spectra themselves are never stored
        scratch = soap_basis.scratch_doubles(self.n_species, self.n_max, self.l_max)
        per_frame = max(1, self.n_cent) * scratch * 8.0 + len(self.sel_ix) * 1100.0
        # memory not held by torch's allocator (cudaMemGetInfo costs ~10 ms per call: not here)
        dev = frames.device()
        free = torch.cuda.get_device_properties(dev).total_memory - torch.cuda.memory_reserved(dev)
        ring = max(1, self.n_cent) * self.delay * (self.n_feat + 1) * 8.0
        budget = max(0.45 * free - ring, 1e9)
        # batches overlap by `delay` frames (computed twice): at least 8 pairs per batch if they fit
        return int(max(self.delay + 1, min(512, budget // per_frame)))

    def process(self, fb: frames.FrameBatch) -> None:
        if self.n_cent == 0:
            return
        box = frames.soap_boxes(fb, self.periodic)
        pos = fb.select(self.sel_all, self.n_all)
        # fused: every spectrum is folded into the distance to the one `delay` frames earlier in
        # the spectrum epilogue (ring in L2); nothing of size (N, F, C) is written
        engine.soap_timesoap(
            pos, self.species_t, self.centres_t, box, self.r_cut, self.n_max, self.l_max,
            delay=self.delay, n_species=self.n_species, out=self.out, col0=fb.seq0,
        )  # fmt: skip

    def finish(self) -> torch.Tensor:
        return self.out


def timesoap_from_trajectory(
    universe: Any,
    soaprcut: float,
    soapnmax: int = 8,
    soaplmax: int = 8,
    delay: int = 1,
    selection: str = "all",
    soap_respectpbc: bool = True,
    centers: str = "all",
    trajslice: slice | None = None,
    batch_frames: int | None = None,
    as_numpy: bool = False,
) -> torch.Tensor | NDArray[np.float64]:
    """timeSOAP straight from positions without materialising the (N, F, C) SOAP tensor.

    Extension for long trajectories (SURVEY.md section 5 "long-context"): SOAP is computed in
    frame batches that overlap by ``delay`` frames and each batch is reduced to timeSOAP on the
    device.  Returns a CUDA float64 tensor (n_centers, n_frames - delay), or the same values as a
    numpy array (like ``timesoap``) with ``as_numpy=True``.
    """
    job = TimesoapJob(universe, soaprcut, soapnmax, soaplmax, delay, selection, soap_respectpbc,
                      centers, trajslice)  # fmt: skip
    if batch_frames is None:
        batch_frames = job.default_batch_frames()
    batch_frames = max(batch_frames, delay + 1)
    for fb in frames.iter_batches(universe, job.fr_idx, batch_frames, overlap=delay):
        job.process(fb)
    out = job.finish()
    return frames.to_host(out) if as_numpy else out
