"""ctypes binding of ``libdynsight_b200.so`` (the C ABI declared in ``include/dynsight_b200.h``).

There is no CPU fallback: if the CUDA library is missing or was not built, importing any compute
entry point raises ``RuntimeError`` (build it with ``python -m dynsight_b200.build`` or
``__graft_entry__.build()``).
"""

from __future__ import annotations

import ctypes
from pathlib import Path

import os

_PKG = Path(__file__).resolve().parent
# DSG_LIB_PATH selects an alternative build of the same library (kernel tuning experiments)
LIB_PATH = Path(os.environ.get("DSG_LIB_PATH", _PKG / "libdynsight_b200.so"))

DSG_F32 = 0
DSG_F64 = 1

_c_int = ctypes.c_int
_c_i64 = ctypes.c_int64
_c_dbl = ctypes.c_double
_c_ptr = ctypes.c_void_p
_c_size = ctypes.c_size_t

# name -> (restype, argtypes); kept in one table so tests can check it against the header.
SIGNATURES: dict[str, tuple] = {
    "dsg_version": (ctypes.c_char_p, []),
    "dsg_last_error": (ctypes.c_char_p, []),
    "dsg_device_arch": (_c_int, []),
    "dsg_launch_count": (ctypes.c_longlong, []),
    "dsg_fp64_fma_probe": (_c_int, [_c_ptr, _c_int, _c_int, _c_ptr, ctypes.POINTER(_c_dbl)]),
    "dsg_neigh_workspace_bytes": (
        _c_int, [_c_int, _c_int, _c_int, _c_int, _c_ptr, _c_dbl, ctypes.POINTER(_c_size)],
    ),
    "dsg_neigh_count": (
        _c_int,
        [_c_ptr, _c_ptr, _c_int, _c_int, _c_int, _c_int, _c_ptr, _c_dbl, _c_int, _c_ptr, _c_size,
         _c_ptr, _c_ptr, _c_ptr, _c_ptr],
    ),
    "dsg_neigh_fill": (
        _c_int,
        [_c_ptr, _c_ptr, _c_int, _c_int, _c_int, _c_int, _c_ptr, _c_dbl, _c_int, _c_ptr, _c_size,
         _c_ptr, _c_ptr, _c_ptr, _c_ptr],
    ),
    "dsg_neigh_rows": (
        _c_int,
        [_c_ptr, _c_ptr, _c_int, _c_int, _c_int, _c_int, _c_ptr, _c_dbl, _c_int, _c_ptr, _c_size,
         _c_ptr, _c_ptr, _c_ptr, _c_i64, _c_ptr, _c_ptr],
    ),
    "dsg_lens_pairs_rows": (
        _c_int, [_c_ptr, _c_ptr, _c_ptr, _c_int, _c_int, _c_int, _c_ptr, _c_i64, _c_i64, _c_ptr],
    ),
    "dsg_lens_direct_supported": (_c_int, [_c_int, _c_int, _c_ptr, _c_dbl, _c_int]),
    "dsg_lens_direct_workspace_bytes": (
        _c_int, [_c_int, _c_int, _c_ptr, _c_dbl, ctypes.POINTER(_c_size)],
    ),
    "dsg_lens_direct": (
        _c_int,
        [_c_ptr, _c_int, _c_int, _c_int, _c_ptr, _c_dbl, _c_int, _c_int, _c_ptr, _c_size, _c_ptr,
         _c_ptr, _c_i64, _c_i64, _c_ptr],
    ),
    "dsg_lens_pairs": (
        _c_int, [_c_ptr, _c_ptr, _c_ptr, _c_int, _c_int, _c_int, _c_ptr, _c_i64, _c_i64, _c_ptr],
    ),
    "dsg_lens_from_two_csr": (_c_int, [_c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_int, _c_ptr, _c_ptr]),
    "dsg_counts_to_f64": (_c_int, [_c_ptr, _c_int, _c_int, _c_ptr, _c_i64, _c_i64, _c_ptr]),
    "dsg_orientational_op": (
        _c_int,
        [_c_ptr, _c_int, _c_ptr, _c_ptr, _c_ptr, _c_int, _c_int, _c_int, _c_ptr, _c_i64, _c_i64,
         _c_int, _c_ptr],
    ),
    "dsg_velocity_alignment": (
        _c_int,
        [_c_ptr, _c_int, _c_int, _c_ptr, _c_ptr, _c_ptr, _c_int, _c_int, _c_ptr, _c_i64, _c_i64,
         _c_int, _c_ptr],
    ),
    "dsg_timesoap": (
        _c_int,
        [_c_ptr, _c_int, _c_int, _c_int, _c_i64, _c_i64, _c_int, _c_ptr, _c_i64, _c_i64, _c_ptr],
    ),
    "dsg_normalize_soap": (_c_int, [_c_ptr, _c_ptr, _c_i64, _c_int, _c_ptr]),
    "dsg_soap_n_features": (_c_int, [_c_int, _c_int, _c_int]),
    "dsg_soap_workspace_bytes": (
        _c_int,
        [_c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_ptr, _c_dbl, ctypes.POINTER(_c_size)],
    ),
    "dsg_soap": (
        _c_int,
        [_c_ptr, _c_int, _c_ptr, _c_ptr, _c_int, _c_int, _c_int, _c_int, _c_ptr, _c_dbl, _c_int,
         _c_int, _c_ptr, _c_ptr, _c_ptr, _c_size, _c_ptr, _c_i64, _c_i64, _c_ptr],
    ),
    "dsg_soap_timesoap_workspace_bytes": (
        _c_int,
        [_c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_ptr, _c_dbl, _c_int,
         ctypes.POINTER(_c_size)],
    ),
    "dsg_soap_timesoap": (
        _c_int,
        [_c_ptr, _c_int, _c_ptr, _c_ptr, _c_int, _c_int, _c_int, _c_int, _c_ptr, _c_dbl, _c_int,
         _c_int, _c_ptr, _c_ptr, _c_int, _c_ptr, _c_size, _c_ptr, _c_i64, _c_i64, _c_ptr, _c_i64,
         _c_i64, _c_ptr],
    ),
    "dsg_rows_csr_workspace_bytes": (_c_int, [_c_int, _c_int, ctypes.POINTER(_c_size)]),
    "dsg_rows_csr_offsets": (
        _c_int, [_c_ptr, _c_int, _c_int, _c_ptr, _c_size, _c_ptr, _c_ptr, _c_ptr],
    ),
    "dsg_rows_to_csr": (
        _c_int, [_c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_int, _c_int, _c_ptr, _c_int, _c_ptr],
    ),
    "dsg_wrap_positions": (_c_int, [_c_ptr, _c_int, _c_int, _c_int, _c_ptr, _c_ptr, _c_ptr]),
    "dsg_spatial_average_rows": (
        _c_int,
        [_c_ptr, _c_ptr, _c_ptr, _c_int, _c_int, _c_int, _c_ptr, _c_i64, _c_i64, _c_ptr, _c_i64,
         _c_i64, _c_int, _c_ptr],
    ),
    "dsg_tcf_workspace_bytes": (_c_int, [_c_int, _c_int, _c_int, ctypes.POINTER(_c_size)]),
    "dsg_self_time_correlation": (
        _c_int,
        [_c_ptr, _c_i64, _c_int, _c_int, _c_int, _c_ptr, _c_size, _c_ptr, _c_ptr, _c_ptr],
    ),
    "dsg_cross_tcf_workspace_bytes": (_c_int, [_c_int, _c_int, _c_int, ctypes.POINTER(_c_size)]),
    "dsg_cross_time_correlation": (
        _c_int,
        [_c_ptr, _c_i64, _c_int, _c_int, _c_int, _c_ptr, _c_size, _c_ptr, _c_ptr, _c_ptr],
    ),
    "dsg_xtc_index": (
        _c_int,
        [ctypes.c_char_p, ctypes.POINTER(_c_i64), ctypes.POINTER(ctypes.c_int32), _c_ptr, _c_i64],
    ),
    "dsg_xtc_read": (
        _c_int,
        [ctypes.c_char_p, _c_ptr, _c_i64, ctypes.c_int32, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_int],
    ),
}  # fmt: skip

_LIB: ctypes.CDLL | None = None


class DsgError(RuntimeError):
    """A non-zero status returned through the C ABI."""

    def __init__(self, code: int, message: str) -> None:
        super().__init__(f"libdynsight_b200 error {code}: {message}")
        self.code = code
        self.message = message


def load() -> ctypes.CDLL:
    """Load the shared library (once).  Raises if it has not been built -- no fallback."""
    global _LIB  # noqa: PLW0603
    if _LIB is not None:
        return _LIB
    if not LIB_PATH.exists():
        msg = (
            f"{LIB_PATH} not found: the CUDA extension is required (no CPU fallback). "
            "Build it with `python -m dynsight_b200.build`."
        )
        raise RuntimeError(msg)
    lib = ctypes.CDLL(str(LIB_PATH))
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing: fail loudly
        fn.restype = restype
        fn.argtypes = argtypes
    _LIB = lib
    return lib


def last_error() -> str:
    return load().dsg_last_error().decode()


def check(status: int) -> None:
    if status != 0:
        raise DsgError(status, last_error())
