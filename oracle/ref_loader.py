"""Load the reference's own hot-path modules by file path (TEST INFRASTRUCTURE ONLY).

This file is part of ``oracle/``: only ``tests/``, ``__graft_entry__.smoke()``, the golden
generator ``tests/golden/make_golden.py`` and ``bench.py``'s CPU-baseline legs may use it.
The product (``dynsight_b200``) never imports it.

``/root/reference`` exists only in the build container (not on the GPU box), so everything
here degrades to ``None`` when the tree is absent; the golden vectors it produced are committed
under ``tests/golden/``.

Modules loaded (reference file:line):
  * ``src/dynsight/_internal/lens/lens.py:17-189``   numba cell list, CSR neighbour list, LENS
  * ``src/dynsight/_internal/timesoap/timesoap.py:13-179``  normalize_soap, soap_distance, timesoap
``saponify.py`` cannot be imported (ase / dscribe are not installed).
"""

from __future__ import annotations

import importlib.util
import os
import sys
from pathlib import Path
from types import ModuleType

REFERENCE_ROOT = Path(os.environ.get("DYNSIGHT_REFERENCE", "/root/reference"))


def reference_available() -> bool:
    return (REFERENCE_ROOT / "src/dynsight/_internal/lens/lens.py").exists()


def _load(name: str, rel: str) -> ModuleType | None:
    path = REFERENCE_ROOT / rel
    if not path.exists():
        return None
    if name in sys.modules:
        return sys.modules[name]
    # numba cache=True needs a writable cache dir (the reference tree is read-only) and the
    # module registered in sys.modules before exec_module (pickling of the cache index).
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/dsg_numba_cache")
    spec = importlib.util.spec_from_file_location(name, path)
    assert spec is not None
    assert spec.loader is not None
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_ref_lens() -> ModuleType | None:
    return _load("_dsg_ref_lens", "src/dynsight/_internal/lens/lens.py")


def load_ref_timesoap() -> ModuleType | None:
    return _load("_dsg_ref_timesoap", "src/dynsight/_internal/timesoap/timesoap.py")
