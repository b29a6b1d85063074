"""Build ``libdynsight_b200.so`` in-tree with nvcc for sm_100a (``python -m dynsight_b200.build``)."""

from __future__ import annotations

import shutil
import subprocess
import sys
from pathlib import Path

_PKG = Path(__file__).resolve().parent
CSRC = _PKG / "csrc"
SOURCES = ["core.cu", "neighbors.cu", "lens.cu", "descriptors.cu", "timesoap.cu", "soap.cu",
           "spatial.cu", "tcf.cu", "xtc.cpp"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-shared",
]  # fmt: skip


def needs_build(out: Path) -> bool:
    if not out.exists():
        return True
    newest = max(p.stat().st_mtime for p in [*CSRC.glob("*.cu"), *CSRC.glob("*.cuh"), *CSRC.glob("*.cpp"),
                                             _PKG.parent / "include" / "dynsight_b200.h"])
    return out.stat().st_mtime < newest


def build_library(force: bool = False, verbose: bool = False) -> Path:
    out = _PKG / "libdynsight_b200.so"
    if not force and not needs_build(out):
        return out
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc, *NVCC_FLAGS, *(["-Xptxas", "-v"] if verbose else []), "-o", str(out),
           *[str(CSRC / s) for s in SOURCES]]
    proc = subprocess.run(cmd, capture_output=True, text=True, check=False)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        msg = f"nvcc failed with exit code {proc.returncode}"
        raise RuntimeError(msg)
    return out


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
