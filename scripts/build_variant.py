"""Build a tuning variant of the library: python scripts/build_variant.py NAME -DFLAG=VALUE ...
-> gpurun_variants/lib_NAME.so (select it at run time with DSG_LIB_PATH)."""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from dynsight_b200.build import CSRC, NVCC_FLAGS, SOURCES  # noqa: E402

name, flags = sys.argv[1], sys.argv[2:]
out = ROOT / "gpurun_variants" / f"lib_{name}.so"
out.parent.mkdir(exist_ok=True)
cmd = ["nvcc", *NVCC_FLAGS, *flags, "-o", str(out), *[str(CSRC / s) for s in SOURCES]]
subprocess.run(cmd, check=True)
print(out)
