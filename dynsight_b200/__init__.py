"""dynsight_b200 -- B200-native (sm_100a) descriptor engine for dynsight's hot path.

Drop-in for ``dynsight.lens``, ``dynsight.soap`` and the ``Trj`` / ``Insight`` descriptor entry
points: per-frame periodic neighbour search -> LENS / coordination number, SOAP power spectrum
-> timeSOAP.  Host code is Python over PyTorch tensors; the arithmetic is hand-written CUDA behind
a C ABI (``include/dynsight_b200.h``, ``libdynsight_b200.so``).  There is no CPU fallback.

    import dynsight_b200 as dynsight
    trj = dynsight.trajectory.Trj.init_from_xtc(xtc, gro)
    lens = trj.get_lens(r_cut=15.0)
"""

from dynsight_b200 import (
    analysis,
    descriptors,
    lens,
    logs,
    soap,
    trajectory,
    universe,
    utilities,
)

__all__ = [
    "analysis",
    "descriptors",
    "lens",
    "logs",
    "soap",
    "trajectory",
    "universe",
    "utilities",
]
__version__ = "0.1.0"
