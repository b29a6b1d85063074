"""timeSOAP kernel time for delays 1..8 on a (100000, 33, 324) spectra block."""
import sys, torch
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dynsight_b200 import engine
soap = torch.rand((100000, 33, 324), device="cuda", dtype=torch.float64)
for delay in (1, 2, 3, 4, 8):
    for it in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = engine.timesoap_device(soap, delay); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"delay {delay}: {ms:.3f} ms  ({soap.numel() * 8 / 1e6 / ms:.0f} GB/s of one pass over the spectra)")
