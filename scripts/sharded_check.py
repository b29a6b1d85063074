"""torchrun --nproc-per-node N scripts/sharded_check.py : the sharded drivers (NCCL) reproduce the
single-GPU results bit for bit (LENS) / to 1e-12 (timeSOAP)."""
import os, sys
from pathlib import Path
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
from golden_cases import fluid_case
from dynsight_b200 import lens, parallel, soap
from dynsight_b200.universe import ArrayUniverse

rank = int(os.environ["RANK"]); local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl")
pos, box, r_cut = fluid_case(n_side=10, n_frames=23)
pos = pos * 3.0; box = box * 3.0   # angstrom-like scale for SOAP (r_cut 4.5)
dims = np.tile(np.concatenate([box, [90, 90, 90]]).astype(np.float32), (pos.shape[0], 1))
uni = ArrayUniverse(pos, dims, ["O"] * pos.shape[1], ["O"] * pos.shape[1])
for delay, sl in ((1, None), (3, slice(2, None, 2))):
    got = parallel.compute_lens_sharded(uni, 4.5, delay, trajslice=sl)
    got_ts = parallel.timesoap_sharded(uni, 4.5, 4, 4, delay, trajslice=sl)
    if rank == 0:
        want = lens.compute_lens_device(uni, 4.5, delay, trajslice=sl)
        want_ts = soap.timesoap_from_trajectory(uni, 4.5, 4, 4, delay, trajslice=sl)
        assert torch.equal(got, want), "LENS mismatch"
        err = float((got_ts - want_ts).abs().max())
        assert err < 1e-10, err
        print(f"delay {delay} slice {sl}: LENS {tuple(got.shape)} bit-exact, timeSOAP max diff {err:.1e}")
dist.barrier()
dist.destroy_process_group()
