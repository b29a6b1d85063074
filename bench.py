#!/usr/bin/env python
"""Benchmark of the dynsight hot path on B200: LENS + SOAP + timeSOAP particle-frames/s.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
    python bench.py --impl reference --gpus N --steps K ...  # the reference CPU path (oracle port)

Workload (BASELINE.json configs[4], the configuration the metric "at 1/2/4/8 B200" is quoted on):
a 100k-atom single-species box at water-oxygen density (L = 144.1 A), LENS(r_cut = 5 A, delay 1)
+ SOAP(r_cut = 5 A, n_max = l_max = 8) + timeSOAP(delay 1).  A *step* is one frame block of
``--frames-per-step`` frames (+ a one-frame halo) per GPU; trajectories are sharded by frame
blocks, every rank works on its own blocks (weak scaling) and receives the halo frame of the next
rank's block over NCCL inside the timed region.  ``value`` = particle-frames of all ranks / the
max-over-ranks device time, inputs resident in HBM; ``e2e`` = the same through the public Python
API (``dynsight_b200.lens.compute_lens`` + ``dynsight_b200.soap.timesoap_from_trajectory``) with
pinned HOST positions in and host results out.  One JSON line is printed by rank 0.
"""

from __future__ import annotations

import argparse
import ctypes
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "LENS+SOAP+timeSOAP particle-frames/s"
UNIT = "particle-frames/s"
RHO = 0.0334  # atoms / A^3 (water oxygen density), BASELINE.md cfg2 / cfg5
R_CUT_LENS = 5.0
R_CUT_SOAP = 5.0
N_MAX = 8
L_MAX = 8
DELAY = 1


def parse_args() -> argparse.Namespace:
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", choices=["b200", "reference"], default="b200")
    p.add_argument("--atoms", type=int, default=100_000)
    p.add_argument("--frames-per-step", type=int, default=32)
    p.add_argument("--pool-blocks", type=int, default=4, help="resident frame blocks cycled through")
    p.add_argument("--no-extra", action="store_true", help="skip the LENS-only cfg3 section")
    p.add_argument("--no-cpu-baseline", action="store_true")
    return p.parse_args()


def soap_flops_per_pf(k_mean: float, n_species: int = 1) -> tuple[float, float]:
    """Algorithmic flops and exps per particle-frame."""
    lm = (L_MAX + 1) ** 2
    sn = n_species * N_MAX
    flops = 2.0 * (k_mean * N_MAX * lm + lm * sn * (sn + 1) / 2)
    exps = k_mean * N_MAX * (L_MAX + 1)
    return flops, exps


def measured_peaks() -> tuple[float, str]:
    path = ROOT / "MEASURED_PEAKS.json"
    if path.exists():
        return float(json.loads(path.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")  # fmt: skip

    def __init__(self, index: int) -> None:
        self.rows: list[list[str]] = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={index}", f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )  # fmt: skip
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self) -> None:
        assert self.proc is not None and self.proc.stdout is not None
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power = [], [], []
        reasons: set[str] = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:  # noqa: PLR2004
                continue
            try:
                sm.append(float(r[0]))
                smax.append(float(r[1]))
                power.append(float(r[2]))
            except ValueError:
                continue
            for name, flag in zip(names, r[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": max(smax) if smax else None,
            "power_w_max": max(power) if power else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


# ------------------------------------------------------------------------------------------------
# CPU reference arm / baseline (the oracle port of the reference's numba + DScribe path)
# ------------------------------------------------------------------------------------------------
def host_trajectory(n_atoms: int, n_frames: int, seed: int) -> tuple[np.ndarray, np.ndarray]:
    rng = np.random.default_rng(seed)
    box_l = (n_atoms / RHO) ** (1.0 / 3.0)
    pos = np.empty((n_frames, n_atoms, 3), dtype=np.float32)
    cur = rng.random((n_atoms, 3)) * box_l
    for f in range(n_frames):
        if f:
            cur = np.mod(cur + rng.normal(0.0, 0.1, cur.shape), box_l)
        pos[f] = cur.astype(np.float32)
    box = np.full(3, np.float32(box_l), dtype=np.float32)
    np.minimum(pos, np.nextafter(box[0], np.float32(0)), out=pos)
    return pos, box


def cpu_reference_step(pos: np.ndarray, box: np.ndarray, n_soap_centres: int, threads: int
                       ) -> dict:  # fmt: skip
    """One bounded sample of the reference CPU path on ``pos`` (2 frames): LENS as compute_lens
    drives it (two neighbour lists + merge per pair, lens.py:263-308), SOAP on a subset of centres
    (C/OpenMP restatement of DScribe -- DScribe is not installed), timeSOAP with numpy."""
    from oracle import lens_oracle, soap_oracle, timesoap_oracle

    lens_oracle.set_num_threads(threads)
    n_atoms = pos.shape[1]
    dims = np.tile(np.concatenate([box, [90, 90, 90]]).astype(np.float32), (pos.shape[0], 1))
    t0 = time.perf_counter()
    lens_oracle.compute_lens_arrays(pos[:2], dims[:2], R_CUT_LENS, delay=1)
    t_lens = time.perf_counter() - t0
    centres = np.linspace(0, n_atoms - 1, n_soap_centres).astype(np.int32)
    basis = soap_oracle.gto_basis(R_CUT_SOAP, N_MAX, L_MAX)
    z = np.full(n_atoms, 8)
    t0 = time.perf_counter()
    spectra = [
        soap_oracle.soap_frame_c(pos[f], z, centres, box.astype(np.float64), R_CUT_SOAP, N_MAX,
                                 L_MAX, basis, n_threads=threads)
        for f in range(2)
    ]  # fmt: skip
    t_soap = (time.perf_counter() - t0) / 2.0
    soap = np.stack(spectra, axis=1)
    t0 = time.perf_counter()
    timesoap_oracle.timesoap(soap, 1)
    t_ts = time.perf_counter() - t0
    per_pf = t_lens / n_atoms + t_soap / n_soap_centres + t_ts / n_soap_centres
    return {
        "pf_per_s": 1.0 / per_pf,
        "lens_pf_per_s": n_atoms / t_lens,
        "soap_pf_per_s": n_soap_centres / t_soap,
        "timesoap_pf_per_s": n_soap_centres / t_ts,
        "seconds": t_lens + 2 * t_soap + t_ts,
    }


def run_reference(args: argparse.Namespace) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the host path
    threads = os.cpu_count() or 1
    pos, box = host_trajectory(args.atoms, 2, seed=0)
    n_c = args.atoms
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_reference_step(pos, box, 500, threads)
    t0 = time.perf_counter()
    res = [cpu_reference_step(pos, box, n_c, threads) for _ in range(args.steps)]
    wall = time.perf_counter() - t0
    value = float(np.mean([r["pf_per_s"] for r in res]))
    sample = (f"per step: LENS on 1 frame pair of {args.atoms} atoms, SOAP on {n_c} centres x 2 "
              f"frames, timeSOAP on {n_c} pairs; combined per-particle-frame cost")  # fmt: skip
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * wall / max(1, args.steps),
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {
            "value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
            "lens_pf_per_s": float(np.mean([r["lens_pf_per_s"] for r in res])),
            "soap_pf_per_s": float(np.mean([r["soap_pf_per_s"] for r in res])),
            "timesoap_pf_per_s": float(np.mean([r["timesoap_pf_per_s"] for r in res])),
            "note": "reference is pure Python (numba + DScribe); oracle port timed "
                    "(C/OpenMP restatement of the numba kernels and of DScribe SOAP-GTO, numpy "
                    "timeSOAP); DScribe itself is not installable here",
        },
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }  # fmt: skip
    print(json.dumps(line), flush=True)


def workload_config(args: argparse.Namespace) -> dict:
    return {
        "workload": (
            f"cfg5 (BASELINE.json configs[4]): {args.atoms}-atom single-species box, rho=0.0334/A^3, "
            f"LENS r_cut=5A delay=1 + SOAP r_cut=5A n_max=8 l_max=8 + timeSOAP delay=1; one step = "
            f"one block of {args.frames_per_step} frames (+1 halo frame) per GPU"
        ),
        "atoms": args.atoms,
        "frames_per_step": args.frames_per_step,
        "frame_sharding": "contiguous frame blocks per rank, 1-frame NCCL halo per step",
        "l2_policy": (
            "each step reads a different resident frame block; SOAP spectra of a block "
            "(>8 GB) exceed L2 (126 MB)"
        ),
    }


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200(args: argparse.Namespace) -> None:
    import torch
    import torch.distributed as dist

    from dynsight_b200 import _lib, engine
    from dynsight_b200 import lens as dlens
    from dynsight_b200 import soap as dsoap
    from dynsight_b200.universe import ArrayUniverse

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    n_atoms, bsz = args.atoms, args.frames_per_step
    box_l = (n_atoms / RHO) ** (1.0 / 3.0)
    box = np.full(3, np.float32(box_l), dtype=np.float32)
    n_pool = args.pool_blocks * bsz
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    # synthetic trajectory generated on the device: uniform start + Gaussian random walk, wrapped
    pool = torch.empty((n_pool, n_atoms, 3), dtype=torch.float32, device=dev)
    cur = torch.rand((n_atoms, 3), device=dev, generator=gen, dtype=torch.float64) * box_l
    top = float(np.nextafter(box[0], np.float32(0)))
    for f in range(n_pool):
        if f:
            cur = torch.remainder(
                cur + 0.1 * torch.randn((n_atoms, 3), device=dev, generator=gen, dtype=torch.float64),
                box_l,
            )
        pool[f] = cur.to(torch.float32).clamp_(max=top)
    species = torch.zeros(n_atoms, dtype=torch.int32, device=dev)
    centres = torch.arange(n_atoms, dtype=torch.int32, device=dev)
    n_feat = 324
    soap_buf = torch.empty((n_atoms, bsz + 1, n_feat), dtype=torch.float64, device=dev)
    lens_out = torch.empty((n_atoms, bsz), dtype=torch.float64, device=dev)
    ts_out = torch.empty((n_atoms, bsz), dtype=torch.float64, device=dev)
    nc_out = torch.empty((n_atoms, bsz + 1), dtype=torch.float64, device=dev)
    block = torch.empty((bsz + 1, n_atoms, 3), dtype=torch.float32, device=dev)
    halo_recv = torch.empty((DELAY, n_atoms, 3), dtype=torch.float32, device=dev)

    ev = {k: [] for k in ("neigh", "lens", "soap", "tsoap")}
    # rows-path capacity: generous fixed estimate, verified after the timed region (no mid-step sync)
    k_est = RHO * 4.18879 * R_CUT_LENS**3
    rows_capacity = [int(1.4 * k_est * n_atoms * (bsz + 1)) + 256 * 4096]
    last_rows = [None]

    def step(b: int, record: bool) -> None:
        lo = (b % args.pool_blocks) * bsz
        block[:bsz].copy_(pool[lo : lo + bsz])
        nxt = ((b + 1) % args.pool_blocks) * bsz
        if world > 1:  # halo = first frame of the next rank's block, over NCCL / NVLink
            send = pool[lo : lo + DELAY].contiguous()
            ops = [
                dist.P2POp(dist.isend, send, (rank - 1) % world),
                dist.P2POp(dist.irecv, halo_recv, (rank + 1) % world),
            ]
            for req in dist.batch_isend_irecv(ops):
                req.wait()
            block[bsz:].copy_(halo_recv)
        else:
            block[bsz:].copy_(pool[nxt : nxt + DELAY])
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(5)] if record else None
        if marks:
            marks[0].record()
        nb = engine.neighbor_rows(block, box, R_CUT_LENS, capacity=rows_capacity[0])
        if marks:
            marks[1].record()
        engine.lens_pairs_rows(nb, DELAY, out=lens_out, col0=0)
        engine.coordination_numbers(nb.counts, out=nc_out, col0=0)
        last_rows[0] = nb
        if marks:
            marks[2].record()
        engine.soap_power_spectrum(block, species, centres, box, R_CUT_SOAP, N_MAX, L_MAX,
                                   n_species=1, out=soap_buf, frame0=0)  # fmt: skip
        if marks:
            marks[3].record()
        engine.timesoap_device(soap_buf, DELAY, out=ts_out, col0=0)
        if marks:
            marks[4].record()
            for k, name in enumerate(("neigh", "lens", "soap", "tsoap")):
                ev[name].append((marks[k], marks[k + 1]))

    def barrier() -> None:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for w in range(args.warmup):
        step(w, record=False)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = engine.launch_count()
    t_start = torch.cuda.Event(enable_timing=True)
    t_stop = torch.cuda.Event(enable_timing=True)
    t_start.record()
    for k in range(args.steps):
        step(args.warmup + k, record=True)
    t_stop.record()
    barrier()
    ms = t_start.elapsed_time(t_stop)
    if last_rows[0].overflowed():
        raise SystemExit("bench.py: neighbour rows capacity overflow -- timing invalid")
    launches = engine.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    pf_per_step = n_atoms * bsz
    value = world * pf_per_step * args.steps / (ms * 1e-3)
    part_ms = {k: float(np.mean([a.elapsed_time(b) for a, b in v])) for k, v in ev.items()}
    k_mean_lens = float(nc_out[:, :bsz].mean().item())

    # ---- fp64 FMA peak of this device ("builder must measure") ----
    scratch = torch.empty(148 * 8 * 256, dtype=torch.float64, device=dev)
    flops = ctypes.c_double(0.0)
    best = 0.0
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.dsg_fp64_fma_probe(scratch.data_ptr(), 148 * 8, 20000,
                                          torch.cuda.current_stream().cuda_stream,
                                          ctypes.byref(flops)))  # fmt: skip
        e1.record()
        torch.cuda.synchronize()
        best = max(best, flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    fp64_peak_tflops = best

    # ---- rooflines ----
    hbm_peak, hbm_src = measured_peaks()
    r_soap = R_CUT_SOAP + math.sqrt(-2.0 * math.log(1e-3))
    k_soap = RHO * 4.0 / 3.0 * math.pi * r_soap**3 + 1.0
    fl, ex = soap_flops_per_pf(k_soap)
    soap_pf = n_atoms * (bsz + 1)
    soap_tflops = fl * soap_pf / (part_ms["soap"] * 1e-3) / 1e12
    soap_bytes = (12 + 8 * n_feat) * soap_pf
    lens_bytes_pf = 20 + 8 * (k_mean_lens + 1)
    lens_gbs = lens_bytes_pf * pf_per_step / ((part_ms["neigh"] + part_ms["lens"]) * 1e-3) / 1e9
    ts_gbs = (8 * n_feat + 8) * pf_per_step / (part_ms["tsoap"] * 1e-3) / 1e9
    roofline = {
        "kernel": "soap_kernel<8,20,0,0,0,false,LIST=true> (+ soap_list_kernel 3.5 %, binning < 1 %)",
        "bound": "fp64",
        "achieved": soap_tflops,
        "peak": fp64_peak_tflops,
        "unit": "TFLOP/s",
        "frac": soap_tflops / fp64_peak_tflops if fp64_peak_tflops else None,
        # dram__bytes_read+write of this kernel from the ncu --set full capture summarised in
        # profiles/r1_soap_kernel_list_summary.txt: 2724 B per centre-frame, scaled to this launch
        "traffic": 2724.0 * soap_pf,
        "traffic_source": "profiles/r1_soap_kernel_list_summary.txt (2724 B per centre-frame)",
        "peak_source": "dsg_fp64_fma_probe measured in this run (not in MEASURED_PEAKS.json)",
        "algorithmic_flops_per_pf": fl,
        "exps_per_pf_not_counted_as_flops": ex,
        "mean_neighbours_incl_self": k_soap,
        "share_of_step": part_ms["soap"] / sum(part_ms.values()),
        "hbm_view": {
            "achieved_gbs": soap_bytes / (part_ms["soap"] * 1e-3) / 1e9, "peak_gbs": hbm_peak,
            "algorithmic_bytes_per_pf": 12 + 8 * n_feat,
        },
    }  # fmt: skip
    rooflines_hbm = [
        {
            "kernel": "cell list + neigh_rows_kernel (single traversal) + lens_rows_run_kernel",
            "bound": "hbm", "achieved": lens_gbs, "peak": hbm_peak, "unit": "GB/s",
            "frac": lens_gbs / hbm_peak, "traffic": 87.0 * n_atoms * (bsz + 1),
            "traffic_source": "profiles/r1_neigh_rows_warp_kernel_summary.txt (87 B per centre-frame, "
                              "neighbour kernel only)",
            "algorithmic_bytes_per_pf": lens_bytes_pf, "mean_neighbours": k_mean_lens,
            "ms_per_step": part_ms["neigh"] + part_ms["lens"], "peak_source": hbm_src,
        },
        {
            "kernel": "timesoap_stream_kernel<12>",
            "bound": "hbm", "achieved": ts_gbs, "peak": hbm_peak, "unit": "GB/s",
            "frac": ts_gbs / hbm_peak, "traffic": None,
            "algorithmic_bytes_per_pf": 8 * n_feat + 8, "ms_per_step": part_ms["tsoap"],
            "peak_source": hbm_src,
        },
    ]  # fmt: skip

    # ---- end to end through the public API: pinned host positions in, host arrays out ----
    host_block = torch.empty((bsz + 1, n_atoms, 3), dtype=torch.float32, pin_memory=True)
    host_block.copy_(block)
    dims = np.tile(np.concatenate([box, [90, 90, 90]]).astype(np.float32), (bsz + 1, 1))
    uni = ArrayUniverse(host_block.numpy(), dims, names=["O"] * n_atoms, types=["O"] * n_atoms)

    def e2e_step() -> int:
        lens_np = dlens.compute_lens(uni, R_CUT_LENS, delay=DELAY)
        ts_np = dsoap.timesoap_from_trajectory(uni, R_CUT_SOAP, N_MAX, L_MAX, delay=DELAY,
                                               as_numpy=True)
        return lens_np.nbytes + ts_np.nbytes

    for _ in range(2):  # warm-up: page-locked result buffers enter the host allocator's cache
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(max(1, min(args.steps, 3))):
        t1 = time.perf_counter()
        d2h = e2e_step()
        if os.environ.get("DSG_BENCH_DEBUG"):
            print(f"e2e step {1e3 * (time.perf_counter() - t1):.2f} ms", file=sys.stderr)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / max(1, min(args.steps, 3))
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = {
        "value": world * pf_per_step / e2e_s,
        "unit": UNIT,
        "h2d_bytes_per_step": 2 * host_block.numel() * 4,
        "d2h_bytes_per_step": d2h,
        "api": "dynsight_b200.lens.compute_lens + dynsight_b200.soap.timesoap_from_trajectory",
        "ms_per_step": e2e_s * 1e3,
    }

    extra = None
    widened = None
    if rank == 0 and not args.no_extra:
        extra = lens_cfg3_section(engine, dev, hbm_peak, hbm_src)
        widened = widened_section(engine, dev, hbm_peak, fp64_peak_tflops)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        hpos = block[:2].cpu().numpy()
        cpu_reference_step(hpos, box, 500, threads)  # warm the thread pool and page in the libs
        reps = [cpu_reference_step(hpos, box, n_atoms, threads) for _ in range(4)]
        secs = sum(r["seconds"] for r in reps)
        cpu_baseline = {
            "value": float(np.mean([r["pf_per_s"] for r in reps])), "unit": UNIT,
            "cores": threads, "kind": "port",
            "sample": (f"4 x (LENS on 1 frame pair of {n_atoms} atoms, SOAP on all {n_atoms} "
                       f"centres x 2 frames, timeSOAP on {n_atoms} pairs): {secs:.1f} s of CPU "
                       "work"),
            "lens_pf_per_s": float(np.mean([r["lens_pf_per_s"] for r in reps])),
            "soap_pf_per_s": float(np.mean([r["soap_pf_per_s"] for r in reps])),
            "timesoap_pf_per_s": float(np.mean([r["timesoap_pf_per_s"] for r in reps])),
        }  # fmt: skip

    if rank == 0:
        line = {
            "metric": METRIC,
            "value": value,
            "unit": UNIT,
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": ms / args.steps,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f64",
            "data": "synthetic",
            "config": workload_config(args),
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "roofline": roofline,
            "rooflines_hbm": rooflines_hbm,
            "breakdown_ms_per_step": part_ms,
            "breakdown_pf_per_s": {
                "lens_from_positions": pf_per_step / ((part_ms["neigh"] + part_ms["lens"]) * 1e-3),
                "soap": soap_pf / (part_ms["soap"] * 1e-3),
                "timesoap": pf_per_step / (part_ms["tsoap"] * 1e-3),
            },
            "cpu_baseline": cpu_baseline,
            "lens_cfg3": extra,
            "widened_8f": widened,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def lens_cfg3_section(engine, dev, hbm_peak: float, hbm_src: str) -> dict:
    """LENS-only throughput on BASELINE.json configs[2]: 1M particles, rho*=0.8, r_cut=1.5."""
    import torch

    n_side = 100
    n = n_side**3
    box_l = (n / 0.8) ** (1.0 / 3.0)
    n_fr = 9
    gen = torch.Generator(device=dev).manual_seed(7)
    grid = (torch.arange(n_side, device=dev, dtype=torch.float32) + 0.5) * (box_l / n_side)
    lat = torch.stack(torch.meshgrid(grid, grid, grid, indexing="ij"), dim=-1).reshape(-1, 3)
    lat = lat[torch.randperm(n, device=dev, generator=gen)]
    top = float(np.nextafter(np.float32(box_l), np.float32(0)))
    cur = lat + 0.2 * torch.randn(lat.shape, device=dev, generator=gen)
    frames = []
    pools = []
    for _ in range(2):  # two resident blocks (each 108 MB of positions) alternate
        for _f in range(n_fr):
            cur = torch.remainder(cur + 0.05 * torch.randn(lat.shape, device=dev, generator=gen), box_l)
            frames.append(cur.clamp(max=top))
        pools.append(torch.stack(frames[-n_fr:]).contiguous())
    box = np.full(3, np.float32(box_l), dtype=np.float32)
    out = torch.empty((n, n_fr - 1), dtype=torch.float64, device=dev)
    times = []
    k_mean = 0.0
    for it in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _, nb = engine.lens_from_positions(pools[it % 2], box, 1.5, 1, out=out, col0=0)
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            times.append(e0.elapsed_time(e1))
        k_mean = float(nb.counts.float().mean().item())
    ms = float(np.mean(times))
    pf = n * (n_fr - 1)
    bytes_pf = 20 + 8 * (k_mean + 1)
    gbs = bytes_pf * pf / (ms * 1e-3) / 1e9
    return {
        "workload": "cfg3 (BASELINE.json configs[2]): 1,000,000 particles, rho*=0.8, r_cut=1.5, "
                    f"{n_fr} frames per launch batch (LENS delay 1)",
        "value": pf / (ms * 1e-3), "unit": UNIT, "ms_per_batch": ms,
        "mean_neighbours": k_mean,
        "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                     "frac": gbs / hbm_peak, "traffic": None,
                     "algorithmic_bytes_per_pf": bytes_pf, "peak_source": hbm_src},
    }  # fmt: skip


def widened_section(engine, dev, hbm_peak: float, fp64_peak_tflops: float) -> dict:
    """Throughput of the rows widened per SURVEY.md section 8f (reported beside the headline)."""
    import torch

    def timed(fn, reps: int = 4) -> float:
        fn()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.mean(ts))

    out: dict = {}
    # spatial average: cfg3-like fluid, scalar descriptor and a 324-component one (SOAP size)
    gen = torch.Generator(device=dev).manual_seed(11)
    for n, nf, n_comp in ((1_000_000, 4, 1), (100_000, 4, 324)):
        box_l = (n / 0.8) ** (1.0 / 3.0)
        pos = (torch.rand((nf, n, 3), device=dev, generator=gen) * box_l).float()
        box = np.full((nf, 3), box_l)
        shape = (n, nf) if n_comp == 1 else (n, nf, n_comp)
        values = torch.rand(shape, device=dev, generator=gen, dtype=torch.float64)
        res = torch.empty_like(values)
        state = {}

        def run(pos=pos, box=box, values=values, res=res, state=state) -> None:
            nr = engine.neighbor_rows(engine.wrap_positions(pos, box), box, 1.5,
                                      inclusive_cutoff=True)
            engine.spatial_average_rows(nr, values, res)
            state["nr"] = nr

        ms = timed(run)
        k = float(state["nr"].counts.float().mean().item())
        bytes_pf = 12 + 4 * (k + 1) + 8 * n_comp * (k + 2)
        gbs = bytes_pf * n * nf / (ms * 1e-3) / 1e9
        out[f"spatial_average_c{n_comp}"] = {
            "workload": f"{n} particles x {nf} frames, rho*=0.8, r_cut=1.5, {n_comp} component(s)",
            "value": n * nf / (ms * 1e-3), "unit": UNIT, "ms": ms, "mean_neighbours": k,
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                         "frac": gbs / hbm_peak, "algorithmic_bytes_per_pf": bytes_pf},
        }  # fmt: skip
        del pos, values, res, state
    # neighbour-list consumers (section 8f row 2): canonical CSR, then psi_6 and phi
    n, nf = 1_000_000, 4
    box_l = (n / 0.8) ** (1.0 / 3.0)
    pos = (torch.rand((nf, n, 3), device=dev, generator=gen) * box_l).float()
    box = np.full((nf, 3), box_l)
    state = {}

    def run_csr() -> None:
        state["nb"] = engine.neighbor_lists(pos, box, 1.5)

    ms_csr = timed(run_csr)
    nb = state["nb"]
    k = float(nb.counts.float().mean().item())
    ms_psi = timed(lambda: engine.orientational_op_device(pos, nb, 6))
    ms_phi = timed(lambda: engine.velocity_alignment_device(pos, nb, displacement=True))
    for name, ms, npf, bytes_pf in (
        ("canonical_csr", ms_csr, n * nf, 12 + 4 + 4 * k),
        ("orientational_op", ms_psi, n * nf, 4 * 2 + 4 * k + 8 * (k + 1) + 8),
        ("velocity_alignment", ms_phi, n * (nf - 1), 4 * 2 + 4 * k + 12 * 2 * (k + 1) + 8),
    ):
        gbs = bytes_pf * npf / (ms * 1e-3) / 1e9
        out[name] = {
            "workload": f"{n} particles x {nf} frames, rho*=0.8, r_cut=1.5, k={k:.1f}",
            "value": npf / (ms * 1e-3), "unit": UNIT, "ms": ms,
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                         "frac": gbs / hbm_peak, "algorithmic_bytes_per_pf": bytes_pf},
        }  # fmt: skip
    del pos, nb, state
    # self time-correlation function, all lags
    n_part, n_frames = 20_000, 2_000
    data = torch.randn((n_part, n_frames), device=dev, generator=gen, dtype=torch.float64).cumsum(1)
    ms = timed(lambda: engine.self_time_correlation_device(data, n_frames), reps=3)
    flops = 2.0 * n_part * (n_frames * (n_frames + 1) / 2)
    tfl = flops / (ms * 1e-3) / 1e12
    out["self_time_correlation"] = {
        "workload": f"{n_part} particles x {n_frames} frames, all {n_frames} lags",
        "value": n_part * n_frames / (ms * 1e-3), "unit": UNIT, "ms": ms,
        "roofline": {"bound": "fp64", "achieved": tfl, "peak": fp64_peak_tflops,
                     "unit": "TFLOP/s",
                     "frac": tfl / fp64_peak_tflops if fp64_peak_tflops else None},
    }  # fmt: skip
    del data
    # SOAP at the cfg4 parameters (BASELINE.json configs[3]): 3 species, l_max = 10, metal density;
    # a bounded sample of the 250k-atom box (the kernels are per-centre, throughput is size-stable)
    n, nf, n_sp, l_max4 = 50_000, 4, 3, 10
    box_l = (n / 0.0857) ** (1.0 / 3.0)
    pos = (torch.rand((nf, n, 3), device=dev, generator=gen, dtype=torch.float64) * box_l).float()
    pos.clamp_(max=float(np.nextafter(np.float32(box_l), np.float32(0))))
    sp3 = (torch.arange(n, device=dev, dtype=torch.int32) * 7919 % n_sp).to(torch.int32)
    cen = torch.arange(n, dtype=torch.int32, device=dev)
    box3 = np.full(3, np.float32(box_l), dtype=np.float32)
    ms = timed(lambda: engine.soap_power_spectrum(pos, sp3, cen, box3, R_CUT_SOAP, N_MAX, l_max4,
                                                  n_species=n_sp), reps=3)  # fmt: skip
    r_s = R_CUT_SOAP + math.sqrt(-2.0 * math.log(1e-3))
    k4 = 0.0857 * 4.0 / 3.0 * math.pi * r_s**3 + 1.0
    lm4, sn4 = (l_max4 + 1) ** 2, n_sp * N_MAX
    flops = 2.0 * (k4 * N_MAX * lm4 + lm4 * sn4 * (sn4 + 1) / 2) * n * nf
    tfl = flops / (ms * 1e-3) / 1e12
    out["soap_cfg4_parameters"] = {
        "workload": f"{n} atoms x {nf} frames, 3 species, rho=0.0857/A^3, r_cut={R_CUT_SOAP}, "
                    f"n_max={N_MAX}, l_max={l_max4} (3300 features, K={k4:.0f})",
        "value": n * nf / (ms * 1e-3), "unit": UNIT, "ms": ms,
        "roofline": {"bound": "fp64", "achieved": tfl, "peak": fp64_peak_tflops,
                     "unit": "TFLOP/s",
                     "frac": tfl / fp64_peak_tflops if fp64_peak_tflops else None},
    }  # fmt: skip
    return out


def main() -> None:
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
