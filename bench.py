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
API (``dynsight_b200.pipeline.descriptors_from_trajectory``: LENS + coordination numbers +
timeSOAP in one pass over the frames) with pinned HOST positions in and host results out.  After
the timed region the outputs of the last step are compared with the CPU oracle (``parity`` in the
JSON line; a mismatch exits non-zero).  One JSON line is printed by rank 0.
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
    p.add_argument("--no-parity", action="store_true", help="skip the oracle check of the outputs")
    p.add_argument("--strong-frames", type=int, default=256,
                   help="frames of the fixed trajectory of the strong-scaling section (N > 1)")
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
            "each step reads a different resident frame block; the working set of a step "
            "(neighbour rows 3 GB, sorted coordinates, rings) exceeds L2 (126 MB)"
        ),
    }


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def traffic_for(kernel: str, atoms: int, frames: int) -> tuple[float | None, str | None]:
    """dram__bytes_read + dram__bytes_write of one launch from the committed ncu capture of this
    very shape (profiles/r2_traffic_bench_shape.json); None when the run uses another shape."""
    path = ROOT / "profiles" / "r2_traffic_bench_shape.json"
    if not path.exists():
        return None, None
    table = json.loads(path.read_text())
    entry = table.get(kernel)
    if not entry or entry.get("atoms") != atoms or entry.get("frames") != frames:
        return None, None
    return float(entry["dram_bytes_per_launch"]), f"{path.name}: {entry['source']}"


def parity_check(engine, block, box, lens_out, nc_out, species, ts_out, n_sample: int) -> dict:
    """The outputs of the last timed step against the CPU oracle (the checker, not the product):
    LENS and coordination numbers of the first two frame pairs for ALL atoms, bit for bit;
    timeSOAP of the first pair on ``n_sample`` evenly spaced centres (1e-5 t + 1e-7), and -- the
    timed step writes no spectra -- the SOAP vectors of those centres from one extra call of the
    same kernels (1e-8 max|p|)."""
    from oracle import lens_oracle, soap_oracle, timesoap_oracle

    n_atoms = block.shape[1]
    hpos = block[:3].cpu().numpy()
    dims = np.tile(np.concatenate([box, [90, 90, 90]]).astype(np.float32), (3, 1))
    lens_oracle.set_num_threads(os.cpu_count() or 1)
    want_lens = lens_oracle.compute_lens_arrays(hpos, dims, R_CUT_LENS, delay=DELAY)
    want_nc = lens_oracle.coord_number_arrays(hpos, dims, R_CUT_LENS)
    got_lens = lens_out[:, :2].cpu().numpy()
    got_nc = nc_out[:, :3].cpu().numpy()
    sample = np.linspace(0, n_atoms - 1, n_sample).astype(np.int32)
    basis = soap_oracle.gto_basis(R_CUT_SOAP, N_MAX, L_MAX)
    z = np.full(n_atoms, 8)
    want_soap = np.stack(
        [soap_oracle.soap_frame_c(hpos[f], z, sample, box.astype(np.float64), R_CUT_SOAP, N_MAX,
                                  L_MAX, basis, n_threads=os.cpu_count() or 1) for f in range(2)],
        axis=1,
    )  # fmt: skip
    torch = __import__("torch")
    idx = torch.as_tensor(sample.astype(np.int64), device=ts_out.device)
    got_soap = engine.soap_power_spectrum(
        block[:2].contiguous(), species, torch.as_tensor(sample, device=ts_out.device), box,
        R_CUT_SOAP, N_MAX, L_MAX, n_species=1,
    ).cpu().numpy()  # fmt: skip
    rel = float((np.abs(got_soap - want_soap) / np.abs(want_soap).max(axis=-1, keepdims=True)).max())
    want_ts = timesoap_oracle.timesoap(want_soap, DELAY)[:, 0]
    got_ts = ts_out[idx, 0].cpu().numpy()
    ts_err = np.abs(got_ts - want_ts)
    res = {
        "checked_against": "oracle/ (C restatement of the numba kernels, C/numpy restatement of "
                           "DScribe SOAP-GTO, numpy timeSOAP)",
        "lens_bit_exact": bool(np.array_equal(got_lens, want_lens)),
        "coord_number_bit_exact": bool(np.array_equal(got_nc, want_nc)),
        "lens_values_checked": int(want_lens.size), "lens_mean": float(want_lens.mean()),
        "soap_centres_checked": int(n_sample) * 2, "soap_max_rel_err": rel, "soap_tol": 1e-8,
        "timesoap_max_abs_err": float(ts_err.max()),
        "timesoap_ok": bool(np.all(ts_err <= 1e-5 * np.abs(want_ts) + 1e-7)),
    }  # fmt: skip
    res["ok"] = bool(res["lens_bit_exact"] and res["coord_number_bit_exact"] and rel < 1e-8
                     and res["timesoap_ok"])  # fmt: skip
    return res


def run_b200(args: argparse.Namespace) -> None:
    import torch
    import torch.distributed as dist

    from dynsight_b200 import _lib, engine, pipeline
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
    lens_out = torch.empty((n_atoms, bsz), dtype=torch.float64, device=dev)
    ts_out = torch.empty((n_atoms, bsz), dtype=torch.float64, device=dev)
    nc_out = torch.empty((n_atoms, bsz + 1), dtype=torch.float64, device=dev)
    block = torch.empty((bsz + 1, n_atoms, 3), dtype=torch.float32, device=dev)
    halo_recv = torch.empty((DELAY, n_atoms, 3), dtype=torch.float32, device=dev)
    if not engine.lens_direct_supported(bsz + 1, n_atoms, box, R_CUT_LENS, True):
        raise SystemExit("bench.py: this shape is outside the pair path of the LENS kernels")

    ev = {k: [] for k in ("lens", "soap_timesoap")}

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
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if record else None
        if marks:
            marks[0].record()
        # LENS + coordination numbers: pair path (half-shell search, partner-frame lookup)
        _, counts = engine.lens_direct(block, box, R_CUT_LENS, DELAY, True, out=lens_out, col0=0)
        engine.coordination_numbers(counts, out=nc_out, col0=0)
        if marks:
            marks[1].record()
        # SOAP fused with timeSOAP: every spectrum is folded into the distance to the previous
        # frame's inside the spectrum epilogue; no (N, F, C) tensor is written
        engine.soap_timesoap(block, species, centres, box, R_CUT_SOAP, N_MAX, L_MAX, delay=DELAY,
                             n_species=1, out=ts_out, col0=0)  # fmt: skip
        if marks:
            marks[2].record()
            for k, name in enumerate(("lens", "soap_timesoap")):
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
    part_ms["soap"] = part_ms["soap_timesoap"]   # one fused kernel chain
    soap_tflops = fl * soap_pf / (part_ms["soap"] * 1e-3) / 1e12
    # second view: every exponential is 8 fp64 operations (2 range reduction, 5 polynomial, 1 table)
    soap_tflops_exp = (fl + 2.0 * 8.0 * ex) * soap_pf / (part_ms["soap"] * 1e-3) / 1e12
    lens_bytes_pf = 20 + 8 * (k_mean_lens + 1)
    lens_gbs = lens_bytes_pf * pf_per_step / (part_ms["lens"] * 1e-3) / 1e9
    del part_ms["soap"]
    soap_ms = part_ms["soap_timesoap"]
    soap_traffic, soap_traffic_src = traffic_for("soap_kernel", n_atoms, bsz + 1)
    lens_traffic, lens_traffic_src = traffic_for("lens_pair_pipeline", n_atoms, bsz + 1)
    roofline = {
        "kernel": "soap_kernel<8,20,0,0,0,SPLIT=false,LIST=true,SEQ=true> (tensor path: phase C, "
                  "Loewdin mixing and the power spectrum as DMMA.8x8x4 products, timeSOAP folded into "
                  "the epilogue; + soap_list_kernel, binning)",
        "pipe_note": "DMMA.8x8x4 (mma.sync.m8n8k4.f64) and DFMA share the fp64 pipe at 32 flop/clk/SMSP "
                     "(scripts/probes/dmma_issue_probe.cu), so the DFMA probe is the peak of both; tile "
                     "padding, harmonics and exponentials are not counted as flops",
        "bound": "fp64",
        "achieved": soap_tflops,
        "peak": fp64_peak_tflops,
        "unit": "TFLOP/s",
        "frac": soap_tflops / fp64_peak_tflops if fp64_peak_tflops else None,
        "traffic": soap_traffic,
        "traffic_source": soap_traffic_src,
        "peak_source": "dsg_fp64_fma_probe (csrc/core.cu: 16 independent DFMA chains per thread, "
                       "148 x 8 blocks x 256 threads), best of 4 in this run; MEASURED_PEAKS.json "
                       "has no fp64 figure (B200 nominal 37-40 TFLOP/s)",
        "algorithmic_flops_per_pf": fl,
        "exps_per_pf": ex,
        "frac_counting_exps_as_16_flops": (soap_tflops_exp / fp64_peak_tflops
                                           if fp64_peak_tflops else None),
        "mean_neighbours_incl_self": k_soap,
        "share_of_step": soap_ms / sum(part_ms.values()),
        "hbm_view": {
            "note": "fused: 12 B of positions in, 8 B of timeSOAP out per particle-frame; the "
                    "2592 B spectrum goes to the warp's ring (L2) instead of HBM",
            "unfused_bytes_per_pf": 12 + 2 * 8 * n_feat + 8, "fused_bytes_per_pf": 20,
            "peak_gbs": hbm_peak,
        },
    }  # fmt: skip
    rooflines_hbm = [
        {
            "kernel": "cell_assign + scan + cell_scatter_local + lens_pair_kernel + "
                      "pair_finalize_kernel (+ counts_to_f64)",
            "bound": "hbm", "achieved": lens_gbs, "peak": hbm_peak, "unit": "GB/s",
            "frac": lens_gbs / hbm_peak, "traffic": lens_traffic,
            "traffic_source": lens_traffic_src,
            "algorithmic_bytes_per_pf": lens_bytes_pf, "mean_neighbours": k_mean_lens,
            "ms_per_step": part_ms["lens"], "peak_source": hbm_src,
        },
    ]  # fmt: skip

    # ---- end to end through the public API: pinned host positions in, host arrays out ----
    host_block = torch.empty((bsz + 1, n_atoms, 3), dtype=torch.float32, pin_memory=True)
    host_block.copy_(block)
    dims = np.tile(np.concatenate([box, [90, 90, 90]]).astype(np.float32), (bsz + 1, 1))
    uni = ArrayUniverse(host_block.numpy(), dims, names=["O"] * n_atoms, types=["O"] * n_atoms)

    def e2e_step() -> int:
        res = pipeline.descriptors_from_trajectory(
            uni, lens_r_cut=R_CUT_LENS, soap_r_cut=R_CUT_SOAP, soapnmax=N_MAX, soaplmax=L_MAX,
            delay=DELAY, coord_number=True,
        )
        return sum(v.nbytes for v in res.values())

    # warm-up (the same W as the device-timed loop, at least 3): the page-locked result buffers
    # enter the host allocator's cache -- a cudaHostAlloc of 77 MB inside the timed region costs
    # 10 ms alone and 40-60 ms when the eight processes of a box allocate at the same time
    for _ in range(max(3, args.warmup)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(args.steps):
        t1 = time.perf_counter()
        d2h = e2e_step()
        if os.environ.get("DSG_BENCH_DEBUG"):
            print(f"rank {rank} e2e step {1e3 * (time.perf_counter() - t1):.2f} ms", file=sys.stderr)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / args.steps
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = {
        "value": world * pf_per_step / e2e_s,
        "unit": UNIT,
        "h2d_bytes_per_step": host_block.numel() * 4,
        "d2h_bytes_per_step": d2h,
        "api": "dynsight_b200.pipeline.descriptors_from_trajectory (LENS + coordination numbers + "
               "timeSOAP, one pass over the frames)",
        "ms_per_step": e2e_s * 1e3,
        "steps_timed": args.steps,
    }

    # ---- parity of what the device-timed loop computed last (rank 0; the oracle is the checker
    # only).  It runs after the end-to-end loop: the oracle's 32 OpenMP threads and its host arrays
    # on rank 0 alone made that rank's end-to-end steps 10 ms slower than the other seven.
    parity = None
    if rank == 0 and not args.no_parity:
        parity = parity_check(engine, block, box, lens_out, nc_out, species, ts_out, 1024)

    strong = None
    if world > 1:
        strong = strong_scaling_section(args, dist, dev, rank, world)

    extra = None
    widened = None
    cfg4 = None
    small = None
    if rank == 0 and not args.no_extra:
        del pool
        torch.cuda.empty_cache()
        extra = lens_cfg3_section(engine, dev, hbm_peak, hbm_src)
        cfg4 = soap_cfg4_section(engine, dev, hbm_peak, fp64_peak_tflops, args.no_parity)
        small = small_config_section(dev)
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
            "reference_numba": reference_numba_figure(),
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
            "parity": parity,
            "roofline": roofline,
            "rooflines_hbm": rooflines_hbm,
            "breakdown_ms_per_step": part_ms,
            "breakdown_pf_per_s": {
                "lens_from_positions": pf_per_step / (part_ms["lens"] * 1e-3),
                "soap_timesoap_fused": soap_pf / (soap_ms * 1e-3),
            },
            "cpu_baseline": cpu_baseline,
            "strong_scaling": strong,
            "lens_cfg3": extra,
            "soap_cfg4": cfg4,
            "small_configs": small,
            "widened_8f": widened,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if parity is not None and not parity["ok"]:
        raise SystemExit("bench.py: the timed outputs differ from the oracle -- see 'parity'")
    if cfg4 is not None and cfg4.get("parity") and not cfg4["parity"]["ok"]:
        raise SystemExit("bench.py: cfg4 outputs differ from the oracle -- see 'soap_cfg4.parity'")


def reference_numba_figure() -> dict | None:
    """The reference's own numba LENS kernels (lens.py:69-189, imported by path, unmodified) timed
    in the BUILD container at this shape -- /root/reference does not exist on the GPU box, so the
    figure is static and labelled (profiles/r2_reference_numba_lens.json, scripts/time_reference_numba.py)."""
    path = ROOT / "profiles" / "r2_reference_numba_lens.json"
    if not path.exists():
        return None
    fig = json.loads(path.read_text())
    fig["note"] = "static: measured in the build container, not on this host"
    return fig


def strong_scaling_section(args, dist, dev, rank: int, world: int) -> dict:
    """One FIXED trajectory through the user-facing sharded drivers
    (``parallel.compute_lens_sharded`` + ``parallel.timesoap_sharded``, results gathered on rank 0):
    total work independent of N.  Every rank generates the same host trajectory from one seed and
    reads its own frame block (+ the halo frame) from it; rank 0 checks a sub-slice against a
    single-GPU run (LENS bit-equal, timeSOAP to 1e-12)."""
    import torch

    from dynsight_b200 import lens as dlens
    from dynsight_b200 import parallel, soap as dsoap
    from dynsight_b200.universe import ArrayUniverse

    n_atoms, n_frames = args.atoms, args.strong_frames
    pos, box = host_trajectory(n_atoms, n_frames, seed=99)
    pinned = torch.empty(pos.shape, dtype=torch.float32, pin_memory=True)
    pinned.copy_(torch.from_numpy(pos))
    dims = np.tile(np.concatenate([box, [90, 90, 90]]).astype(np.float32), (n_frames, 1))
    uni = ArrayUniverse(pinned.numpy(), dims, names=["O"] * n_atoms, types=["O"] * n_atoms)

    def run() -> tuple:
        lens = parallel.compute_lens_sharded(uni, R_CUT_LENS, delay=DELAY)
        ts = parallel.timesoap_sharded(uni, R_CUT_SOAP, N_MAX, L_MAX, delay=DELAY)
        return lens, ts

    run()  # warm-up (allocator, NCCL channels)
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    lens, ts = run()
    torch.cuda.synchronize()
    dist.barrier()
    secs = time.perf_counter() - t0
    t = torch.tensor([secs], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    secs = float(t.item())
    check = None
    if rank == 0:
        sl = slice(0, 9)
        ref_l = dlens.compute_lens_device(uni, R_CUT_LENS, DELAY, trajslice=sl)
        ref_t = dsoap.timesoap_from_trajectory(uni, R_CUT_SOAP, N_MAX, L_MAX, DELAY, trajslice=sl)
        # the last block boundary as well: pairs that straddle two ranks
        b = n_frames * (world - 1) // world
        sl2 = slice(b - 4, b + 5)
        ref_l2 = dlens.compute_lens_device(uni, R_CUT_LENS, DELAY, trajslice=sl2)
        ref_t2 = dsoap.timesoap_from_trajectory(uni, R_CUT_SOAP, N_MAX, L_MAX, DELAY, trajslice=sl2)
        check = {
            "lens_bit_equal": bool(torch.equal(lens[:, :8], ref_l)
                                   and torch.equal(lens[:, b - 4 : b + 4], ref_l2)),
            "timesoap_max_abs_diff": float(max((ts[:, :8] - ref_t).abs().max().item(),
                                               (ts[:, b - 4 : b + 4] - ref_t2).abs().max().item())),
            "shape": [list(lens.shape), list(ts.shape)],
        }  # fmt: skip
        check["ok"] = bool(check["lens_bit_equal"] and check["timesoap_max_abs_diff"] <= 1e-12)
    pf = n_atoms * (n_frames - DELAY)
    return {
        "workload": f"fixed trajectory: {n_atoms} atoms x {n_frames} frames, LENS + timeSOAP through "
                    "parallel.compute_lens_sharded + parallel.timesoap_sharded, host positions in, "
                    "results gathered on rank 0",
        "scaling": "strong", "n_gpus": world, "seconds": secs, "value": pf / secs, "unit": UNIT,
        "check_vs_single_gpu": check,
    }  # fmt: skip


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
    pf = n * (n_fr - 1)

    def timed(fn) -> float:
        times = []
        for it in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(pools[it % 2])
            e1.record()
            torch.cuda.synchronize()
            if it >= 2:
                times.append(e0.elapsed_time(e1))
        return float(np.mean(times))

    state = {}

    def pair_path(p) -> None:
        state["counts"] = engine.lens_direct(p, box, 1.5, 1, True, out=out, col0=0)[1]

    def rows_path(p) -> None:
        engine.lens_from_positions(p, box, 1.5, 1, out=out, col0=0)

    ms = timed(pair_path)
    res_pair = out.clone()
    ms_rows = timed(rows_path)
    same = bool(torch.equal(res_pair, out))
    k_mean = float(state["counts"].float().mean().item())
    bytes_pf = 20 + 8 * (k_mean + 1)
    gbs = bytes_pf * pf / (ms * 1e-3) / 1e9
    traffic, traffic_src = traffic_for("lens_pair_pipeline_cfg3", n, n_fr)
    return {
        "workload": "cfg3 (BASELINE.json configs[2]): 1,000,000 particles, rho*=0.8, r_cut=1.5, "
                    f"{n_fr} frames per launch batch (LENS delay 1)",
        "value": pf / (ms * 1e-3), "unit": UNIT, "ms_per_batch": ms,
        "kernels": "pair path: cell_assign, scan, cell_scatter_local, lens_pair_kernel, "
                   "pair_finalize_kernel",
        "mean_neighbours": k_mean,
        "rows_path": {"value": pf / (ms_rows * 1e-3), "ms_per_batch": ms_rows,
                      "kernels": "neigh_rows_thread_kernel + lens_rows_thread_kernel (what "
                                 "centres != selection still uses)",
                      "bit_equal_to_pair_path": same},
        "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                     "frac": gbs / hbm_peak, "traffic": traffic, "traffic_source": traffic_src,
                     "algorithmic_bytes_per_pf": bytes_pf, "peak_source": hbm_src},
    }  # fmt: skip


def soap_cfg4_section(engine, dev, hbm_peak: float, fp64_peak_tflops: float, no_parity: bool
                      ) -> dict:  # fmt: skip
    """BASELINE.json configs[3] at its stated size: 250,000 atoms, 3 species, r_cut = 5 A, n_max = 8,
    l_max = 10 (3 300 features), SOAP streamed into timeSOAP over 33 frames -- the spectra of the
    trajectory (218 GB) never exist.  Sampled centres are compared with the oracle."""
    import torch

    n, n_sp, l_max4, n_frames = 250_000, 3, 10, 33
    rho4 = 0.0857
    box_l = (n / rho4) ** (1.0 / 3.0)
    gen = torch.Generator(device=dev).manual_seed(44)
    top = float(np.nextafter(np.float32(box_l), np.float32(0)))
    cur = torch.rand((n, 3), device=dev, generator=gen, dtype=torch.float64) * box_l
    pos = torch.empty((n_frames, n, 3), dtype=torch.float32, device=dev)
    for f in range(n_frames):
        cur = torch.remainder(cur + 0.05 * torch.randn((n, 3), device=dev, generator=gen,
                                                       dtype=torch.float64), box_l)  # fmt: skip
        pos[f] = cur.to(torch.float32).clamp_(max=top)
    sp3 = (torch.arange(n, device=dev, dtype=torch.int64) * 7919 % n_sp).to(torch.int32)
    cen = torch.arange(n, dtype=torch.int32, device=dev)
    box3 = np.full(3, np.float32(box_l), dtype=np.float32)
    ts_out = torch.empty((n, n_frames - DELAY), dtype=torch.float64, device=dev)
    # primitive sums between the accumulate and the spectrum kernel: 23 KB per centre-frame; the
    # spectra (26.4 KB per centre-frame) only ever exist in the warps' rings
    batch = 17   # 2 batches of 17 frames (one shared): 5.8 GB of primitive sums per frame

    def run() -> int:
        done = 0
        start = 0
        while start < n_frames - DELAY:
            stop = min(n_frames, start + batch)
            engine.soap_timesoap(pos[start:stop], sp3, cen, box3, R_CUT_SOAP, N_MAX, l_max4,
                                 delay=DELAY, n_species=n_sp, out=ts_out, col0=start)  # fmt: skip
            done += stop - start
            start = stop - DELAY
        return done

    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    soap_frames = run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    r_s = R_CUT_SOAP + math.sqrt(-2.0 * math.log(1e-3))
    k4 = rho4 * 4.0 / 3.0 * math.pi * r_s**3 + 1.0
    lm4, sn4 = (l_max4 + 1) ** 2, n_sp * N_MAX
    flops_pf = 2.0 * (k4 * N_MAX * lm4 + lm4 * sn4 * (sn4 + 1) / 2)
    tfl = flops_pf * n * soap_frames / (ms * 1e-3) / 1e12
    n_feat4 = sn4 * (sn4 + 1) // 2 * (l_max4 + 1)
    parity = None
    if not no_parity:
        from oracle import soap_oracle, timesoap_oracle

        sample = np.arange(0, n, n // 48, dtype=np.int32)
        hpos = pos[:2].cpu().numpy()
        z = np.asarray([26, 28, 29])[sp3.cpu().numpy()]   # Fe, Ni, Cu: ascending atomic numbers
        basis = soap_oracle.gto_basis(R_CUT_SOAP, N_MAX, l_max4)
        want = np.stack(
            [soap_oracle.soap_frame_c(hpos[f], z, sample, box3.astype(np.float64), R_CUT_SOAP,
                                      N_MAX, l_max4, basis, n_threads=os.cpu_count() or 1)
             for f in range(2)], axis=1,
        )  # fmt: skip
        got = engine.soap_power_spectrum(pos[:2], sp3, torch.as_tensor(sample, device=dev), box3,
                                         R_CUT_SOAP, N_MAX, l_max4, n_species=n_sp).cpu().numpy()  # fmt: skip
        rel = float((np.abs(got - want) / np.abs(want).max(axis=-1, keepdims=True)).max())
        want_t = timesoap_oracle.timesoap(want, DELAY)[:, 0]
        got_t = ts_out[:: n // 48, 0].cpu().numpy()
        t_err = np.abs(got_t - want_t)
        parity = {"soap_centres_checked": int(sample.size) * 2, "soap_max_rel_err": rel,
                  "timesoap_max_abs_err": float(t_err.max()),
                  "ok": bool(rel < 1e-8 and np.all(t_err <= 1e-5 * np.abs(want_t) + 1e-7))}  # fmt: skip
    return {
        "workload": f"cfg4 (BASELINE.json configs[3]): {n} atoms, {n_sp} species, rho={rho4}/A^3, "
                    f"r_cut={R_CUT_SOAP}, n_max={N_MAX}, l_max={l_max4} ({n_feat4} features, "
                    f"K={k4:.0f}), {n_frames} frames through the fused SOAP -> timeSOAP kernels in "
                    f"batches of {batch} frames (the 218 GB of spectra never exist)",
        "value": n * (n_frames - DELAY) / (ms * 1e-3), "unit": UNIT,
        "value_note": "timeSOAP particle-frames per second including the SOAP of every frame "
                      f"(batches overlap by one frame: {soap_frames} SOAP frames for "
                      f"{n_frames - DELAY} pairs)",
        "soap_pf_per_s": n * soap_frames / (ms * 1e-3), "ms_total": ms,
        "roofline": {"bound": "fp64", "achieved": tfl, "peak": fp64_peak_tflops,
                     "unit": "TFLOP/s",
                     "frac": tfl / fp64_peak_tflops if fp64_peak_tflops else None,
                     "algorithmic_flops_per_pf": flops_pf,
                     "hbm_view_gbs": (12 + 8 * n_feat4) * n * soap_frames / (ms * 1e-3) / 1e9,
                     "hbm_peak_gbs": hbm_peak},
        "parity": parity,
    }  # fmt: skip


def small_config_section(dev) -> dict:
    """The small-N, launch-latency regime through the public API (per-call wall time, host arrays
    in and out): cfg1 (BASELINE.json configs[0]) and cfg2 (configs[1]), plus XTC ingest."""
    import tempfile

    import torch

    from dynsight_b200 import io as dio
    from dynsight_b200 import soap as dsoap
    from dynsight_b200.trajectory import Trj
    from dynsight_b200.universe import ArrayUniverse

    out: dict = {}
    rng = np.random.default_rng(0)

    def wall(fn, reps: int = 3) -> float:
        fn()
        ts = []
        for _ in range(reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        return float(np.min(ts))

    # cfg1: 2,048-particle LJ fluid, 200 frames, r_cut = 1.5 sigma
    n, nf = 2048, 200
    box_l = (n / 0.8) ** (1 / 3)
    cur = rng.random((n, 3)) * box_l
    pos = np.empty((nf, n, 3), dtype=np.float32)
    for f in range(nf):
        cur = np.mod(cur + rng.normal(0, 0.05, cur.shape), box_l)
        pos[f] = np.minimum(cur, np.nextafter(np.float32(box_l), np.float32(0)))
    dims = np.tile(np.asarray([box_l] * 3 + [90] * 3, dtype=np.float32), (nf, 1))
    trj = Trj(ArrayUniverse(pos, dims, ["A"] * n, ["A"] * n))
    s_lens = wall(lambda: trj.get_lens(r_cut=1.5))
    s_nc = wall(lambda: trj.get_coord_number(r_cut=1.5))
    out["cfg1"] = {
        "workload": f"{n} particles x {nf} frames, rho*=0.8, r_cut=1.5 through Trj.get_lens / "
                    "Trj.get_coord_number",
        "get_lens_ms": s_lens * 1e3, "get_coord_number_ms": s_nc * 1e3,
        "lens_pf_per_s": n * (nf - 1) / s_lens, "coord_number_pf_per_s": n * nf / s_nc,
    }  # fmt: skip
    # cfg2: 10,000-atom water-oxygen box, 1,000 frames, SOAP(5 A, 8, 8) + timeSOAP
    n, nf = 10_000, 1000
    box_l = (n / RHO) ** (1 / 3)
    cur = rng.random((n, 3)) * box_l
    pos = np.empty((nf, n, 3), dtype=np.float32)
    for f in range(nf):
        cur = np.mod(cur + rng.normal(0, 0.1, cur.shape), box_l)
        pos[f] = np.minimum(cur, np.nextafter(np.float32(box_l), np.float32(0)))
    dims = np.tile(np.asarray([box_l] * 3 + [90] * 3, dtype=np.float32), (nf, 1))
    uni = ArrayUniverse(pos, dims, ["O"] * n, ["O"] * n)
    s_ts = wall(lambda: dsoap.timesoap_from_trajectory(uni, R_CUT_SOAP, N_MAX, L_MAX, as_numpy=True),
                reps=2)  # fmt: skip
    out["cfg2"] = {
        "workload": f"{n} atoms x {nf} frames, SOAP(5 A, 8, 8) + timeSOAP through "
                    "soap.timesoap_from_trajectory (SOAP never materialised: 25.9 GB)",
        "seconds": s_ts, "value": n * (nf - 1) / s_ts, "unit": UNIT,
    }  # fmt: skip
    # XTC ingest: compressed synthetic file (tests/xtc_writer.py), native threaded decoder
    try:
        sys.path.insert(0, str(ROOT / "tests"))
        import xtc_writer

        n_mol, nf = 20_000, 64
        with tempfile.TemporaryDirectory() as tmp:
            path = Path(tmp) / "synthetic.xtc"
            n_bytes = xtc_writer.write_synthetic_waters(path, n_mol, nf, seed=1)
            s_dec = wall(lambda: dio.read_xtc(path), reps=3)
            pos_x, _ = dio.read_xtc(path)
            pinned = torch.empty(pos_x.shape, dtype=torch.float32, pin_memory=True)

            def decode_and_upload() -> None:
                p, _ = dio.read_xtc(path)
                pinned.copy_(torch.from_numpy(p))
                pinned.to(dev, non_blocking=True)

            s_up = wall(decode_and_upload, reps=3)
        out["xtc_ingest"] = {
            "workload": f"{3 * n_mol} atoms x {nf} frames, compressed XTC ({n_bytes / 1e6:.1f} MB), "
                        f"{os.cpu_count()} host threads",
            "decode_frames_per_s": nf / s_dec, "decode_mb_per_s_compressed": n_bytes / 1e6 / s_dec,
            "decode_mb_per_s_positions": pos_x.nbytes / 1e6 / s_dec,
            "decode_plus_h2d_frames_per_s": nf / s_up,
        }  # fmt: skip
    except Exception as exc:  # noqa: BLE001  (the ingest line is informative only)
        out["xtc_ingest"] = {"unavailable": repr(exc)}
    return out


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
