"""Turn the ncu reports of scripts/gpu_profiles_r2.sh (gpurun_out/*.ncu-rep, read here with `ncu -i`)
into the committed evidence under profiles/: per-kernel summaries, the launch shares of the default
bench command, profiles/r2_traffic_bench_shape.json (what bench.py quotes as `traffic`) and the SASS
listings of the dominant kernels.

usage: python scripts/summarize_profiles.py
"""
import csv
import io
import json
import subprocess
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "profiles"
SRC = ROOT / "gpurun_out"

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__cycles_active.avg",
]


def raw_rows(rep: Path):
    txt = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True,
                         text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    return rows[0], rows[1], rows[2:]


def to_bytes(value: str, unit: str) -> float:
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    return float(value) * scale


def summarize(rep_name: str, out_name: str, title: str, per_unit: dict | None = None):
    rep = SRC / rep_name
    if not rep.exists():
        print("missing", rep)
        return {}
    hdr, units, rows = raw_rows(rep)
    ki = hdr.index("Kernel Name")
    lines = [title, ""]
    result = {}
    for r in rows:
        name = r[ki]
        lines.append(name)
        for i, h in enumerate(hdr):
            stall = "issue_stalled" in h and h.endswith("per_issue_active.ratio")
            if h in METRICS or (stall and r[i] and float(r[i]) >= 0.05):
                lines.append(f"  {h:92s} {r[i]:>18s} {units[i]}")
        rd = to_bytes(r[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")])
        wr = to_bytes(r[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
        ms = float(r[hdr.index("gpu__time_duration.sum")])
        unit_t = units[hdr.index("gpu__time_duration.sum")]
        ms *= {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}.get(unit_t, 1.0)
        inst = float(r[hdr.index("smsp__inst_executed.sum")])
        if per_unit:
            n = per_unit["units"]
            lines.append(f"  -> per {per_unit['what']}: {(rd + wr) / n:.1f} B DRAM, {inst / n:.1f} warp "
                         f"instructions, {ms * 1e6 / n:.3f} ns")
        lines.append("")
        result.setdefault(name, []).append({"dram": rd + wr, "ms": ms, "inst": inst})
    (OUT / out_name).write_text("\n".join(lines))
    print("wrote", out_name)
    return result


def launch_shares(csv_name: str, out_csv: str, out_txt: str):
    src = SRC / csv_name
    if not src.exists():
        print("missing", src)
        return
    text = src.read_text()
    (OUT / out_csv).write_text(text)
    body = text[text.index('"ID"'):]
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for row in csv.DictReader(io.StringIO(body)):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(row["Metric Unit"], 1e-6)
        k = row["Kernel Name"].split("(")[0]
        tot[k] += v
        cnt[k] += 1
    all_ms = sum(tot.values())
    lines = [f"launch shares of `python bench.py --no-extra` under ncu (the library's own kernels, first "
             f"{sum(cnt.values())} launches = warm-up + timed device steps + end-to-end steps + parity "
             f"block, {all_ms:.1f} ms of kernel time; cold-cache, serialised: compare shares, not "
             f"absolutes)", ""]
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        lines.append(f"{100 * v / all_ms:6.2f} %  {v:10.3f} ms  {cnt[k]:5d} launches  {k}")
    (OUT / out_txt).write_text("\n".join(lines) + "\n")
    print("wrote", out_txt)


def sass(pattern: str, out_name: str):
    lib = ROOT / "dynsight_b200" / "libdynsight_b200.so"
    names = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True).stdout
    chunks = names.split("\t\tFunction : ")
    for ch in chunks[1:]:
        fn = ch.split("\n", 1)[0].strip()
        if pattern in fn:
            ops = defaultdict(int)
            for ln in ch.splitlines():
                parts = ln.strip().split()
                if len(parts) > 1 and parts[0].startswith("/*") and parts[0].endswith("*/"):
                    op = parts[1] if not parts[1].startswith("@") else parts[2]
                    ops[op.split(".")[0].rstrip(";")] += 1
            hist = ", ".join(f"{k} {v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:24])
            (OUT / out_name).write_text(
                f"cuobjdump -sass of {fn} (libdynsight_b200.so, sm_100a)\nstatic instruction mix: {hist}\n"
                f"(no UTC*MMA / UTMALDG / LDTM: tcgen05 has no float64 kind; DMMA.8x8x4 is the fp64 "
                f"tensor instruction of sm_100a)\n\n" + "\t\tFunction : " + ch)
            print("wrote", out_name, "DMMA" in ch and "(DMMA present)" or "")
            return
    print("no kernel matches", pattern)


def main():
    table = {}
    r = summarize("r2_soap_fused_bench_shape.ncu-rep", "r2_soap_fused_bench_shape_summary.txt",
                  "ncu --set full --clock-control none, scripts/profile_target.py fused 100000 33 (the bench "
                  "shape: 100 000 atoms x 33 frames, one launch = 3.3e6 centre-frames, K = 93.7): "
                  "soap_kernel<8,20,0,0,0,SPLIT=false,LIST=true,SEQ=true>, tensor-path phase C + epilogue",
                  {"what": "centre-frame", "units": 3.3e6})
    for name, v in r.items():
        table["soap_kernel"] = {"atoms": 100000, "frames": 33, "dram_bytes_per_launch": v[0]["dram"],
                                "source": "r2_soap_fused_bench_shape_summary.txt (ncu --set full, "
                                          "dram__bytes_read.sum + dram__bytes_write.sum)"}
    r = summarize("r2_pair_pipeline_bench_shape.ncu-rep", "r2_pair_pipeline_bench_shape_summary.txt",
                  "ncu --set full --clock-control none, scripts/profile_target.py pair 100000 33 (bench shape, "
                  "LENS r_cut 5 A): the seven launches of one dsg_lens_direct call",
                  {"what": "particle-frame", "units": 3.3e6})
    if r:
        table["lens_pair_pipeline"] = {
            "atoms": 100000, "frames": 33,
            "dram_bytes_per_launch": sum(x["dram"] for v in r.values() for x in v),
            "source": "r2_pair_pipeline_bench_shape_summary.txt (sum over the launches of one call)"}
    r = summarize("r2_pair_pipeline_cfg3.ncu-rep", "r2_pair_pipeline_cfg3_summary.txt",
                  "ncu --set full --clock-control none, scripts/profile_target.py pair_fluid 1000000 9 (cfg3: "
                  "1e6 particles, rho* = 0.8, r_cut = 1.5): the seven launches of one dsg_lens_direct call",
                  {"what": "particle-frame", "units": 9e6})
    if r:
        table["lens_pair_pipeline_cfg3"] = {
            "atoms": 1000000, "frames": 9,
            "dram_bytes_per_launch": sum(x["dram"] for v in r.values() for x in v),
            "source": "r2_pair_pipeline_cfg3_summary.txt (sum over the launches of one call)"}
    summarize("r2_soap_cfg4_kernels.ncu-rep", "r2_soap_cfg4_kernels_summary.txt",
              "ncu --set full --clock-control none, scripts/profile_target.py fused3 50000 8 (cfg4 parameters: "
              "3 species, l_max = 10, K = 239): soap_list_kernel, soap_kernel<10,...,SPLIT,LIST>, "
              "soap_spectrum_kernel<SEQ>", {"what": "centre-frame", "units": 4e5})
    if table:
        (OUT / "r2_traffic_bench_shape.json").write_text(json.dumps(table, indent=1) + "\n")
        print("wrote r2_traffic_bench_shape.json")
    launch_shares("r2_launches_default_bench.csv", "r2_launches_default_bench.csv",
                  "r2_launch_shares_default_bench.txt")
    sass("soap_kernelILi8ELi20ELi0ELi0ELi0ELb0ELb1ELb1E", "r2_sass_soap_kernel_fused.txt")
    sass("lens_pair_kernelIfE", "r2_sass_lens_pair_kernel.txt")
    sass("soap_kernelILi10ELi21ELi13ELi5ELi0ELb1ELb1ELb0E", "r2_sass_soap_kernel_split_l10.txt")


if __name__ == "__main__":
    sys.exit(main())
