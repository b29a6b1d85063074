#!/bin/bash
# round 2 evidence: ncu launch list of the default bench command, --set full of the dominant kernels
# at the bench's own shapes (numbers under ncu are never bench values)
mkdir -p gpurun_out
P=gpurun_out
python scripts/profile_target.py fused 100000 33 > $P/p_fused_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:soap_kernel -s 2 -c 1 \
    -o $P/r2_soap_fused_bench_shape python scripts/profile_target.py fused 100000 33 > $P/p_fused_ncu.log 2>&1
python scripts/profile_target.py pair 100000 33 > $P/p_pair_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"cell_assign|scan_|cell_scatter|frame_finalize|lens_pair|pair_finalize" -s 14 -c 7 \
    -o $P/r2_pair_pipeline_bench_shape python scripts/profile_target.py pair 100000 33 > $P/p_pair_ncu.log 2>&1
python scripts/profile_target.py pair_fluid 1000000 9 > $P/p_pairf_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"cell_assign|scan_|cell_scatter|frame_finalize|lens_pair|pair_finalize" -s 14 -c 7 \
    -o $P/r2_pair_pipeline_cfg3 python scripts/profile_target.py pair_fluid 1000000 9 > $P/p_pairf_ncu.log 2>&1
python scripts/profile_target.py fused3 50000 8 > $P/p_fused3_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"soap_kernel|soap_spectrum|soap_list" -s 6 -c 6 \
    -o $P/r2_soap_cfg4_kernels python scripts/profile_target.py fused3 50000 8 > $P/p_fused3_ncu.log 2>&1
python bench.py --no-extra > $P/p_bench_plain.json 2> $P/p_bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none \
    -k regex:"soap|lens_|pair_|cell_|scan_|frame_finalize|counts_to|neigh_|timesoap|fp64_fma" -c 1500 --csv \
    --log-file $P/r2_launches_default_bench.csv python bench.py --no-extra > $P/p_bench_ncu.log 2>&1
ls -la $P/*.ncu-rep | tail -5
