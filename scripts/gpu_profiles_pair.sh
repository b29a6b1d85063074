P=gpurun_out
python scripts/profile_target.py pair 100000 33 > $P/p_pair_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"cell_assign|scan_|cell_scatter|frame_finalize|lens_pair|pair_finalize" -s 14 -c 7 \
    -f -o $P/r2_pair_pipeline_bench_shape python scripts/profile_target.py pair 100000 33 > $P/p_pair_ncu.log 2>&1
python scripts/profile_target.py pair_fluid 1000000 9 > $P/p_pairf_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"cell_assign|scan_|cell_scatter|frame_finalize|lens_pair|pair_finalize" -s 14 -c 7 \
    -f -o $P/r2_pair_pipeline_cfg3 python scripts/profile_target.py pair_fluid 1000000 9 > $P/p_pairf_ncu.log 2>&1
ls -la $P/r2_pair_pipeline*.ncu-rep
