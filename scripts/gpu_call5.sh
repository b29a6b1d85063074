#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/c5_pytest.log 2>&1
echo "exit $?" >> gpurun_out/c5_pytest.log
for t in "pair_fluid 1000000 9" "pair 100000 33"; do
  timeout 300 python scripts/profile_target.py $t 2>&1 | tail -1 >> gpurun_out/c5_timing.log
done
python scripts/profile_target.py pair_fluid 1000000 9 > gpurun_out/c5_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"lens_pair_kernel|cell_assign|cell_scatter_local|pair_finalize|scan_" -s 7 -c 7 \
    -o gpurun_out/r2_pair_pipeline python scripts/profile_target.py pair_fluid 1000000 9 > gpurun_out/c5_ncu.log 2>&1
tail -15 gpurun_out/c5_pytest.log; cat gpurun_out/c5_timing.log; tail -3 gpurun_out/c5_ncu.log
