#!/bin/bash
# round 2, GPU call 1: pair-path parity + timings, SOAP shared-load probe, ncu of the pair kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_lens.py -x -q -m gpu -k "pair_path or cfg3_size or cfg5_size or delay_zero" > gpurun_out/c1_pytest_pair.log 2>&1
echo "pair tests exit $?" >> gpurun_out/c1_pytest_pair.log
for v in default pair4 pair5 pair8; do
  if [ $v = default ]; then unset DSG_LIB_PATH; else export DSG_LIB_PATH=$PWD/gpurun_variants/lib_$v.so; fi
  echo "== $v" >> gpurun_out/c1_timing.log
  timeout 300 python scripts/profile_target.py pair_fluid 1000000 9 >> gpurun_out/c1_timing.log 2>&1
  timeout 300 python scripts/profile_target.py pair 100000 33 >> gpurun_out/c1_timing.log 2>&1
done
unset DSG_LIB_PATH
echo "== rows path" >> gpurun_out/c1_timing.log
timeout 300 python scripts/profile_target.py neigh_fluid 1000000 9 >> gpurun_out/c1_timing.log 2>&1
timeout 300 python scripts/profile_target.py neigh 100000 33 >> gpurun_out/c1_timing.log 2>&1
echo "== soap default" >> gpurun_out/c1_timing.log
timeout 300 python scripts/profile_target.py soap 100000 8 >> gpurun_out/c1_timing.log 2>&1
echo "== soap halfloads probe" >> gpurun_out/c1_timing.log
DSG_LIB_PATH=$PWD/gpurun_variants/lib_soap_halfloads.so timeout 300 python scripts/profile_target.py soap 100000 8 >> gpurun_out/c1_timing.log 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/c1_pytest_all.log 2>&1
echo "all tests exit $?" >> gpurun_out/c1_pytest_all.log
python scripts/profile_target.py pair_fluid 1000000 9 > gpurun_out/c1_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:lens_pair_kernel -s 1 -c 1 \
    -o gpurun_out/r2_pair_kernel python scripts/profile_target.py pair_fluid 1000000 9 > gpurun_out/c1_ncu.log 2>&1
tail -3 gpurun_out/c1_pytest_pair.log; cat gpurun_out/c1_timing.log; tail -5 gpurun_out/c1_pytest_all.log
