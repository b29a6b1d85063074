#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_soap.py tests/test_gpu_timesoap.py tests/test_gpu_api.py -q -m gpu > gpurun_out/c10_pytest.log 2>&1
echo "exit $?" >> gpurun_out/c10_pytest.log
: > gpurun_out/c10_timing.log
for t in "fused 100000 33" "soap3 50000 8" "fused3 50000 8"; do
  timeout 300 python scripts/profile_target.py $t 2>&1 | tail -1 >> gpurun_out/c10_timing.log
done
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/c10_bench.json 2> gpurun_out/c10_bench.err
echo "bench exit $?" >> gpurun_out/c10_bench.err
tail -4 gpurun_out/c10_pytest.log; cat gpurun_out/c10_timing.log; tail -3 gpurun_out/c10_bench.err; head -c 300 gpurun_out/c10_bench.json
