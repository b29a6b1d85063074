#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_soap.py tests/test_gpu_timesoap.py -q -m gpu > gpurun_out/c8_pytest.log 2>&1
echo "exit $?" >> gpurun_out/c8_pytest.log
: > gpurun_out/c8_timing.log
for t in "soap 100000 33" "fused 100000 33" "soap3 50000 8" "fused3 50000 8"; do
  timeout 300 python scripts/profile_target.py $t 2>&1 | tail -2 >> gpurun_out/c8_timing.log
done
timeout 1200 python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/c8_bench.json 2> gpurun_out/c8_bench.err
echo "bench exit $?" >> gpurun_out/c8_bench.err
tail -4 gpurun_out/c8_pytest.log; cat gpurun_out/c8_timing.log; tail -3 gpurun_out/c8_bench.err; head -c 400 gpurun_out/c8_bench.json
