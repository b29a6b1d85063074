#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_soap.py tests/test_gpu_timesoap.py -q -m gpu > gpurun_out/c9_pytest.log 2>&1
echo "exit $?" >> gpurun_out/c9_pytest.log
: > gpurun_out/c9_timing.log
for t in "soap 100000 33" "fused 100000 33" "fused 10000 200" "soap 10000 200"; do
  timeout 300 python scripts/profile_target.py $t 2>&1 | tail -1 >> gpurun_out/c9_timing.log
done
tail -4 gpurun_out/c9_pytest.log; cat gpurun_out/c9_timing.log
