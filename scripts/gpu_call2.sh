#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_soap.py tests/test_gpu_timesoap.py tests/test_gpu_api.py -x -q -m gpu > gpurun_out/c2_pytest.log 2>&1
echo "exit $?" >> gpurun_out/c2_pytest.log
timeout 300 python scripts/profile_target.py soap 100000 8 > gpurun_out/c2_timing.log 2>&1
timeout 300 python scripts/profile_target.py soap3 50000 4 >> gpurun_out/c2_timing.log 2>&1
timeout 300 python scripts/fuzz_soap.py >> gpurun_out/c2_fuzz.log 2>&1
tail -3 gpurun_out/c2_pytest.log; cat gpurun_out/c2_timing.log; tail -5 gpurun_out/c2_fuzz.log
