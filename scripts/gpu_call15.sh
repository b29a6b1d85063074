#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_soap.py tests/test_gpu_timesoap.py -x -q > gpurun_out/c15_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c15_pytest.log
tail -5 gpurun_out/c15_pytest.log
{
echo "== default (MMA)"; python scripts/profile_target.py fused 100000 33 | tail -2
python scripts/profile_target.py soap 20000 8 | tail -1
python scripts/profile_target.py fused3 50000 8 | tail -1
echo "== nomma"; DSG_LIB_PATH=gpurun_variants/lib_nomma.so python scripts/profile_target.py fused 100000 33 | tail -2
DSG_LIB_PATH=gpurun_variants/lib_nomma.so python scripts/profile_target.py soap 20000 8 | tail -1
DSG_LIB_PATH=gpurun_variants/lib_nomma.so python scripts/profile_target.py fused3 50000 8 | tail -1
} > gpurun_out/c15_timing.log 2>&1
cat gpurun_out/c15_timing.log
