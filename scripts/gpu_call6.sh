#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_lens.py tests/test_gpu_api.py -q -m gpu -x > gpurun_out/c6_pytest.log 2>&1
echo "exit $?" >> gpurun_out/c6_pytest.log
bash scripts/gpu_variants.sh pair_fluid 1000000 9 > /dev/null
cp gpurun_out/variants_pair_fluid.log gpurun_out/c6_variants_fluid.log
bash scripts/gpu_variants.sh pair 100000 33 > /dev/null
cp gpurun_out/variants_pair.log gpurun_out/c6_variants_cfg5.log
tail -5 gpurun_out/c6_pytest.log; cat gpurun_out/c6_variants_fluid.log gpurun_out/c6_variants_cfg5.log
