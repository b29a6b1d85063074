#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/c7_pytest.log 2>&1
echo "exit $?" >> gpurun_out/c7_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c7_smoke.log 2>&1
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/c7_bench.json 2> gpurun_out/c7_bench.err
echo "bench exit $?" >> gpurun_out/c7_bench.err
tail -25 gpurun_out/c7_pytest.log; cat gpurun_out/c7_smoke.log; tail -5 gpurun_out/c7_bench.err; head -c 600 gpurun_out/c7_bench.json
