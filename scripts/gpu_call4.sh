#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/c4_pytest.log 2>&1
echo "exit $?" >> gpurun_out/c4_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c4_smoke.log 2>&1
echo "exit $?" >> gpurun_out/c4_smoke.log
for t in "pair_fluid 1000000 9" "pair 100000 33" "soap 100000 8"; do
  timeout 300 python scripts/profile_target.py $t 2>&1 | tail -1 >> gpurun_out/c4_timing.log
done
DSG_LIB_PATH=$PWD/gpurun_variants/lib_w2b6.so timeout 300 python scripts/profile_target.py soap 100000 8 2>&1 | tail -1 >> gpurun_out/c4_timing.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err
echo "bench exit $?" >> gpurun_out/c4_bench.err
tail -4 gpurun_out/c4_pytest.log; cat gpurun_out/c4_smoke.log; cat gpurun_out/c4_timing.log; tail -5 gpurun_out/c4_bench.err; head -c 1500 gpurun_out/c4_bench.json
