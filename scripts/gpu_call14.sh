#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/c14_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c14_pytest.log
tail -5 gpurun_out/c14_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c14_smoke.log 2>&1; echo "smoke exit $?"
python bench.py > gpurun_out/c14_bench.json 2> gpurun_out/c14_bench.err; echo "bench exit $?"
tail -c 600 gpurun_out/c14_bench.err
