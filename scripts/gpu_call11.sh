#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-extra > gpurun_out/c11_bench2.json 2> gpurun_out/c11_bench2.err
echo "exit $?" >> gpurun_out/c11_bench2.err
tail -5 gpurun_out/c11_bench2.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/c11_bench2.json').read().strip().splitlines()[-1])
    for k in ('value','ms_per_step','e2e','strong_scaling','parity'): print(k, d.get(k))
except Exception as e: print("parse fail", e)
PY
