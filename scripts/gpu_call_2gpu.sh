#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-extra > gpurun_out/bench2_r2.json 2> gpurun_out/bench2_r2.err
echo "exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench2_r2.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','strong_scaling','parity'): print(k, d.get(k))
PY
