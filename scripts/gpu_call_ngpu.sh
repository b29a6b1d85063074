#!/bin/bash
# usage: gpu_call_ngpu.sh N  -- the driver's launch line for N ranks
N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-extra > gpurun_out/bench${N}_r2.json 2> gpurun_out/bench${N}_r2.err
echo "exit $?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench${N}_r2.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','strong_scaling'): print(k, d.get(k))
print('parity ok', d['parity']['ok'])
PY
