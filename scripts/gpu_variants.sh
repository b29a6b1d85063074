#!/bin/bash
# usage: gpu_variants.sh TAG variant... ("default" = the in-tree library): fused cfg5 + soap + fused3 timings
mkdir -p gpurun_out
tag=$1; shift
log=gpurun_out/${tag}_timing.log; : > $log
for v in "$@"; do
  if [ $v = default ]; then unset DSG_LIB_PATH; else export DSG_LIB_PATH=$PWD/gpurun_variants/lib_$v.so; fi
  echo "== $v" >> $log
  timeout 300 python scripts/profile_target.py fused 100000 33 2>&1 | tail -1 >> $log
  timeout 300 python scripts/profile_target.py soap 20000 8 2>&1 | tail -1 >> $log
  timeout 300 python scripts/profile_target.py fused3 50000 8 2>&1 | tail -1 >> $log
done
unset DSG_LIB_PATH
cat $log
