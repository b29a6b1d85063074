#!/bin/bash
# time one profile_target stage with every library variant under gpurun_variants/
mkdir -p gpurun_out
out=gpurun_out/variants_$1.log
: > $out
for lib in default gpurun_variants/lib_*.so; do
  if [ $lib = default ]; then unset DSG_LIB_PATH; else export DSG_LIB_PATH=$PWD/$lib; fi
  echo "== $lib" >> $out
  timeout 300 python scripts/profile_target.py "$@" 2>&1 | tail -2 >> $out
done
cat $out
