#!/bin/bash
mkdir -p gpurun_out
python scripts/profile_target.py soap 20000 8 > gpurun_out/c16_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:soap_kernel -s 2 -c 1 \
    -o gpurun_out/r2_soap_mma python scripts/profile_target.py soap 20000 8 > gpurun_out/c16_ncu.log 2>&1
tail -3 gpurun_out/c16_ncu.log
