#!/bin/bash
set -x
mkdir -p gpurun_out
for v in build/variants/libws_*.so; do
  n=$(basename $v .so)
  BCP_LIB=$v BCP_KERNEL=ws timeout -k 5 200 python scripts/profile_target.py 4096 20000 4 > gpurun_out/ws7_time_$n.log 2>&1; echo $n; tail -1 gpurun_out/ws7_time_$n.log
done
