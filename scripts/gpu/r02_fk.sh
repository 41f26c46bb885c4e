#!/bin/bash
set -x
mkdir -p gpurun_out
for v in build/variants/libws_*.so; do
  n=$(basename $v .so)
  BCP_LIB=$v BCP_KERNEL=ws timeout -k 5 200 python scripts/profile_target.py 4096 20000 4 > gpurun_out/fk_time_$n.log 2>&1; echo $n; tail -1 gpurun_out/fk_time_$n.log
done
BCP_LIB=build/variants/libws_fk_r.so BCP_KERNEL=ws timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:mcmc_ws_kernel -s 2 -c 1 \
    -f -o gpurun_out/r02_mcmc_ws_fk python scripts/profile_target.py 4096 20000 3 > gpurun_out/fk_ncu.log 2>&1
