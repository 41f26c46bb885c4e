#!/bin/bash
set -x
mkdir -p gpurun_out
for v in build/variants/libws_*.so; do
  nm=$(basename $v .so)
  for n in 704 2048 4096; do
    BCP_LIB=$v BCP_KERNEL=ws timeout -k 5 200 python scripts/profile_target.py $n 20000 3 > gpurun_out/ws12_${nm}_$n.log 2>&1; echo "$nm $n $(tail -1 gpurun_out/ws12_${nm}_$n.log)"
  done
done
BCP_LIB=build/variants/libws_pf.so BCP_KERNEL=ws timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:mcmc_ws_kernel -s 2 -c 1 \
    -f -o gpurun_out/r02_mcmc_ws_704 python scripts/profile_target.py 704 20000 3 > gpurun_out/ws12_ncu.log 2>&1
