#!/bin/bash
set -x
mkdir -p gpurun_out
for v in build/variants/libws_*.so; do
  n=$(basename $v .so)
  BCP_LIB=$v BCP_KERNEL=ws timeout -k 5 200 python scripts/profile_target.py 4096 20000 4 > gpurun_out/ws4_time_$n.log 2>&1; echo $n; tail -2 gpurun_out/ws4_time_$n.log
done
BCP_KERNEL=spec timeout -k 5 200 python scripts/profile_target.py 4096 20000 4 > gpurun_out/ws4_time_spec.log 2>&1; tail -2 gpurun_out/ws4_time_spec.log
BCP_LIB=build/variants/libws_r0a.so timeout -k 5 400 python scripts/fuzz_parity.py 25 13 > gpurun_out/ws4_fuzz_r0a.log 2>&1; tail -3 gpurun_out/ws4_fuzz_r0a.log
BCP_LIB=build/variants/libws_r0a_nr.so BCP_KERNEL=ws timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:mcmc_ws_kernel -s 2 -c 1 \
    -f -o gpurun_out/r02_mcmc_ws4_r0anr python scripts/profile_target.py 4096 20000 3 > gpurun_out/ws4_ncu.log 2>&1
