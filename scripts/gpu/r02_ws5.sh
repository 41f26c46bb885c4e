#!/bin/bash
set -x
mkdir -p gpurun_out
BCP_KERNEL=ws timeout -k 5 180 python __graft_entry__.py smoke > gpurun_out/ws5_smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -20 gpurun_out/ws5_smoke.log; exit 1; }
timeout -k 5 600 python scripts/fuzz_parity.py 40 51 > gpurun_out/ws5_fuzz.log 2>&1; echo "fuzz rc=$?"; tail -5 gpurun_out/ws5_fuzz.log
BCP_KERNEL=ws timeout -k 5 200 python scripts/profile_target.py 4096 20000 4 > gpurun_out/ws5_time_main.log 2>&1; tail -1 gpurun_out/ws5_time_main.log
for v in build/variants/libws_*.so; do
  n=$(basename $v .so)
  BCP_LIB=$v BCP_KERNEL=ws timeout -k 5 200 python scripts/profile_target.py 4096 20000 4 > gpurun_out/ws5_time_$n.log 2>&1; echo $n; tail -1 gpurun_out/ws5_time_$n.log
done
BCP_KERNEL=ws timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:mcmc_ws_kernel -s 2 -c 1 \
    -f -o gpurun_out/r02_mcmc_ws5 python scripts/profile_target.py 4096 20000 3 > gpurun_out/ws5_ncu.log 2>&1
