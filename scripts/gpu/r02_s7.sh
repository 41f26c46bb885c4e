#!/bin/bash
set -x
mkdir -p gpurun_out
BCP_KERNEL=spec timeout -k 5 200 python scripts/profile_target.py 4096 20000 4 > gpurun_out/s7_time_base.log 2>&1; tail -1 gpurun_out/s7_time_base.log
for v in build/variants/libspec_*.so; do
  n=$(basename $v .so)
  BCP_LIB=$v BCP_KERNEL=spec timeout -k 5 200 python scripts/profile_target.py 4096 20000 4 > gpurun_out/s7_time_$n.log 2>&1; echo $n; tail -1 gpurun_out/s7_time_$n.log
  BCP_LIB=$v BCP_KERNEL=spec timeout -k 5 200 python scripts/profile_target.py 704 20000 3 > gpurun_out/s7_time704_$n.log 2>&1; tail -1 gpurun_out/s7_time704_$n.log
done
BCP_KERNEL=spec timeout -k 5 200 python scripts/profile_target.py 704 20000 3 > gpurun_out/s7_time704_base.log 2>&1; tail -1 gpurun_out/s7_time704_base.log
BCP_LIB=build/variants/libspec_spread.so timeout -k 5 600 python scripts/fuzz_parity.py 25 41 > gpurun_out/s7_fuzz_spread.log 2>&1; tail -2 gpurun_out/s7_fuzz_spread.log
