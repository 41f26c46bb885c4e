#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 400 python scripts/fuzz_parity.py 12 16 big > gpurun_out/tab32_fuzz.log 2>&1; echo "fuzz rc=$?"; tail -1 gpurun_out/tab32_fuzz.log
timeout -k 5 200 python scripts/time_e2e.py > gpurun_out/tab32_e2e_base.log 2>&1; echo base; grep "rep" gpurun_out/tab32_e2e_base.log | tail -2
BCP_KERNEL=tab timeout -k 5 120 python scripts/profile_target.py 4096 100000 3 > gpurun_out/tab32_base.log 2>&1; echo "base steady: $(grep '^kernel' gpurun_out/tab32_base.log | tr '\n' ' ')"
for v in build/variants/libtab_n*.so; do
  nm=$(basename $v .so)
  BCP_LIB=$v timeout -k 5 200 python scripts/time_e2e.py > gpurun_out/tab32_e2e_$nm.log 2>&1; echo $nm; grep "rep" gpurun_out/tab32_e2e_$nm.log | tail -2
  BCP_LIB=$v BCP_KERNEL=tab timeout -k 5 120 python scripts/profile_target.py 4096 100000 3 > gpurun_out/tab32_$nm.log 2>&1; echo "$nm steady: $(grep '^kernel' gpurun_out/tab32_$nm.log | tr '\n' ' ')"
done
