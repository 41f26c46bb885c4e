#!/bin/bash
mkdir -p gpurun_out
BCP_KERNEL=tab timeout -k 5 120 python scripts/profile_target.py 4096 100000 2 > gpurun_out/tab26_base.log 2>&1; echo "base: $(grep '^kernel' gpurun_out/tab26_base.log | tr '\n' ' ')"
for v in build/variants/libtab_*.so; do
  nm=$(basename $v .so)
  BCP_LIB=$v BCP_KERNEL=tab timeout -k 5 100 python scripts/profile_target.py 4096 100000 2 > gpurun_out/tab26_${nm}.log 2>&1; echo "$nm: $(grep '^kernel' gpurun_out/tab26_${nm}.log | tr '\n' ' ')"
done
