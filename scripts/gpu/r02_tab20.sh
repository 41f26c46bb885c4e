#!/bin/bash
mkdir -p gpurun_out
BCP_KERNEL=tab timeout -k 5 120 python scripts/profile_target.py 4096 100000 3 > gpurun_out/tab20_base.log 2>&1; echo "base rc=$? $(tail -1 gpurun_out/tab20_base.log)"
for v in build/variants/libtab_*.so; do
  nm=$(basename $v .so)
  BCP_LIB=$v BCP_KERNEL=tab timeout -k 5 100 python scripts/profile_target.py 4096 100000 3 > gpurun_out/tab20_${nm}.log 2>&1; echo "$nm rc=$? $(tail -1 gpurun_out/tab20_${nm}.log)"
done
