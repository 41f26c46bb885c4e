#!/bin/bash
mkdir -p gpurun_out
for v in build/variants/libtab_*.so; do
  nm=$(basename $v .so)
  for n in 4096; do
    BCP_SPEC_STATS=1 BCP_LIB=$v BCP_KERNEL=tab timeout -k 5 100 python scripts/profile_target.py $n 20000 3 > gpurun_out/tab19_${nm}_$n.log 2>&1; echo "$nm $n rc=$? $(tail -2 gpurun_out/tab19_${nm}_$n.log | tr '\n' ' ')"
  done
done
