#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 300 python scripts/fuzz_parity.py 6 6 big > gpurun_out/tab8_fuzz.log 2>&1; echo "fuzz rc=$?"; tail -2 gpurun_out/tab8_fuzz.log
for n in 704 4096; do
    BCP_SPEC_STATS=1 BCP_KERNEL=tab timeout -k 5 120 python scripts/profile_target.py $n 20000 3 > gpurun_out/tab8_${n}_tab.log 2>&1; echo "$n tab rc=$? $(tail -1 gpurun_out/tab8_${n}_tab.log | tr '\n' ' ')"
done
for v in build/variants/libtab_*.so; do
  nm=$(basename $v .so)
  for n in 704 4096; do
    BCP_LIB=$v BCP_KERNEL=tab timeout -k 5 100 python scripts/profile_target.py $n 20000 2 > gpurun_out/tab8_${nm}_$n.log 2>&1; echo "$nm $n rc=$? $(tail -1 gpurun_out/tab8_${nm}_$n.log | tr '\n' ' ')"
  done
done
