#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 300 python scripts/fuzz_parity.py 10 5 big > gpurun_out/tab7_fuzz.log 2>&1; echo "fuzz rc=$?"; tail -4 gpurun_out/tab7_fuzz.log
for n in 704 4096; do
    BCP_SPEC_STATS=1 BCP_KERNEL=tab timeout -k 5 120 python scripts/profile_target.py $n 20000 3 > gpurun_out/tab7_${n}_tab.log 2>&1; echo "$n tab rc=$? $(tail -2 gpurun_out/tab7_${n}_tab.log | tr '\n' ' ')"
done
for v in build/variants/libtab_*.so; do
  nm=$(basename $v .so)
  for n in 704 4096; do
    BCP_SPEC_STATS=1 BCP_LIB=$v BCP_KERNEL=tab timeout -k 5 100 python scripts/profile_target.py $n 20000 2 > gpurun_out/tab7_${nm}_$n.log 2>&1; echo "$nm $n rc=$? $(tail -2 gpurun_out/tab7_${nm}_$n.log | tr '\n' ' ')"
  done
done
BCP_KERNEL=tab timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:mcmc_tab_kernel -s 1 -c 1 \
    -f -o gpurun_out/r02_mcmc_tab_704 python scripts/profile_target.py 704 20000 3 > gpurun_out/tab7_ncu.log 2>&1; echo rc=$?
