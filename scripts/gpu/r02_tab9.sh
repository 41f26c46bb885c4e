#!/bin/bash
mkdir -p gpurun_out
for v in build/variants/libtab_*.so; do
  nm=$(basename $v .so)
  for n in 4096; do
    BCP_LIB=$v BCP_KERNEL=tab timeout -k 5 100 python scripts/profile_target.py $n 20000 2 > gpurun_out/tab9_${nm}_$n.log 2>&1; echo "$nm $n rc=$? $(tail -1 gpurun_out/tab9_${nm}_$n.log | tr '\n' ' ')"
  done
done
BCP_LIB=build/variants/libtab_noacc.so BCP_KERNEL=tab timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:mcmc_tab_kernel -s 1 -c 1 \
    -f -o gpurun_out/r02_mcmc_tab_noacc python scripts/profile_target.py 704 20000 3 > gpurun_out/tab9_ncu.log 2>&1; echo rc=$?
