#!/bin/bash
# steady-state comparison of the Metropolis kernels on C3 (third launch of 100 000 steps)
mkdir -p gpurun_out
for n in 704 2048 2368 4096 4736; do
  for k in tab ws spec; do
    BCP_KERNEL=$k timeout -k 5 120 python scripts/profile_target.py $n 100000 3 > gpurun_out/tab22_${n}_$k.log 2>&1; echo "$n $k $(tail -1 gpurun_out/tab22_${n}_$k.log)"
  done
done
