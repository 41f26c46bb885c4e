#!/bin/bash
# first run of the table-driven kernel: parity fuzz on shapes it takes, then timings against spec / ws
mkdir -p gpurun_out
timeout -k 5 300 python scripts/fuzz_parity.py 10 1 big > gpurun_out/tab1_fuzz.log 2>&1; echo "fuzz rc=$?"; tail -15 gpurun_out/tab1_fuzz.log
for n in 704 4096; do
  for k in tab spec; do
    BCP_SPEC_STATS=1 BCP_KERNEL=$k timeout -k 5 120 python scripts/profile_target.py $n 20000 3 > gpurun_out/tab1_${n}_$k.log 2>&1; echo "$n $k rc=$? $(tail -2 gpurun_out/tab1_${n}_$k.log | tr '\n' ' ')"
  done
done
