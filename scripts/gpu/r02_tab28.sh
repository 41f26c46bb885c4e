#!/bin/bash
mkdir -p gpurun_out
for n in 8192 9472 14208 18944; do
  for k in tab spec fast; do
    BCP_KERNEL=$k timeout -k 5 120 python scripts/profile_target.py $n 50000 3 > gpurun_out/tab28_${n}_$k.log 2>&1; echo "$n $k $(tail -1 gpurun_out/tab28_${n}_$k.log)"
  done
done
BCP_SPEC_STATS=1 timeout -k 5 300 python scripts/probe_c2.py > gpurun_out/tab28_c2.log 2>&1; grep "chains" gpurun_out/tab28_c2.log | head -12; grep "stats" gpurun_out/tab28_c2.log | head -4
