#!/bin/bash
set -x
mkdir -p gpurun_out
for n in 704 1184 2048; do for k in spec ws; do
  BCP_KERNEL=$k timeout -k 5 200 python scripts/profile_target.py $n 20000 3 > gpurun_out/ws9_${n}_$k.log 2>&1; echo "$n $k $(tail -1 gpurun_out/ws9_${n}_$k.log)"
done; done
