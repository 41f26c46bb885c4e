#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 400 python scripts/fuzz_parity.py 12 13 big > gpurun_out/tab23_fuzz.log 2>&1; echo "fuzz rc=$?"; tail -1 gpurun_out/tab23_fuzz.log
for n in 704 4096; do
  BCP_SPEC_STATS=1 timeout -k 5 120 python scripts/profile_target.py $n 100000 3 > gpurun_out/tab23_auto_$n.log 2>&1; echo "auto $n $(grep 'trial\|^kernel' gpurun_out/tab23_auto_$n.log | tr '\n' ' ')"
done
timeout -k 5 300 python scripts/profile_score.py > gpurun_out/tab23_score.log 2>&1; tail -3 gpurun_out/tab23_score.log
