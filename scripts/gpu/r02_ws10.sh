#!/bin/bash
set -x
mkdir -p gpurun_out
timeout -k 5 180 python __graft_entry__.py smoke > gpurun_out/ws10_smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -20 gpurun_out/ws10_smoke.log; exit 1; }
timeout -k 5 900 python scripts/fuzz_parity.py 60 71 > gpurun_out/ws10_fuzz.log 2>&1; echo "fuzz rc=$?"; tail -3 gpurun_out/ws10_fuzz.log
for n in 704 1184 2048 2368 4096; do for k in spec ws; do
  BCP_KERNEL=$k timeout -k 5 200 python scripts/profile_target.py $n 20000 3 > gpurun_out/ws10_${n}_$k.log 2>&1; echo "$n $k $(tail -1 gpurun_out/ws10_${n}_$k.log)"
done; done
timeout -k 5 400 python scripts/probe_c2.py > gpurun_out/ws10_c2.log 2>&1; cat gpurun_out/ws10_c2.log
timeout -k 5 300 python scripts/profile_score.py 1000000 > gpurun_out/ws10_score.log 2>&1; tail -3 gpurun_out/ws10_score.log
