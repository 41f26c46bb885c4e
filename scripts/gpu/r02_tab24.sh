#!/bin/bash
mkdir -p gpurun_out
for ps in 1024 4096 8192 16384; do
  BCP_PILOT_STEPS=$ps timeout -k 5 120 python scripts/profile_target.py 4096 100000 2 > gpurun_out/tab24_p$ps.log 2>&1; echo "pilot $ps: $(grep '^kernel' gpurun_out/tab24_p$ps.log | tr '\n' ' ')"
done
BCP_KERNEL=spec timeout -k 5 120 python scripts/profile_target.py 4096 100000 2 > gpurun_out/tab24_spec.log 2>&1; echo "spec: $(grep '^kernel' gpurun_out/tab24_spec.log | tr '\n' ' ')"
BCP_KERNEL=tab timeout -k 5 120 python scripts/profile_target.py 4096 100000 2 > gpurun_out/tab24_tab.log 2>&1; echo "tab forced: $(grep '^kernel' gpurun_out/tab24_tab.log | tr '\n' ' ')"
