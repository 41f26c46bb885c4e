#!/bin/bash
mkdir -p gpurun_out
BCP_KERNEL=tab timeout -k 5 120 python scripts/profile_target.py 4096 100000 3 > gpurun_out/tab34_base.log 2>&1; echo "steady: $(grep '^kernel' gpurun_out/tab34_base.log | tr '\n' ' ')"
timeout -k 5 400 python scripts/fuzz_parity.py 10 17 big > gpurun_out/tab34_fuzz.log 2>&1; echo "fuzz rc=$?"; tail -1 gpurun_out/tab34_fuzz.log
