#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 400 python scripts/fuzz_parity.py 24 11 big > gpurun_out/tab16_fuzz.log 2>&1; echo "fuzz rc=$?"; tail -2 gpurun_out/tab16_fuzz.log
for n in 704 4096; do
    BCP_SPEC_STATS=1 BCP_KERNEL=tab timeout -k 5 120 python scripts/profile_target.py $n 20000 3 > gpurun_out/tab16_${n}_tab.log 2>&1; echo "$n tab rc=$? $(tail -2 gpurun_out/tab16_${n}_tab.log | tr '\n' ' ')"
done
BCP_TAB_TRACE=1 BCP_SPEC_STATS=1 BCP_LIB=build/variants/libtab_trace.so BCP_KERNEL=tab timeout -k 5 100 python scripts/profile_target.py 704 20000 1 > gpurun_out/tab10_trace.log 2>&1; echo rc=$?
