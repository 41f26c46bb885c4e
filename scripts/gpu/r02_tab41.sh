#!/bin/bash
mkdir -p gpurun_out
python scripts/fuzz_parity.py 40 31 big 2>&1 | tail -1
python scripts/probe_c2_auto.py 2>&1 | tail -4
python scripts/profile_target.py 4096 100000 3 2>&1 | tail -3
python scripts/probe_e2e.py 2>&1 | tail -2
BCP_TAB_TRACE=1 BCP_SPEC_STATS=1 BCP_LIB=build/variants/libtab_trace.so BCP_KERNEL=tab timeout -k 5 100 python scripts/profile_target.py 8192 20000 2 C2 > gpurun_out/tab40_trace.log 2>&1
