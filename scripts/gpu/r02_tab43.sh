#!/bin/bash
mkdir -p gpurun_out
BCP_TAB_TRACE=1 BCP_SPEC_STATS=1 BCP_LIB=build/variants/libtab_trace.so BCP_KERNEL=tab timeout -k 5 100 python scripts/profile_target.py 4096 100000 2 C3 > gpurun_out/tab43_trace.log 2>&1; echo rc=$?; grep "^kernel" gpurun_out/tab43_trace.log
