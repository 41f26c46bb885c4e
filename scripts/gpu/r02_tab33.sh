#!/bin/bash
mkdir -p gpurun_out
BCP_TAB_TRACE=1 BCP_SPEC_STATS=1 BCP_LIB=build/variants/libtab_trace.so BCP_KERNEL=tab timeout -k 5 100 python scripts/profile_target.py 4096 10000 1 > gpurun_out/tab33_trace.log 2>&1; echo rc=$?; tail -1 gpurun_out/tab33_trace.log
