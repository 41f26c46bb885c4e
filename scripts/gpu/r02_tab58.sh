#!/bin/bash
mkdir -p gpurun_out
BCP_TAB_TRACE=1 BCP_SPEC_STATS=1 BCP_LIB=build/variants/libtab_trace.so BCP_KERNEL=tab python scripts/profile_target.py 4096 100000 2 C3 > gpurun_out/tab58_trace_c3.log 2>&1
BCP_TAB_TRACE=1 BCP_SPEC_STATS=1 BCP_LIB=build/variants/libtab_trace.so BCP_KERNEL=tab python scripts/profile_target.py 8192 100000 2 C2 > gpurun_out/tab58_trace_c2.log 2>&1
grep -c "^trace" gpurun_out/tab58_trace_c3.log gpurun_out/tab58_trace_c2.log
