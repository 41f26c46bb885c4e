#!/bin/bash
# timing of three consecutive 100 000-step launches (first = burn-in), fuzz, burn-in trace
python scripts/profile_target.py 4096 100000 3 2>&1 | tail -4
python scripts/fuzz_parity.py 40 77 big 2>&1 | tail -3
BCP_LIB=biceps_b200/lib/variants/libbiceps_b200_trace.so BCP_TAB_TRACE=1 BCP_SPEC_STATS=1 python scripts/profile_target.py 4096 2000 1 > gpurun_out/tab33_trace.log 2>&1
