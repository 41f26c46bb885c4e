#!/bin/bash
mkdir -p gpurun_out
BCP_SPEC_STATS=1 BCP_LIB=build/variants/libtab_stats.so BCP_KERNEL=tab timeout -k 5 100 python scripts/profile_target.py 704 20000 3 > gpurun_out/tab11_stats.log 2>&1; echo rc=$?; tail -6 gpurun_out/tab11_stats.log
