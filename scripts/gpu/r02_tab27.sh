#!/bin/bash
mkdir -p gpurun_out
BCP_SPEC_STATS=1 BCP_LIB=build/variants/libtab_stats.so BCP_KERNEL=tab timeout -k 5 120 python scripts/profile_target.py 4096 10000 8 > gpurun_out/tab27_decay.log 2>&1; grep "^kernel\|stats" gpurun_out/tab27_decay.log
BCP_KERNEL=spec timeout -k 5 120 python scripts/profile_target.py 4096 10000 8 > gpurun_out/tab27_decay_spec.log 2>&1; grep "^kernel" gpurun_out/tab27_decay_spec.log | tr '\n' ' '
