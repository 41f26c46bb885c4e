#!/bin/bash
mkdir -p gpurun_out
for v in build/variants/libtab_d*.so; do
BCP_LIB=$v FUZZ_ONLY=9 timeout -k 5 400 python scripts/fuzz_parity.py 24 9 big > gpurun_out/tab15_$(basename $v .so).log 2>&1; echo "$v fuzz rc=$?"; grep "tab call\|fuzz:" gpurun_out/tab15_$(basename $v .so).log | head -4
done
