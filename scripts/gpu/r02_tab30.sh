#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 400 python scripts/fuzz_parity.py 24 15 big > gpurun_out/tab30_fuzz.log 2>&1; echo "fuzz rc=$?"; tail -1 gpurun_out/tab30_fuzz.log
for n in 704 4096; do
  BCP_KERNEL=tab timeout -k 5 120 python scripts/profile_target.py $n 100000 3 > gpurun_out/tab30_$n.log 2>&1; echo "spec2 $n: $(grep '^kernel' gpurun_out/tab30_$n.log | tr '\n' ' ')"
  BCP_LIB=build/variants/libtab_nospec.so BCP_KERNEL=tab timeout -k 5 120 python scripts/profile_target.py $n 100000 3 > gpurun_out/tab30_nospec_$n.log 2>&1; echo "nospec $n: $(grep '^kernel' gpurun_out/tab30_nospec_$n.log | tr '\n' ' ')"
done
BCP_KERNEL=tab timeout -k 5 600 python -m pytest tests/test_sampler_gpu.py -x -q -k "tab" > gpurun_out/tab30_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/tab30_pytest.log
