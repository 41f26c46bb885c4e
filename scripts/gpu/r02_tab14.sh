#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 400 python scripts/fuzz_parity.py 24 9 big > gpurun_out/tab14_fuzz.log 2>&1; echo "fuzz rc=$?"; tail -3 gpurun_out/tab14_fuzz.log
BCP_KERNEL=tab timeout -k 5 600 python -m pytest tests/test_sampler_gpu.py -x -q -k "tab" > gpurun_out/tab14_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/tab14_pytest.log
for n in 704 4096; do
    BCP_SPEC_STATS=1 BCP_KERNEL=tab timeout -k 5 120 python scripts/profile_target.py $n 20000 3 > gpurun_out/tab14_${n}_tab.log 2>&1; echo "$n tab rc=$? $(tail -2 gpurun_out/tab14_${n}_tab.log | tr '\n' ' ')"
done
for v in build/variants/libtab_*.so; do
  nm=$(basename $v .so)
  for n in 4096; do
    BCP_SPEC_STATS=1 BCP_LIB=$v BCP_KERNEL=tab timeout -k 5 100 python scripts/profile_target.py $n 20000 3 > gpurun_out/tab14_${nm}_$n.log 2>&1; echo "$nm $n rc=$? $(tail -2 gpurun_out/tab14_${nm}_$n.log | tr '\n' ' ')"
  done
done
