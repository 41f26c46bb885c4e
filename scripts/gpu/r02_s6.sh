#!/bin/bash
set -x
mkdir -p gpurun_out
BCP_KERNEL=spec timeout -k 5 200 python scripts/profile_target.py 4096 20000 4 > gpurun_out/s6_time_base.log 2>&1; tail -1 gpurun_out/s6_time_base.log
for v in build/variants/libspec_*.so; do
  n=$(basename $v .so)
  BCP_LIB=$v BCP_KERNEL=spec timeout -k 5 200 python scripts/profile_target.py 4096 20000 4 > gpurun_out/s6_time_$n.log 2>&1; echo $n; tail -1 gpurun_out/s6_time_$n.log
done
BCP_LIB=build/variants/libspec_rc.so timeout -k 5 600 python scripts/fuzz_parity.py 40 31 > gpurun_out/s6_fuzz_rc.log 2>&1; tail -3 gpurun_out/s6_fuzz_rc.log
timeout -k 5 600 python -m pytest tests/test_precompute_gpu.py tests/test_api_gpu.py tests/test_full_size_gpu.py -x -q > gpurun_out/s6_pytest_pre.log 2>&1; tail -5 gpurun_out/s6_pytest_pre.log
timeout 600 python scripts/bench_stages.py > gpurun_out/s6_stages.json 2> gpurun_out/s6_stages.err; tail -2 gpurun_out/s6_stages.err
BCP_LIB=build/variants/libspec_rc.so BCP_KERNEL=spec timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:mcmc_spec_kernel -s 2 -c 1 \
    -f -o gpurun_out/r02_mcmc_spec_rc python scripts/profile_target.py 4096 20000 3 > gpurun_out/s6_ncu.log 2>&1
