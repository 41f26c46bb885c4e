#!/bin/bash
set -x
mkdir -p gpurun_out
BCP_KERNEL=spec timeout -k 5 200 python scripts/profile_target.py 4096 20000 4 > gpurun_out/s5_time_base.log 2>&1; tail -1 gpurun_out/s5_time_base.log
for v in build/variants/libspec_*.so; do
  n=$(basename $v .so)
  BCP_LIB=$v BCP_KERNEL=spec timeout -k 5 200 python scripts/profile_target.py 4096 20000 4 > gpurun_out/s5_time_$n.log 2>&1; echo $n; tail -1 gpurun_out/s5_time_$n.log
done
BCP_LIB=build/variants/libspec_pg.so timeout -k 5 400 python scripts/fuzz_parity.py 20 21 > gpurun_out/s5_fuzz_pg.log 2>&1; tail -2 gpurun_out/s5_fuzz_pg.log
timeout -k 5 600 python -m pytest tests/test_mbar_gpu.py tests/test_api_gpu.py -x -q > gpurun_out/s5_pytest_mbar.log 2>&1; tail -3 gpurun_out/s5_pytest_mbar.log
timeout 600 python scripts/bench_stages.py > gpurun_out/s5_stages.json 2> gpurun_out/s5_stages.err; tail -2 gpurun_out/s5_stages.err
