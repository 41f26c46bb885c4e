#!/bin/bash
# first run of the warp-specialised kernel: smoke, fuzz parity, timing against spec; then the stage-kernel captures
set -x
mkdir -p gpurun_out
export BCP_KERNEL=ws
timeout -k 5 180 python __graft_entry__.py smoke > gpurun_out/ws_smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -20 gpurun_out/ws_smoke.log; unset BCP_KERNEL; }
if [ -n "$BCP_KERNEL" ]; then
  unset BCP_KERNEL
  timeout -k 5 600 python scripts/fuzz_parity.py 40 3 > gpurun_out/ws_fuzz.log 2>&1; echo "fuzz rc=$?"; tail -5 gpurun_out/ws_fuzz.log
  for k in spec ws; do
    BCP_KERNEL=$k timeout -k 5 200 python scripts/profile_target.py 4096 20000 4 > gpurun_out/ws_time_$k.log 2>&1; tail -4 gpurun_out/ws_time_$k.log
  done
  BCP_KERNEL=ws timeout -k 5 900 python -m pytest tests/test_sampler_gpu.py tests/test_full_size_gpu.py -x -q -k "ws or not kernel" > gpurun_out/ws_pytest.log 2>&1; tail -5 gpurun_out/ws_pytest.log
  BCP_KERNEL=ws timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:mcmc_ws_kernel -s 2 -c 1 \
    -f -o gpurun_out/r02_mcmc_ws_4096 python scripts/profile_target.py 4096 20000 3 > gpurun_out/ws_ncu.log 2>&1
fi
# stage kernels: one ncu pass without source import (small report) + CSV of the raw page made on the box
timeout 900 ncu --set full --clock-control none \
  -k 'regex:row_kernel|col_partial_kernel|ukn_kernel|mbar_pass_reg_kernel|mbar_pop_reg_kernel|logz_kernel' \
  -f -o /tmp/r02_before_stage_kernels python scripts/profile_stages.py > gpurun_out/r02_ncu_stages.log 2>&1
ncu -i /tmp/r02_before_stage_kernels.ncu-rep --page raw --csv > gpurun_out/r02_before_stage_kernels.raw.csv 2>/dev/null
timeout 600 python scripts/bench_stages.py > gpurun_out/r02_stages_before.json 2> gpurun_out/r02_stages_before.err
du -sh gpurun_out; ls -la gpurun_out
