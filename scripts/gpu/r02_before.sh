#!/bin/bash
# round-2 "before" evidence: baseline timings of the tree as round 1 left it + ncu full captures of the stage kernels
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02_before_smi.txt
timeout 300 python scripts/profile_target.py 4096 20000 4 > gpurun_out/r02_spec_base.log 2>&1
timeout 600 python scripts/bench_stages.py > gpurun_out/r02_stages_before.json 2> gpurun_out/r02_stages_before.err
timeout 300 python scripts/profile_stages.py > gpurun_out/r02_profile_stages.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on \
  -k 'regex:row_kernel|col_partial_kernel|ukn_kernel|mbar_pass_reg_kernel|mbar_pop_reg_kernel|logz_kernel' \
  -f -o gpurun_out/r02_before_stage_kernels python scripts/profile_stages.py > gpurun_out/r02_ncu_stages.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jsd_kernel -c 2 \
  -f -o gpurun_out/r02_jsd python scripts/bench_convergence.py > gpurun_out/r02_ncu_jsd.log 2>&1
BCP_PF_KERNEL=generic timeout 600 ncu --set full --clock-control none --import-source on -k regex:mcmc_pf_kernel -s 1 -c 1 \
  -f -o gpurun_out/r02_mcmc_pf_generic python scripts/profile_pf.py 8192 100 > gpurun_out/r02_ncu_pfgen.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail
