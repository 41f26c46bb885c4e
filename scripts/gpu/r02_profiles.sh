#!/bin/bash
# round-2 evidence: launch list of the bench, one ncu --set full capture per kernel (raw pages exported as CSV on the
# box: the reports themselves stay small or stay on the box)
set -x
mkdir -p gpurun_out
T=/tmp/ncu; mkdir -p $T
NCU="ncu --set full --clock-control none"
raw() { ncu -i $1 --page raw --csv > $2 2>/dev/null; }
# 1. the headline kernel with source (kept: ~10 MB)
BCP_KERNEL=spec timeout 400 $NCU --import-source on -k regex:mcmc_spec_kernel -s 2 -c 1 -f -o gpurun_out/r02_mcmc_spec_4096 python scripts/profile_target.py 4096 20000 3 > gpurun_out/p_spec.log 2>&1
# 2. stage kernels (K1/K2/K4, logZ, K6, K7 passes, populations)
timeout 900 $NCU -k 'regex:row_kernel|col_stats_kernel|col_finish|ref_coef|gamma_prep|ukn_kernel|mbar_pass_reg_kernel|mbar_pop_reg_kernel|mbar_finish|logz_kernel' \
  -f -o $T/stages python scripts/profile_stages.py > gpurun_out/p_stages.log 2>&1
raw $T/stages.ncu-rep gpurun_out/r02_stage_kernels.raw.csv
# 3. the score pipeline as it runs (ukn_traj_kernel on chain-ordered snapshots, MBAR on real u_kn)
timeout 900 $NCU -k 'regex:ukn_traj_kernel|mbar_pass_reg_kernel|mbar_pop_reg_kernel' -c 12 -f -o $T/score python scripts/profile_score.py 100000 > gpurun_out/p_score.log 2>&1
raw $T/score.ncu-rep gpurun_out/r02_score_kernels.raw.csv
# 4. convergence kernels
timeout 600 $NCU -k 'regex:jsd_kernel|hist_kernel|autocorr_partial_kernel' -c 6 -f -o $T/conv python scripts/bench_convergence.py > gpurun_out/p_conv.log 2>&1
raw $T/conv.ncu-rep gpurun_out/r02_conv_kernels.raw.csv
# 5. PF kernels on C4
timeout 600 $NCU -k regex:mcmc_pf2_kernel -s 1 -c 1 -f -o $T/pf2 python scripts/profile_pf.py 32768 500 > gpurun_out/p_pf2.log 2>&1
raw $T/pf2.ncu-rep gpurun_out/r02_mcmc_pf2_32k.raw.csv
BCP_PF_KERNEL=generic timeout 600 $NCU -k regex:mcmc_pf_kernel -s 1 -c 1 -f -o $T/pfg python scripts/profile_pf.py 8192 100 > gpurun_out/p_pfg.log 2>&1
raw $T/pfg.ncu-rep gpurun_out/r02_mcmc_pf_generic.raw.csv
# 6. reference-stream and thread-per-chain kernels
timeout 400 $NCU -k regex:mcmc_fast_kernel -s 1 -c 1 -f -o $T/fast python scripts/profile_target.py 303104 2000 3 > gpurun_out/p_fast.log 2>&1
raw $T/fast.ncu-rep gpurun_out/r02_mcmc_fast_303k.raw.csv
# 7. launch list of the bench (after a plain run that exits 0)
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/p_bench_plain.json 2> gpurun_out/p_bench_plain.err && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/p_bench_ncu.log 2>&1
timeout 600 python scripts/bench_stages.py > gpurun_out/r02_stages.json 2> gpurun_out/p_stages2.err
du -sh gpurun_out; ls -la gpurun_out | tail -20
