#!/bin/bash
# end-of-round run: smoke, the full GPU test suite, the default bench, the reference arm, then the evidence for the table-driven
# kernel (ncu --set full capture of the headline launch, launch list of the bench, hand-off trace)
mkdir -p gpurun_out
timeout -k 5 300 python __graft_entry__.py smoke > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/final_smoke.log
timeout -k 5 1800 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/final_pytest.log
timeout -k 5 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/final_bench.err
timeout -k 5 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo "ref bench rc=$?"
BCP_KERNEL=tab timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:mcmc_tab_kernel -s 2 -c 1 \
    -f -o gpurun_out/r02_mcmc_tab_4096 python scripts/profile_target.py 4096 20000 3 > gpurun_out/final_ncu_tab.log 2>&1; echo "ncu rc=$?"
timeout -k 5 900 python bench.py --steps 2 --warmup 3 > gpurun_out/final_bench_plain.json 2> gpurun_out/final_bench_plain.err && \
timeout -k 5 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02b_bench_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/final_bench_ncu.log 2>&1; echo "launch list rc=$?"
