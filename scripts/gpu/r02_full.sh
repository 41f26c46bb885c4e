#!/bin/bash
# full -m gpu suite + smoke + the default bench
set -x
mkdir -p gpurun_out
timeout -k 5 2400 python -m pytest tests -x -q -m gpu > gpurun_out/full_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/full_pytest.log
timeout -k 5 300 python __graft_entry__.py smoke > gpurun_out/full_smoke.log 2>&1; tail -2 gpurun_out/full_smoke.log
timeout -k 5 900 python bench.py --steps 3 --warmup 3 > gpurun_out/full_bench.json 2> gpurun_out/full_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/full_bench.err
timeout -k 5 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/full_bench_ref.json 2> gpurun_out/full_bench_ref.err; echo "ref rc=$?"
