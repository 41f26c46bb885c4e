#!/bin/bash
# full GPU test suite + smoke + default bench with the table-driven kernel in the automatic choice
mkdir -p gpurun_out
timeout -k 5 300 python __graft_entry__.py smoke > gpurun_out/full2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/full2_smoke.log
timeout -k 5 1500 python -m pytest tests -m gpu -x -q > gpurun_out/full2_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/full2_pytest.log
timeout -k 5 900 python bench.py > gpurun_out/full2_bench.json 2> gpurun_out/full2_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/full2_bench.err
