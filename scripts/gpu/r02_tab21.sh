#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 400 python scripts/fuzz_parity.py 20 12 big > gpurun_out/tab21_fuzz.log 2>&1; echo "fuzz rc=$?"; tail -2 gpurun_out/tab21_fuzz.log
BCP_KERNEL=tab timeout -k 5 600 python -m pytest tests/test_sampler_gpu.py -x -q -k "tab" > gpurun_out/tab21_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/tab21_pytest.log
