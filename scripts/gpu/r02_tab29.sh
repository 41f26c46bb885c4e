#!/bin/bash
mkdir -p gpurun_out
BCP_SPEC_STATS=1 timeout -k 5 300 python scripts/probe_c2.py > gpurun_out/tab29_c2.log 2>&1; grep "chains" gpurun_out/tab29_c2.log | head -6; grep "stats" gpurun_out/tab29_c2.log | head -3
timeout -k 5 400 python scripts/fuzz_parity.py 16 14 big > gpurun_out/tab29_fuzz.log 2>&1; echo "fuzz rc=$?"; tail -1 gpurun_out/tab29_fuzz.log
BCP_KERNEL=tab timeout -k 5 600 python -m pytest tests/test_sampler_gpu.py -x -q -k "tab" > gpurun_out/tab29_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/tab29_pytest.log
