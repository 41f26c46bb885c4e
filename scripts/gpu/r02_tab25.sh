#!/bin/bash
mkdir -p gpurun_out
BCP_SPEC_STATS=1 timeout -k 5 120 python scripts/profile_target.py 4096 100000 2 > gpurun_out/tab25_auto.log 2>&1; echo "auto: $(grep 'trial\|pilot\|^kernel' gpurun_out/tab25_auto.log | tr '\n' ' ')"
BCP_SPEC_STATS=1 timeout -k 5 300 python scripts/probe_c1.py > gpurun_out/tab25_c1.log 2>&1; tail -12 gpurun_out/tab25_c1.log
timeout -k 5 600 python -m pytest tests/test_api_gpu.py tests/test_full_size_gpu.py tests/test_mbar_gpu.py -x -q -m gpu > gpurun_out/tab25_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/tab25_pytest.log
