#!/bin/bash
mkdir -p gpurun_out
for n in 3072 4096; do
  BCP_SPEC_STATS=1 timeout -k 5 120 python scripts/profile_target.py $n 100000 3 > gpurun_out/tab18_auto_$n.log 2>&1; echo "auto $n rc=$?"; grep "pilot\|trial\|^kernel" gpurun_out/tab18_auto_$n.log | tr '\n' ' '; echo
done
BCP_SPEC_STATS=1 timeout -k 5 300 python -m pytest tests/test_api_gpu.py tests/test_full_size_gpu.py -x -q -m gpu > gpurun_out/tab18_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/tab18_pytest.log
