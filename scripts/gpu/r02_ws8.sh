#!/bin/bash
set -x
mkdir -p gpurun_out
BCP_KERNEL=ws timeout -k 5 180 python __graft_entry__.py smoke > gpurun_out/ws8_smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -20 gpurun_out/ws8_smoke.log; exit 1; }
timeout -k 5 900 python scripts/fuzz_parity.py 60 61 > gpurun_out/ws8_fuzz.log 2>&1; echo "fuzz rc=$?"; tail -3 gpurun_out/ws8_fuzz.log
for k in spec ws; do
  BCP_KERNEL=$k timeout -k 5 300 python scripts/probe_snap.py > gpurun_out/ws8_snap_$k.log 2>&1; cat gpurun_out/ws8_snap_$k.log
done
BCP_KERNEL=ws timeout -k 5 900 python -m pytest tests/test_sampler_gpu.py tests/test_full_size_gpu.py tests/test_api_gpu.py -x -q -k "ws or not kernel" > gpurun_out/ws8_pytest.log 2>&1; tail -3 gpurun_out/ws8_pytest.log
