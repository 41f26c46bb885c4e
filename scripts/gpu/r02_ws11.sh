#!/bin/bash
set -x
mkdir -p gpurun_out
for n in 2560 3072 3584; do for k in spec ws; do
  BCP_KERNEL=$k timeout -k 5 200 python scripts/profile_target.py $n 20000 3 > gpurun_out/ws11_${n}_$k.log 2>&1; echo "$n $k $(tail -1 gpurun_out/ws11_${n}_$k.log)"
done; done
timeout -k 5 2400 python -m pytest tests -x -q -m gpu > gpurun_out/ws11_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/ws11_pytest.log
timeout -k 5 900 python bench.py --steps 3 --warmup 3 > gpurun_out/ws11_bench.json 2> gpurun_out/ws11_bench.err; echo "bench rc=$?"
