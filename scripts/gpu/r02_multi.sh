#!/bin/bash
# multi-GPU bench as the driver launches it (N = number of visible GPUs)
set -x
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout -k 5 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "rc=$?"; tail -5 gpurun_out/bench_${N}gpu.err
timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/check_dist_mbar.py > gpurun_out/dist_mbar_${N}gpu.log 2>&1; tail -2 gpurun_out/dist_mbar_${N}gpu.log
