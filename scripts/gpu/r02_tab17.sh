#!/bin/bash
mkdir -p gpurun_out
BCP_KERNEL=tab timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:mcmc_tab_kernel -s 1 -c 1 \
    -f -o gpurun_out/r02_mcmc_tab_704 python scripts/profile_target.py 704 20000 3 > gpurun_out/tab17_ncu.log 2>&1; echo rc=$?
