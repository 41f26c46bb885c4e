#!/bin/bash
mkdir -p gpurun_out
BCP_SPEC_STATS=1 timeout -k 5 200 python scripts/time_e2e.py > gpurun_out/tab31_e2e_auto.log 2>&1; grep "rep\|trial" gpurun_out/tab31_e2e_auto.log | tail -8
BCP_KERNEL=tab timeout -k 5 200 python scripts/time_e2e.py > gpurun_out/tab31_e2e_tab.log 2>&1; echo forced tab; grep "rep" gpurun_out/tab31_e2e_tab.log | tail -3
BCP_TAB=0 timeout -k 5 200 python scripts/time_e2e.py > gpurun_out/tab31_e2e_notab.log 2>&1; echo no tab; grep "rep" gpurun_out/tab31_e2e_notab.log | tail -3
