#!/bin/bash
# final build: steady-state comparison of the kernels, hand-off traces (C3 4096 chains, C2 8192 chains)
mkdir -p gpurun_out
for n in 704 2048 2368 4096 4736; do
  line="$n"
  for k in tab ws spec; do
    ms=$(BCP_KERNEL=$k python scripts/profile_target.py $n 100000 3 2>&1 | grep '^kernel' | tail -1 | sed 's/.*(\([0-9.]*\),.*/\1/')
    line="$line $ms"
  done
  echo "$line"
done > gpurun_out/tab56_steady.txt
cat gpurun_out/tab56_steady.txt
BCP_TAB_TRACE=1 BCP_SPEC_STATS=1 BCP_LIB=build/variants/libtab_trace.so BCP_KERNEL=tab python scripts/profile_target.py 4096 100000 2 C3 > gpurun_out/tab56_trace_c3.log 2>&1
BCP_TAB_TRACE=1 BCP_SPEC_STATS=1 BCP_LIB=build/variants/libtab_trace.so BCP_KERNEL=tab python scripts/profile_target.py 8192 100000 2 C2 > gpurun_out/tab56_trace_c2.log 2>&1
grep -c "^trace" gpurun_out/tab56_trace_c3.log gpurun_out/tab56_trace_c2.log
