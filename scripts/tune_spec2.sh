#!/bin/bash
# Schedule lottery for the speculative kernel: semantically neutral variants (macro knobs of csrc/sampler_spec.cu and
# ptxas flags) built into build/variants/lib_<name>.so; time them on the GPU box with
#   for v in build/variants/lib_*.so; do BCP_LIB=$v BCP_KERNEL=spec python scripts/probe_spec_l.py; done
set -e
cd "$(dirname "$0")/.."
python __graft_entry__.py build > /dev/null
mkdir -p build/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
OBJS=$(ls biceps_b200/lib/*.o | grep -v sampler_spec.o)
build() { name=$1; shift
  nvcc $FLAGS "$@" -c biceps_b200/csrc/sampler_spec.cu -o build/variants/spec_$name.o
  nvcc -shared -o build/variants/lib_$name.so $OBJS build/variants/spec_$name.o -gencode arch=compute_100a,code=sm_100a
  rm build/variants/spec_$name.o; echo built $name; }
build base &
build d3 -DSPEC_DEPTH=3 &
build bf -DSPEC_FINISH_BRANCHFREE=1 &
build fence -DSPEC_STEP_FENCE=1 &
wait
build d3_fence -DSPEC_DEPTH=3 -DSPEC_STEP_FENCE=1 &
build bf_fence -DSPEC_FINISH_BRANCHFREE=1 -DSPEC_STEP_FENCE=1 &
build d3_bf -DSPEC_DEPTH=3 -DSPEC_FINISH_BRANCHFREE=1 &
build ptxO2 -Xptxas -O2 &
wait
build r176 -maxrregcount=176 &
build r200 -maxrregcount=200 &
build evalfirst_fence -DSPEC_EVAL_FIRST=1 -DSPEC_STEP_FENCE=1 &
build novol_bf -DSPEC_LDS_VOLATILE=0 -DSPEC_FINISH_BRANCHFREE=1 &
wait
