#!/bin/bash
# Builds variants of the speculative Metropolis kernel (macro knobs in csrc/sampler_spec.cu) into
# build/variants/lib_<name>.so; run them on the GPU box with BCP_LIB=<path> python scripts/profile_target.py
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
build evalfirst -DSPEC_EVAL_FIRST=1 &
build novol -DSPEC_LDS_VOLATILE=0 &
build mb2 -DSPEC_MINBLOCKS=2 &
wait
build mb3 -DSPEC_MINBLOCKS=3 &
build evalfirst_novol -DSPEC_EVAL_FIRST=1 -DSPEC_LDS_VOLATILE=0 &
build mb3_evalfirst -DSPEC_MINBLOCKS=3 -DSPEC_EVAL_FIRST=1 &
wait
