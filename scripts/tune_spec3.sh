#!/bin/bash
# variant builds of the speculative kernel into build/variants/libspec_<name>.so (see tune_ws.sh)
set -e
cd "$(dirname "$0")/.."
python __graft_entry__.py build > /dev/null
mkdir -p build/variants
rm -f build/variants/libspec_*.so
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
OBJS=$(ls biceps_b200/lib/*.o | grep -v sampler_spec.o)
build() { name=$1; shift
  nvcc $FLAGS "$@" -c biceps_b200/csrc/sampler_spec.cu -o build/variants/spec_$name.o
  nvcc -shared -o build/variants/libspec_$name.so $OBJS build/variants/spec_$name.o -gencode arch=compute_100a,code=sm_100a
  rm build/variants/spec_$name.o; echo built $name; }
for v in "$@"; do
  name=${v%%:*}; flags=${v#*:}
  build $name $flags &
done
wait
