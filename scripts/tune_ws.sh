#!/bin/bash
# Variant builds of the warp-specialised kernel (knobs of csrc/sampler_ws.cu) into build/variants/libws_<name>.so; time
# them on the GPU box with   for v in build/variants/libws_*.so; do BCP_LIB=$v BCP_KERNEL=ws python scripts/profile_target.py 4096 20000 4; done
set -e
cd "$(dirname "$0")/.."
python __graft_entry__.py build > /dev/null
mkdir -p build/variants
rm -f build/variants/libws_*.so
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
OBJS=$(ls biceps_b200/lib/*.o | grep -v sampler_ws.o)
build() { name=$1; shift
  nvcc $FLAGS "$@" -c biceps_b200/csrc/sampler_ws.cu -o build/variants/ws_$name.o
  nvcc -shared -o build/variants/libws_$name.so $OBJS build/variants/ws_$name.o -gencode arch=compute_100a,code=sm_100a
  rm build/variants/ws_$name.o; echo built $name; }
for v in "$@"; do
  name=${v%%:*}; flags=${v#*:}
  build $name $flags &
done
wait
