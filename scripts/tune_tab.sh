#!/bin/bash
# Variant builds of the table-driven kernel (knobs of csrc/sampler_tab.cu) into build/variants/libtab_<name>.so; time them on
# the GPU box with   for v in build/variants/libtab_*.so; do BCP_LIB=$v BCP_KERNEL=tab python scripts/profile_target.py 4096 20000 4; done
# usage: tune_tab.sh name:"-DTAB_X=1 ..." ...
set -e
cd "$(dirname "$0")/.."
python __graft_entry__.py build > /dev/null
mkdir -p build/variants
rm -f build/variants/libtab_*.so
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
OBJS=$(ls biceps_b200/lib/*.o | grep -v sampler_tab.o)
build() { name=$1; shift
  nvcc $FLAGS "$@" -c biceps_b200/csrc/sampler_tab.cu -o build/variants/tab_$name.o 2>/dev/null
  nvcc -shared -o build/variants/libtab_$name.so $OBJS build/variants/tab_$name.o -gencode arch=compute_100a,code=sm_100a
  rm build/variants/tab_$name.o; echo built $name; }
for v in "$@"; do
  name=${v%%:*}; flags=${v#*:}
  build $name $flags &
done
wait
