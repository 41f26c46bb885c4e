#!/bin/bash
for v in norep rep norep rep; do
  echo "$v: $(BCP_LIB=build/variants/libtab_$v.so python scripts/profile_target.py 4096 100000 3 2>&1 | grep '^kernel' | tail -2 | tr '\n' ' ')"
done
BCP_LIB=build/variants/libtab_rep.so python scripts/fuzz_parity.py 30 61 big 2>&1 | tail -1
BCP_LIB=build/variants/libtab_rep.so python scripts/probe_c2_auto.py 2>&1 | tail -2
