#!/bin/bash
for v in k0 k1 k0 k1; do
  L=build/variants/libtab_$v.so
  c3=$(BCP_LIB=$L python scripts/profile_target.py 4096 100000 3 2>&1 | grep '^kernel' | tail -1)
  c2=$(BCP_LIB=$L python scripts/probe_c2_auto.py 2>&1 | tail -1)
  echo "$v | C3 $c3 | C2 $c2"
done
BCP_LIB=build/variants/libtab_k1.so python scripts/fuzz_parity.py 30 71 big 2>&1 | tail -1
