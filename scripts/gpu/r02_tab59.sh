#!/bin/bash
for v in cu1 cu2 cu4 sp2; do
  L=build/variants/libtab_$v.so
  c3=$(BCP_LIB=$L python scripts/profile_target.py 4096 100000 3 2>&1 | grep '^kernel' | tail -1)
  c2=$(BCP_LIB=$L python scripts/probe_c2_auto.py 2>&1 | tail -1)
  e2=$(BCP_LIB=$L python scripts/probe_e2e.py 2>&1 | tail -1 | sed 's/.*sample/sample/' | cut -c1-60)
  echo "$v | C3 $c3 | C2 $c2 | $e2"
done
