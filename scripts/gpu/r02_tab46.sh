#!/bin/bash
for v in st_norep st_rep; do
  echo "== $v"; BCP_SPEC_STATS=1 BCP_TAB_STATS=1 BCP_LIB=build/variants/libtab_$v.so python scripts/profile_target.py 4096 100000 2 2>&1 | grep -v "^trace" | tail -3
done
