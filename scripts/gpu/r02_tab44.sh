#!/bin/bash
python scripts/fuzz_parity.py 60 51 big 2>&1 | tail -1
python scripts/fuzz_parity.py 30 52 2>&1 | tail -1
BCP_SPEC_STATS=1 python scripts/profile_target.py 4096 100000 3 2>&1 | grep -v trace | tail -6
python scripts/probe_c2_auto.py 2>&1 | tail -3
python scripts/probe_e2e.py 2>&1 | tail -2
