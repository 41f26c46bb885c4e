#!/bin/bash
BCP_SPEC_STATS=1 python scripts/probe_c2_auto.py 2>&1 | grep -v "spec stats" | tail -12
BCP_SPEC_STATS=1 python scripts/probe_e2e.py 2>&1 | grep -v "spec stats" | tail -12
python scripts/fuzz_parity.py 30 5 2>&1 | tail -1
python scripts/fuzz_parity.py 20 6 big 2>&1 | tail -1
