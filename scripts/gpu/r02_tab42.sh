#!/bin/bash
python scripts/probe_c2_auto.py 2>&1 | tail -3
python scripts/profile_target.py 4096 100000 3 2>&1 | tail -2
python scripts/fuzz_parity.py 30 41 big 2>&1 | tail -1
