#!/bin/bash
BCP_SPEC_STATS=1 python scripts/probe_c2_auto.py 2>&1 | grep -v "spec stats" | tail -12
BCP_SPEC_STATS=1 python scripts/probe_e2e.py 2>&1 | grep -v "spec stats" | tail -6
python -m pytest tests/test_full_size_gpu.py -x -q -m gpu 2>&1 | tail -3
