#!/bin/bash
python scripts/profile_target.py 4096 100000 3 2>&1 | grep '^kernel' | tail -1
python scripts/probe_c2_auto.py 2>&1 | tail -1
python scripts/fuzz_parity.py 20 91 big 2>&1 | tail -1
python -m pytest tests/test_full_size_gpu.py tests/test_sampler_gpu.py -x -q -m gpu 2>&1 | tail -2
