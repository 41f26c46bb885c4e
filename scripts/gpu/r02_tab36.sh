#!/bin/bash
python scripts/probe_burnin.py 4096 1024 30 2>&1 | tail -3
python scripts/probe_burnin.py 4096 1024 16 C5 2>&1 | tail -3
