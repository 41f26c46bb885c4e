#!/bin/bash
BCP_SPEC_STATS=1 python scripts/probe_e2e.py 2>&1 | tail -24
