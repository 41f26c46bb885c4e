#!/bin/bash
for c in 4 6 8 12 16; do
  echo "cuts $c: $(BCP_CUTS=$c python scripts/probe_e2e.py 2>&1 | tail -2 | sed 's/.*sample/sample/' | tr '\n' ' ')"
done
