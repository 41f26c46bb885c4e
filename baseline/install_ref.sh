#!/bin/bash
# Installs the UNMODIFIED reference (vvoelz/biceps v2.0, /root/reference) into baseline/_ref (git-ignored, travels
# to the GPU box with the snapshot) for `bench.py --impl reference`.  Offline: no index, no dependency resolution.
# The reference imports mdtraj / pymbar / matplotlib / natsort at module top (Restraint.py:5, toolbox.py:7-8,
# Analysis.py:5-12); none of them is in this image and none is used by PosteriorSampler.sample, so the import shims
# of oracle/stubs (a dozen lines, no algorithm) are placed next to the package.
set -e
cd "$(dirname "$0")/.."
SRC=${BICEPS_REFERENCE_ROOT:-/root/reference}
[ -d "$SRC/biceps" ] || { echo "no reference tree at $SRC"; exit 1; }
TMP=$(mktemp -d)
cp -r "$SRC" "$TMP/ref"          # the build writes egg-info into the source tree; /root/reference is read-only
rm -rf baseline/_ref
python -m pip install --quiet --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target baseline/_ref "$TMP/ref"
cp -r oracle/stubs/* baseline/_ref/
rm -rf "$TMP"
echo "reference installed: $(ls baseline/_ref | tr '\n' ' ')"
