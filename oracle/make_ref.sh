#!/bin/bash
# TEST / BENCH INFRASTRUCTURE ONLY.
# Stages the UNMODIFIED reference package (mess42/pyrate, pure Python) under oracle/_ref/
# so that it can travel to the GPU box (oracle/_ref/ is git-ignored, not gpurun-ignored):
# `bench.py --impl reference` and bench.py's cpu_baseline leg then time the reference's own
# OpticalSystem.seqtrace (raytracer/optical_system.py:73-94) on the box's host cores
# (kind "reference"), instead of the NumPy restatement oracle/pyrate_np.py (kind "port").
# Only the Python sources and three small data files (yaml configuration, name lists) are staged -- not the 30 MB
# refractive-index database, which the trace path does not read.  Nothing is edited; the
# three import shims of oracle/refshim.py are applied at import time.
# No reference file ever enters the git history.
set -e
SRC=${1:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
DST=$HERE/_ref
if [ ! -d "$SRC/pyrateoptics" ]; then
  echo "make_ref: no reference tree at $SRC (keeping whatever is staged in $DST)"
  exit 0
fi
rm -rf "$DST"
mkdir -p "$DST"
(cd "$SRC" && find pyrateoptics \( -name '*.py' -o -name '*.yaml' -o -name '*.json' \) \
    -not -path '*refractiveindex.info-database*' -print0 | xargs -0 cp --parents -t "$DST")
(cd "$SRC" && git rev-parse HEAD 2>/dev/null || echo unknown) > "$DST/SOURCE_COMMIT"
echo "make_ref: staged $(find "$DST" -name '*.py' | wc -l) files in $DST"
