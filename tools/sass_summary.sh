#!/bin/bash
# SASS evidence of the built library (runs on the build host: no GPU needed).
#   bash tools/sass_summary.sh > profiles/r02_sass_summary.txt
LIB=${1:-pyrate_b200/_lib/libpyrate_b200.so}
TMP=$(mktemp)
cuobjdump -sass "$LIB" > "$TMP"
echo "# SASS evidence of $LIB (cuobjdump -sass, build of $(git rev-parse --short HEAD)+)"
echo "# nvcc -gencode arch=compute_100a,code=sm_100a; cubins:"
cuobjdump -lelf "$LIB" | sed 's/^/#   /'
for op in "UBLKCP.S.G" "UBLKCP.G.S" "SYNCS.ARRIVE.TRANS64" "SYNCS.PHASECHK.TRANS64.TRYWAIT" "VOTE.ALL" "VOTE.ANY" "DFMA" "DMUL" "MUFU.RSQ64H" "MUFU.RCP64H" "LDS.128" "STS.128" "STG.E.EF" "ATOMG" "CALL" "STL" "LDL"; do
  printf "%-40s %s\n" "$op" "$(grep -c "[^A-Z.]$op" "$TMP")"
done
printf "%-40s %s\n" "UTCMMA\|UTCHMMA\|HMMA\|IMMA" "$(grep -c "UTCMMA\|UTCHMMA\|HMMA\|IMMA" "$TMP")"
echo "# 1-D bulk TMA (UBLKCP) both directions + mbarrier transactions (SYNCS): the row streams of the record-streaming kernels."
echo "# No tensor-core instructions (UTC*MMA / HMMA): there is no dense contraction on this path (HBM-bound FP64 streaming)."
echo "# per kernel (SASS instructions):"
awk '/Function :/{name=$3} /^[ \t]+\/\*[0-9a-f]+\*\/ /{n[name]++} END{for (k in n) print n[k], k}' "$TMP" | sort -n | while read cnt name; do printf "%6d  %s\n" "$cnt" "$(echo "$name" | c++filt | sed 's/void pyr:://; s/pyr:://g')"; done
rm -f "$TMP"
