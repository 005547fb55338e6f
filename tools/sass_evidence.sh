#!/bin/bash
# SASS listing kept under profiles/: which of the shipped sm_100a kernels carry TMA, mbarrier, tensor-memory and bulk-copy
# instructions (cuobjdump reads the .so here; no GPU needed).   bash tools/sass_evidence.sh > profiles/r02_sass_evidence.txt
SO=adept_b200/libadept_b200.so
S=$(mktemp)
cuobjdump -sass $SO > $S
echo "# SASS evidence, cuobjdump -sass $SO (sm_100a cubins), round 2 final tree"
echo "# mnemonic            occurrences"
for m in UTMALDG UTMASTG UBLKCP UTMACMDFLUSH SYNCS LDTM STTM UTCATOMSWS DFMA DADD DMUL FFMA FADD MUFU.RCP64H BAR.SYNC ATOMS; do
  printf "%-20s %s\n" $m $(grep -c "[[:space:]]$m" $S)
done
echo "# kernels carrying tensor-memory instructions (row-sum accumulators / parked outputs of the nx = 4096 x-advection; no tcgen05.mma: the path has no contraction)"
awk '/Function :/ {fn=$3} /LDTM|STTM/ {c[fn]++} END {for (f in c) print c[f], f}' $S | sort -rn
echo "# kernels carrying TMA tensor loads / stores or bulk copies"
awk '/Function :/ {fn=$3} /UTMALDG|UTMASTG|UBLKCP/ {c[fn]++} END {for (f in c) print c[f], f}' $S | sort -rn
rm -f $S
