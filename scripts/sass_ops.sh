#!/bin/bash
# Per-kernel counts of the SASS mnemonics that prove what the library runs on (sm_100a): tcgen05 MMAs (UTC*MMA),
# TMA loads/stores (UTMALDG/UTMASTG), bulk copies (UBLKCP), TMEM loads/stores (LDTM/STTM), packed fp32 FMAs (FFMA2),
# grid-wide reductions (RED), cluster barriers / distributed shared memory.   usage: scripts/sass_ops.sh > profiles/r2_sass_ops.txt
LIB="$(dirname "$0")/../mdbn_b200/csrc/libmdbn_b200.so"
echo "# cuobjdump -sass $(basename $LIB) ($(date -u +%Y-%m-%d)), instruction counts per kernel"
cuobjdump -sass "$LIB" | awk '
/Function :/ { fn=$3; next }
/^[ \t]+\/\*[0-9a-f]+\*\// {
  n[fn]++
  if ($0 ~ /UTC[A-Z]*MMA/) a[fn,"UTCxMMA"]++
  if ($0 ~ /UTMALDG/) a[fn,"UTMALDG"]++
  if ($0 ~ /UTMASTG/) a[fn,"UTMASTG"]++
  if ($0 ~ /UBLKCP/) a[fn,"UBLKCP"]++
  if ($0 ~ /LDTM/) a[fn,"LDTM"]++
  if ($0 ~ /STTM/) a[fn,"STTM"]++
  if ($0 ~ /FFMA2/) a[fn,"FFMA2"]++
  if ($0 ~ / FFMA /) a[fn,"FFMA"]++
  if ($0 ~ /RED\.E/) a[fn,"RED"]++
  if ($0 ~ /SYNCS/) a[fn,"SYNCS(mbarrier)"]++
  if ($0 ~ /UCGABAR|CGABAR/) a[fn,"CGABAR(cluster)"]++
  if ($0 ~ /LDGSTS/) a[fn,"LDGSTS(cp.async)"]++
}
END {
  split("UTCxMMA UTMALDG UTMASTG UBLKCP LDTM STTM FFMA2 FFMA RED SYNCS(mbarrier) CGABAR(cluster) LDGSTS(cp.async)", keys, " ")
  for (f in n) {
    line = sprintf("%-110s instr=%6d", f, n[f])
    for (i = 1; i <= 12; i++) if (a[f,keys[i]] > 0) line = line sprintf("  %s=%d", keys[i], a[f,keys[i]])
    print line
  }
}' | sort | c++filt 2>/dev/null | sed 's/CUtensorMap_st/TMap/g'
