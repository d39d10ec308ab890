#!/bin/bash
# Opcode histogram of the tcgen05 / TMA / TMEM instructions in the shipped library (runs here, no GPU):
#   UTCHMMA = tcgen05.mma, UTMALDG = cp.async.bulk.tensor (TMA load), LDTM = tcgen05.ld, UTCBAR = tcgen05.commit,
#   SYNCS = mbarrier ops, ELECT = elect.sync.   usage: tools/sass_histogram.sh > profiles/rNN_sass_opcodes.txt
cd "$(dirname "$0")/.."
echo "# cuobjdump -sass shapy_b200/libshapy_b200.so ($(stat -c %s shapy_b200/libshapy_b200.so) bytes), per kernel"
cuobjdump -sass shapy_b200/libshapy_b200.so 2>/dev/null | awk '/Function :/ {fn=$3} /(UTCHMMA|UTMALDG|UTMASTG|UTCBAR|LDTM|UTCCP|ELECT|UTMAPF|UBLKCP|SYNCS|STG\.E\.ENL2\.256|LDG\.E\.ENL2\.256)/ { for(i=1;i<=NF;i++) if ($i ~ /^(UTCHMMA|UTMALDG|UTMASTG|UTCBAR|LDTM|UTCCP|ELECT|UTMAPF|UBLKCP|SYNCS|STG\.E\.ENL2\.256|LDG\.E\.ENL2\.256)/) {op=$i; sub(/;$/,"",op); if (op !~ /256/) {split(op,a,"."); op=a[1]} c[fn" "op]++} } END {for (k in c) print k, c[k]}' | sort | c++filt | sed 's/shapy:://'
