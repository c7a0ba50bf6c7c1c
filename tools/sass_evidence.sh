#!/bin/bash
# Per kernel: how many tcgen05 / TMEM / bulk-copy / mbarrier / warp-level MMA instructions its SASS holds
# (B200_PROFILING.md "What proves a Blackwell-native kernel").  usage: tools/sass_evidence.sh [lib.so]
LIB=${1:-keypoint_diffusion_b200/libkpdiff_b200.so}
cuobjdump -sass "$LIB" | awk '
/Function :/ { fn=$3; next }
{ for (i=1;i<=NF;i++) if ($i ~ /^(UTC[A-Z]*MMA|UTCBAR|UTCATOMSWS|UBLKCP|UTMALDG|UTMASTG|LDTM|STTM|HMMA|LDGSTS)/) { m=$i; sub(/\..*/,"",m); c[fn" "m]++; break } }
END { for (k in c) print k, c[k] }' | sort | awk '{ if ($1!=last) { if (last!="") print line; line=$1":"; last=$1 } line=line" "$2"="$3 } END { print line }' | c++filt 2>/dev/null
