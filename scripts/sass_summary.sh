#!/bin/bash
# Per-instantiation SASS summary of the matcher kernels of libmimc3cu.so (packed FP32, local-memory spill ops, vector /
# bulk loads): python -m mimc3_b200.build && bash scripts/sass_summary.sh > profiles/rNN_sass_summary.txt
cd "$(dirname "$0")/.."
LIB=${1:-mimc3_b200/libmimc3cu.so}
echo "cuobjdump -sass $LIB (nvcc $(nvcc --version | grep release | sed 's/.*release //'), sm_100a); instruction counts per kernel"
printf "%-52s %6s %6s %6s %6s %5s %5s %8s %8s %7s %7s %7s\n" kernel instr FFMA2 FADD2 LDS STL LDL LDG.128 LDG.32 LDGSTS REDUX UTMALDG
cuobjdump -sass "$LIB" 2>/dev/null | awk '
/Function :/ { if (name != "") out(); name=$3; n=ff=fa=lds=stl=ldl=l128=l32=rdx=tma=gsts=0; next }
/^ +\/\*[0-9a-f][0-9a-f][0-9a-f][0-9a-f][0-9a-f]*\*\// { n++; if ($0 ~ /FFMA2/) ff++; if ($0 ~ /FADD2/) fa++; if ($0 ~ / LDS/) lds++; if ($0 ~ / STL/) stl++; if ($0 ~ / LDL/) ldl++;
   if ($0 ~ /LDGSTS/) gsts++; else if ($0 ~ /LDG\.E\.128/) l128++; else if ($0 ~ / LDG/) l32++; if ($0 ~ /REDUX/) rdx++; if ($0 ~ /UTMALDG|UBLKCP/) tma++ }
function out() { printf "%s %d %d %d %d %d %d %d %d %d %d %d\n", name, n, ff, fa, lds, stl, ldl, l128, l32, gsts, rdx, tma }
END { out() }' | while read name n ff fa lds stl ldl l128 l32 gsts rdx tma; do
  dn=$(echo "$name" | c++filt | sed -E 's/\(anonymous namespace\):://g; s/\(.*//; s/^void //')
  case "$dn" in *match*|*sat_*|*conv2*|*cluster*|*sweep*) printf "%-52s %6d %6d %6d %6d %5d %5d %8d %8d %7d %7d %7d\n" "$dn" $n $ff $fa $lds $stl $ldl $l128 $l32 $gsts $rdx $tma;; esac
done | sort
