#!/bin/bash
# Per-kernel counts of the SASS mnemonics that prove (or disprove) tcgen05 / TMEM / TMA use and register
# spills in the shipped library:  tools/sass_summary.sh > profiles/sass_summary.txt
#   UTCHMMA = tcgen05.mma kind::f16, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor load/store,
#   UBLKCP = cp.async.bulk (1-D bulk copy), LDL/STL = local-memory (spill / stack) traffic
LIB=${1:-aasist_b200/csrc/libaasist_b200.so}
echo "# $(basename $LIB)  sha256=$(sha256sum $LIB | cut -c1-16)  $(date -u +%Y-%m-%dT%H:%MZ)"
printf "%-58s %8s %6s %6s %8s %8s %7s %5s %5s\n" kernel UTCHMMA LDTM STTM UTMALDG UTMASTG UBLKCP LDL STL
cuobjdump -sass "$LIB" | c++filt | awk '
  BEGIN { n = split("UTCHMMA LDTM STTM UTMALDG UTMASTG UBLKCP LDL STL", w, " ") }
  /Function :/ { if (name != "") out(); name = $0; sub(/.*Function : /, "", name); sub(/\(.*/, "", name);
                 sub(/^void /, "", name); gsub(/aasist::/, "", name); delete c; next }
  { for (i = 1; i <= n; i++) if (index($0, w[i] ".") || index($0, w[i] " ")) c[w[i]]++ }
  function out() { printf "%-58s %8d %6d %6d %8d %8d %7d %5d %5d\n", substr(name, 1, 58), c["UTCHMMA"], c["LDTM"], c["STTM"],
                   c["UTMALDG"], c["UTMASTG"], c["UBLKCP"], c["LDL"], c["STL"] }
  END { if (name != "") out() }' | sort
