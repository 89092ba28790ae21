#!/bin/bash
# Static summary of the shipped library (no GPU needed): registers / stack / shared memory per kernel (cuobjdump -res-usage) and,
# per kernel, the number of SASS instructions with a histogram of the mnemonics that matter for this path (ALU pipe: LOP3, SHF,
# IADD3, POPC; FMA pipe: IMAD; SHFL; LDS/STS; LDG/STG; local-memory LDL/STL; bulk copies UBLKCP; tensor-core UTC*MMA / HMMA).
# Usage: bash profiles/scripts/sass_summary.sh > profiles/<tag>_sass_summary.txt
LIB=${1:-astar_pairwise_aligner_b200/libastarpa_c.so}
echo "== $LIB: resources per kernel"
cuobjdump -res-usage "$LIB" | grep -A1 "^ Function" | grep -v "^--" | paste - - | sed -e 's/ Function //' -e 's/CONSTANT.*//' | c++filt | sort
echo
echo "== SASS instructions per kernel: total | LOP3 SHF IADD3 POPC | IMAD | SHFL | LDS STS | LDG STG ATOM/RED | LDL STL | BAR | UBLKCP | MMA"
cuobjdump -sass "$LIB" | awk '
  /Function :/ { if (name != "") flush(); name = $3; delete c; total = 0; next }
  /^ +\/\*[0-9a-f]+\*\/ +[A-Z@]/ {
      ins = $2; if (ins ~ /^@/) ins = $3; sub(/\..*/, "", ins); sub(/;$/, "", ins); c[ins]++; total++ }
  END { flush() }
  function flush() {
      printf "%s\t%d | %d %d %d %d | %d | %d | %d %d | %d %d %d | %d %d | %d | %d | %d\n", name, total, c["LOP3"], c["SHF"], c["IADD3"] + c["IADD"], c["POPC"],
             c["IMAD"], c["SHFL"], c["LDS"], c["STS"], c["LDG"], c["STG"], c["ATOM"] + c["ATOMG"] + c["RED"] + c["ATOMS"], c["LDL"], c["STL"],
             c["BAR"], c["UBLKCP"], c["HMMA"] + c["IMMA"] + c["UTCHMMA"] + c["UTCIMMA"] + c["UTCQMMA"]
  }' | c++filt | sort
