#!/bin/bash
# Opcode histogram of one kernel of an object file / shared library.
# usage: tools/sass_hist.sh <obj|so> '<demangled-name substring>'
cuobjdump -sass "$1" | c++filt | awk -v pat="$2" '
/Function :/ { on = index($0, pat) > 0; if (on) print $0; next }
on && /^[ \t]+\/\*[0-9a-f]+\*\/[ \t]+[A-Z@]/ { s=$0; sub(/^[ \t]+\/\*[0-9a-f]+\*\/[ \t]+/, "", s); if (s ~ /^@/) sub(/^@!?U?P[0-9T]+[ \t]+/, "", s); split(s, a, /[ ;]/); op=a[1]; split(op,b,"."); cnt[b[1]]++; tot++ }
END { for (k in cnt) printf "%6d %s\n", cnt[k], k | "sort -rn"; close("sort -rn"); print tot, "TOTAL" }'
