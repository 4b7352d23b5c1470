#!/bin/bash
# Opcode histogram of one kernel in an object / library: scripts/sass_hist.sh <file> <substring of mangled name>
cuobjdump -sass "$1" | awk -v pat="$2" '
/Function : /{ on = (index($3, pat) > 0) }
on && /^ +\/\*[0-9a-f]+\*\/ /{ op=$2; if (op ~ /^@/) op=$3; sub(/\..*/,"",op); sub(/;$/,"",op); c[op]++; n++ }
END{ for (o in c) print c[o], o; print n, "TOTAL" }' | sort -rn
