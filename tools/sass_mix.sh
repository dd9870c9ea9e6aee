#!/bin/bash
# instruction mix of one kernel in an object file: tools/sass_mix.sh obj.o <mangled-name-substring>
cuobjdump -sass "$1" 2>/dev/null | awk -v pat="$2" '/Function :/{on = index($0, pat) > 0} on' | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" | sed -E 's#^\s+/\*[0-9a-f]+\*/\s+##' | sed -E 's/^@!?U?P[0-9T]+\s+//' | awk '{split($1,a,"."); c[a[1]]++; n++} END{for(k in c) printf "%6d %s\n", c[k], k; printf "%6d TOTAL\n", n}' | sort -rn | head -${3:-25}
