#!/bin/bash
# static SASS opcode histogram per kernel:  tools/sass_hist.sh <regex of kernel names>
cuobjdump -sass ffsubsync_b200/_lib/libffsubsync_b200.so | awk -v pat="$1" '
/Function :/{f=$3; sub(/_ZN[0-9]*_GLOBAL__N__[0-9a-f_]*_cu_[0-9a-f]*/,"",f)}
/^ +\/\*[0-9a-f]+\*\//{ op=$2; if (op ~ /^@/) op=$3; split(op,a,"."); if (f ~ pat) c[f" "a[1]]++ }
END{for(k in c) if (c[k]>30) print k, c[k]}' | sort -k1,1 -k3,3nr
