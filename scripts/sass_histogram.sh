#!/bin/bash
# SASS opcode histogram of the tensor-core kernels in the built library: the mnemonics that prove tcgen05 / TMEM / TMA.
# usage: scripts/sass_histogram.sh > profiles/r2_sass_histogram.txt
set -e
LIB=${1:-mipsfusion_b200/libmipsfusion_b200.so}
for K in field_bwd_tc2_kernel field_fwd_tc3_kernel field_bwd_tc_kernel adam_sharded_kernel; do
  cuobjdump -sass "$LIB" 2>/dev/null | awk -v k="$K" '
    /Function :/ { on = index($0, k) > 0; if (on) name = $0 }
    on && /^\s+\/\*[0-9a-f]{4}\*\// { op = $2; sub(/;$/, "", op); split(op, a, "."); h[name][a[1]]++; n[name]++ }
    END { for (f in h) { printf "\n%s  (%d instructions)\n", f, n[f];
            m = 0; for (o in h[f]) { line[m++] = sprintf("%7d %s", h[f][o], o) }
            asort(line); for (i = m; i >= 1 && i > m - 28; --i) print line[i]; delete line } }'
done
