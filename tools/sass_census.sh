#!/bin/bash
# Static SASS census of libbvio.so: DMMA / DFMA / local-memory / barrier instruction counts per kernel.
# Runs without a GPU (cuobjdump only).  Output format: profiles/r01b_sass.md.
SO=${1:-anticipated-vins-mono_b200/csrc/libbvio.so}
cuobjdump -sass "$SO" | awk '/Function :/{fn=$3} /DMMA/{d[fn]++} /DFMA/{f[fn]++} /LDL|STL/{s[fn]++} /BAR.SYNC|BAR.ARV/{b[fn]++}
  END{for(k in f) printf "%s %d %d %d %d\n", k, d[k], f[k], s[k], b[k]}' | while read k d f s b; do
  echo "| \`$(echo "$k" | c++filt | sed 's/(anonymous namespace):://; s/bvio:://g; s/(.*//; s/^void //')\` | $d | $f | $s | $b |"
done | sort
