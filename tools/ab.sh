#!/bin/bash
# A/B of library builds under build/variants/*.so (development loop): alternating, 2 rounds; args: variant names
for r in 1 2; do for v in "$@"; do
  echo -n "$v: "; BVIO_LIB_PATH=$PWD/build/variants/$v.so QB_NOTEST=1 bash tools/quick_bench.sh
done; done
