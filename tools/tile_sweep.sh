#!/bin/bash
# usage: tile_sweep.sh <batch> T...
B=$1; shift
for T in "$@"; do
  BVIO_TILES=$T python bench.py --steps 5 --warmup 3 --no-cpu --stream-frames 0 --batch $B > gpurun_out/ts_$T.json 2> gpurun_out/ts_$T.err
  python -c "
import json;d=json.loads(open('gpurun_out/ts_$T.json').read().strip().splitlines()[-1]);print('B=$B T=$T',d['value'],d['ms_per_step'],d['roofline']['kernel_ms_per_pass'])"
done
