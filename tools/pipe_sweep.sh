#!/bin/bash
# e2e of bvio_optimize_batch for a few sub-batch sizes (BVIO_PIPE_WINDOWS: tuning knob of ba_api.cu)
for P in 148 198 296 592; do
  BVIO_PIPE_WINDOWS=$P python bench.py --steps 3 --warmup 3 --no-cpu --stream-frames 0 > gpurun_out/ps_$P.json 2> gpurun_out/ps_$P.err
  python -c "
import json;d=json.loads(open('gpurun_out/ps_$P.json').read().strip().splitlines()[-1]);print('per=$P e2e',d['e2e']['value'],'resident',d['value'])"
done
