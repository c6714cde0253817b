#!/bin/bash
# single-window latency: solve kernel with 512 (default for small batches) vs 256 threads
python -m pytest tests/test_gpu_ba.py tests/test_slider.py -m gpu -q -x 2>&1 | tail -2
for mode in wide narrow; do
  if [ $mode = narrow ]; then export BVIO_SOLVE_NARROW=1; else unset BVIO_SOLVE_NARROW; fi
  python bench.py --steps 3 --warmup 3 --no-cpu --stream-frames 150 --batch 1 > gpurun_out/lat_$mode.json 2> gpurun_out/lat_$mode.err
  python -c "
import json;d=json.loads(open('gpurun_out/lat_$mode.json').read().strip().splitlines()[-1]);k=d['roofline']['kernel_ms_per_pass'];s=d['stream'];print('$mode', 'B=1 ms_per_solve', d['ms_per_step'], k, 'stream optimize p50', s['optimize_call_ms']['p50'], 'frame', s['frame_call_ms']['p50'])"
done
