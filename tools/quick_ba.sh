#!/bin/bash
# quick GPU loop: BA + marginalization parity tests, then a short bench (no CPU leg, no stream leg)
python -m pytest tests/test_gpu_ba.py tests/test_gpu_marg.py -m gpu -x -q 2>&1 | tail -15
python bench.py --steps ${1:-5} --warmup 3 --no-cpu --stream-frames 0 > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/quick_bench.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['roofline']['kernel_ms_per_pass'],d['config']['final_cost_check'], d['selector']['value'])"
tail -3 gpurun_out/quick_bench.err
