#!/bin/bash
# one short resident + e2e measurement and the BA parity tests (development loop; not a bench value of record)
timeout 300 python bench.py --steps 5 --warmup 3 --stream-frames 0 --no-cpu --no-latency --pool 8 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('resident', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],3), {k[:14]: round(v,4) for k,v in d['roofline']['kernel_ms_per_pass'].items()})"
[ -n "$QB_NOTEST" ] || timeout 600 python -m pytest tests/test_gpu_ba.py tests/test_zz_gpu_vs_reference.py -m gpu -q -x 2>&1 | tail -2
