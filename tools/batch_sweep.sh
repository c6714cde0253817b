#!/bin/bash
# GN iterations/s of the resident solve vs batch size (B = 1, 64, 1024) -- not a bench line
for B in 1 8 64 148 592 1024; do
  python bench.py --steps 3 --warmup 3 --no-cpu --stream-frames 0 --batch $B > gpurun_out/bs_$B.json 2> gpurun_out/bs_$B.err
  python -c "
import json;d=json.loads(open('gpurun_out/bs_$B.json').read().strip().splitlines()[-1]);k=d['roofline']['kernel_ms_per_pass'];print('| $B | %.0f | %.3f | %.0f | %.3f | %.3f | %.3f |' % (d['value'], d['ms_per_step'], d['e2e']['value'], k['ba_linearize_kernel'], k['ba_solve_kernel'], k['ba_cost_kernel']))"
done
