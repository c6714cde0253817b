#!/bin/bash
# one ncu --set full capture (with source correlation) of the linearize kernel at the benched batch
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-latency --pool 4 --stream-frames 0"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"ba_linearize" -s 150 -c 1 -f -o gpurun_out/prof_lin_ws $B > gpurun_out/prof_lin_ws.log 2>&1
tail -3 gpurun_out/prof_lin_ws.log; ls -la gpurun_out/prof_lin_ws.ncu-rep
