#!/bin/bash
# One gpurun call that produces the ncu artefacts profiles/r02_* is built from (tools/profile_summary.py reads them here).
# Numbers printed by bench.py under ncu are never bench values.  --pool 4: the pool construction (4 one-window solves +
# marginalizations, ~160 launches) is skipped with -s; everything after it is the benchmark proper.
set -x
cd "${GRAFT_REPO_ROOT:-.}"
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-latency --pool 4"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 160 -c 1200 --csv --log-file gpurun_out/launches_r02.csv $B --stream-frames 12 > gpurun_out/prof_launch.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"ba_linearize|ba_imu|ba_solve|ba_cost|ba_dogleg" -s 190 -c 5 -o gpurun_out/prof_ba_r02 $B --stream-frames 0 > gpurun_out/prof_ba.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"ba_linearize|ba_imu|ba_solve|ba_cost|ba_dogleg" -s 4 -c 4 -o gpurun_out/prof_ba1_r02 $B --stream-frames 0 > gpurun_out/prof_ba1.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"sel_" -s 0 -c 5 -o gpurun_out/prof_sel_r02 $B --stream-frames 0 > gpurun_out/prof_sel.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"ba_marg" -s 8 -c 2 -o gpurun_out/prof_marg_r02 $B --stream-frames 16 > gpurun_out/prof_marg.log 2>&1
ls -la gpurun_out | tail -8
mkdir -p gpurun_out/prof_out
cp profiles/traffic.json gpurun_out/prof_out/ 2>/dev/null
python tools/profile_summary.py r02 --outdir gpurun_out/prof_out --launches gpurun_out/launches_r02.csv \
  --rep gpurun_out/prof_ba_r02.ncu-rep --rep gpurun_out/prof_ba1_r02.ncu-rep --rep gpurun_out/prof_sel_r02.ncu-rep --rep gpurun_out/prof_marg_r02.ncu-rep \
  --note "Round 2. BA (first block): bench.py --steps 2 --warmup 1, 592 windows of 11 kf / ~1550 features with real n = 75 marginalized priors per launch, traditional dogleg. BA (second block): the same kernels at B = 1 (latency mode: 512-thread solve, one CTA per IMU factor), L ~ 1550. Selector: N=2000, H=10, kappa=150 (persistent cooperative kernel). Marginalization: closed-loop stream window (L~160, n=75): ba_marg_factors_kernel + ba_marginalize_kernel."
# per-instruction stall profile of the marginalization kernel's hottest lines (source page)
ncu -i gpurun_out/prof_marg_r02.ncu-rep --page source --csv > gpurun_out/prof_out/r02_marg_source.csv 2>/dev/null
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/prof_out/r02_marg_source.csv')))
hdr = None
for i, r in enumerate(rows):
    if '# Samples' in ''.join(r) or 'Warp Stall Sampling (All Samples)' in r:
        hdr = i; break
print('source rows', len(rows), 'hdr', hdr)
PY
rm -f gpurun_out/prof_sel_r02.ncu-rep gpurun_out/prof_ba1_r02.ncu-rep
du -sh gpurun_out
