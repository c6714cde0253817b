#!/bin/bash
# One gpurun call that produces the ncu artefacts profiles/ is built from (tools/profile_summary.py reads them here).
# Numbers printed by bench.py under ncu are never bench values.
set -x
cd "${GRAFT_REPO_ROOT:-.}"
B="python bench.py --steps 2 --warmup 1 --no-cpu"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r01b.csv $B --stream-frames 12 > gpurun_out/prof_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ba_linearize|ba_solve|ba_cost" -s 9 -c 3 -o gpurun_out/prof_ba_r01b $B --stream-frames 0 > gpurun_out/prof_ba.log 2>&1
ncu --set full --clock-control none -k regex:"sel_" -s 24 -c 6 -o gpurun_out/prof_sel_r01b $B --stream-frames 0 > gpurun_out/prof_sel.log 2>&1
ncu --set full --clock-control none -k regex:"ba_marginalize" -s 2 -c 1 -o gpurun_out/prof_marg_r01b $B --stream-frames 16 > gpurun_out/prof_marg.log 2>&1
ls -la gpurun_out | tail -8
# summarise on the box (the .ncu-rep files of --set full are tens of MB each; gpurun_out/ travels back only below 64 MiB)
mkdir -p gpurun_out/prof_out
cp profiles/traffic.json gpurun_out/prof_out/ 2>/dev/null
python tools/profile_summary.py r01b --outdir gpurun_out/prof_out --launches gpurun_out/launches_r01b.csv \
  --rep gpurun_out/prof_ba_r01b.ncu-rep --rep gpurun_out/prof_sel_r01b.ncu-rep --rep gpurun_out/prof_marg_r01b.ncu-rep \
  --note "Round-1 kernels after the DMMA linearize rewrite. BA: bench.py --steps 2 --warmup 1 (592 windows of 11 kf / 1500 features per launch). Selector: N=2000, H=10, kappa=150 (persistent cooperative kernel). Marginalization: closed-loop stream window (L~160, n=75)."
rm -f gpurun_out/prof_sel_r01b.ncu-rep gpurun_out/prof_marg_r01b.ncu-rep
du -sh gpurun_out
