#!/usr/bin/env python
"""Turn the ncu artefacts a gpurun call left in gpurun_out/ into the tracked summaries under profiles/.
  python tools/profile_summary.py <tag> --launches gpurun_out/launches.csv --rep gpurun_out/prof_ba.ncu-rep [--rep ...]
Writes profiles/<tag>_launches.md, profiles/<tag>_kernels.md and updates profiles/traffic.json."""
import argparse, collections, csv, io, json, os, subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM (smem limit)"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard (smem)"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard (global)"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait (fixed latency)"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe"),
]


def kname(full):
    """'void ba_linearize_kernel<1>(BaBatch)' -> 'ba_linearize_kernel'"""
    n = full.replace("(anonymous namespace)::", "").split("(")[0].strip()
    if n.startswith("void "):
        n = n[5:]
    return n.split("<")[0].split("::")[-1]


def launches_md(path, out):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(list)
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        agg[kname(row["Kernel Name"])].append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list ({os.path.basename(path)}): gpu__time_duration.sum, --clock-control none\n\n")
        f.write("Cold-cache, serialised per-launch times: compare SHARES, not absolutes.\n\n")
        f.write("| kernel | launches | total us | mean us | share |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| {k} | {len(v)} | {sum(v)/1e3:.1f} | {sum(v)/len(v)/1e3:.2f} | {sum(v)/tot:.3f} |\n")


def rep_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--launches")
    ap.add_argument("--rep", action="append", default=[])
    ap.add_argument("--note", default="")
    ap.add_argument("--outdir", default=os.path.join(ROOT, "profiles"),
                    help="where to write (on the GPU box: a directory under gpurun_out/, the only one that travels back)")
    a = ap.parse_args()
    pd = a.outdir
    os.makedirs(pd, exist_ok=True)
    if a.launches:
        launches_md(a.launches, os.path.join(pd, f"{a.tag}_launches.md"))
    traffic_path = os.path.join(pd, "traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    seen_traffic = set()
    if a.rep:
        with open(os.path.join(pd, f"{a.tag}_kernels.md"), "w") as f:
            f.write(f"# ncu --set full --clock-control none ({a.tag})\n\n{a.note}\n\n")
            for rep in a.rep:
                hdr, units, rows = rep_rows(rep)
                idx = {h: i for i, h in enumerate(hdr)}
                seen = collections.Counter()
                for r in rows:
                    name = kname(r[idx["Kernel Name"]])
                    seen[name] += 1
                    if seen[name] > 1:
                        continue
                    f.write(f"## {name}  (grid {r[idx['Grid Size']]}, block {r[idx['Block Size']]}; from {os.path.basename(rep)})\n\n")
                    f.write("| metric | value | unit |\n|---|---:|---|\n")
                    for k, label in KEYS:
                        if k in idx:
                            f.write(f"| {label} (`{k}`) | {r[idx[k]]} | {units[idx[k]]} |\n")
                    f.write("\n")
                    try:
                        def b(k):
                            v, u = float(r[idx[k]].replace(",", "")), units[idx[k]].lower()
                            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
                        if name not in seen_traffic:        # the first --rep that holds a kernel wins (list the headline capture first)
                            traffic[name] = int(b("dram__bytes_read.sum") + b("dram__bytes_write.sum"))
                            seen_traffic.add(name)
                    except Exception:
                        pass
        json.dump(traffic, open(traffic_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
