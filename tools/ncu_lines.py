#!/usr/bin/env python
"""Join an `ncu --page source --csv` SASS export with `nvdisasm -g -c` line info and print the hottest
CUDA source lines (instructions executed + stall samples).  Usage:
  ncu_lines.py <sass.csv> <nvdisasm listing> <mangled-or-substring kernel name> [source file] [top N]"""
import csv, re, sys, collections

def main():
    sass_csv, listing, kname = sys.argv[1:4]
    src = sys.argv[4] if len(sys.argv) > 4 else None
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    # address -> line from the listing
    addr2line, cur, infn = {}, None, False
    for ln in open(listing):
        if ln.startswith(".text."):
            infn = kname in ln
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if m:
            addr2line[int(m.group(1), 16)] = cur
    rows = list(csv.reader(open(sass_csv)))
    hdr = rows[1]
    ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    base = None
    agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    for r in rows[2:]:
        try:
            a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
        except ValueError:
            continue
        if base is None:
            base = a
        key = addr2line.get(a - base)
        e = agg[key]
        e[0] += int(r[ii] or 0)
        e[1] += int(r[isamp] or 0)
        for c in stall_cols:
            v = int(r[c] or 0)
            if v:
                e[2][hdr[c]] += v
    ti = sum(v[0] for v in agg.values()) or 1
    ts = sum(v[1] for v in agg.values()) or 1
    lines = open(src).read().split("\n") if src else None
    print(f"total warp-instructions {ti}, samples {ts}")
    for key, v in sorted(agg.items(), key=lambda kv: -(kv[1][0] if "--by-inst" in sys.argv else kv[1][1]))[:top]:
        text = ""
        if lines and key and key[1] - 1 < len(lines):
            text = lines[key[1] - 1].strip()[:90]
        st = ",".join(f"{k[6:]}:{n * 100 // max(v[1], 1)}" for k, n in v[2].most_common(3))
        print(f"{v[0] / ti * 100:5.1f}%i {v[1] / ts * 100:5.1f}%s  {str(key):28s} [{st}] {text}")

if __name__ == "__main__":
    main()
