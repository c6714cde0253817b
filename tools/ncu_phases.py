#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` SASS export by named source-line ranges (kernel phases).
Usage: ncu_phases.py <sass.csv> <nvdisasm -g -c listing> <kernel substring> <file> name:lo-hi [name:lo-hi ...]"""
import csv, re, sys, collections

def main():
    sass_csv, listing, kname, fname = sys.argv[1:5]
    ranges = []
    for a in sys.argv[5:]:
        n, r = a.split(":"); lo, hi = r.split("-"); ranges.append((n, int(lo), int(hi)))
    addr2line, cur, infn = {}, None, False
    for ln in open(listing):
        if ln.startswith(".text."):
            infn = kname in ln; continue
        if not infn: continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if m: addr2line[int(m.group(1), 16)] = cur
    rows = list(csv.reader(open(sass_csv)))
    hdr = rows[1]
    ia, ii, isamp, isrc = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    def phase(k):
        if k is None: return "none"
        f, l = k
        if f != fname: return "inl:" + f
        for n, lo, hi in ranges:
            if lo <= l < hi: return n
        return "other"
    ph = collections.defaultdict(lambda: [0, 0, collections.Counter(), collections.Counter()])
    base = None
    for r in rows[2:]:
        try: a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
        except ValueError: continue
        if base is None: base = a
        p = ph[phase(addr2line.get(a - base))]
        n = int(r[ii] or 0); p[0] += n; p[1] += int(r[isamp] or 0)
        for c in stall_cols:
            v = int(r[c] or 0)
            if v: p[2][hdr[c][6:]] += v
        toks = r[isrc].split()
        op = (toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")).split(".")[0]
        p[3][op] += n
    ti = sum(v[0] for v in ph.values()) or 1; ts = sum(v[1] for v in ph.values()) or 1
    print(f"total warp-instructions {ti}, samples {ts}")
    for n, v in sorted(ph.items(), key=lambda kv: -kv[1][1]):
        print(f"{n:26s} inst {v[0]/ti*100:5.1f}%  samples {v[1]/ts*100:5.1f}%  stalls " +
              ",".join(f"{k}:{c*100//max(v[1],1)}" for k, c in v[2].most_common(3)) + "  ops " +
              ",".join(f"{o}:{c*100//max(v[0],1)}" for o, c in v[3].most_common(6)))

if __name__ == "__main__":
    main()
