"""Print the numbers of a bench.py JSON line that the docs quote (development helper)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
if d.get("impl") == "reference":
    print("reference arm:", round(d["value"], 1), d["unit"], "cores", d["cpu_baseline"]["cores"])
    sys.exit(0)
print({k: d[k] for k in ("value", "n_gpus", "steps", "ms_per_step", "gpu_launches")}, "clocks", d["clocks"])
print("e2e", round(d["e2e"]["value"]), "lm", round(d["lm"]["value"]))
r = d["roofline"]
print("roofline frac", round(r["frac"], 4), "traffic", r["traffic"], "alg bytes", r["algorithmic_bytes_per_launch"], "fp64 frac", round(r["fp64"]["frac"], 4))
print("ms per pass", {k: round(v, 4) for k, v in r["kernel_ms_per_pass"].items()})
if d.get("cpu_baseline"):
    print("cpu_baseline", round(d["cpu_baseline"]["value"], 1), "cores", d["cpu_baseline"]["cores"])
s = d["selector"]
print("selector ms", round(s["ms_per_select"], 3), "h13", round(s["h13"]["ms_per_select"], 3), "e2e", round(s["e2e"]["ms_per_select"], 3))
if s.get("strong_scaling"):
    for c in s["strong_scaling"]:
        print("  N", c["N"], "H", c["H"], "one", round(c["one_gpu_ms"], 2), "sharded", round(c["sharded_ms"], 2), c["sharded"].get("transport", "")[:24])
for b in d.get("single_window") or []:
    print("single L", b["L"], "e2e p50", round(b["e2e_ms_p50"], 3), "cpu", round(b.get("cpu_oracle_ms", 0), 2), {k: round(v * 1e3, 1) for k, v in b["kernel_ms_per_pass"].items()})
for v in d.get("stream") or []:
    print(v["label"][:50], v["frames"], "opt", round(v["optimize_ms"]["p50"], 3), "marg", round(v["marginalize_ms"]["p50"], 3), "sel", round(v["select_ms"]["p50"], 3),
          "call p50/p99", round(v["frame_call_ms"]["p50"], 3), round(v["frame_call_ms"]["p99"], 3), "cpu", (v.get("cpu_oracle") or {}).get("frame_ms_p50"))
