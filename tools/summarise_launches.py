"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv).
    python tools/summarise_launches.py gpurun_out/r02_launches.csv "<command>" > profiles/r02_ncu_launches_summary.json"""
import csv, json, sys
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = {}
for r in rd:
    if len(r) <= iv or r[im] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", ""))
    u = r[iu]
    us = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
    name = r[ik].split("(")[0]
    t = tot.setdefault(name, [0, 0.0])
    t[0] += 1
    t[1] += us
total = sum(t[1] for t in tot.values())
out = {"command": sys.argv[2] if len(sys.argv) > 2 else "", "note": "cold-cache, serialised per-launch times: compare shares, not absolutes",
       "total_us": round(total, 1), "launches": sum(t[0] for t in tot.values()),
       "kernels": [{"kernel": k, "launches": t[0], "total_us": round(t[1], 1), "share": round(t[1] / total, 4),
                    "avg_us": round(t[1] / t[0], 2)} for k, t in sorted(tot.items(), key=lambda kv: -kv[1][1])]}
print(json.dumps(out, indent=1))
