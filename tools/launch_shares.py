"""Kernel shares of an ncu launch list (gpu__time_duration.sum per launch).  python tools/launch_shares.py <launches.csv> "<command>" """
import collections, csv, json, sys
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
h = rows[0]; ik, ig, iv = h.index("Kernel Name"), h.index("Grid Size"), h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    try: v = float(r[iv].replace(",", ""))
    except ValueError: continue
    unit = r[h.index("Metric Unit")]
    us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
    k = (r[ik][:60], r[ig])
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
out = {"command": sys.argv[2] if len(sys.argv) > 2 else "", "note": "per-launch times are cold-cache and serialised under ncu: compare shares, not absolutes",
       "kernels": [{"kernel": k[0], "grid": k[1], "launches": a[0], "total_us": round(a[1], 1), "share": round(a[1] / tot, 4), "avg_us": round(a[1] / a[0], 1)}
                   for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])]}
print(json.dumps(out, indent=1))
