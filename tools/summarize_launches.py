"""Summarise an ncu launch list (csv with gpu__time_duration.sum) per kernel name."""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr = None
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if hdr is None:
        if "Kernel Name" in r:
            hdr = r
        continue
    if len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", d["Kernel Name"])
    v = float(d["Metric Value"].replace(",", ""))
    unit = d.get("Metric Unit", "ns")
    v = v * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1e-3)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print("total kernel time %.2f ms over %d launches" % (tot / 1e3, sum(v[0] for v in agg.values())))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print("%6.2f%%  %9.1f us  %5d x  %8.1f us/launch  %s" % (100 * t / tot, t, n, t / n, k[:110]))
