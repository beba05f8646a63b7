"""Per-launch table from an `ncu --metrics gpu__time_duration.sum --csv` log: the last N launches whose name matches."""
import csv, re, sys
path, pat, n = sys.argv[1], sys.argv[2], int(sys.argv[3])
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = []
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(r["Metric Unit"], v)
        rows.append((r["Kernel Name"], v))
sel = [(k, v) for k, v in rows if re.search(pat, k)][-n:]
agg = {}
for k, v in sel:
    k = re.sub(r"\(.*", "", k).replace("void psgla::", "")
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v for _, v in sel)
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-45s n=%3d total %9.1f us  avg %8.1f us  %5.1f%%" % (k, c, t, t / c, 100 * t / tot))
print("TOTAL %.1f us over %d launches" % (tot, len(sel)))
if "--list" in sys.argv:
    for k, v in sel:
        print("  %-45s %8.1f" % (re.sub(r"\(.*", "", k).replace("void psgla::", ""), v))
