"""Aggregate an ncu --csv launch list (gpu__time_duration.sum) by kernel name: count, total us, share."""
import collections, csv, re, sys
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v = v / 1000 if u in ("ns", "nsecond") else v * (1000 if u in ("ms", "msecond") else 1)
        rows.append((re.sub(r"\(.*", "", r["Kernel Name"])[:90], v))
agg = collections.defaultdict(lambda: [0, 0.0])
for k, v in rows:
    agg[k][0] += 1; agg[k][1] += v
tot = sum(v for _, v in rows)
print(f"{len(rows)} launches, {tot:.1f} us total")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{t:10.1f} us {100 * t / tot:5.1f}% x{c:4d}  {k}")
