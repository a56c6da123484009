"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel and grid."""
import collections, csv, re, sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
agg, tot = collections.OrderedDict(), 0.0
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    v = v / 1000 if unit == "ns" else v * 1000 if unit == "ms" else v
    a = agg.setdefault((name, row["Grid Size"]), [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
print(f"total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches")
for (name, grid), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{t:10.1f} us {100*t/tot:5.1f}%  {c:4d}x  {t/c:8.2f} us/launch  {name[:64]} grid={grid}")
