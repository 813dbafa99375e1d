"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections, csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
tot = collections.defaultdict(lambda: [0, 0.0])
n = 0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"at::native::.*?(\w+)(<.*)?$", r"torch:\1", name)[:64]
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    v = v / 1e3 if unit in ("nsecond", "ns") else (v * 1e3 if unit in ("msecond", "ms") else v)
    tot[name][0] += 1; tot[name][1] += v; n += 1
S = sum(v[1] for v in tot.values())
print("launches %d  total %.0f us" % (n, S))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print("%9.0f us %5.1f%%  n=%5d  avg=%7.1f  %s" % (v[1], 100 * v[1] / S, v[0], v[1] / v[0], k))
