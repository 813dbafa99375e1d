"""Summarises an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list by
kernel name: launches, total / average device time, share, and (when captured) DRAM bytes per launch.
Usage: python tools/summarize_launches.py <csv> [top_n]"""
import collections, csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
tot = collections.defaultdict(lambda: [0, 0.0, 0.0])     # launches, us, dram bytes
n = 0
TO_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"at::native::.*?(\w+)(<.*)?$", r"torch:\1", name)[:64]
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    m = row.get("Metric Name")
    if m == "gpu__time_duration.sum":
        v = v / 1e3 if unit in ("nsecond", "ns") else (v * 1e3 if unit in ("msecond", "ms") else v)
        tot[name][0] += 1; tot[name][1] += v; n += 1
    elif m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot[name][2] += v * TO_BYTES.get(unit, 1.0)
S = sum(v[1] for v in tot.values())
print("launches %d  total %.0f us" % (n, S))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    extra = "  dram %8.2f MB/launch (%7.1f MB total)" % (v[2] / v[0] / 1e6, v[2] / 1e6) if v[2] > 0 else ""
    print("%9.0f us %5.1f%%  n=%5d  avg=%7.1f  %s%s" % (v[1], 100 * v[1] / S, v[0], v[1] / v[0], k, extra))
