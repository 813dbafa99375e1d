"""Blackwell-native evidence from the built library: per kernel, the count of the SASS mnemonics that prove tcgen05 /
TMEM / TMA (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, cp.async.bulk.tensor -> UTMALDG /
UTMASTG / UTMAREDG, cp.async.bulk -> UBLKCP) next to the legacy HMMA (mma.sync) count, plus a short excerpt around the
first UTC*MMA of each kernel.  Usage: python tools/sass_summary.py [lib.so] > profiles/r2_sass_summary.txt"""
import collections, os, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "spmm_b200", "libspmm_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
PAT = ["UTCHMMA", "UTCQMMA", "UTCMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "UBLKCP", "HMMA", "FFMA2", "MUFU"]
cur, counts, excerpt = None, collections.OrderedDict(), {}
lines = sass.splitlines()
for i, l in enumerate(lines):
    m = re.match(r"\s*Function : (\S+)", l)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for p in PAT:
        if re.search(r"\b" + p + r"(\.|\b)", l):
            counts[cur][p] += 1
            if p.startswith("UTC") and p != "UTCBAR" and cur not in excerpt:
                excerpt[cur] = [re.sub(r"/\*[0-9a-f]{4}\*/\s*", "", x.split("/* 0x")[0]).strip() for x in lines[max(0, i - 3):i + 4] if "/*" in x]
print("# SASS mnemonic counts per kernel (%s, cuobjdump -sass)" % os.path.basename(lib))
print("%-64s %s" % ("kernel", " ".join("%8s" % p for p in PAT)))
for k, c in counts.items():
    if sum(c.values()):
        print("%-64s %s" % (k[:64], " ".join("%8d" % c[p] for p in PAT)))
print()
for k, ex in excerpt.items():
    print("## %s" % k)
    for x in ex:
        print("    " + x)
