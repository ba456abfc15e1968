"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list:  python tools/summarize_launches.py file.csv"""
import csv, re, sys
from collections import defaultdict
rows = []
with open(sys.argv[1], newline="") as fh:
    lines = [l for l in fh if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rd:
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    u = r[iu]
    v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}.get(u, 1e-6)
    name = re.sub(r"\(.*", "", r[ik])
    tot[name] += v; cnt[name] += 1
allt = sum(tot.values())
print(f"# {sum(cnt.values())} launches, {allt:.3f} ms (cold-cache, serialised: shares, not absolutes)")
for k in sorted(tot, key=tot.get, reverse=True):
    print(f"{k[:70]:70s} {cnt[k]:6d} {tot[k]:10.3f} ms {100 * tot[k] / allt:5.1f}%")
