"""Per-kernel time shares of an ncu launch list (--metrics gpu__time_duration.sum --csv).   python tools/launch_share.py <csv>"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.defaultdict(float), collections.Counter()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
    n = re.sub(r"\(.*", "", r[ki])[:90]
    tot[n] += v
    cnt[n] += 1
s = sum(tot.values())
for n, v in sorted(tot.items(), key=lambda x: -x[1]):
    print(f"{100 * v / s:6.2f}% {v:9.1f} us n={cnt[n]:3d} avg {v / cnt[n]:7.1f}  {n}")
