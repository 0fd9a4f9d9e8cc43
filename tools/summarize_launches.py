"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`): launches, total and
mean device time and share per kernel.  Usage: python tools/summarize_launches.py launches.csv"""
import collections
import csv
import re
import sys

rows = list(csv.DictReader(l for l in open(sys.argv[1]) if l.startswith('"')))
agg = collections.OrderedDict()
for r in rows:
    n = r["Kernel Name"]
    m = re.search(r"(kv_\w+)(<.*?>)?\(", n)
    if m:
        short = "b200kv::" + m.group(1) + (m.group(2) or "") + f" grid={r['Grid Size']} block={r['Block Size']}"
    else:
        short = "[torch] " + re.sub(r"\(.*", "", n).replace("void ", "")[:70]
    d = agg.setdefault(short, [0, 0.0])
    d[0] += 1
    d[1] += float(r["Metric Value"])
tot = sum(v[1] for v in agg.values())
ours = sum(v[1] for k, v in agg.items() if k.startswith("b200kv::"))
print(f"{len(rows)} launches, {tot / 1e6:.3f} ms device time; b200kv kernels {ours / 1e6:.3f} ms ({100 * ours / tot:.1f}%)")
for k, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{c:6d} launches {ns / 1e6:10.3f} ms total {ns / c / 1e3:10.1f} us mean {100 * ns / tot:6.2f}%  {k}")
