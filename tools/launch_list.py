"""Mean duration per kernel from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
h = rows[0]
k, v, u = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
d = collections.defaultdict(list)
for r in rows[1:]:
    try:
        x = float(r[v].replace(",", ""))
    except ValueError:
        continue
    x = x / 1000.0 if r[u] in ("ns", "nsecond") else x * 1000.0 if r[u] in ("ms", "msecond") else x
    d[r[k].split("(")[0].split("::")[-1]].append(x)
tot = sum(sum(x) for x in d.values())
for n, x in d.items():
    print(f"   launch list: {n:28s} n={len(x):3d} mean {sum(x) / len(x):8.2f} us  share {100 * sum(x) / tot:5.1f} %")
