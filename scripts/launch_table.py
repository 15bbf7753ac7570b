"""Per-kernel table of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
    python scripts/launch_table.py gpurun_out/<tag>_launches.csv [top]"""
import csv
import re
import sys
from collections import OrderedDict

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
cols = rows[hdr]
ki, vi = cols.index("Kernel Name"), cols.index("Metric Value")
order = []
for r in rows[hdr + 2:]:
    if len(r) <= vi:
        continue
    short = re.sub(r"\(.*", "", r[ki])
    short = re.sub(r"^void (advk::)?", "", short)
    order.append((short, float(r[vi].replace(",", "")) / 1000.0))
tot = sum(v for _, v in order)
print("total %.1f us in %d launches" % (tot, len(order)))
agg = OrderedDict()
for k, v in order:
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%7.1f us %5.1f%% %3d x %6.1f  %s" % (v, 100 * v / tot, c, v / c, k[:100]))
