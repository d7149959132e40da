"""Development aid: condense an `ncu --csv` launch list (one row per metric) into one line per launch."""
import csv
import sys
from collections import OrderedDict

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H, data = rows[hdr], rows[hdr + 1:]
L = OrderedDict()
for r in data:
    d = dict(zip(H, r))
    k = (d["ID"], d["Kernel Name"].split("::")[-1].split("(")[0])
    L.setdefault(k, {})[d["Metric Name"]] = d["Metric Value"]
limit = int(sys.argv[2]) if len(sys.argv) > 2 else 10 ** 9
for i, (k, v) in enumerate(L.items()):
    if i >= limit:
        break
    print(k[0], k[1], {a.split("__")[-1][:30]: b for a, b in v.items()})
