#!/usr/bin/env python
"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list.

    python tools/summarize_launches.py gpurun_out/launches.csv [skip_first_n_launches]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]
unit = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}
d = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(h) or not r[0].isdigit() or int(r[0]) < skip:
        continue
    k = r[h.index("Kernel Name")].split("(")[0][-70:]
    d.setdefault(k, []).append(float(r[-1].replace(",", "")) * unit[r[h.index("Metric Unit")]])
tot = sum(sum(v) for v in d.values())
for k, v in d.items():
    print(f"{k:70s} n={len(v):4d} avg={sum(v) / len(v):10.1f} us  total={sum(v) / 1e3:9.3f} ms {100 * sum(v) / tot:5.1f}%")
print(f"total {tot / 1e3:.3f} ms")
