"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean and total (us)."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
per = collections.defaultdict(list)
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    v = v / 1000 if r[ui] == "ns" else v * 1000 if r[ui] == "ms" else v
    per[r[ki][:64]].append(v)
for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:64s} n={len(v):4d} avg={sum(v) / len(v):9.2f} us  total={sum(v):10.1f} us")
