"""Per-source-line stall samples from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`.
usage: ncu_lines.py file.csv [kernel-substring] [top-n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2] if len(sys.argv) > 2 else ""
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
path, func, data = "", "", {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        path = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        func = r[1]
    elif r[0] not in ("", "Line No") and len(r) > 6 and r[0].isdigit():
        try:
            w = int(r[4])
        except ValueError:
            continue
        key = (func[:60], path, int(r[0]))
        prev = data.get(key, (0, ""))
        data[key] = (prev[0] + w, r[1][:120])
funcs = sorted({k[0] for k in data})
for f in funcs:
    if want not in f:
        continue
    items = [(v[0], k[1], k[2], v[1]) for k, v in data.items() if k[0] == f]
    tot = sum(i[0] for i in items) or 1
    print(f"== {f}: {tot} samples")
    for w, p, l, s in sorted(items, reverse=True)[:topn]:
        print(f"{w:6d} {100 * w / tot:5.1f}%  {p}:{l}: {s}")
