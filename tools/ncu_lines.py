"""Per-source-line stall samples from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`.
usage: ncu_lines.py file.csv [kernel-substring] [top-n] [--no-barrier]
With --no-barrier the ranking ignores samples of warps parked at a barrier (idle helpers), which is
what the critical path of a one-worker-per-chain kernel looks like; the dominant stall reasons of
each line are printed next to it."""
import csv
import sys

args = [a for a in sys.argv[1:] if not a.startswith("--")]
nobar = "--no-barrier" in sys.argv
rows = list(csv.reader(open(args[0])))
want = args[1] if len(args) > 1 else ""
topn = int(args[2]) if len(args) > 2 else 30
path, func, data, hdr = "", "", {}, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        path = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        func = r[1]
    elif r[0] == "Line No":
        hdr = r
        stall_cols = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "(" not in h]
        i_samp = hdr.index("# Samples")
    elif r[0].isdigit() and hdr and len(r) > 6:
        try:
            w = int(r[i_samp])
        except ValueError:
            continue
        st = {}
        for i, name in stall_cols:
            try:
                st[name] = int(r[i])
            except (ValueError, IndexError):
                st[name] = 0
        key = (func[:60], path, int(r[0]))
        prev = data.get(key, [0, "", {}])
        prev[0] += w
        prev[1] = r[1][:100]
        for k, v in st.items():
            prev[2][k] = prev[2].get(k, 0) + v
        data[key] = prev
for f in sorted({k[0] for k in data}):
    if want not in f:
        continue
    items = []
    for k, v in data.items():
        if k[0] != f:
            continue
        w = v[0] - (v[2].get("barrier", 0) if nobar else 0)
        items.append((w, k[1], k[2], v[1], v[2]))
    tot = sum(i[0] for i in items) or 1
    print(f"== {f}: {tot} samples" + (" (barrier-parked warps excluded)" if nobar else ""))
    for w, p, l, s, st in sorted(items, key=lambda t: -t[0])[:topn]:
        top = sorted(((v, k) for k, v in st.items() if v and not (nobar and k == "barrier")), reverse=True)[:3]
        print(f"{w:6d} {100 * w / tot:5.1f}%  {p}:{l}: {s}   [" + ", ".join(f"{k} {v}" for v, k in top) + "]")
