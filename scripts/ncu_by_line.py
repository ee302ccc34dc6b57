"""Aggregate `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv` by CUDA source line:
share of warp-stall samples and of executed warp instructions per line (top N) and the stall mix."""
import csv
import sys

path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 50
fname, hdr, out = "", None, []
for r in csv.reader(open(path)):
    if not r:
        continue
    if r[0] == "File Name":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = {}
        for j, k in enumerate(r):
            hdr.setdefault(k, j)
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
        continue   # SASS rows have an empty line number
    try:
        s, ins = int(r[hdr["# Samples"]] or 0), int(r[hdr["Instructions Executed"]] or 0)
    except ValueError:
        continue
    st = {k[6:]: int(r[j]) for k, j in hdr.items() if k.startswith("stall_") and "Not Issued" not in k and r[j].isdigit()}
    out.append((s, ins, fname, r[0], r[1].strip()[:96], st))
ts, ti = sum(o[0] for o in out) or 1, sum(o[1] for o in out) or 1
print(f"total samples {ts}, warp instructions {ti}")
agg = {}
for o in out:
    for k, v in o[5].items():
        agg[k] = agg.get(k, 0) + v
print("stall mix:", ", ".join(f"{k} {100 * v / ts:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]))
for s, ins, f, line, src, st in sorted(out, key=lambda o: -o[0])[:top]:
    main = max(st.items(), key=lambda kv: kv[1])[0] if st else ""
    print(f"{100 * s / ts:5.1f}%s {100 * ins / ti:5.1f}%i {f}:{line:>4} [{main:>12}] {src}")
