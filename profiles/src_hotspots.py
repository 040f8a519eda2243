"""Aggregate `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv` per CUDA source line:
share of executed warp instructions, of stall samples, and the dominant stall reasons."""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hdr = None
cur_file = ""
agg = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        i_s, i_i = hdr.index("# Samples"), hdr.index("Instructions Executed")
        stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        i_conf = hdr.index("L1 Wavefronts Shared Excessive")
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    key = (cur_file, line)
    a = agg.setdefault(key, dict(src=r[1], s=0, i=0, st={}, conf=0))
    try:
        a["s"] += int(r[i_s] or 0)
        a["i"] += int(r[i_i] or 0)
        a["conf"] += int(r[i_conf] or 0)
        for i, h in stall_cols:
            v = int(r[i] or 0)
            if v:
                a["st"][h] = a["st"].get(h, 0) + v
    except ValueError:
        pass
ts = sum(a["s"] for a in agg.values()) or 1
ti = sum(a["i"] for a in agg.values()) or 1
print(f"total samples {ts}, warp instructions {ti}")
for (f, line), a in sorted(agg.items(), key=lambda kv: -kv[1]["s"])[:top]:
    st = sorted(a["st"].items(), key=lambda kv: -kv[1])[:3]
    sts = " ".join(f"{k[6:]}={100 * v / max(a['s'], 1):.0f}%" for k, v in st)
    print(f"{f}:{line:<4} smp {100 * a['s'] / ts:5.1f}%  ins {100 * a['i'] / ti:5.1f}%  bankx {a['conf']:>9}  {sts:40s} | {a['src'].strip()[:90]}")
