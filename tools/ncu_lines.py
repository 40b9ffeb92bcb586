"""Per-source-line stall samples / instruction counts from `ncu --page source --csv --print-source cuda,sass`.

    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv ; python tools/ncu_lines.py src.csv [top]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
files = {}
cur = None
hdr = None
out = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        i_s = hdr.index("# Samples"); i_i = hdr.index("Instructions Executed")
        stall_cols = [(j, h) for j, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) < len(hdr) or r[0] == "":
        continue
    def num(x):
        try:
            return int(x)
        except ValueError:
            return 0
    st = sorted(((num(r[j]), h[6:]) for j, h in stall_cols), reverse=True)[:3]
    out.append((num(r[i_s]), num(r[i_i]), cur, num(r[0]), r[1].strip()[:90], st))
tot_s = sum(o[0] for o in out); tot_i = sum(o[1] for o in out)
print(f"total samples {tot_s}  warp-instructions {tot_i}")
for o in sorted(out, reverse=True)[:top]:
    print(f"{o[0]:7d} {100 * o[0] / tot_s:5.1f}%  inst {o[1]:9d} {100 * o[1] / tot_i:5.1f}%  {o[2]}:{o[3]:<5d} {o[4]}   {' '.join(f'{n}={v}' for v, n in o[5] if v)}")
