"""Instruction counts / stall samples per SASS region (split at RET/EXIT), labelled by the source lines inside.

    python tools/ncu_funcs.py src.csv     (src.csv from: ncu -i X --page source --csv --print-source cuda,sass)
"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hdr = None
cur_file, cur_line = None, 0
sass = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No":
        hdr = r; i_s = hdr.index("# Samples"); i_i = hdr.index("Instructions Executed")
        stall_cols = [(j, h[6:]) for j, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[0] != "":
        cur_line = int(r[0]); continue
    def num(x):
        try: return int(x)
        except ValueError: return 0
    if not r[2].startswith("0x"):
        continue
    sass.append((int(r[2], 16), r[3].strip(), num(r[i_s]), num(r[i_i]), cur_file, cur_line, {h: num(r[j]) for j, h in stall_cols if num(r[j])}))
sass.sort()
dedup = {}
for s_ in sass:            # the cuda,sass view repeats a SASS row under every inlined source line it belongs to
    dedup.setdefault(s_[0], s_)
sass = [dedup[a] for a in sorted(dedup)]
regions, cur = [], []
for s in sass:
    cur.append(s)
    op = s[1].split()[0] if not s[1].startswith("@") else s[1].split()[1]
    if op.startswith("RET") or op.startswith("EXIT"):
        regions.append(cur); cur = []
if cur: regions.append(cur)
tot_i = sum(s[3] for s in sass); tot_s = sum(s[2] for s in sass)
print(f"total warp-instructions {tot_i}, samples {tot_s}, sass rows {len(sass)}")
for reg in regions:
    ni = sum(s[3] for s in reg); ns = sum(s[2] for s in reg)
    if ni < tot_i * 0.002 and ns < tot_s * 0.002: continue
    lines = Counter()
    for s in reg:
        if s[4] == "fast_forward.cu": lines[s[5] // 10 * 10] += s[3]
    ops = Counter()
    for s in reg:
        op = s[1].split()[0] if not s[1].startswith("@") else s[1].split()[1]
        ops[op.split(".")[0]] += s[3]
    print(f"region 0x{reg[0][0] & 0xfffff:05x}+{len(reg):5d} sass: inst {ni:10d} {100 * ni / tot_i:5.1f}%  samples {ns:6d} {100 * ns / tot_s:5.1f}%  lines~{[l for l, _ in lines.most_common(4)]}")
    st = Counter()
    for s_ in reg:
        st.update(s_[6])
    tot_st = max(sum(st.values()), 1)
    print("      stalls:  " + " ".join(f"{k}={100 * v / tot_st:.0f}%" for k, v in st.most_common(7)))
    print("      top ops: " + " ".join(f"{o}={100 * n / max(ni, 1):.0f}%" for o, n in ops.most_common(12)))
