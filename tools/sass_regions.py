"""Static SASS op mix per subroutine (split at RET / EXIT) of one kernel in a cuobjdump -sass dump.

    cuobjdump -sass build/fast_forward.o > /tmp/ff.sass ; python tools/sass_regions.py /tmp/ff.sass 'ILi1ELb0'
"""
import re
import sys
from collections import Counter

txt = open(sys.argv[1]).read()
key = sys.argv[2]
parts = re.split(r'\n\s*Function : ', txt)
for p in parts[1:]:
    name = p.split('\n')[0]
    if key not in name:
        continue
    regs, cur = [], []
    for line in p.split('\n'):
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)', line)
        if m:
            cur.append(m.group(3))
            if m.group(3).startswith(('RET', 'EXIT')):
                regs.append(cur); cur = []
    if cur:
        regs.append(cur)
    for i, r in enumerate(regs):
        if len(r) < 40:
            continue
        ops = Counter(o.split('.')[0] for o in r)
        print(f"region {i}: {len(r)} instr: " + " ".join(f"{o}={n}" for o, n in ops.most_common(16)))
