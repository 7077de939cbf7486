#!/usr/bin/env python
"""Static instruction count of the main loop of a kernel from `cuobjdump -sass` output: the widest backward branch
delimits the loop; prints the count and an opcode histogram of the instructions inside it.
usage: cuobjdump -sass file.cubin | python tools/sass_loop.py [function-substring]"""
import re
import sys
from collections import Counter

want = sys.argv[1] if len(sys.argv) > 1 else None
funcs, cur = {}, None
for ln in sys.stdin:
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and cur is not None:
        funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
for name, ins in funcs.items():
    if want and want not in name:
        continue
    best = None
    for addr, text in ins:
        m = re.search(r"\bBRA(?:\.\w+)*\s+(?:!?U?P\d+,\s*)?`?\(?(?:\.L_x_\d+|0x([0-9a-f]+))", text)
        if m and m.group(1):
            tgt = int(m.group(1), 16)
            if tgt < addr and (best is None or addr - tgt > best[1] - best[0]):
                best = (tgt, addr)
    if not best:
        print(name, "no backward branch found", len(ins))
        continue
    body = [t for a, t in ins if best[0] <= a <= best[1]]
    ops = Counter(re.sub(r"^@!?U?P\w+\s+", "", t).split()[0].split(".")[0] for t in body)
    print(f"{name}: {len(ins)} instructions, main loop {len(body)} ({best[0]:#x}..{best[1]:#x})")
    print("  " + "  ".join(f"{k} {v}" for k, v in ops.most_common(28)))
