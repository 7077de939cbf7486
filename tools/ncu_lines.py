#!/usr/bin/env python
"""Per-source-line / per-opcode breakdown of executed instructions for one kernel of an .ncu-rep.

  python tools/ncu_lines.py gpurun_out/prof.ncu-rep k_render_pathILb0ELb0 [top]

Joins `ncu --page source --csv` (one row per SASS instruction, in address order) with `nvdisasm -g` line info of
the shipped cubin (same order).  Needs the .so the profile was taken with (built with -lineinfo).
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "aurora_rendering_engine_b200", "lib", "libare_b200.so")


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", SO], cwd=tmp, check=True, capture_output=True)
    seq = []
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        fn = cur = None
        for ln in txt.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+),", ln)
            if m:
                fn = m.group(1)
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (m.group(1).split("/")[-1], int(m.group(2)))
                continue
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m and fn and kern in fn:
                seq.append((m.group(2).strip(), cur))
        if seq:
            break
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kern.split("IL")[0].replace("_ZN4areb13", "")],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = next(r for r in rows if "Instructions Executed" in r)
    data = rows[rows.index(hdr) + 1:]
    ie, te = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    if len(data) != len(seq):
        print(f"warning: {len(data)} profiled instructions vs {len(seq)} disassembled", file=sys.stderr)
    agg, aggt, byfile, byop = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
    tot = tott = 0
    for (txt, cur), r in zip(seq, data):
        n, t = int(r[ie]), int(r[te])
        tot += n
        tott += t
        agg[cur] += n
        aggt[cur] += t
        byfile[cur[0] if cur else None] += n
        w = txt.split()
        byop[w[1] if w[0].startswith("@") else w[0]] += n
    dump = os.environ.get("NCU_LINES_DUMP")  # full listing in address order: count, threads / instruction, source line, SASS
    if dump:
        with open(dump, "w") as f:
            for (txt, cur), r in zip(seq, data):
                n, t = int(r[ie]), int(r[te])
                f.write(f"{n:12d} {t / max(1, n):5.1f}  {str(cur[0]) + ':' + str(cur[1]) if cur else '-':24s} {txt}\n")
    print(f"warp instructions {tot}  avg active threads {tott / max(1, tot):.2f}")
    for k, v in byfile.most_common():
        print(f"  {str(k):34s} {100 * v / tot:5.1f}%")
    print("-- lines")
    for k, v in agg.most_common(top):
        print(f"  {str(k):34s} {100 * v / tot:5.1f}%  thr/inst {aggt[k] / max(1, v):5.1f}")
    print("-- opcodes")
    for k, v in byop.most_common(24):
        print(f"  {k:24s} {100 * v / tot:5.1f}%")


if __name__ == "__main__":
    main()
