#!/usr/bin/env python
"""Executed-instruction breakdown of one kernel of an .ncu-rep by source line and by opcode, joined with the line info of
a stand-alone CUBIN (the scene-specialised kernel is compiled by NVRTC, so its code is not inside libare_b200.so:
`capi.bake_probe(..., cubin_path=...)` reproduces the CUBIN from the same generated source).

  python tools/ncu_cubin_lines.py gpurun_out/prof.ncu-rep /tmp/baked.cubin k_render_baked [top]
"""
import collections
import csv
import re
import subprocess
import sys


def main():
    rep, cubin, kern = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 45
    txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    seq, fn, cur = [], None, None
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
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kern], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = next(r for r in rows if "Instructions Executed" in r)
    data = rows[rows.index(hdr) + 1:]
    ci, ct, cs = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("Source")
    if len(data) != len(seq):
        print(f"warning: {len(data)} profiled instructions vs {len(seq)} in the cubin — line attribution may be off", file=sys.stderr)
    by_line, by_op, by_file = collections.Counter(), collections.Counter(), collections.Counter()
    thr_line = collections.Counter()
    smp_line, smp_reason = collections.Counter(), collections.Counter()
    line_reason = collections.defaultdict(collections.Counter)
    c_smp = hdr.index("# Samples") if "# Samples" in hdr else -1
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = tot_thr = tot_smp = 0
    for i, r in enumerate(data):
        n, t = int(r[ci]), int(r[ct])
        tot += n
        tot_thr += t
        op = re.sub(r"^@!?U?P\w+\s+", "", r[cs].strip()).split()[0].split(".")[0]
        by_op[op] += n
        line = seq[i][1] if i < len(seq) else None
        by_line[line] += n
        thr_line[line] += t
        by_file[line[0] if line else None] += n
        if c_smp >= 0:
            k = int(r[c_smp] or 0)
            smp_line[line] += k
            tot_smp += k
            for ci_, h in stall_cols:
                v = int(r[ci_] or 0)
                if v:
                    smp_reason[h] += v
                    line_reason[line][h] += v
    print(f"warp instructions {tot}  avg active threads {tot_thr / max(1, tot):.2f}")
    for f, n in by_file.most_common():
        print(f"  {str(f):34s} {100.0 * n / tot:5.1f}%")
    print("-- opcodes")
    print("  " + "  ".join(f"{k} {100.0 * v / tot:.1f}%" for k, v in by_op.most_common(30)))
    print("-- lines")
    for line, n in by_line.most_common(top):
        print(f"  {str(line):36s} {100.0 * n / tot:5.1f}%  thr/inst {thr_line[line] / max(1, n):5.1f}")


    if tot_smp:
        print(f"-- warp-state samples {tot_smp}: " + "  ".join(f"{k[6:]} {100.0 * v / tot_smp:.1f}%" for k, v in smp_reason.most_common(10)))
        print("-- lines by samples (where warps wait)")
        for line, k in smp_line.most_common(top):
            rs = "  ".join(f"{h[6:]} {100.0 * v / k:.0f}%" for h, v in line_reason[line].most_common(3))
            print(f"  {str(line):36s} {100.0 * k / tot_smp:5.1f}% of samples  {100.0 * by_line[line] / tot:5.1f}% of instr   {rs}")


if __name__ == "__main__":
    main()
