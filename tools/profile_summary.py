#!/usr/bin/env python
"""Summarise one kernel of an .ncu-rep into profiles/<name>.json + .md (the numbers bench.py and DESIGN.md cite).

  python tools/profile_summary.py gpurun_out/prof.ncu-rep profiles/r01_render_path_full --width 2048 --height 2048 --spp 64
"""
import argparse
import csv
import json
import subprocess

KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "lts__t_bytes.sum": "l2_bytes",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__occupancy_limit_registers": "occupancy_limit_registers_blocks",
    "launch__grid_size": "grid_size",
    "launch__block_size": "block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slot_utilisation_pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "active_threads_per_instruction",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "pipe_xu_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_pct",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active": "pipe_tensor_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__cycles_elapsed.max": "sm_cycles",
    "smsp__warps_eligible.avg.per_cycle_active": "eligible_warps_per_cycle",
}
STALLS = "smsp__average_warps_issue_stalled_"


def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(v) * m.get(unit, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("out")
    ap.add_argument("--width", type=int)
    ap.add_argument("--height", type=int)
    ap.add_argument("--spp", type=int)
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {"kernel": vals[hdr.index("Kernel Name")], "width": a.width, "height": a.height, "spp_per_step": a.spp, "note": a.note, "source": a.rep}
    stalls = {}
    for h, u, v in zip(hdr, units, vals):
        if h in KEYS:
            k = KEYS[h]
            if k in ("dram_read", "dram_write", "l2_bytes"):
                d[k + "_bytes"] = to_bytes(v, u)
            elif k == "duration":
                d["duration_ms"] = float(v) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[u]
            else:
                try:
                    d[k] = float(v)
                except ValueError:
                    d[k] = v
        elif h.startswith(STALLS) and h.endswith("_per_issue_active.ratio"):
            stalls[h[len(STALLS):-len("_per_issue_active.ratio")]] = float(v)
    d["dram_bytes"] = d.get("dram_read_bytes", 0) + d.get("dram_write_bytes", 0)
    d["stall_warps_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1]))
    json.dump(d, open(a.out + ".json", "w"), indent=1)
    with open(a.out + ".md", "w") as f:
        f.write(f"# ncu --set full summary: `{d['kernel']}`\n\n{a.note}\n\n")
        f.write(f"source capture: `{a.rep}` (ncu --set full --clock-control none --import-source on; per-launch values, replayed, cold cache)\n\n")
        f.write("| metric | value |\n|---|---|\n")
        for k, v in d.items():
            if k in ("kernel", "note", "source", "stall_warps_per_issue"):
                continue
            f.write(f"| {k} | {v:,.4g} |\n" if isinstance(v, float) else f"| {k} | {v} |\n")
        f.write("\n## warps stalled per issued instruction (by reason)\n\n| reason | warps |\n|---|---|\n")
        for k, v in d["stall_warps_per_issue"].items():
            f.write(f"| {k} | {v:.3f} |\n")
    print("wrote", a.out + ".json/.md")


if __name__ == "__main__":
    main()
