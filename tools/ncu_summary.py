#!/usr/bin/env python
"""Text summary of one .ncu-rep capture (raw page): duration, DRAM traffic, issue/occupancy, pipes, stall mix.
Usage: python tools/ncu_summary.py <file.ncu-rep> [title] [kernel-name substring: the LAST matching launch is summarised]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
if len(sys.argv) > 3:
    ki = hdr.index("Kernel Name")
    match = [r for r in rows[2:] if len(r) > ki and sys.argv[3] in r[ki]]
    if not match:
        sys.exit(f"no launch of a kernel matching {sys.argv[3]!r} in {rep}")
    vals = match[-1]
d = dict(zip(hdr, vals))
u = dict(zip(hdr, units))


def g(k, default="n/a"):
    return d.get(k, default)


print(f"# {title}")
print(f"kernel            : {g('Kernel Name')}")
print(f"grid x block      : {g('launch__grid_size')} x {g('launch__block_size')}, {g('launch__registers_per_thread')} regs/thread, "
      f"{g('launch__shared_mem_per_block_dynamic')} {u.get('launch__shared_mem_per_block_dynamic','')} dyn smem/CTA")
print(f"occupancy limits  : regs {g('launch__occupancy_limit_registers')} / smem {g('launch__occupancy_limit_shared_mem')} / warps {g('launch__occupancy_limit_warps')} CTAs per SM")
print(f"duration          : {g('gpu__time_duration.sum')} {u.get('gpu__time_duration.sum','')}  (under ncu: cold caches, serialised)")
rd, wr = float(g('dram__bytes_read.sum', 0)), float(g('dram__bytes_write.sum', 0))
print(f"DRAM traffic      : read {rd:.2f} + write {wr:.2f} = {rd + wr:.2f} {u.get('dram__bytes_read.sum','')}  "
      f"(read {g('dram__bytes_read.sum.per_second')} {u.get('dram__bytes_read.sum.per_second','')})")
print(f"warp instructions : {float(g('smsp__inst_executed.sum', 0)) / 1e6:.1f} M")
print(f"issue active      : {float(g('smsp__issue_active.avg.pct_of_peak_sustained_active', 0)):.1f} % of SMSP cycles")
print(f"warps active      : {float(g('sm__warps_active.avg.pct_of_peak_sustained_active', 0)):.1f} % of peak")
pipes = ["fma", "alu", "fp64", "lsu", "xu", "tensor_subpipe_dmma", "tma"]
print("pipe utilisation  : " + ", ".join(
    f"{p} {float(g(f'sm__inst_executed_pipe_{p}.avg.pct_of_peak_sustained_active', 0)):.1f}%" for p in pipes))
tot = float(g("smsp__pcsamp_sample_count", 0)) or 1.0
stalls = [(k.replace("smsp__pcsamp_warps_issue_stalled_", ""), float(v)) for k, v in d.items()
          if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued")]
print("warp stall samples: " + ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in sorted(stalls, key=lambda kv: -kv[1])[:8]))
