#!/usr/bin/env python
"""Aggregate ncu warp-stall samples per CUDA source line.

ncu's CSV source page is SASS-level; this joins it (by instruction order) with `nvdisasm -g`
line markers of the same cubin.  Usage:
  ncu -i prof.ncu-rep --page source --csv > src.csv
  cuobjdump -xelf all libplsvo_b200.so ; nvdisasm -g -c align_kernel.sm_100a.cubin > align.sass
  python tools/ncu_lines.py src.csv align.sass <kernel-name-substring> [top]
"""
import csv
import re
import sys
from collections import defaultdict


def main():
    src_csv, sass, key = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(src_csv)))
    hdr = rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    body = [r for r in rows[2:] if len(r) == len(hdr)]
    # nvdisasm: pick the .text section whose name contains key
    lines = open(sass).read().split("\n")
    insts = []  # (file, line) per instruction, in order
    in_sec = False
    cur = ("?", 0)
    for ln in lines:
        if ln.startswith(".text."):
            in_sec = key in ln
            continue
        if ln.startswith(".section") or ln.startswith(".text"):
            in_sec = False
        if not in_sec:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.search(r"/\*[0-9a-f]{4,}\*/", ln) and ";" in ln:
            insts.append((cur, ln.strip()))
    print(f"# {len(body)} profiled SASS rows, {len(insts)} disassembled instructions", file=sys.stderr)
    n = min(len(body), len(insts))
    agg = defaultdict(lambda: defaultdict(float))
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for i in range(n):
        r = body[i]
        loc = insts[i][0]
        agg[loc]["samples"] += float(r[col["# Samples"]] or 0)
        agg[loc]["inst"] += float(r[col["Instructions Executed"]] or 0)
        for h in stall_cols:
            agg[loc][h] += float(r[col[h]] or 0)
    tot = sum(v["samples"] for v in agg.values())
    toti = sum(v["inst"] for v in agg.values())
    print(f"total samples {tot:.0f}, warp-instructions {toti:.0f}")
    for loc, v in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        stalls = sorted(((h[6:], v[h]) for h in stall_cols if v[h] > 0), key=lambda x: -x[1])[:3]
        st = " ".join(f"{k}={x:.0f}" for k, x in stalls)
        print(f"{loc[0]}:{loc[1]:<5d} samples {v['samples']:8.0f} ({100*v['samples']/tot:5.1f}%)  inst {v['inst']:12.0f} ({100*v['inst']/toti:5.1f}%)  {st}")


if __name__ == "__main__":
    main()
