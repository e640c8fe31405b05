#!/usr/bin/env python
"""profiles/r02_traffic.json from one `ncu --set full` capture of bench.py's device-resident leg.

    python tools/traffic_from_ncu.py gpurun_out/prof.ncu-rep 1024 300 80 "ncu --set full ... (command)"
Records dram__bytes_read.sum + dram__bytes_write.sum of the last sparse_img_align_kernel launch in the report together
with the sha256 of the kernel sources (bench.py refuses the number when the sources have changed since)."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

rep, B, n_pts, n_segs = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
src = sys.argv[5] if len(sys.argv) > 5 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
ki = hdr.index("Kernel Name")
row = [r for r in rows[2:] if "sparse_img_align" in r[ki]][-1]
d = dict(zip(hdr, row))
u = dict(zip(hdr, units))
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
rd = float(d["dram__bytes_read.sum"]) * scale[u["dram__bytes_read.sum"]]
wr = float(d["dram__bytes_write.sum"]) * scale[u["dram__bytes_write.sum"]]
out = {"sparse_img_align_kernel": {
    "dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr, "batch": B, "n_pts": n_pts, "n_segs": n_segs,
    "duration_under_ncu": d["gpu__time_duration.sum"] + " " + u["gpu__time_duration.sum"],
    "kernel_source_sha256": bench.kernel_source_hash(), "source": src}}
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w"), indent=1)
print(json.dumps(out))
