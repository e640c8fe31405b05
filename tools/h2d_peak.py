"""Pinned host->device copy rate of this box for the copy sizes the end-to-end leg issues (one array of one 256-pair chunk
is 0.5 - 5 MB; a whole lean batch of 1024 pairs is 51 MB): the ceiling of bench.py's `e2e` is bytes / this rate."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import plsvo_b200  # noqa: E402
from plsvo_b200 import numa  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.init()
info = {"before": numa.describe(0)}
if "--bind" in sys.argv:
    numa.bind_to_device(0)
    info["after_bind"] = numa.describe(0)
try:
    info["nodes"] = {n: open(f"/sys/devices/system/node/{n}/cpulist").read().strip() for n in sorted(os.listdir("/sys/devices/system/node")) if n.startswith("node")}
except Exception as ex:
    info["nodes"] = repr(ex)
res = {}
for mb in (0.25, 1, 4.9, 16, 51.4, 256):
    n = int(mb * 1e6)
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    reps = 20
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    e.record()
    torch.cuda.synchronize()
    up = n * reps / (s.elapsed_time(e) * 1e-3) / 1e9
    s.record()
    for _ in range(reps):
        h.copy_(d, non_blocking=True)
    e.record()
    torch.cuda.synchronize()
    down = n * reps / (s.elapsed_time(e) * 1e-3) / 1e9
    res[f"{mb} MB"] = {"h2d_gbs": round(up, 2), "d2h_gbs": round(down, 2)}
# the same 51.4 MB copy once the buffer has left the CPU caches (a generator has just written the bench's inputs: they sit
# modified in the last-level cache; frames delivered by a capture DMA would not)
import numpy as np

n = int(51.4e6)
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h.fill_(7)
d = torch.empty(n, dtype=torch.uint8, device=dev)
evict = np.ones(1 << 29, np.uint8)
rates = []
for trial in range(3):
    evict += 1  # read-modify-write of 512 MB: pushes everything else out of the caches
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    per = []
    for rep in range(4):
        s.record()
        d.copy_(h, non_blocking=True)
        e.record()
        torch.cuda.synchronize()
        per.append(round(n / (s.elapsed_time(e) * 1e-3) / 1e9, 2))
    rates.append(per)
res["51.4 MB after cache eviction, 4 consecutive copies x 3 trials (h2d_gbs)"] = rates
h.fill_(9)  # written again by the CPU: back in the cache
per = []
for rep in range(4):
    s.record()
    d.copy_(h, non_blocking=True)
    e.record()
    torch.cuda.synchronize()
    per.append(round(n / (s.elapsed_time(e) * 1e-3) / 1e9, 2))
res["51.4 MB right after a CPU write (h2d_gbs)"] = per
print(json.dumps({"device": torch.cuda.get_device_name(0), "placement": info, "pinned_copy_rate": res}))
