"""Depth-filter seed kernels: kernel time (CUDA events, plsvo_last_kernel_ms) at per-frame and at throughput sizes, with exactness
against the oracle.  PLSVO_LIB selects an A/B build (-DPLSVO_SERIAL_EPI_SEARCH: every seed's epipolar search walked by its own
thread); PLSVO_SEEDS_PER_WARP overrides the seeds-per-warp choice."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np
import plsvo_b200
from plsvo_b200 import abi, api, synth
import oracle_lib

ctx = api.default_context()
df = api.DepthFilter(ctx)
for n in (500, 4000, 25000, 200000):
    row = {"lib": os.path.basename(os.environ.get("PLSVO_LIB", "default")), "spw": os.environ.get("PLSVO_SEEDS_PER_WARP", "auto"), "n": n}
    for kind in ("point", "line"):
        d = (synth.make_seed_batch if kind == "point" else synth.make_line_seed_batch)(n=n if kind == "point" else n // 2, n_ref=8, n_cur=8, n_pyr_levels=3, seed=5 + n, device="cuda")
        run = (df.updatePointSeeds if kind == "point" else df.updateLineSeeds)
        for _ in range(3):
            out = run(d)
        ts = []
        for _ in range(7):
            out = run(d)
            ts.append(ctx.last_kernel_ms())
        ms = float(np.median(ts))
        row[kind + "_kernel_ms"] = round(ms, 4)
        row[kind + "_seeds_per_s"] = round(d.n / (ms * 1e-3))
        if n <= 25000:
            ref = (oracle_lib.seed_update if kind == "point" else oracle_lib.line_seed_update)(abi, d, 16)
            up = ref.status == abi.SEED_UPDATED
            row[kind + "_exact"] = bool(np.array_equal(out.status, ref.status) and np.array_equal(out.depth[up], ref.depth[up]))
    print(json.dumps(row), flush=True)
