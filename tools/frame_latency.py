"""Per-frame latency of the loops either side of the hot path at the sizes ONE frame produces, on the reference's own
objects: the reference's loop (one thread, as the reference runs it) vs the B200 binding (pl-svo_b200/host/plsvo_shim_next).

  Reprojector match loop : N map points + N/2 map segments seen from 3 keyframes, one current frame
                           reference: Matcher::findMatchDirect per candidate      binding: b200::DirectMatcher (one device call)
  DepthFilter::updateSeeds: S point seeds + S/2 line seeds of 3 keyframes, one current frame
                           reference: DepthFilter::updateSeeds                     binding: b200::DepthFilterB200 (two device calls)

    python tools/frame_latency.py [reps] > profiles/r02_frame_latency.txt      (needs oracle/_ref: built where /root/reference exists)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np
import plsvo_b200
from plsvo_b200 import abi, synth
import oracle_lib

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 15


def pct(x):
    x = np.asarray(x) * 1e3
    return f"p50 {np.percentile(x, 50):7.3f} ms   p90 {np.percentile(x, 90):7.3f} ms"


print("per-frame latency on reference-typed objects, %d repetitions, 1 x B200 vs one host thread" % reps)
for n in (300, 1200):
    d = synth.make_match_batch(n=n, n_ref=3, n_cur=1, n_pyr_levels=3, seed=8100 + n, device="cuda")
    tr, ts = [], []
    for _ in range(3):
        oracle_lib.shimref_match_scene(abi, d, 3)
    for _ in range(reps):
        r = oracle_lib.ref_match_scene(abi, d, 3)
        tr.append(oracle_lib.last_loop_seconds(abi, shim=False))
        s = oracle_lib.shimref_match_scene(abi, d, 3)
        ts.append(oracle_lib.last_loop_seconds(abi, shim=True))
    same = bool(np.array_equal(s.pt_found, r.pt_found) and np.array_equal(s.seg_found, r.seg_found) and
                np.array_equal(s.pt_px[r.pt_found > 0], r.pt_px[r.pt_found > 0]))
    print(f"reprojector match loop, {n} point + {n // 2} segment candidates: reference {pct(tr)} | DirectMatcher {pct(ts)} | "
          f"speed-up (p50) {np.median(tr) / np.median(ts):5.1f}x | results identical: {same}")
for n in (500, 2000):
    pts = synth.make_seed_batch(n=n, n_ref=3, n_cur=1, n_pyr_levels=3, seed=8200 + n, device="cuda")
    lines = synth.make_line_seed_batch(n=n // 2, n_ref=3, n_cur=1, n_pyr_levels=3, seed=8200 + n, device="cuda")
    tr, ts = [], []
    for _ in range(3):
        oracle_lib.shimref_seed_scene(abi, pts, lines, None, None, True)
    for _ in range(reps):
        r = oracle_lib.ref_seed_scene(abi, pts, lines, None, None, True)
        tr.append(oracle_lib.last_loop_seconds(abi, shim=False))
        s = oracle_lib.shimref_seed_scene(abi, pts, lines, None, None, True)
        ts.append(oracle_lib.last_loop_seconds(abi, shim=True))
    agree = float((s.pt_fate == r.pt_fate).mean()), float((s.seg_fate == r.seg_fate).mean())
    print(f"DepthFilter::updateSeeds, {n} point + {n // 2} line seeds: reference {pct(tr)} | DepthFilterB200 {pct(ts)} | "
          f"speed-up (p50) {np.median(tr) / np.median(ts):5.1f}x | seed fates equal: {agree[0]:.4f} / {agree[1]:.4f}")
