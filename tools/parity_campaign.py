#!/usr/bin/env python
"""GPU-vs-reference parity over many seeded C2-shaped pairs (run on the GPU box).

For every seed: generate a batch on the GPU, run plsvo::SparseImgAlign on the GPU and on the CPU checker
(oracle/_ref = the reference's own translation units when the library travelled with the repo, else the oracle
restatement, which is bit-identical to it), and compare: pose tolerance (1e-5 rad / 1e-4 relative translation),
per-level iteration counts, n_tracked, killed segments, status.  Every pair that differs is listed with the
level at which its iteration count first differs.

    python tools/parity_campaign.py [n_seeds] [batch] [first_seed] > profiles/r02_parity_8192_pairs.txt
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np  # noqa: E402

import plsvo_b200  # noqa: E402
from plsvo_b200 import abi, synth  # noqa: E402
import oracle_lib  # noqa: E402


def main():
    n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    seed0 = int(sys.argv[3]) if len(sys.argv) > 3 else 4100
    n_pts = int(os.environ.get("PARITY_PTS", 300))
    n_segs = int(os.environ.get("PARITY_SEGS", 80))
    threads = int(os.environ.get("PARITY_THREADS", min(64, os.cpu_count() or 1)))
    use_ref = oracle_lib.ref_available() and not os.environ.get("PARITY_USE_ORACLE")
    print(f"checker: {'oracle/_ref (reference translation units)' if use_ref else 'oracle restatement'}, {threads} host threads; "
          f"variant {os.environ.get('PLSVO_VARIANT', 'default')}; lib {os.environ.get('PLSVO_LIB', 'default')}")
    print(f"workload: VGA, levels 4->2, {n_pts} points + {n_segs} segments, <=30 GN iterations per level, B={B} per seed")
    tot = bad = itd = intd = 0
    max_rot = max_rel = 0.0
    flagged = 0
    for seed in range(seed0, seed0 + n_seeds):
        t0 = time.time()
        d = synth.make_align_batch(batch=B, n_pts=n_pts, n_segs=n_segs, device="cuda", seed=seed)
        gpu = plsvo_b200.SparseImgAlign(4, 2, 30).run(d)
        ref = (oracle_lib.ref_align if use_ref else oracle_lib.align)(abi, d, n_threads=threads)
        ang, rel = synth.pose_error(gpu.T_cur_w, ref.T_cur_w)
        ang, rel = np.asarray(ang), np.asarray(rel)
        ok = (ang <= 1e-5) & (rel <= 1e-4)
        it_same = (gpu.iters == ref.iters).all(axis=1)
        int_same = (gpu.n_tracked == ref.n_tracked) & (gpu.seg_killed == ref.seg_killed).all(axis=1) & ((gpu.status & 3) == (ref.status & 3))
        flagged += int((gpu.status >> 2).astype(bool).sum())
        tot += B
        bad += int((~ok).sum())
        itd += int((~it_same).sum())
        intd += int((~int_same).sum())
        max_rot, max_rel = max(max_rot, float(ang.max())), max(max_rel, float(rel.max()))
        print(f"seed {seed}: {B} pairs, out of tolerance {int((~ok).sum())}, iteration counts differ {int((~it_same).sum())}, "
              f"integer outputs differ {int((~int_same).sum())}, max rot {ang.max():.3e} rad, max rel t {rel.max():.3e}, "
              f"median rot {np.median(ang):.2e} ({time.time() - t0:.1f}s)", flush=True)
        for b in np.nonzero(~it_same | ~ok | ~int_same)[0]:
            lv = [l for l in range(gpu.iters.shape[1]) if gpu.iters[b, l] != ref.iters[b, l]]
            print(f"    pair {b}: gpu iters {gpu.iters[b].tolist()} ref iters {ref.iters[b].tolist()} first differing level "
                  f"{max(lv) if lv else '-'} rot {ang[b]:.2e} relt {rel[b]:.2e} n_tracked {int(gpu.n_tracked[b])}/{int(ref.n_tracked[b])} "
                  f"status {int(gpu.status[b])}/{int(ref.status[b])}")
    print(f"TOTAL {tot} pairs over {n_seeds} seeds: {tot - bad} inside 1e-5 rad / 1e-4 rel-t ({100.0 * (tot - bad) / tot:.3f} %), "
          f"{itd} with different per-level iteration counts, {intd} with different integer outputs, "
          f"{flagged} with a chi2-order flag; max rot {max_rot:.3e} rad, max rel t {max_rel:.3e}")


if __name__ == "__main__":
    main()
