#!/usr/bin/env python
"""CPU experiment behind DESIGN.md §2 "exact chi2": if the kernel reproduces the reference's sequential float
chi2 exactly, how often does the accept/rollback decision still flip because the normal equations are formed
differently (five in-patch sums + rank-2 update; fp32 FMA sums or double sums)?

Runs the oracle three times on the same seeded C2-shaped pairs: flags=0 (the reference's arithmetic),
flags=2 (h_mode 1: fp32 in-patch sums) and flags=4 (h_mode 2: double in-patch sums); chi2 is summed in the
reference's order in all three.

    python tools/emulate_kernel_sums.py [n_batches] [batch]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

import plsvo_b200  # noqa: E402,F401
from plsvo_b200 import abi, synth  # noqa: E402
import oracle_lib  # noqa: E402


def main():
    n_batches = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    threads = os.cpu_count() or 1
    tot = 0
    stats = {1: [0, 0, 0.0, 0.0], 2: [0, 0, 0.0, 0.0]}
    for seed in range(6000, 6000 + n_batches):
        t0 = time.time()
        d = synth.make_align_batch(cam=synth.VGA, batch=B, n_pts=300, n_segs=80, seed=seed)
        t1 = time.time()
        ref = oracle_lib.align(abi, d, n_threads=threads, flags=0)
        tot += B
        line = f"seed {seed} (gen {t1 - t0:.1f}s):"
        for mode in (1, 2):
            o = oracle_lib.align(abi, d, n_threads=threads, flags=mode << 1)
            ang, rel = synth.pose_error(torch.tensor(o.T_cur_w), torch.tensor(ref.T_cur_w))
            ang, rel = np.asarray(ang), np.asarray(rel)
            bad = int(((ang > 1e-5) | (rel > 1e-4)).sum())
            nd = int((o.iters != ref.iters).any(axis=1).sum())
            st = stats[mode]
            st[0] += bad
            st[1] += nd
            st[2] = max(st[2], float(ang.max()))
            st[3] = max(st[3], float(rel.max()))
            line += f"  h_mode {mode}: out-of-tol {bad}, iters differ {nd}, max rot {ang.max():.2e}, max relt {rel.max():.2e};"
        print(line, flush=True)
    for mode, name in ((1, "fp32 FMA in-patch sums"), (2, "double in-patch sums")):
        st = stats[mode]
        print(f"TOTAL {tot} pairs, {name}: out of tolerance {st[0]}, iteration counts differ {st[1]}, max rot {st[2]:.3e}, max rel t {st[3]:.3e}")


if __name__ == "__main__":
    main()
