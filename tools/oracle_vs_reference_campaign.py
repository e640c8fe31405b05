#!/usr/bin/env python
"""Large seeded campaign: the oracle restatement (oracle/plsvo_oracle.cpp) against the reference's own translation units
compiled in place (oracle/_ref), bit for bit, on every row.  CPU only.

    python tools/oracle_vs_reference_campaign.py > profiles/r01_oracle_vs_reference.txt
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np  # noqa: E402

import plsvo_b200  # noqa: E402,F401
from plsvo_b200 import abi, synth  # noqa: E402
import oracle_lib  # noqa: E402


def same(a, b):
    return bool(np.array_equal(a, b) or (np.isnan(a) == np.isnan(b)).all() and np.array_equal(np.nan_to_num(a), np.nan_to_num(b)))


def main():
    if not oracle_lib.build_ref():
        raise SystemExit("needs oracle/_ref (built where /root/reference exists)")
    oracle_lib.build()
    th = os.cpu_count() or 1
    print("reference build :", oracle_lib.load_ref(abi).plsvo_ref_describe().decode())
    t0 = time.time()
    # ---- SparseImgAlign::run ----
    n = bad = 0
    for seed in range(20000, 20016):
        cam = (synth.VGA, synth.QVGA, synth.HD720)[seed % 3]
        d = synth.make_align_batch(cam=cam, batch=64 if cam is not synth.HD720 else 16, n_pts=300, n_segs=80, seed=seed,
                                   motion_t=(0.03, 0.08)[seed % 2], motion_r=(0.01, 0.03)[seed % 2])
        o, r = oracle_lib.align(abi, d, n_threads=th), oracle_lib.ref_align(abi, d, n_threads=th)
        ok = np.array([all(same(getattr(o, f)[b], getattr(r, f)[b]) for f in ("T_cur_w", "n_tracked", "H", "seg_killed", "iters", "status"))
                       for b in range(d.batch)])
        n += d.batch
        bad += int((~ok).sum())
    print(f"SparseImgAlign::run            : {n} frame pairs (VGA/QVGA/720p, 300 points + 80 segments), not bit-identical: {bad}")
    # ---- pose_optimizer ----
    n = bad = 0
    for seed in range(21000, 21008):
        d = synth.make_poseopt_batch(batch=256, seed=seed, outlier_frac=(0.1, 0.3)[seed % 2])
        for n_ref in (-1, 3):
            p = abi.poseopt_params(2.0, 10, n_ref)
            o, r = oracle_lib.poseopt(abi, d, p, n_threads=th), oracle_lib.ref_poseopt(abi, d, p, n_threads=th)
            ok = np.array([all(same(getattr(o, f)[b], getattr(r, f)[b]) for f in ("T_f_w", "cov", "estimated_scale", "error_init",
                                                                                  "error_final", "num_obs_pt", "num_obs_ls",
                                                                                  "pt_outlier", "seg_outlier")) for b in range(d.batch)])
            n += d.batch
            bad += int((~ok).sum())
    print(f"pose_optimizer (9- and 10-arg) : {n} frames (300 points + 80 lines), not bit-identical: {bad}")
    # ---- findMatchDirect ----
    n = bad = 0
    for seed in range(22000, 22006):
        d = synth.make_match_batch(n=5000, seed=seed, n_pyr_levels=(3, 5)[seed % 2])
        o, r = oracle_lib.match_direct(abi, d, th), oracle_lib.ref_match_direct(abi, d)
        ok = (o.success == r.success) & (o.search_level == r.search_level) & (o.px_cur == r.px_cur).all(axis=1)
        n += d.n
        bad += int((~ok).sum())
    print(f"Matcher::findMatchDirect       : {n} candidates (25 % edgelets), not bit-identical: {bad}")
    # ---- structure optimisation ----
    n = bad = 0
    for seed in range(23000, 23004):
        d = synth.make_structopt_batch(n_points=8000, n_segs=2000, seed=seed)
        o, r = oracle_lib.structopt(abi, d, th), oracle_lib.ref_structopt(abi, d)
        n += 10000
        bad += int((~(o.pt_pos == r.pt_pos).all(axis=1)).sum() + (~((o.seg_spos == r.seg_spos).all(axis=1) & (o.seg_epos == r.seg_epos).all(axis=1))).sum())
    print(f"Point/LineSeg::optimize        : {n} 3D features, not bit-identical: {bad}")
    # ---- depth filter ----
    n = bad = erased = 0
    for seed in range(24000, 24006):
        d = synth.make_seed_batch(n=4000, seed=seed, n_pyr_levels=(3, 5)[seed % 2])
        o, r = oracle_lib.seed_update(abi, d, th), oracle_lib.ref_seed_update(abi, d)
        alive = r.status == 0
        ok = np.ones(d.n, bool)
        for f in ("a", "b", "mu", "sigma2"):
            x, y = getattr(o, f), getattr(r, f)
            ok &= (x == y) | (np.isnan(x) & np.isnan(y))
        n += int(alive.sum())
        erased += int((~alive).sum())
        bad += int((~ok & alive).sum())
    print(f"DepthFilter::updatePointSeeds  : {n} seeds compared ({erased} erased by the reference as converged), not bit-identical: {bad}")
    n = bad = erased = 0
    for seed in range(25000, 25004):
        d = synth.make_line_seed_batch(n=3000, seed=seed)
        o, r = oracle_lib.line_seed_update(abi, d, th), oracle_lib.ref_line_seed_update(abi, d)
        alive = r.status == 0
        ok = np.ones(d.n, bool)
        for f in ("a", "b", "mu", "sigma2", "mu_e", "sigma2_e"):
            x, y = getattr(o, f), getattr(r, f)
            ok &= (x == y) | (np.isnan(x) & np.isnan(y))
        n += int(alive.sum())
        erased += int((~alive).sum())
        bad += int((~ok & alive).sum())
    print(f"DepthFilter::updateLineSeeds   : {n} line seeds compared ({erased} erased), not bit-identical: {bad}")
    print(f"wall time {time.time() - t0:.0f} s on {th} threads")


if __name__ == "__main__":
    main()
