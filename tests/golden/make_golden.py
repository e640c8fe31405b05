"""Generates tests/golden/*.npz: small frozen inputs + the outputs of THE REFERENCE ITSELF run here
(oracle/_ref: the reference's src/sparse_img_align.cpp, pose_optimizer.cpp and feature.cpp compiled
unmodified against stand-in third-party headers — oracle/ref_harness.cpp), plus the oracle's
per-iteration traces, which the reference class does not expose.  The reference ships no golden
vectors of its own.  /root/reference cannot travel to the GPU box, so the vectors are
committed; they pin the oracle restatement (bit-exact) and give the GPU tests a reference-derived
target.  Needs /root/reference; run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np  # noqa: E402

import plsvo_b200  # noqa: E402,F401
from plsvo_b200 import abi, synth  # noqa: E402
import oracle_lib  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ALIGN_FIELDS = ["T_ref_w", "T_cur_w", "T_cur_w_gt", "pt_px", "pt_f", "pt_pos", "seg_spx", "seg_epx", "seg_sf", "seg_ef",
                "seg_spos", "seg_epos", "seg_length"]
PO_FIELDS = ["T_f_w", "T_f_w_gt", "pt_f", "pt_pos", "pt_level", "seg_line", "seg_spos", "seg_epos", "seg_level"]


def align_case():
    d = synth.make_align_batch(cam=synth.QVGA, batch=3, n_pts=64, n_segs=16, max_level=3, min_level=1, seed=9001,
                               margin=32, motion_t=0.02, motion_r=0.006)
    params = abi.align_params(3, 1, 30)
    out = oracle_lib.ref_align(abi, d, params)   # the reference's own translation units
    port = oracle_lib.align(abi, d, params)      # oracle restatement: only for the roofline counters below
    traces = [oracle_lib.align_trace(abi, d, b, params) for b in range(d.batch)]
    save = {f: getattr(d, f) for f in ALIGN_FIELDS}
    for l, v in d.ref_pyr.items():
        save[f"ref_pyr_{l}"] = v
    for l, v in d.cur_pyr.items():
        save[f"cur_pyr_{l}"] = v
    save.update(out_T_cur_w=out.T_cur_w, out_n_tracked=out.n_tracked, out_H=out.H, out_seg_killed=out.seg_killed,
                out_iters=out.iters, out_status=out.status, out_patch_iters=port.patch_iters, out_patch_levels=port.patch_levels,
                source=np.array(oracle_lib.load_ref(abi).plsvo_ref_describe().decode()))
    for b, t in enumerate(traces):
        save[f"trace_{b}"] = t
    np.savez_compressed(os.path.join(HERE, "align_qvga.npz"), **save)


def poseopt_case():
    d = synth.make_poseopt_batch(batch=6, n_pts=96, n_segs=24, seed=9002)
    save = {f: getattr(d, f) for f in PO_FIELDS}
    for tag, n_ref in (("9arg", -1), ("10arg", 3)):
        out = oracle_lib.ref_poseopt(abi, d, abi.poseopt_params(2.0, 10, n_ref))  # the reference's own pose_optimizer.cpp
        for f in ("T_f_w", "cov", "estimated_scale", "error_init", "error_final", "num_obs_pt", "num_obs_ls", "pt_outlier",
                  "seg_outlier", "status"):
            save[f"out_{tag}_{f}"] = getattr(out, f)
        # GN pass counts are not observable from outside the reference function: taken from the oracle restatement
        save[f"out_{tag}_iters"] = oracle_lib.poseopt(abi, d, abi.poseopt_params(2.0, 10, n_ref)).iters
    save["source"] = np.array(oracle_lib.load_ref(abi).plsvo_ref_describe().decode())
    np.savez_compressed(os.path.join(HERE, "poseopt.npz"), **save)


if __name__ == "__main__":
    oracle_lib.build()
    if not oracle_lib.build_ref():
        raise SystemExit("oracle/_ref could not be built (needs /root/reference)")
    align_case()
    poseopt_case()
    print("golden fixtures written to", HERE)
