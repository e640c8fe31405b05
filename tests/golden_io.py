"""Load tests/golden/*.npz back into synth.AlignData / synth.PoseOptData objects."""
import os

import numpy as np

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_align(synth):
    z = np.load(os.path.join(HERE, "align_qvga.npz"))
    levels = sorted(int(k.split("_")[-1]) for k in z.files if k.startswith("ref_pyr_"))
    d = synth.AlignData(
        cam=synth.QVGA, max_level=max(levels), min_level=min(levels),
        ref_pyr={l: np.ascontiguousarray(z[f"ref_pyr_{l}"]) for l in levels},
        cur_pyr={l: np.ascontiguousarray(z[f"cur_pyr_{l}"]) for l in levels},
        **{f: np.ascontiguousarray(z[f]) for f in ("T_ref_w", "T_cur_w", "T_cur_w_gt", "pt_px", "pt_f", "pt_pos", "seg_spx",
                                                    "seg_epx", "seg_sf", "seg_ef", "seg_spos", "seg_epos", "seg_length")})
    return d, z


def load_poseopt(synth):
    z = np.load(os.path.join(HERE, "poseopt.npz"))
    d = synth.PoseOptData(fx=synth.VGA.fx, **{f: np.ascontiguousarray(z[f]) for f in (
        "T_f_w", "T_f_w_gt", "pt_f", "pt_pos", "pt_level", "seg_line", "seg_spos", "seg_epos", "seg_level")})
    return d, z
