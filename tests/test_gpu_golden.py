"""GPU results against the committed golden fixtures (no oracle needed at run time)."""
import numpy as np
import pytest

from golden_io import load_align, load_poseopt

pytestmark = pytest.mark.gpu


def test_gpu_align_matches_golden(pkg, synth):
    d, z = load_align(synth)
    out = pkg.SparseImgAlign(3, 1, 30).run(d)
    ang, rel = synth.pose_error(out.T_cur_w, z["out_T_cur_w"])
    assert ang.max() <= 1e-5 and rel.max() <= 1e-4
    np.testing.assert_array_equal(out.n_tracked, z["out_n_tracked"])
    np.testing.assert_array_equal(out.seg_killed, z["out_seg_killed"])
    np.testing.assert_array_equal(out.patch_levels, z["out_patch_levels"])
    same = (out.iters == z["out_iters"]).all(axis=1)
    np.testing.assert_array_equal(out.patch_iters[same], z["out_patch_iters"][same])
    H = z["out_H"]
    assert (np.abs(out.H - H)[same] / np.abs(H).max(axis=1, keepdims=True)[same]).max() < 1e-4


@pytest.mark.parametrize("tag,n_ref", [("9arg", None), ("10arg", 3)])
def test_gpu_poseopt_matches_golden(pkg, synth, tag, n_ref):
    d, z = load_poseopt(synth)
    out = pkg.pose_optimizer.optimizeGaussNewton(2.0, 10, False, d, n_iter_ref=n_ref)
    ang, rel = synth.pose_error(out.T_f_w, z[f"out_{tag}_T_f_w"])
    assert ang.max() <= 1e-5 and rel.max() <= 1e-4
    for f in ("num_obs_pt", "num_obs_ls", "pt_outlier", "seg_outlier", "iters"):
        np.testing.assert_array_equal(getattr(out, f), z[f"out_{tag}_{f}"])
    np.testing.assert_allclose(out.error_final, z[f"out_{tag}_error_final"], rtol=1e-6)
    np.testing.assert_allclose(out.estimated_scale, z[f"out_{tag}_estimated_scale"], rtol=1e-12)
