"""CPU tests that pin the oracle (the reference ships no tests or golden vectors — SURVEY.md §4):
closed-form checks of the restated third-party pieces, analytic properties of both optimisers,
cross-agreement with the independent NumPy restatement, and the frozen golden fixtures."""
import ctypes as C

import numpy as np
import pytest
import scipy.linalg

from golden_io import load_align, load_poseopt


def _vec(a):
    return np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(C.POINTER(C.c_double))


# ---- restated Sophus / Eigen pieces against closed forms -----------------------------------------
def test_se3_exp_matches_matrix_exponential(oracle, abi, synth):
    lib = oracle.load(abi)
    rng = np.random.default_rng(0)
    for scale in (1e-12, 1e-6, 1e-2, 1.0):
        xi = rng.normal(size=6) * scale
        out = np.zeros(7)
        lib.plsvo_oracle_se3_exp(_vec(xi), _vec(out))
        M = np.zeros((4, 4))
        w = xi[3:]
        M[:3, :3] = [[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]
        M[:3, 3] = xi[:3]
        E = scipy.linalg.expm(M)
        import torch

        R = synth.quat_to_R(torch.tensor(out[:4])).numpy()
        np.testing.assert_allclose(R, E[:3, :3], atol=1e-13)
        np.testing.assert_allclose(out[4:], E[:3, 3], atol=1e-13 * max(1.0, scale))


def test_se3_group_laws(oracle, abi):
    lib = oracle.load(abi)
    rng = np.random.default_rng(1)
    a, b, ab, inv, e = (np.zeros(7) for _ in range(5))
    lib.plsvo_oracle_se3_exp(_vec(rng.normal(size=6) * 0.3), _vec(a))
    lib.plsvo_oracle_se3_exp(_vec(rng.normal(size=6) * 0.3), _vec(b))
    lib.plsvo_oracle_se3_mul(_vec(a), _vec(b), _vec(ab))
    lib.plsvo_oracle_se3_inverse(_vec(ab), _vec(inv))
    lib.plsvo_oracle_se3_mul(_vec(ab), _vec(inv), _vec(e))
    np.testing.assert_allclose(e, [0, 0, 0, 1, 0, 0, 0], atol=1e-15)
    assert abs(np.linalg.norm(ab[:4]) - 1) < 1e-15


def test_ldlt_solve_and_inverse(oracle, abi):
    lib = oracle.load(abi)
    rng = np.random.default_rng(2)
    for cond in (1.0, 1e4, 1e8):
        J = rng.normal(size=(40, 6)) * np.geomspace(1, np.sqrt(cond), 6)
        A = J.T @ J
        b = rng.normal(size=6)
        x, inv = np.zeros(6), np.zeros(36)
        lib.plsvo_oracle_solve6(_vec(A), _vec(b), _vec(x))
        lib.plsvo_oracle_inverse6(_vec(A), _vec(inv))
        np.testing.assert_allclose(x, np.linalg.solve(A, b), rtol=1e-7 * cond**0.5)
        np.testing.assert_allclose(inv.reshape(6, 6) @ A, np.eye(6), atol=1e-8 * cond**0.5)
    # Eigen's LDLT of the zero matrix solves to zero (not NaN): the GN loop then stops on norm_max(x) <= eps
    x = np.ones(6)
    lib.plsvo_oracle_solve6(_vec(np.zeros(36)), _vec(np.ones(6)), _vec(x))
    assert np.all(x == 0)


# ---- analytic properties ---------------------------------------------------------------------------
def test_poseopt_noiseless_recovers_ground_truth(oracle, abi, synth):
    # points only: exact Jacobians -> quadratic convergence to the ground truth
    d = synth.make_poseopt_batch(batch=4, n_pts=120, n_segs=0, seed=11, noise_px=0.0, outlier_frac=0.0)
    out = oracle.poseopt(abi, d)
    ang, rel = synth.pose_error(out.T_f_w, d.T_f_w_gt)
    assert ang.max() < 1e-9 and rel.max() < 1e-8  # a wrong Jacobian or SE3 update cannot reach this
    assert out.pt_outlier.sum() == 0
    assert (out.error_final < 1e-6).all()
    # with lines the reference's Jacobian uses ds for both endpoint rows (pose_optimizer.cpp:157-158):
    # still a fixed point at the ground truth, but only linear convergence within the 10 iterations
    d = synth.make_poseopt_batch(batch=4, n_pts=120, n_segs=30, seed=11, noise_px=0.0, outlier_frac=0.0)
    out = oracle.poseopt(abi, d)
    ang0, _ = synth.pose_error(d.T_f_w, d.T_f_w_gt)
    ang, _ = synth.pose_error(out.T_f_w, d.T_f_w_gt)
    assert (ang < 1e-2 * ang0).all()


def test_poseopt_rejects_outliers(oracle, abi, synth):
    # small initial error (as after sparse alignment): the MAD scale is taken once, at the initial pose
    d = synth.make_poseopt_batch(batch=4, n_pts=200, n_segs=0, seed=12, noise_px=0.2, outlier_frac=0.15,
                                 pert_t=0.002, pert_r=0.001)
    out = oracle.poseopt(abi, d)
    frac = out.pt_outlier.mean()
    assert 0.08 < frac < 0.25
    ang0, _ = synth.pose_error(d.T_f_w, d.T_f_w_gt)
    ang, _ = synth.pose_error(out.T_f_w, d.T_f_w_gt)
    assert ang.max() < 5e-4 and np.median(ang) < 0.5 * np.median(ang0)


def test_align_zero_motion_is_a_fixed_point(oracle, abi, synth):
    # points only: identical images -> zero residuals, Jres = 0, x = 0 at the coarsest level
    d = synth.make_align_batch(cam=synth.QVGA, batch=2, n_pts=40, n_segs=0, max_level=3, min_level=1, seed=21, margin=32,
                               motion_t=0.0, motion_r=0.0)
    out = oracle.align(abi, d, abi.align_params(3, 1, 30))
    ang, rel = synth.pose_error(out.T_cur_w, d.T_cur_w_gt)
    assert ang.max() < 1e-7 and (out.status == 0).all()
    tr = oracle.align_trace(abi, d, 0, abi.align_params(3, 1, 30))
    assert np.abs(tr[0, 41:47]).max() == 0.0 and tr[0, 2] == 0.0 and np.abs(tr[0, 47:53]).max() == 0.0
    # with segments the reference divides by the mean absolute residual (H += H_*w/res_,
    # sparse_img_align.cpp:681): a perfectly matching segment gives inf/NaN, solve() fails, stop_ is
    # raised and stays raised for the remaining levels -> the pose is left untouched.
    d = synth.make_align_batch(cam=synth.QVGA, batch=2, n_pts=40, n_segs=8, max_level=3, min_level=1, seed=21, margin=32,
                               motion_t=0.0, motion_r=0.0)
    out = oracle.align(abi, d, abi.align_params(3, 1, 30))
    ang, rel = synth.pose_error(out.T_cur_w, d.T_cur_w_gt)
    assert ang.max() < 1e-12 and (out.status == 2).all()
    assert (out.iters[:, 1:4] == 1).all()


def test_align_reduces_pose_error(oracle, abi, synth):
    d = synth.make_align_batch(cam=synth.QVGA, batch=4, n_pts=80, n_segs=16, max_level=3, min_level=1, seed=22, margin=32,
                               motion_t=0.015, motion_r=0.005)
    out = oracle.align(abi, d, abi.align_params(3, 1, 30))
    a0, r0 = synth.pose_error(d.T_cur_w, d.T_cur_w_gt)
    a1, r1 = synth.pose_error(out.T_cur_w, d.T_cur_w_gt)
    assert np.median(a1) < 0.3 * np.median(a0)
    assert (out.n_tracked == (80 * 16 + 16) // 16).all()  # every patch tracked: n_meas/16 (:94)


def test_align_early_out_and_invalid_features(oracle, abi, synth):
    d = synth.make_align_batch(cam=synth.QVGA, batch=2, n_pts=20, n_segs=4, max_level=3, min_level=1, seed=23, margin=32)
    d.pt_count = np.array([0, 20], np.int32)
    d.seg_count = np.array([0, 4], np.int32)
    out = oracle.align(abi, d, abi.align_params(3, 1, 30))
    assert out.status[0] == 1 and out.n_tracked[0] == 0
    np.testing.assert_array_equal(out.T_cur_w[0], d.T_cur_w[0])
    d.pt_count = None
    d.seg_count = None
    d.pt_valid = np.zeros((2, 20), np.uint8)
    d.seg_valid = np.zeros((2, 4), np.uint8)
    out = oracle.align(abi, d, abi.align_params(3, 1, 30))  # lists non-empty, all feat3D NULL: H = 0 -> x = 0
    assert (out.n_tracked == 0).all() and (out.status == 0).all()
    ang, _ = synth.pose_error(out.T_cur_w, d.T_cur_w)
    assert ang.max() < 1e-15


# ---- two independent restatements agree -------------------------------------------------------------
def test_cpp_and_numpy_restatements_agree_align(oracle, abi, synth):
    import np_oracle

    d, _ = load_align(synth)
    for b in range(d.batch):
        tr = oracle.align_trace(abi, d, b, abi.align_params(3, 1, 30))
        res = np_oracle.align_pair(d, b, 3, 1, 30)
        assert len(res["trace"]) == len(tr)
        for rec, ref in zip(res["trace"], tr):
            assert rec["level"] == int(ref[0]) and rec["iter"] == int(ref[1])
            assert rec["n_meas"] == int(ref[3]) and rec["accepted"] == bool(ref[4])
            H = ref[5:41].reshape(6, 6)
            np.testing.assert_allclose(rec["H"], H, rtol=0, atol=1e-9 * np.abs(H).max())
            np.testing.assert_allclose(rec["Jres"], ref[41:47], rtol=0, atol=1e-8 * np.abs(ref[41:47]).max() + 1e-6)
            np.testing.assert_allclose(rec["chi2"], ref[2], rtol=2e-5)
            np.testing.assert_allclose(rec["x"], ref[47:53], rtol=0, atol=1e-6 * np.abs(ref[47:53]).max() + 1e-12)


def test_cpp_and_numpy_restatements_agree_poseopt(oracle, abi, synth):
    import np_oracle
    import torch

    d, _ = load_poseopt(synth)
    out = oracle.poseopt(abi, d)
    for b in range(d.batch):
        res = np_oracle.poseopt_frame(d, b)
        R = synth.quat_to_R(torch.tensor(out.T_f_w[b, :4])).numpy()
        np.testing.assert_allclose(res["R"], R, atol=1e-10)
        np.testing.assert_allclose(res["t"], out.T_f_w[b, 4:], atol=1e-10)
        assert res["iters"] == out.iters[b, 0]
        np.testing.assert_array_equal(res["pt_outlier"], out.pt_outlier[b].astype(bool))
        np.testing.assert_array_equal(res["seg_outlier"], out.seg_outlier[b].astype(bool))
        np.testing.assert_allclose(res["cov"].ravel(), out.cov[b], rtol=1e-6, atol=1e-12)
        np.testing.assert_allclose(res["scale"], out.estimated_scale[b], rtol=1e-12)


# ---- frozen golden vectors ----------------------------------------------------------------------------
def test_oracle_reproduces_golden_align(oracle, abi, synth):
    d, z = load_align(synth)
    out = oracle.align(abi, d, abi.align_params(3, 1, 30), n_threads=2)
    np.testing.assert_array_equal(out.T_cur_w, z["out_T_cur_w"])
    np.testing.assert_array_equal(out.n_tracked, z["out_n_tracked"])
    np.testing.assert_array_equal(out.iters, z["out_iters"])
    np.testing.assert_array_equal(out.seg_killed, z["out_seg_killed"])
    np.testing.assert_array_equal(out.H, z["out_H"])
    for b in range(d.batch):
        np.testing.assert_array_equal(oracle.align_trace(abi, d, b, abi.align_params(3, 1, 30)), z[f"trace_{b}"])


@pytest.mark.parametrize("tag,n_ref", [("9arg", -1), ("10arg", 3)])
def test_oracle_reproduces_golden_poseopt(oracle, abi, synth, tag, n_ref):
    d, z = load_poseopt(synth)
    out = oracle.poseopt(abi, d, abi.poseopt_params(2.0, 10, n_ref))
    for f in ("T_f_w", "cov", "estimated_scale", "error_init", "error_final", "num_obs_pt", "num_obs_ls", "pt_outlier",
              "seg_outlier", "iters"):
        np.testing.assert_array_equal(getattr(out, f), z[f"out_{tag}_{f}"])
