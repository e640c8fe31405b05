"""GPU parity of the pose-optimiser kernel against the CPU oracle, through the C ABI."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROT_TOL = 1e-5
TRANS_TOL = 1e-4


def _check(synth, gpu, ref):
    ang, rel = synth.pose_error(gpu.T_f_w, ref.T_f_w)
    assert ang.max() <= ROT_TOL, f"rotation parity {ang.max():.3e}"
    assert rel.max() <= TRANS_TOL, f"translation parity {rel.max():.3e}"
    np.testing.assert_array_equal(gpu.status, ref.status)
    np.testing.assert_array_equal(gpu.iters, ref.iters)
    np.testing.assert_array_equal(gpu.pt_outlier, ref.pt_outlier)
    np.testing.assert_array_equal(gpu.seg_outlier, ref.seg_outlier)
    np.testing.assert_array_equal(gpu.num_obs_pt, ref.num_obs_pt)
    np.testing.assert_array_equal(gpu.num_obs_ls, ref.num_obs_ls)
    ok = ref.status == 0
    np.testing.assert_allclose(gpu.estimated_scale[ok], ref.estimated_scale[ok], rtol=1e-12)
    np.testing.assert_allclose(gpu.error_init[ok], ref.error_init[ok], rtol=1e-9)
    np.testing.assert_allclose(gpu.error_final[ok], ref.error_final[ok], rtol=1e-6)
    scale = np.abs(ref.cov).max(axis=1, keepdims=True) + 1e-300
    assert (np.abs(gpu.cov - ref.cov)[ok] / scale[ok]).max() < 1e-6


def test_poseopt_9arg(pkg, abi, synth, oracle):
    data = synth.make_poseopt_batch(batch=64, n_pts=300, n_segs=80, seed=5000)
    gpu = pkg.pose_optimizer.optimizeGaussNewton(2.0, 10, False, data)
    ref = oracle.poseopt(abi, data, abi.poseopt_params(2.0, 10, -1), n_threads=8)
    _check(synth, gpu, ref)
    ang0, _ = synth.pose_error(data.T_f_w, data.T_f_w_gt)
    ang, _ = synth.pose_error(gpu.T_f_w, data.T_f_w_gt)
    assert np.median(ang) < 0.5 * np.median(ang0)


def test_poseopt_10arg_refinement(pkg, abi, synth, oracle):
    data = synth.make_poseopt_batch(batch=32, n_pts=300, n_segs=80, seed=5100)
    gpu = pkg.pose_optimizer.optimizeGaussNewton(2.0, 10, False, data, n_iter_ref=3)
    ref = oracle.poseopt(abi, data, abi.poseopt_params(2.0, 10, 3), n_threads=8)
    _check(synth, gpu, ref)


def test_poseopt_points_only_and_lines_dead(pkg, abi, synth, oracle):
    data = synth.make_poseopt_batch(batch=16, n_pts=200, n_segs=0, seed=5200)
    gpu = pkg.pose_optimizer.optimizeGaussNewton(2.0, 10, False, data)
    ref = oracle.poseopt(abi, data, abi.poseopt_params(2.0, 10, -1))
    _check(synth, gpu, ref)


def test_poseopt_edge_cases(pkg, abi, synth, oracle):
    data = synth.make_poseopt_batch(batch=10, n_pts=64, n_segs=24, seed=5300)
    rng = np.random.default_rng(9)
    data.pt_valid = (rng.uniform(size=(10, 64)) > 0.25).astype(np.uint8)
    data.seg_valid = (rng.uniform(size=(10, 24)) > 0.25).astype(np.uint8)
    data.pt_count = rng.integers(8, 65, 10).astype(np.int32)
    data.seg_count = rng.integers(0, 25, 10).astype(np.int32)
    data.pt_valid[2] = 0
    data.seg_valid[2] = 0  # no observations at all: early return, outputs untouched
    gpu = pkg.pose_optimizer.optimizeGaussNewton(2.0, 10, False, data)
    ref = oracle.poseopt(abi, data, abi.poseopt_params(2.0, 10, -1))
    _check(synth, gpu, ref)
    assert gpu.status[2] == 1
    np.testing.assert_array_equal(gpu.T_f_w[2], data.T_f_w[2])
