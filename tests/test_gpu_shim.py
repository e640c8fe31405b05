"""The C++ shim keeps the reference's two call signatures (pl-svo_b200/host/plsvo_shim.h).  The test
harness builds Frame objects and calls them exactly like FrameHandlerMono::processFrame does; results
must agree with the oracle on the same inputs (same kernels as the batch ABI, B = 1)."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "pl-svo_b200", "host", "libplsvo_shim.so")

dp = C.POINTER(C.c_double)
up = C.POINTER(C.c_ubyte)
ip = C.POINTER(C.c_int)


def _d(a):
    return np.ascontiguousarray(a, np.float64).ctypes.data_as(dp)


@pytest.fixture(scope="module")
def shim():
    assert os.path.exists(SHIM), "libplsvo_shim.so not built (python __graft_entry__.py build)"
    lib = C.CDLL(SHIM)
    lib.plsvo_shim_test_align.restype = C.c_longlong
    lib.plsvo_shim_test_poseopt.restype = C.c_int
    return lib


def test_shim_sparse_img_align_run(shim, abi, synth, oracle, gen_device):
    d = synth.make_align_batch(batch=3, n_pts=150, n_segs=40, device=gen_device, seed=8100, keep_levels_only=False)
    ref = oracle.align(abi, d, n_threads=3)
    nl = 5
    for b in range(d.batch):
        keep = [np.ascontiguousarray(d.ref_pyr[l][b]) for l in range(nl)] + [np.ascontiguousarray(d.cur_pyr[l][b]) for l in range(nl)]
        refl = (up * nl)(*[k.ctypes.data_as(up) for k in keep[:nl]])
        curl = (up * nl)(*[k.ctypes.data_as(up) for k in keep[nl:]])
        T_out, killed, fisher = np.zeros(7), np.zeros(d.n_segs, np.uint8), np.zeros(36)
        cam = d.cam
        n = shim.plsvo_shim_test_align(
            C.c_int(cam.width), C.c_int(cam.height), C.c_double(cam.fx), C.c_double(cam.fy), C.c_double(cam.cx), C.c_double(cam.cy),
            C.c_int(nl), refl, curl, _d(d.T_ref_w[b]), _d(d.T_cur_w[b]), C.c_int(d.n_pts), _d(d.pt_px[b]), _d(d.pt_f[b]),
            _d(d.pt_pos[b]), None, C.c_int(d.n_segs), _d(d.seg_spx[b]), _d(d.seg_epx[b]), _d(d.seg_sf[b]), _d(d.seg_ef[b]),
            _d(d.seg_spos[b]), _d(d.seg_epos[b]), _d(d.seg_length[b]), None, C.c_int(4), C.c_int(2), C.c_int(30),
            T_out.ctypes.data_as(dp), killed.ctypes.data_as(up), fisher.ctypes.data_as(dp))
        assert n == ref.n_tracked[b]
        ang, rel = synth.pose_error(T_out[None], ref.T_cur_w[b][None])
        assert ang.max() <= 1e-5 and rel.max() <= 1e-4
        np.testing.assert_array_equal(killed, ref.seg_killed[b])
        np.testing.assert_allclose(fisher * (5e-4 * 255 * 255), ref.H[b], rtol=1e-3, atol=1e-6 * np.abs(ref.H[b]).max())


@pytest.mark.parametrize("n_ref", [-1, 3])
def test_shim_pose_optimizer(shim, abi, synth, oracle, n_ref):
    d = synth.make_poseopt_batch(batch=3, n_pts=150, n_segs=40, seed=8200)
    ref = oracle.poseopt(abi, d, abi.poseopt_params(2.0, 10, n_ref))
    for b in range(d.batch):
        T_out, cov, sc = np.zeros(7), np.zeros(36), np.zeros(5)
        po, so = np.zeros(d.n_pts, np.uint8), np.zeros(d.n_segs, np.uint8)
        lvl = np.ascontiguousarray(d.pt_level[b], np.int32)
        slvl = np.ascontiguousarray(d.seg_level[b], np.int32)
        shim.plsvo_shim_test_poseopt(
            C.c_double(d.fx), _d(d.T_f_w[b]), C.c_int(d.n_pts), _d(d.pt_f[b]), _d(d.pt_pos[b]), lvl.ctypes.data_as(ip), None,
            C.c_int(d.n_segs), _d(d.seg_line[b]), _d(d.seg_spos[b]), _d(d.seg_epos[b]), slvl.ctypes.data_as(ip), None,
            C.c_double(2.0), C.c_int(10), C.c_int(n_ref), T_out.ctypes.data_as(dp), cov.ctypes.data_as(dp),
            sc.ctypes.data_as(dp), po.ctypes.data_as(up), so.ctypes.data_as(up))
        ang, rel = synth.pose_error(T_out[None], ref.T_f_w[b][None])
        assert ang.max() <= 1e-5 and rel.max() <= 1e-4
        np.testing.assert_array_equal(po, ref.pt_outlier[b])
        np.testing.assert_array_equal(so, ref.seg_outlier[b])
        np.testing.assert_allclose(sc[0], ref.estimated_scale[b], rtol=1e-12)
        np.testing.assert_allclose(sc[2], ref.error_final[b], rtol=1e-6)
        assert int(sc[3]) == ref.num_obs_pt[b] and int(sc[4]) == ref.num_obs_ls[b]
        np.testing.assert_allclose(cov, ref.cov[b], rtol=1e-5, atol=1e-9 * np.abs(ref.cov[b]).max())


# ---- the reference's OWN Frame / Feature / SE3 types through the shim (oracle/shimref_harness.cpp) ------------------
@pytest.fixture(scope="module")
def shimref(oracle, abi):
    """oracle/_ref/libplsvo_shimref.so: plsvo_shim.cpp compiled in -DPLSVO_SHIM_WITH_REFERENCE_HEADERS mode against the
    reference's real class definitions (built where /root/reference exists, travels prebuilt to the GPU box)."""
    if not oracle.build_shimref():
        pytest.skip("oracle/_ref/libplsvo_shimref.so is not built and /root/reference is absent")
    oracle.load_shimref(abi)
    return oracle


def test_reference_typed_frames_through_the_shim_align(shimref, pkg, abi, synth, gen_device):
    d = synth.make_align_batch(batch=6, n_pts=200, n_segs=60, device=gen_device, seed=8300)
    d.seg_valid = np.ones((6, 60), np.uint8)
    d.seg_valid[:, ::7] = 0
    d.pt_valid = np.ones((6, 200), np.uint8)
    d.pt_valid[:, ::11] = 0
    got = shimref.shimref_align(abi, d)                 # reference objects -> shim -> C ABI (B = 1 per pair) -> CUDA
    direct = pkg.SparseImgAlign(4, 2, 30).run(d)        # the C ABI called directly on the whole batch
    cpu = shimref.align(abi, d, n_threads=4)            # the oracle
    # same kernel, same arrays — except that the poses make a trip through the reference's SE3 (the quaternion is
    # re-normalised: last-bit input differences), so the poses agree to round-off, not bit for bit.  The kernel's
    # accept / rollback decisions are taken on the reference's own float chi2, so the exact
    # quantities agree on every pair.
    np.testing.assert_array_equal(got.n_tracked, direct.n_tracked)
    np.testing.assert_array_equal(got.seg_killed, direct.seg_killed)
    ang, rel = synth.pose_error(got.T_cur_w, direct.T_cur_w)
    assert ang.max() <= 1e-9 and rel.max() <= 1e-8
    Hs = got.H * (5e-4 * 255 * 255)
    for b in range(d.batch):
        assert np.allclose(Hs[b], direct.H[b], rtol=1e-6, atol=1e-9 * np.abs(direct.H[b]).max())
    ang, rel = synth.pose_error(got.T_cur_w, cpu.T_cur_w)
    assert ang.max() <= 1e-5 and rel.max() <= 1e-4
    np.testing.assert_array_equal(got.n_tracked, cpu.n_tracked)
    np.testing.assert_array_equal(got.seg_killed, cpu.seg_killed)


@pytest.mark.parametrize("n_ref", [-1, 3])
def test_reference_typed_frames_through_the_shim_poseopt(shimref, pkg, abi, synth, n_ref):
    d = synth.make_poseopt_batch(batch=6, n_pts=200, n_segs=60, seed=8400)
    p = abi.poseopt_params(2.0, 10, n_ref)
    got = shimref.shimref_poseopt(abi, d, p)
    direct = pkg.pose_optimizer.optimizeGaussNewton(2.0, 10, False, d, n_iter_ref=None if n_ref < 0 else n_ref)
    for f in ("num_obs_pt", "num_obs_ls", "pt_outlier", "seg_outlier"):
        np.testing.assert_array_equal(getattr(got, f), getattr(direct, f), err_msg=f)
    np.testing.assert_allclose(got.T_f_w, direct.T_f_w, rtol=0, atol=1e-9)
    np.testing.assert_allclose(got.cov, direct.cov, rtol=1e-6, atol=1e-12)
    for f in ("estimated_scale", "error_init", "error_final"):
        np.testing.assert_allclose(getattr(got, f), getattr(direct, f), rtol=1e-9, err_msg=f)
