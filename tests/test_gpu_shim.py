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
