"""Loader of the CPU oracle (oracle/libplsvo_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (pl-svo_b200/) never imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libplsvo_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "plsvo_oracle.cpp")
    hdr = os.path.join(_HERE, "..", "include", "plsvo_b200.h")
    stale = (not os.path.exists(LIB_PATH)) or any(
        os.path.exists(f) and os.path.getmtime(f) > os.path.getmtime(LIB_PATH) for f in (src, hdr)
    )
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"] + ([] if force else []))
    return LIB_PATH


def load(abi):
    """abi = the pl-svo_b200.abi module (struct definitions are shared with the product header)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    lib = C.CDLL(LIB_PATH)
    P = C.POINTER
    lib.plsvo_oracle_align_batch.restype = C.c_int
    lib.plsvo_oracle_align_batch.argtypes = [P(abi.AlignBatch), P(abi.AlignParams), P(abi.AlignResult), C.c_int, C.c_int]
    lib.plsvo_oracle_align_trace.restype = C.c_int
    lib.plsvo_oracle_align_trace.argtypes = [P(abi.AlignBatch), P(abi.AlignParams), C.c_int, P(C.c_double), C.c_int, P(C.c_int)]
    lib.plsvo_oracle_trace_stride.restype = C.c_int
    lib.plsvo_oracle_poseopt_batch.restype = C.c_int
    lib.plsvo_oracle_poseopt_batch.argtypes = [P(abi.PoseOptBatch), P(abi.PoseOptParams), P(abi.PoseOptResult), C.c_int]
    for name, n_in in (("plsvo_oracle_se3_exp", 1), ("plsvo_oracle_se3_inverse", 1), ("plsvo_oracle_se3_mul", 2),
                       ("plsvo_oracle_solve6", 2), ("plsvo_oracle_inverse6", 1)):
        fn = getattr(lib, name)
        fn.restype = None
        fn.argtypes = [P(C.c_double)] * (n_in + 1)
    lib.plsvo_oracle_half_sample.restype = None
    lib.plsvo_oracle_half_sample.argtypes = [C.POINTER(C.c_uint8), C.c_int, C.c_int, C.c_size_t, C.POINTER(C.c_uint8), C.c_size_t]
    lib.plsvo_oracle_align2d.restype = C.c_int
    lib.plsvo_oracle_align2d.argtypes = [C.POINTER(C.c_uint8), C.c_int, C.c_int, C.c_size_t, C.POINTER(C.c_uint8), C.POINTER(C.c_uint8), C.c_int, C.POINTER(C.c_double)]
    lib.plsvo_oracle_align1d.restype = C.c_int
    lib.plsvo_oracle_align1d.argtypes = [C.POINTER(C.c_uint8), C.c_int, C.c_int, C.c_size_t, C.POINTER(C.c_float), C.POINTER(C.c_uint8),
                                         C.POINTER(C.c_uint8), C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.plsvo_oracle_hardware_threads.restype = C.c_int
    _lib = lib
    return lib


def align(abi, data, params=None, n_threads: int = 1, flags: int = 0):
    """Run the oracle's SparseImgAlign restatement on an AlignData batch -> abi.AlignOut."""
    lib = load(abi)
    params = params or abi.align_params(data.max_level, data.min_level)
    batch, keep = abi.make_align_batch(data)
    out = abi.AlignOut(data.batch, data.n_segs)
    rc = lib.plsvo_oracle_align_batch(C.byref(batch), C.byref(params), C.byref(out.struct), n_threads, flags)
    if rc != 0:
        raise RuntimeError(f"oracle align failed rc={rc}")
    return out


def align_trace(abi, data, pair: int, params=None, max_records: int = 512):
    lib = load(abi)
    params = params or abi.align_params(data.max_level, data.min_level)
    batch, keep = abi.make_align_batch(data)
    stride = lib.plsvo_oracle_trace_stride()
    rec = np.zeros((max_records, stride))
    n = C.c_int(0)
    rc = lib.plsvo_oracle_align_trace(C.byref(batch), C.byref(params), pair,
                                      rec.ctypes.data_as(C.POINTER(C.c_double)), max_records, C.byref(n))
    if rc != 0:
        raise RuntimeError(f"oracle trace failed rc={rc}")
    return rec[: n.value]


def poseopt(abi, data, params=None, n_threads: int = 1):
    lib = load(abi)
    params = params or abi.poseopt_params()
    batch, keep = abi.make_poseopt_batch(data)
    out = abi.PoseOptOut(data.batch, data.n_pts, data.n_segs)
    rc = lib.plsvo_oracle_poseopt_batch(C.byref(batch), C.byref(params), C.byref(out.struct), n_threads)
    if rc != 0:
        raise RuntimeError(f"oracle poseopt failed rc={rc}")
    return out


def pyramid(abi, img0, n_levels: int):
    """frame_utils::createImgPyramid restated: u8 [B,H,W] -> list of levels."""
    lib = load(abi)
    img0 = np.ascontiguousarray(img0, np.uint8)
    levels = [img0]
    u8p = C.POINTER(C.c_uint8)
    for _ in range(1, n_levels):
        prev = levels[-1]
        B, H, W = prev.shape
        out = np.empty((B, H // 2, W // 2), np.uint8)
        for b in range(B):
            lib.plsvo_oracle_half_sample(prev[b].ctypes.data_as(u8p), W, H, prev.strides[1], out[b].ctypes.data_as(u8p), out.strides[1])
        levels.append(out)
    return levels


def align2d(abi, cur_pyr, image_index, level, border, ref, px, n_iter):
    """feature_alignment::align2D restated, looped over features -> (converged [n] bool, px [n,2])."""
    lib = load(abi)
    u8p, dp = C.POINTER(C.c_uint8), C.POINTER(C.c_double)
    out = np.array(px, np.float64, copy=True)
    conv = np.zeros(len(image_index), bool)
    border = np.ascontiguousarray(border, np.uint8)
    ref = np.ascontiguousarray(ref, np.uint8)
    for i in range(len(image_index)):
        im = np.ascontiguousarray(cur_pyr[int(level[i])][int(image_index[i])])
        p = out[i].copy()
        conv[i] = bool(lib.plsvo_oracle_align2d(im.ctypes.data_as(u8p), im.shape[1], im.shape[0], im.strides[0],
                                                border[i].ctypes.data_as(u8p), ref[i].ctypes.data_as(u8p), n_iter,
                                                p.ctypes.data_as(dp)))
        out[i] = p
    return conv, out


# ---- oracle/_ref: the reference's own translation units (oracle/ref_harness.cpp, Makefile target `ref`) ----
REF_LIB_PATH = os.path.join(_HERE, "_ref", "libplsvo_ref.so")
REFERENCE_ROOT = os.environ.get("PLSVO_REFERENCE_ROOT", "/root/reference")
_ref_lib = None


def build_ref(force: bool = False) -> str | None:
    """Build oracle/_ref/libplsvo_ref.so when the reference sources are present (authoring container only).
    Returns the path, or None when neither the sources nor a prebuilt library exist."""
    srcs = [os.path.join(REFERENCE_ROOT, "src", f) for f in ("sparse_img_align.cpp", "pose_optimizer.cpp", "feature.cpp", "feature_alignment.cpp", "matcher.cpp", "config.cpp", "feature3D_impl.cpp", "depth_filter.cpp")]
    if all(os.path.exists(s) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "ref", f"REFERENCE={REFERENCE_ROOT}"] + (["-B"] if force else ["-s"]))
    return REF_LIB_PATH if os.path.exists(REF_LIB_PATH) else None


SHIMREF_LIB_PATH = os.path.join(_HERE, "_ref", "libplsvo_shimref.so")
_shimref_lib = None


def build_shimref(force: bool = False) -> str | None:
    """oracle/_ref/libplsvo_shimref.so: reference-typed Frame / Feature objects driven through the B200 shim compiled in
    reference-headers mode (oracle/shimref_harness.cpp).  Needs /root/reference and the built CUDA library to build."""
    if os.path.exists(os.path.join(REFERENCE_ROOT, "src", "feature.cpp")) and os.path.exists(
            os.path.join(_HERE, "..", "pl-svo_b200", "csrc", "libplsvo_b200.so")):
        subprocess.check_call(["make", "-C", _HERE, "shimref", f"REFERENCE={REFERENCE_ROOT}"] + (["-B"] if force else ["-s"]))
    return SHIMREF_LIB_PATH if os.path.exists(SHIMREF_LIB_PATH) else None


SHIMREF_CPU_LIB_PATH = os.path.join(_HERE, "_ref", "libplsvo_shimref_cpu.so")
_shimref_cpu_lib = None


def build_shimref_cpu(force: bool = False) -> str | None:
    """oracle/_ref/libplsvo_shimref_cpu.so: the same harness + shim with the C ABI answered by the CPU oracle
    (oracle/abi_on_oracle.cpp) — checks the shim's packing of reference objects without a GPU."""
    if os.path.exists(os.path.join(REFERENCE_ROOT, "src", "feature.cpp")):
        subprocess.check_call(["make", "-C", _HERE, "shimref-cpu", f"REFERENCE={REFERENCE_ROOT}"] + (["-B"] if force else ["-s"]))
    return SHIMREF_CPU_LIB_PATH if os.path.exists(SHIMREF_CPU_LIB_PATH) else None


def load_shimref(abi, cpu: bool = False):
    global _shimref_lib, _shimref_cpu_lib
    if cpu:
        if _shimref_cpu_lib is None:
            lib = C.CDLL(SHIMREF_CPU_LIB_PATH)
            P = C.POINTER
            lib.plsvo_shimref_align_batch.restype = C.c_int
            lib.plsvo_shimref_align_batch.argtypes = [P(abi.AlignBatch), P(abi.AlignParams), P(abi.AlignResult)]
            lib.plsvo_shimref_poseopt_batch.restype = C.c_int
            lib.plsvo_shimref_poseopt_batch.argtypes = [P(abi.PoseOptBatch), P(abi.PoseOptParams), P(abi.PoseOptResult)]
            _shimref_cpu_lib = lib
        return _shimref_cpu_lib
    if _shimref_lib is None:
        lib = C.CDLL(SHIMREF_LIB_PATH)
        P = C.POINTER
        lib.plsvo_shimref_align_batch.restype = C.c_int
        lib.plsvo_shimref_align_batch.argtypes = [P(abi.AlignBatch), P(abi.AlignParams), P(abi.AlignResult)]
        lib.plsvo_shimref_poseopt_batch.restype = C.c_int
        lib.plsvo_shimref_poseopt_batch.argtypes = [P(abi.PoseOptBatch), P(abi.PoseOptParams), P(abi.PoseOptResult)]
        _shimref_lib = lib
    return _shimref_lib


def shimref_align(abi, data, params=None, cpu: bool = False):
    """plsvo::SparseImgAlign(...).run(ref, cur) of the B200 shim on reference-typed frames -> abi.AlignOut
    (T_cur_w, n_tracked, seg_killed; H holds getFisherInformation() = H / (5e-4 * 255^2))."""
    lib = load_shimref(abi, cpu)
    params = params or abi.align_params(data.max_level, data.min_level)
    batch, keep = abi.make_align_batch(data)
    out = abi.AlignOut(data.batch, data.n_segs)
    rc = lib.plsvo_shimref_align_batch(C.byref(batch), C.byref(params), C.byref(out.struct))
    if rc != 0:
        raise RuntimeError(f"shimref align failed rc={rc}")
    return out


def shimref_poseopt(abi, data, params=None, cpu: bool = False):
    lib = load_shimref(abi, cpu)
    params = params or abi.poseopt_params()
    batch, keep = abi.make_poseopt_batch(data)
    out = abi.PoseOptOut(data.batch, data.n_pts, data.n_segs)
    rc = lib.plsvo_shimref_poseopt_batch(C.byref(batch), C.byref(params), C.byref(out.struct))
    if rc != 0:
        raise RuntimeError(f"shimref poseopt failed rc={rc}")
    return out


def ref_available() -> bool:
    return os.path.exists(REF_LIB_PATH)


def load_ref(abi):
    global _ref_lib
    if _ref_lib is not None:
        return _ref_lib
    if not os.path.exists(REF_LIB_PATH):
        raise FileNotFoundError(f"{REF_LIB_PATH} is not built (needs {REFERENCE_ROOT}; run `make -C oracle ref`)")
    lib = C.CDLL(REF_LIB_PATH)
    P = C.POINTER
    lib.plsvo_ref_align_batch.restype = C.c_int
    lib.plsvo_ref_align_batch.argtypes = [P(abi.AlignBatch), P(abi.AlignParams), P(abi.AlignResult), C.c_int]
    lib.plsvo_ref_poseopt_batch.restype = C.c_int
    lib.plsvo_ref_poseopt_batch.argtypes = [P(abi.PoseOptBatch), P(abi.PoseOptParams), P(abi.PoseOptResult), C.c_int]
    lib.plsvo_ref_describe.restype = C.c_char_p
    u8p, dp = P(C.c_uint8), P(C.c_double)
    lib.plsvo_ref_align2d.restype = C.c_int
    lib.plsvo_ref_align2d.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, u8p, u8p, C.c_int, dp]
    lib.plsvo_ref_align1d.restype = C.c_int
    lib.plsvo_ref_align1d.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, P(C.c_float), u8p, u8p, C.c_int, dp, dp]
    _ref_lib = lib
    return lib


def ref_align(abi, data, params=None, n_threads: int = 1):
    """SparseImgAlign::run of the reference's own sparse_img_align.cpp on an AlignData batch -> abi.AlignOut."""
    lib = load_ref(abi)
    params = params or abi.align_params(data.max_level, data.min_level)
    batch, keep = abi.make_align_batch(data)
    out = abi.AlignOut(data.batch, data.n_segs)
    rc = lib.plsvo_ref_align_batch(C.byref(batch), C.byref(params), C.byref(out.struct), n_threads)
    if rc != 0:
        raise RuntimeError(f"reference align failed rc={rc}")
    return out


def ref_poseopt(abi, data, params=None, n_threads: int = 1):
    """pose_optimizer::optimizeGaussNewton of the reference's own pose_optimizer.cpp -> abi.PoseOptOut."""
    lib = load_ref(abi)
    params = params or abi.poseopt_params()
    batch, keep = abi.make_poseopt_batch(data)
    out = abi.PoseOptOut(data.batch, data.n_pts, data.n_segs)
    rc = lib.plsvo_ref_poseopt_batch(C.byref(batch), C.byref(params), C.byref(out.struct), n_threads)
    if rc != 0:
        raise RuntimeError(f"reference poseopt failed rc={rc}")
    return out


def ref_align2d(abi, cur_pyr, image_index, level, border, ref, px, n_iter):
    """feature_alignment::align2D of the reference's own feature_alignment.cpp, looped over features."""
    lib = load_ref(abi)
    u8p, dp = C.POINTER(C.c_uint8), C.POINTER(C.c_double)
    out = np.array(px, np.float64, copy=True)
    conv = np.zeros(len(image_index), bool)
    border = np.ascontiguousarray(border, np.uint8)
    ref = np.ascontiguousarray(ref, np.uint8)
    for i in range(len(image_index)):
        im = np.ascontiguousarray(cur_pyr[int(level[i])][int(image_index[i])])
        p = out[i].copy()
        conv[i] = bool(lib.plsvo_ref_align2d(im.ctypes.data_as(u8p), im.shape[1], im.shape[0], im.strides[0],
                                             border[i].ctypes.data_as(u8p), ref[i].ctypes.data_as(u8p), n_iter,
                                             p.ctypes.data_as(dp)))
        out[i] = p
    return conv, out


def _align1d_loop(fn, cur_pyr, image_index, level, dirs, border, ref, px, n_iter):
    u8p, dp, fp = C.POINTER(C.c_uint8), C.POINTER(C.c_double), C.POINTER(C.c_float)
    out = np.array(px, np.float64, copy=True)
    conv = np.zeros(len(image_index), bool)
    h_inv = np.zeros(len(image_index))
    border = np.ascontiguousarray(border, np.uint8)
    ref = np.ascontiguousarray(ref, np.uint8)
    dirs = np.ascontiguousarray(dirs, np.float32)
    for i in range(len(image_index)):
        im = np.ascontiguousarray(cur_pyr[int(level[i])][int(image_index[i])])
        p = out[i].copy()
        h = C.c_double(0)
        conv[i] = bool(fn(im.ctypes.data_as(u8p), im.shape[1], im.shape[0], im.strides[0], dirs[i].ctypes.data_as(fp),
                          border[i].ctypes.data_as(u8p), ref[i].ctypes.data_as(u8p), n_iter, p.ctypes.data_as(dp), C.byref(h)))
        out[i] = p
        h_inv[i] = h.value
    return conv, out, h_inv


def align1d(abi, cur_pyr, image_index, level, dirs, border, ref, px, n_iter):
    """feature_alignment::align1D restated -> (converged [n], px [n,2], h_inv [n])."""
    return _align1d_loop(load(abi).plsvo_oracle_align1d, cur_pyr, image_index, level, dirs, border, ref, px, n_iter)


def ref_align1d(abi, cur_pyr, image_index, level, dirs, border, ref, px, n_iter):
    """feature_alignment::align1D of the reference's own feature_alignment.cpp."""
    return _align1d_loop(load_ref(abi).plsvo_ref_align1d, cur_pyr, image_index, level, dirs, border, ref, px, n_iter)


def match_direct(abi, data, n_threads: int = 1):
    """Matcher::findMatchDirect restated, over a synth.MatchData batch -> abi.MatchOut."""
    lib = load(abi)
    lib.plsvo_oracle_match_direct_batch.restype = C.c_int
    lib.plsvo_oracle_match_direct_batch.argtypes = [C.POINTER(abi.MatchBatch), C.POINTER(abi.MatchResult), C.c_int]
    b, keep = abi.make_match_batch(data)
    out = abi.MatchOut(data.n)
    rc = lib.plsvo_oracle_match_direct_batch(C.byref(b), C.byref(out.struct), n_threads)
    if rc != 0:
        raise RuntimeError(f"oracle match_direct failed rc={rc}")
    return out


def ref_match_direct(abi, data):
    """Matcher::findMatchDirect of the reference's own matcher.cpp -> abi.MatchOut."""
    lib = load_ref(abi)
    lib.plsvo_ref_match_direct_batch.restype = C.c_int
    lib.plsvo_ref_match_direct_batch.argtypes = [C.POINTER(abi.MatchBatch), C.POINTER(abi.MatchResult)]
    b, keep = abi.make_match_batch(data)
    out = abi.MatchOut(data.n)
    rc = lib.plsvo_ref_match_direct_batch(C.byref(b), C.byref(out.struct))
    if rc != 0:
        raise RuntimeError(f"reference match_direct failed rc={rc}")
    return out


def structopt(abi, data, n_threads: int = 1):
    """Point::optimize / LineSeg::optimize restated -> abi.StructOptOut."""
    lib = load(abi)
    lib.plsvo_oracle_structopt_batch.restype = C.c_int
    lib.plsvo_oracle_structopt_batch.argtypes = [C.POINTER(abi.StructOptBatch), C.POINTER(abi.StructOptResult), C.c_int]
    b, keep = abi.make_structopt_batch(data)
    out = abi.StructOptOut(b.n_points, b.n_segs)
    rc = lib.plsvo_oracle_structopt_batch(C.byref(b), C.byref(out.struct), n_threads)
    if rc != 0:
        raise RuntimeError(f"oracle structopt failed rc={rc}")
    return out


def ref_structopt(abi, data):
    """Point::optimize / LineSeg::optimize of the reference's own feature3D_impl.cpp -> abi.StructOptOut (iters not observable)."""
    lib = load_ref(abi)
    lib.plsvo_ref_structopt_batch.restype = C.c_int
    lib.plsvo_ref_structopt_batch.argtypes = [C.POINTER(abi.StructOptBatch), C.POINTER(abi.StructOptResult)]
    b, keep = abi.make_structopt_batch(data)
    out = abi.StructOptOut(b.n_points, b.n_segs)
    rc = lib.plsvo_ref_structopt_batch(C.byref(b), C.byref(out.struct))
    if rc != 0:
        raise RuntimeError(f"reference structopt failed rc={rc}")
    return out


def seed_update(abi, data, n_threads: int = 1):
    """DepthFilter::updatePointSeeds body restated, over a synth.SeedData batch -> abi.SeedOut."""
    lib = load(abi)
    lib.plsvo_oracle_seed_update_batch.restype = C.c_int
    lib.plsvo_oracle_seed_update_batch.argtypes = [C.POINTER(abi.SeedBatch), C.POINTER(abi.SeedResult), C.c_int]
    b, keep = abi.make_seed_batch(data)
    out = abi.SeedOut(data.n)
    rc = lib.plsvo_oracle_seed_update_batch(C.byref(b), C.byref(out.struct), n_threads)
    if rc != 0:
        raise RuntimeError(f"oracle seed_update failed rc={rc}")
    return out


def ref_seed_update(abi, data):
    """DepthFilter::updatePointSeeds of the reference's own depth_filter.cpp + matcher.cpp -> abi.SeedOut
    (a, b, mu, sigma2 only; status is -1 where the reference erased the seed, 0 otherwise)."""
    lib = load_ref(abi)
    lib.plsvo_ref_seed_update_batch.restype = C.c_int
    lib.plsvo_ref_seed_update_batch.argtypes = [C.POINTER(abi.SeedBatch), C.POINTER(abi.SeedResult)]
    b, keep = abi.make_seed_batch(data)
    out = abi.SeedOut(data.n)
    rc = lib.plsvo_ref_seed_update_batch(C.byref(b), C.byref(out.struct))
    if rc != 0:
        raise RuntimeError(f"reference seed_update failed rc={rc}")
    return out


def line_seed_update(abi, data, n_threads: int = 1):
    """DepthFilter::updateLineSeeds body restated -> abi.LineSeedOut."""
    lib = load(abi)
    lib.plsvo_oracle_line_seed_update_batch.restype = C.c_int
    lib.plsvo_oracle_line_seed_update_batch.argtypes = [C.POINTER(abi.LineSeedBatch), C.POINTER(abi.LineSeedResult), C.c_int]
    b, keep = abi.make_line_seed_batch(data)
    out = abi.LineSeedOut(data.n)
    rc = lib.plsvo_oracle_line_seed_update_batch(C.byref(b), C.byref(out.line_struct), n_threads)
    if rc != 0:
        raise RuntimeError(f"oracle line_seed_update failed rc={rc}")
    return out


def ref_line_seed_update(abi, data):
    """DepthFilter::updateLineSeeds of the reference's own depth_filter.cpp + matcher.cpp -> abi.LineSeedOut (state only)."""
    lib = load_ref(abi)
    lib.plsvo_ref_line_seed_update_batch.restype = C.c_int
    lib.plsvo_ref_line_seed_update_batch.argtypes = [C.POINTER(abi.LineSeedBatch), C.POINTER(abi.LineSeedResult)]
    b, keep = abi.make_line_seed_batch(data)
    out = abi.LineSeedOut(data.n)
    rc = lib.plsvo_ref_line_seed_update_batch(C.byref(b), C.byref(out.line_struct))
    if rc != 0:
        raise RuntimeError(f"reference line_seed_update failed rc={rc}")
    return out


def _optimize_structure(fn, abi, data, pt_last, seg_last, max_n_pts, max_n_segs, frame_id):
    from plsvo_b200 import abi as _abi  # noqa: F401

    batch, keep = abi.make_structopt_batch(data)
    out = abi.StructOptOut(data.pt_pos.shape[0], data.seg_spos.shape[0])
    pl = np.ascontiguousarray(pt_last, dtype=np.int32).copy()
    sl = np.ascontiguousarray(seg_last, dtype=np.int32).copy()
    fn.restype = C.c_int
    fn.argtypes = [C.POINTER(abi.StructOptBatch), C.POINTER(abi.StructOptResult), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                   C.c_int, C.c_int, C.c_int]
    rc = fn(C.byref(batch), C.byref(out.struct), pl.ctypes.data_as(C.POINTER(C.c_int32)), sl.ctypes.data_as(C.POINTER(C.c_int32)),
            max_n_pts, max_n_segs, frame_id)
    if rc != 0:
        raise RuntimeError(f"optimize_structure failed rc={rc}")
    return out, pl, sl


def ref_optimize_structure(abi, data, pt_last, seg_last, max_n_pts, max_n_segs, frame_id=77):
    """FrameHandlerBase::optimizeStructure with the reference's own Point::optimize / LineSeg::optimize (oracle/_ref)."""
    return _optimize_structure(load_ref(abi).plsvo_ref_optimize_structure, abi, data, pt_last, seg_last, max_n_pts, max_n_segs, frame_id)


def shimref_optimize_structure(abi, data, pt_last, seg_last, max_n_pts, max_n_segs, frame_id=77, cpu: bool = False):
    """The same call on reference-typed objects through the shim's plsvo::b200::optimizeStructure (C ABI: GPU, or the
    oracle-backed adapter with cpu=True)."""
    return _optimize_structure(load_shimref(abi, cpu).plsvo_shimref_optimize_structure, abi, data, pt_last, seg_last, max_n_pts,
                               max_n_segs, frame_id)


# ---- reference-typed scenes for the loops either side of the hot path (oracle/next_scenes.h) ---------------------------
class SceneMatchOut(C.Structure):
    _fields_ = [("pt_found", C.POINTER(C.c_uint8)), ("pt_px", C.POINTER(C.c_double)), ("pt_level", C.POINTER(C.c_int32)),
                ("pt_A", C.POINTER(C.c_double)), ("pt_ref", C.POINTER(C.c_int32)), ("seg_found", C.POINTER(C.c_uint8)),
                ("seg_spx", C.POINTER(C.c_double)), ("seg_epx", C.POINTER(C.c_double)), ("seg_level", C.POINTER(C.c_int32)),
                ("seg_A", C.POINTER(C.c_double)), ("seg_ref", C.POINTER(C.c_int32))]


class SceneSeedOut(C.Structure):
    _fields_ = [("pt_fate", C.POINTER(C.c_int32)), ("pt_state", C.POINTER(C.c_float)), ("pt_xyz", C.POINTER(C.c_double)),
                ("pt_cb_sigma2", C.POINTER(C.c_double)), ("n_pt_marks", C.POINTER(C.c_int32)), ("pt_marks", C.POINTER(C.c_double)),
                ("seg_fate", C.POINTER(C.c_int32)), ("seg_state", C.POINTER(C.c_float)), ("seg_xyz", C.POINTER(C.c_double)),
                ("seg_cb_sigma2", C.POINTER(C.c_double)), ("n_seg_marks", C.POINTER(C.c_int32)), ("seg_marks", C.POINTER(C.c_double))]


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class _Rec:
    pass


def _match_scene(fn, abi, data, n_obs):
    n, m = data.n, data.n // 2
    r = _Rec()
    r.pt_found, r.pt_px, r.pt_level = np.zeros(n, np.uint8), np.full((n, 2), np.nan), np.full(n, -99, np.int32)
    r.pt_A, r.pt_ref = np.full((n, 4), np.nan), np.full(n, -99, np.int32)
    r.seg_found, r.seg_spx, r.seg_epx = np.zeros(m, np.uint8), np.full((m, 2), np.nan), np.full((m, 2), np.nan)
    r.seg_level, r.seg_A, r.seg_ref = np.full(m, -99, np.int32), np.full((m, 4), np.nan), np.full(m, -99, np.int32)
    out = SceneMatchOut(_p(r.pt_found, C.c_uint8), _p(r.pt_px, C.c_double), _p(r.pt_level, C.c_int32), _p(r.pt_A, C.c_double),
                        _p(r.pt_ref, C.c_int32), _p(r.seg_found, C.c_uint8), _p(r.seg_spx, C.c_double), _p(r.seg_epx, C.c_double),
                        _p(r.seg_level, C.c_int32), _p(r.seg_A, C.c_double), _p(r.seg_ref, C.c_int32))
    b, keep = abi.make_match_batch(data)
    fn.restype = C.c_int
    fn.argtypes = [C.POINTER(abi.MatchBatch), C.c_int, C.POINTER(SceneMatchOut)]
    rc = fn(C.byref(b), n_obs, C.byref(out))
    if rc != 0:
        raise RuntimeError(f"match scene failed rc={rc}")
    return r


def ref_match_scene(abi, data, n_obs=3):
    """Reprojector-style pass with the reference's own Matcher::findMatchDirect + Point/LineSeg::getCloseViewObs (oracle/_ref)
    over map points / segments observed in n_obs keyframes."""
    return _match_scene(load_ref(abi).plsvo_ref_match_scene, abi, data, n_obs)


def shimref_match_scene(abi, data, n_obs=3, cpu: bool = False):
    """The same pass answered by plsvo::b200::DirectMatcher (one C-ABI call per frame: GPU, or the oracle-backed adapter)."""
    return _match_scene(load_shimref(abi, cpu).plsvo_shimref_match_scene, abi, data, n_obs)


def _seed_scene(fn, abi, pts, lines, pt_age, seg_age, is_keyframe):
    n, m = pts.n, (lines.n if lines is not None else 0)
    r = _Rec()
    r.pt_fate, r.pt_state, r.pt_xyz = np.full(n, -1, np.int32), np.full((n, 4), np.nan, np.float32), np.full((n, 3), np.nan)
    r.pt_cb_sigma2, r.n_pt_marks, r.pt_marks = np.full(n, np.nan), np.zeros(1, np.int32), np.full((n + 1, 2), np.nan)
    r.seg_fate, r.seg_state, r.seg_xyz = np.full(m + 1, -1, np.int32), np.full((m + 1, 6), np.nan, np.float32), np.full((m + 1, 6), np.nan)
    r.seg_cb_sigma2, r.n_seg_marks, r.seg_marks = np.full((m + 1, 2), np.nan), np.zeros(1, np.int32), np.full((m + 1, 4), np.nan)
    out = SceneSeedOut(_p(r.pt_fate, C.c_int32), _p(r.pt_state, C.c_float), _p(r.pt_xyz, C.c_double), _p(r.pt_cb_sigma2, C.c_double),
                       _p(r.n_pt_marks, C.c_int32), _p(r.pt_marks, C.c_double), _p(r.seg_fate, C.c_int32), _p(r.seg_state, C.c_float),
                       _p(r.seg_xyz, C.c_double), _p(r.seg_cb_sigma2, C.c_double), _p(r.n_seg_marks, C.c_int32), _p(r.seg_marks, C.c_double))
    pb, keep = abi.make_seed_batch(pts)
    lb, lkeep = abi.make_line_seed_batch(lines) if lines is not None else (None, None)
    pa = np.ascontiguousarray(pt_age, np.int32) if pt_age is not None else None
    sa = np.ascontiguousarray(seg_age, np.int32) if seg_age is not None else None
    fn.restype = C.c_int
    fn.argtypes = [C.POINTER(abi.SeedBatch), C.POINTER(abi.LineSeedBatch), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int,
                   C.POINTER(SceneSeedOut)]
    rc = fn(C.byref(pb), C.byref(lb) if lb is not None else None, _p(pa, C.c_int32) if pa is not None else None,
            _p(sa, C.c_int32) if sa is not None else None, int(is_keyframe), C.byref(out))
    if rc != 0:
        raise RuntimeError(f"seed scene failed rc={rc}")
    r.seg_fate, r.seg_state, r.seg_xyz, r.seg_cb_sigma2 = r.seg_fate[:m], r.seg_state[:m], r.seg_xyz[:m], r.seg_cb_sigma2[:m]
    r.pt_marks, r.seg_marks = r.pt_marks[: r.n_pt_marks[0]], r.seg_marks[: r.n_seg_marks[0]]
    return r


def ref_seed_scene(abi, pts, lines=None, pt_age=None, seg_age=None, is_keyframe=False):
    """DepthFilter::updateSeeds of the reference's own depth_filter.cpp (real ageing, convergence, callbacks, detector marks)."""
    return _seed_scene(load_ref(abi).plsvo_ref_seed_scene, abi, pts, lines, pt_age, seg_age, is_keyframe)


def shimref_seed_scene(abi, pts, lines=None, pt_age=None, seg_age=None, is_keyframe=False, cpu: bool = False):
    """The same update through plsvo::b200::DepthFilterB200 (two C-ABI calls per frame: GPU, or the oracle-backed adapter)."""
    return _seed_scene(load_shimref(abi, cpu).plsvo_shimref_seed_scene, abi, pts, lines, pt_age, seg_age, is_keyframe)


def last_loop_seconds(abi, shim: bool, cpu: bool = False) -> float:
    """Wall-clock seconds of the loop under test in the last *_match_scene / *_seed_scene call (scene construction excluded)."""
    if shim:
        fn = load_shimref(abi, cpu).plsvo_shimref_last_loop_seconds
    else:
        fn = load_ref(abi).plsvo_ref_last_loop_seconds
    fn.restype = C.c_double
    fn.argtypes = []
    return float(fn())
