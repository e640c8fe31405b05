"""Host-side mirror of the reference's interface for the hot path, over the C ABI.

The reference's boundary is two C++ symbols:

    plsvo::SparseImgAlign(max_level, min_level, n_iter, method, display, verbose).run(ref, cur)
        include/plsvo/sparse_img_align.h:56-70, src/sparse_img_align.cpp:40-95
    plsvo::pose_optimizer::optimizeGaussNewton(reproj_thresh, n_iter[, n_iter_ref], verbose, frame, ...)
        include/plsvo/pose_optimizer.h:47-64

This module keeps the same names and argument meaning for *batches* of frame pairs / frames held
in flat arrays (synth.AlignData / synth.PoseOptData).  The C++ shim that keeps the exact
FramePtr signatures lives in pl-svo_b200/host/.  All compute happens in libplsvo_b200.so on
the GPU; nothing here falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C

from . import abi


class PlsvoError(RuntimeError):
    pass


class Context:
    """A device context (plsvo_ctx): one CUDA device + stream + reusable device buffers."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self.lib = abi.load_library()
        h = C.c_void_p()
        rc = self.lib.plsvo_ctx_create(device, C.c_void_p(stream or 0), C.byref(h))
        if rc != abi.OK:
            msg = self.lib.plsvo_last_error(None)
            raise PlsvoError(f"plsvo_ctx_create failed rc={rc}: {msg.decode() if msg else ''}")
        self.handle = h

    def check(self, rc: int, what: str):
        if rc != abi.OK:
            msg = self.lib.plsvo_last_error(self.handle)
            raise PlsvoError(f"{what} failed rc={rc}: {msg.decode() if msg else ''}")

    @property
    def stream(self) -> int:
        return int(self.lib.plsvo_ctx_stream(self.handle) or 0)

    def sync(self):
        self.check(self.lib.plsvo_sync(self.handle), "plsvo_sync")

    def last_kernel_ms(self) -> float:
        """Device time of the kernel of the last pyramid / align2D / align1D call (plsvo_last_kernel_ms)."""
        ms = C.c_float(0)
        self.check(self.lib.plsvo_last_kernel_ms(self.handle, C.byref(ms)), "plsvo_last_kernel_ms")
        return float(ms.value)

    def launch_count(self) -> int:
        return int(self.lib.plsvo_launch_count(self.handle))

    def close(self):
        if getattr(self, "handle", None):
            self.lib.plsvo_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx: Context | None = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


class SparseImgAlign:
    """Batched counterpart of plsvo::SparseImgAlign (src/sparse_img_align.cpp:40-52)."""

    GaussNewton = 0
    LevenbergMarquardt = 1  # accepted for signature parity; the reference only ever passes GaussNewton

    def __init__(self, max_level: int, min_level: int, n_iter: int, method: int = 0,
                 display: bool = False, verbose: bool = False, ctx: Context | None = None, eps: float = 1e-6):
        if method != self.GaussNewton:
            raise PlsvoError("only Method::GaussNewton is on the hot path (frame_handler_mono.cpp:272-273)")
        self.params = abi.align_params(max_level, min_level, n_iter, eps)
        self.ctx = ctx or default_context()
        self.last = None

    # three-leg form (device-resident between legs)
    def upload(self, data):
        batch, self._keep = abi.make_align_batch(data)
        self._shape = (data.batch, data.n_segs)
        self.ctx.check(self.ctx.lib.plsvo_align_upload(self.ctx.handle, C.byref(batch)), "plsvo_align_upload")

    def launch(self):
        self.ctx.check(self.ctx.lib.plsvo_align_launch(self.ctx.handle, C.byref(self.params)), "plsvo_align_launch")

    def download(self) -> abi.AlignOut:
        out = abi.AlignOut(*self._shape)
        self.ctx.check(self.ctx.lib.plsvo_align_download(self.ctx.handle, C.byref(out.struct)), "plsvo_align_download")
        self.last = out
        return out

    def run(self, data) -> abi.AlignOut:
        """run(ref_frames, cur_frames) for a whole batch: returns poses, n_tracked (the reference's
        return value, sparse_img_align.cpp:94), H, killed-segment flags."""
        batch, keep = abi.make_align_batch(data)
        out = abi.AlignOut(data.batch, data.n_segs)
        self.ctx.check(
            self.ctx.lib.plsvo_align_batch_run(self.ctx.handle, C.byref(batch), C.byref(self.params), C.byref(out.struct)),
            "plsvo_align_batch_run",
        )
        self.last = out
        return out

    def getFisherInformation(self):
        """H_ / (5e-4 * 255^2), sparse_img_align.cpp:97-102 (per pair)."""
        if self.last is None:
            raise PlsvoError("run() has not been called")
        return self.last.H.reshape(-1, 6, 6) / (5e-4 * 255 * 255)


class pose_optimizer:
    """Namespace mirror of plsvo::pose_optimizer (include/plsvo/pose_optimizer.h:47-64)."""

    @staticmethod
    def optimizeGaussNewton(reproj_thresh: float, n_iter: int, verbose: bool, data, n_iter_ref: int | None = None,
                            ctx: Context | None = None) -> abi.PoseOptOut:
        """9-argument overload when n_iter_ref is None, 10-argument overload otherwise."""
        ctx = ctx or default_context()
        params = abi.poseopt_params(reproj_thresh, n_iter, -1 if n_iter_ref is None else n_iter_ref)
        batch, keep = abi.make_poseopt_batch(data)
        out = abi.PoseOptOut(data.batch, data.n_pts, data.n_segs)
        ctx.check(
            ctx.lib.plsvo_poseopt_batch_run(ctx.handle, C.byref(batch), C.byref(params), C.byref(out.struct)),
            "plsvo_poseopt_batch_run",
        )
        return out


def track(align_data, poseopt_data, max_level: int = 4, min_level: int = 2, n_iter: int = 30, reproj_thresh: float = 2.0,
          po_n_iter: int = 10, po_n_iter_ref: int | None = None, chained: bool = True, ctx: Context | None = None):
    """FrameHandlerMono::processFrame's two hot-path calls back to back (src/frame_handler_mono.cpp:272-274, :327-329):
    SparseImgAlign::run on every pair, then pose_optimizer::optimizeGaussNewton on every frame, the pose staying on the
    device in between (chained=True: the pose optimiser starts from the aligned pose of the same batch index).
    Returns (AlignOut, PoseOptOut)."""
    ctx = ctx or default_context()
    ap = abi.align_params(max_level, min_level, n_iter)
    pp = abi.poseopt_params(reproj_thresh, po_n_iter, -1 if po_n_iter_ref is None else po_n_iter_ref)
    ab, keep_a = abi.make_align_batch(align_data)
    pb, keep_p = abi.make_poseopt_batch(poseopt_data)
    if chained:
        pb.T_f_w = abi._f64p()
    ao = abi.AlignOut(align_data.batch, align_data.n_segs)
    po = abi.PoseOptOut(poseopt_data.batch, poseopt_data.n_pts, poseopt_data.n_segs)
    ctx.check(ctx.lib.plsvo_track_batch_run(ctx.handle, C.byref(ab), C.byref(ap), C.byref(pb), C.byref(pp), C.byref(ao.struct),
                                            C.byref(po.struct)), "plsvo_track_batch_run")
    return ao, po


def createImgPyramid(img_level_0, n_levels: int, ctx: Context | None = None):
    """Batched frame_utils::createImgPyramid (src/frame.cpp:171-180): u8 images [B,H,W] -> list of levels
    (level 0 is the input array itself), each the truncating 2x2 half-sample of the previous one."""
    import numpy as np

    ctx = ctx or default_context()
    img = np.ascontiguousarray(img_level_0, dtype=np.uint8)
    B, H, W = img.shape
    b = abi.PyramidBatch(B, W, H, n_levels, img.ctypes.data_as(C.POINTER(C.c_uint8)), img.strides[1], img.strides[0])
    r = abi.PyramidResult()
    levels = [img]
    for l in range(1, n_levels):
        out = np.empty((B, H >> l, W >> l), np.uint8)
        levels.append(out)
        r.level[l] = out.ctypes.data_as(C.POINTER(C.c_uint8))
        r.pitch[l] = out.strides[1]
        r.stride[l] = out.strides[0]
    ctx.check(ctx.lib.plsvo_pyramid_batch_run(ctx.handle, C.byref(b), C.byref(r)), "plsvo_pyramid_batch_run")
    return levels


class feature_alignment:
    """Namespace mirror of plsvo::feature_alignment (include/plsvo/feature_alignment.h:49-55)."""

    @staticmethod
    def align2D(cur_pyr, image_index, level, ref_patch_with_border, ref_patch, n_iter, cur_px_estimate,
                width: int, height: int, ctx: Context | None = None):
        """Batched align2D (src/feature_alignment.cpp:160-290): returns (converged [n] bool, px [n,2])."""
        import numpy as np

        ctx = ctx or default_context()
        image_index = np.ascontiguousarray(image_index, np.int32)
        level = np.ascontiguousarray(level, np.int32)
        border = np.ascontiguousarray(ref_patch_with_border, np.uint8)
        ref = np.ascontiguousarray(ref_patch, np.uint8)
        px = np.ascontiguousarray(cur_px_estimate, np.float64)
        b, keep = abi.make_align2d_batch(cur_pyr, image_index, level, border, ref, px, n_iter, width, height)
        out_px = np.zeros_like(px)
        conv = np.zeros(len(image_index), np.uint8)
        r = abi.Align2DResult(out_px.ctypes.data_as(C.POINTER(C.c_double)), conv.ctypes.data_as(C.POINTER(C.c_uint8)))
        ctx.check(ctx.lib.plsvo_align2d_batch_run(ctx.handle, C.byref(b), C.byref(r)), "plsvo_align2d_batch_run")
        return conv.astype(bool), out_px

    @staticmethod
    def align1D(cur_pyr, image_index, level, dir, ref_patch_with_border, ref_patch, n_iter, cur_px_estimate,
                width: int, height: int, ctx: Context | None = None):
        """Batched align1D (src/feature_alignment.cpp:40-157): returns (converged [n] bool, px [n,2], h_inv [n])."""
        import numpy as np

        ctx = ctx or default_context()
        image_index = np.ascontiguousarray(image_index, np.int32)
        level = np.ascontiguousarray(level, np.int32)
        border = np.ascontiguousarray(ref_patch_with_border, np.uint8)
        ref = np.ascontiguousarray(ref_patch, np.uint8)
        px = np.ascontiguousarray(cur_px_estimate, np.float64)
        d = np.ascontiguousarray(dir, np.float32)
        feats, keep = abi.make_align2d_batch(cur_pyr, image_index, level, border, ref, px, n_iter, width, height)
        b = abi.Align1DBatch(feats, d.ctypes.data_as(C.POINTER(C.c_float)))
        out_px = np.zeros_like(px)
        conv = np.zeros(len(image_index), np.uint8)
        h_inv = np.zeros(len(image_index), np.float64)
        r = abi.Align1DResult(out_px.ctypes.data_as(C.POINTER(C.c_double)), conv.ctypes.data_as(C.POINTER(C.c_uint8)),
                              h_inv.ctypes.data_as(C.POINTER(C.c_double)))
        ctx.check(ctx.lib.plsvo_align1d_batch_run(ctx.handle, C.byref(b), C.byref(r)), "plsvo_align1d_batch_run")
        return conv.astype(bool), out_px, h_inv


class Matcher:
    """Batched counterpart of plsvo::Matcher::findMatchDirect (include/plsvo/matcher.h:104-107, src/matcher.cpp:159-211)
    for candidates whose reference observation has already been chosen (getCloseViewObs stays host-side list logic)."""

    def __init__(self, align_max_iter: int = 10, ctx: Context | None = None):
        self.align_max_iter = align_max_iter  # Matcher::Options::align_max_iter
        self.ctx = ctx or default_context()

    def findMatchDirect(self, data) -> abi.MatchOut:
        """data: synth.MatchData-like batch -> px_cur (refined, level-0 pixels), success flags, search levels."""
        data.n_iter = self.align_max_iter
        b, keep = abi.make_match_batch(data)
        out = abi.MatchOut(data.n)
        self.ctx.check(self.ctx.lib.plsvo_match_direct_batch_run(self.ctx.handle, C.byref(b), C.byref(out.struct)),
                       "plsvo_match_direct_batch_run")
        return out


def optimizeStructure(data, ctx: Context | None = None) -> abi.StructOptOut:
    """Batched Point::optimize / LineSeg::optimize (src/feature3D_impl.cpp:36-174) as FrameHandlerBase::optimizeStructure
    (src/frame_handler_base.cpp:202-237) applies them: data = synth.StructOptData-like CSR observation lists."""
    ctx = ctx or default_context()
    b, keep = abi.make_structopt_batch(data)
    out = abi.StructOptOut(b.n_points, b.n_segs)
    ctx.check(ctx.lib.plsvo_structopt_batch_run(ctx.handle, C.byref(b), C.byref(out.struct)), "plsvo_structopt_batch_run")
    return out


class DepthFilter:
    """Batched counterpart of plsvo::DepthFilter::updatePointSeeds (src/depth_filter.cpp:270-365): visibility test,
    Matcher::findEpipolarMatchDirect, computeTau and the Gaussian x Beta update of every seed; seed ageing, point creation
    and the detector's occupancy grid stay on the host."""

    def __init__(self, ctx: Context | None = None):
        self.ctx = ctx or default_context()

    def updateLineSeeds(self, data) -> abi.LineSeedOut:
        """Batched body of DepthFilter::updateLineSeeds (src/depth_filter.cpp:367-471): data = synth.LineSeedData-like."""
        b, keep = abi.make_line_seed_batch(data)
        out = abi.LineSeedOut(data.n)
        self.ctx.check(self.ctx.lib.plsvo_line_seed_update_batch_run(self.ctx.handle, C.byref(b), C.byref(out.line_struct)),
                       "plsvo_line_seed_update_batch_run")
        return out

    def updatePointSeeds(self, data) -> abi.SeedOut:
        b, keep = abi.make_seed_batch(data)
        out = abi.SeedOut(data.n)
        self.ctx.check(self.ctx.lib.plsvo_seed_update_batch_run(self.ctx.handle, C.byref(b), C.byref(out.struct)),
                       "plsvo_seed_update_batch_run")
        return out
