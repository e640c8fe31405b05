// plsvo_shim.h — the reference's two entry points, source-compatible with their callers, over the
// B200 C ABI (include/plsvo_b200.h).
//
//   plsvo::SparseImgAlign                     replaces include/plsvo/sparse_img_align.h:46-70
//   plsvo::pose_optimizer::optimizeGaussNewton (x2) replaces include/plsvo/pose_optimizer.h:47-64
//
// src/frame_handler_mono.cpp:272-274, :327-329 and :418-420 compile against this header unchanged.
// Build modes: with -DPLSVO_SHIM_WITH_REFERENCE_HEADERS the reference's own Frame / Feature /
// Sophus / OpenCV types are used; otherwise the stand-ins of plsvo_compat.h.
#pragma once
#ifdef PLSVO_SHIM_WITH_REFERENCE_HEADERS
#include <plsvo/feature.h>
#include <plsvo/feature3D.h>
#include <plsvo/frame.h>
#include <plsvo/global.h>
#include <vikit/pinhole_camera.h>
#else
#include "plsvo_compat.h"
#endif
#include <cstddef>

struct plsvo_ctx;  // include/plsvo_b200.h

namespace plsvo {

/// Optimize the pose of the frame by minimizing the photometric error of feature patches.
class SparseImgAlign {
 public:
  enum Method { GaussNewton, LevenbergMarquardt };  // vk::NLLSSolver::Method, named at the call site
#ifdef PLSVO_SHIM_WITH_REFERENCE_HEADERS
  cv::Mat resimg_;  // public member of the reference class (display only; never filled, display is false)
#endif
  SparseImgAlign(int max_level, int min_level, int n_iter, Method method, bool display, bool verbose);
  /// returns n_meas_/16 (src/sparse_img_align.cpp:94); writes cur_frame->T_f_w_; nulls feat3D of
  /// reference-frame segments the optimisation rejects (:687-688)
  size_t run(FramePtr ref_frame, FramePtr cur_frame);
  /// H_/(5e-4*255^2) (:97-102), 36 doubles row-major
  void getFisherInformation(double out36[36]) const;
#ifdef PLSVO_SHIM_WITH_REFERENCE_HEADERS
  Eigen::Matrix<double, 6, 6> getFisherInformation();
#endif

 private:
  int max_level_, min_level_, n_iter_;
  double H_[36];
};

namespace pose_optimizer {
void optimizeGaussNewton(const double reproj_thresh, const size_t n_iter, const bool verbose, FramePtr& frame,
                         double& estimated_scale, double& error_init, double& error_final, size_t& num_obs_pt,
                         size_t& num_obs_ls);
void optimizeGaussNewton(const double reproj_thresh, const size_t n_iter, const size_t n_iter_ref, const bool verbose,
                         FramePtr& frame, double& estimated_scale, double& error_init, double& error_final,
                         size_t& num_obs_pt, size_t& num_obs_ls);
}  // namespace pose_optimizer

#ifdef PLSVO_SHIM_WITH_REFERENCE_HEADERS
namespace b200 {
/// Drop-in body of FrameHandlerBase::optimizeStructure (src/frame_handler_base.cpp:202-237; called at
/// src/frame_handler_mono.cpp:340): the same selection of the max_n_pts / max_n_segs least recently optimised 3-D
/// features of `frame` (nth_element on last_structure_optim_), then ALL their Point::optimize / LineSeg::optimize
/// calls (src/feature3D_impl.cpp:36-174) in one plsvo_structopt_batch_run instead of a loop, and the write-back of
/// pos_ / spos_ / epos_ / last_structure_optim_.  The reference function becomes the one line
///     plsvo::b200::optimizeStructure(frame, max_n_pts, max_iter, max_n_segs, max_iter_segs);
/// Returns the C-ABI status (PLSVO_OK = 0); on failure the map is left untouched.
int optimizeStructure(FramePtr frame, size_t max_n_pts, int max_iter, size_t max_n_segs, int max_iter_segs);
}  // namespace b200
#endif

/// process-wide device context used by the shim (SparseImgAlign is stack-constructed per call at
/// frame_handler_mono.cpp:272, so the device state cannot live in the object)
int shim_set_device(int device);
const char* shim_last_error();

/// The shim's process-wide device context under the shim's lock, for the shim's own translation units
/// (plsvo_shim.cpp, plsvo_shim_next.cpp): holds the lock for its lifetime; ctx() is NULL without a device.
class ShimSession {
 public:
  ShimSession();
  ~ShimSession();
  ShimSession(const ShimSession&) = delete;
  ShimSession& operator=(const ShimSession&) = delete;
  ::plsvo_ctx* ctx() const { return ctx_; }
  /// records plsvo_last_error(ctx) for shim_last_error() and prints it with `what`; returns rc
  int fail(int rc, const char* what);
  /// Page-locked scratch shared by the shim's calls (plsvo_host_alloc): what a call packs for the device goes here, so that
  /// the library's host<->device copies are asynchronous DMA instead of staged pageable copies.  reserve() makes room for
  /// `bytes` and rewinds; take() hands out 256-byte aligned pieces (NULL when the reservation is exhausted).  The
  /// contents are only valid while this session is alive.
  bool scratch_reserve(size_t bytes);
  void* scratch_take(size_t bytes);

 private:
  ::plsvo_ctx* ctx_;
};

}  // namespace plsvo
