// ref_harness.cpp — drives the reference's OWN translation units from flat batch arrays.
//
// TEST INFRASTRUCTURE, NOT THE PRODUCT (same rule as plsvo_oracle.cpp: only tests/, smoke() and
// bench.py's CPU legs may load the library this builds).
//
// oracle/Makefile target `ref` compiles, unmodified and where they lie,
//     /root/reference/src/sparse_img_align.cpp   /root/reference/src/pose_optimizer.cpp
//     /root/reference/src/feature.cpp            /root/reference/src/feature_alignment.cpp
//     /root/reference/src/matcher.cpp            /root/reference/src/config.cpp            
//     /root/reference/src/feature3D_impl.cpp     (Point::optimize / LineSeg::optimize)
//     /root/reference/src/depth_filter.cpp       (DepthFilter::updatePointSeeds)
// against the reference's own headers (/root/reference/include/plsvo/*.h) and the stand-in
// third-party headers in oracle/refdeps/ (Eigen, Sophus, rpg_vikit, OpenCV core, boost — absent
// from the image and from /root/reference), links this file, and writes oracle/_ref/libplsvo_ref.so.
// Nothing under /root/reference is copied into the repo; the .so is git-ignored and travels to
// the GPU box like any other built artefact.
//
// This file only (1) defines the handful of out-of-line members of plsvo::Frame / Point /
// LineSeg that live in reference sources we do not build (frame.cpp, feature3D.cpp: OpenCV image
// processing and map bookkeeping, not on the path), (2) turns a plsvo_align_batch /
// plsvo_poseopt_batch into Frame / PointFeat / LineFeat / Point / LineSeg objects, (3) makes the
// two calls exactly as src/frame_handler_mono.cpp:272-274 and :327-329 make them, and (4) copies
// the mutated state back out.  It is the checker for oracle/plsvo_oracle.cpp, which stays the
// self-contained restatement.

#include <plsvo/feature.h>
#include <plsvo/feature_alignment.h>
#include <plsvo/feature3D.h>
#include <plsvo/depth_filter.h>
#include <plsvo/frame.h>
#include <plsvo/matcher.h>
#include <plsvo/config.h>
#include <plsvo/pose_optimizer.h>
#include <plsvo/sparse_img_align.h>
#include <vikit/pinhole_camera.h>

#include <atomic>
#include <cstring>
#include <algorithm>
#include <deque>
#include <memory>
#include <thread>
#include <vector>

#include "../include/plsvo_b200.h"
#include "next_scenes.h"

// ---- out-of-line members the unbuilt reference sources would provide --------------------------
namespace plsvo {

int Frame::frame_counter_ = 0;

// src/frame.cpp:38-47 builds the pyramid with OpenCV; here the caller supplies the levels.
Frame::Frame(vk::AbstractCamera* cam, const cv::Mat&, double timestamp)
    : id_(0), timestamp_(timestamp), cam_(cam), key_pts_(5), is_keyframe_(false), v_kf_(NULL) {}

// src/frame.cpp:49-53: a frame owns its features.
Frame::~Frame() {
  for (PointFeat* f : pt_fts_) delete f;
  for (LineFeat* f : seg_fts_) delete f;
}

// Point / LineSeg constructors and getCloseViewObs (closest viewing direction among obs_, > 60 degrees rejected) are the
// reference's own: src/feature3D.cpp is compiled in (oracle/Makefile, REF_SRCS).

}  // namespace plsvo

namespace {

using plsvo::FramePtr;
using Eigen::Quaterniond;
using Eigen::Vector2d;
using Eigen::Vector3d;
using Sophus::SE3;

SE3 pose_from7(const double* p) { return SE3(Quaterniond(p[3], p[0], p[1], p[2]), Vector3d(p[4], p[5], p[6])); }
void pose_to7(const SE3& T, double* p) {
  const Quaterniond& q = T.unit_quaternion();
  p[0] = q.x(), p[1] = q.y(), p[2] = q.z(), p[3] = q.w();
  p[4] = T.translation()[0], p[5] = T.translation()[1], p[6] = T.translation()[2];
}
Vector3d v3(const double* p) { return Vector3d(p[0], p[1], p[2]); }
Vector2d v2(const double* p) { return Vector2d(p[0], p[1]); }

// SparseImgAlign with its protected solver state readable and GN iterations counted per level.
struct AlignProbe : plsvo::SparseImgAlign {
  int iters[PLSVO_MAX_LEVELS] = {0};
  AlignProbe(int max_level, int min_level, int n_iter)
      : plsvo::SparseImgAlign(max_level, min_level, n_iter, plsvo::SparseImgAlign::GaussNewton, false, false) {}
  void startIteration() override {
    if (level_ >= 0 && level_ < PLSVO_MAX_LEVELS) ++iters[level_];
  }
  const Eigen::Matrix<double, 6, 6>& H() const { return H_; }
};

template <class F>
void parallel_for(int n, int n_threads, F&& body) {
  if (n_threads <= 1 || n <= 1) {
    for (int i = 0; i < n; ++i) body(i);
    return;
  }
  std::atomic<int> next(0);
  std::vector<std::thread> pool;
  for (int t = 0; t < std::min(n_threads, n); ++t)
    pool.emplace_back([&] {
      for (int i = next++; i < n; i = next++) body(i);
    });
  for (auto& t : pool) t.join();
}

void align_one(const plsvo_align_batch* B, const plsvo_align_params* P, const plsvo_align_result* out, int b) {
  const int np = B->pt_count ? B->pt_count[b] : B->n_pts;
  const int ns = B->seg_count ? B->seg_count[b] : B->n_segs;
  const size_t po = (size_t)b * B->n_pts, so = (size_t)b * B->n_segs;

  vk::PinholeCamera cam(B->cam.width, B->cam.height, B->cam.fx, B->cam.fy, B->cam.cx, B->cam.cy);
  FramePtr ref(new plsvo::Frame(&cam, cv::Mat(), 0.0));
  FramePtr cur(new plsvo::Frame(&cam, cv::Mat(), 1.0));
  ref->img_pyr_.resize(P->max_level + 1);
  cur->img_pyr_.resize(P->max_level + 1);
  for (int l = P->min_level; l <= P->max_level; ++l) {
    const int cols = B->cam.width >> l, rows = B->cam.height >> l;
    ref->img_pyr_[l] = cv::Mat(rows, cols, CV_8U, const_cast<uint8_t*>(B->ref_img[l] + (size_t)b * B->img_stride[l]), B->img_pitch[l]);
    cur->img_pyr_[l] = cv::Mat(rows, cols, CV_8U, const_cast<uint8_t*>(B->cur_img[l] + (size_t)b * B->img_stride[l]), B->img_pitch[l]);
  }
  ref->T_f_w_ = pose_from7(B->T_ref_w + 7 * (size_t)b);
  cur->T_f_w_ = pose_from7(B->T_cur_w + 7 * (size_t)b);

  std::vector<std::unique_ptr<plsvo::Point>> points;
  std::vector<std::unique_ptr<plsvo::LineSeg>> lines;
  for (int i = 0; i < np; ++i) {
    const bool valid = !B->pt_valid || B->pt_valid[po + i];
    plsvo::Point* p3 = NULL;
    if (valid) {
      points.emplace_back(new plsvo::Point(v3(B->pt_pos + 3 * (po + i))));
      p3 = points.back().get();
    }
    ref->pt_fts_.push_back(new plsvo::PointFeat(ref.get(), p3, v2(B->pt_px + 2 * (po + i)), v3(B->pt_f + 3 * (po + i)), 0));
  }
  std::vector<plsvo::LineFeat*> segs;
  for (int j = 0; j < ns; ++j) {
    const bool valid = !B->seg_valid || B->seg_valid[so + j];
    plsvo::LineSeg* l3 = NULL;
    if (valid) {
      lines.emplace_back(new plsvo::LineSeg(v3(B->seg_spos + 3 * (so + j)), v3(B->seg_epos + 3 * (so + j))));
      l3 = lines.back().get();
    }
    plsvo::LineFeat* f = new plsvo::LineFeat(ref.get(), l3, v2(B->seg_spx + 2 * (so + j)), v2(B->seg_epx + 2 * (so + j)),
                                             v3(B->seg_sf + 3 * (so + j)), v3(B->seg_ef + 3 * (so + j)), 0);
    f->length = B->seg_length[so + j];  // the ABI carries LineFeat::length explicitly
    ref->seg_fts_.push_back(f);
    segs.push_back(f);
  }

  // src/frame_handler_mono.cpp:272-274
  AlignProbe img_align(P->max_level, P->min_level, P->n_iter);
  img_align.eps_ = P->eps;  // the reference hard-codes 1e-6 (sparse_img_align.cpp:51); the ABI carries it
  const bool empty = (np == 0 && ns == 0);
  const size_t n_tracked = img_align.run(ref, cur);

  if (out->T_cur_w) {
    if (empty)
      std::memcpy(out->T_cur_w + 7 * (size_t)b, B->T_cur_w + 7 * (size_t)b, 7 * sizeof(double));
    else
      pose_to7(cur->T_f_w_, out->T_cur_w + 7 * (size_t)b);
  }
  if (out->n_tracked) out->n_tracked[b] = (int64_t)n_tracked;
  if (out->H)
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < 6; ++j) out->H[36 * (size_t)b + 6 * i + j] = empty ? 0.0 : img_align.H()(i, j);
  if (out->seg_killed) {
    for (int j = 0; j < B->n_segs; ++j) out->seg_killed[so + j] = 0;
    for (int j = 0; j < ns; ++j) {
      const bool valid = !B->seg_valid || B->seg_valid[so + j];
      out->seg_killed[so + j] = (valid && segs[j]->feat3D == NULL) ? 1 : 0;
    }
  }
  if (out->iters)
    for (int l = 0; l < PLSVO_MAX_LEVELS; ++l) out->iters[(size_t)b * PLSVO_MAX_LEVELS + l] = img_align.iters[l];
  if (out->status) out->status[b] = (empty ? 1 : 0) | (img_align.stop_ ? 2 : 0);
  if (out->patch_iters) out->patch_iters[b] = 0;   // not observable from outside the reference class
  if (out->patch_levels) out->patch_levels[b] = 0;
}

void poseopt_one(const plsvo_poseopt_batch* B, const plsvo_poseopt_params* P, const plsvo_poseopt_result* out, int b) {
  const int np = B->pt_count ? B->pt_count[b] : B->n_pts;
  const int ns = B->seg_count ? B->seg_count[b] : B->n_segs;
  const size_t po = (size_t)b * B->n_pts, so = (size_t)b * B->n_segs;

  vk::PinholeCamera cam(640, 480, B->fx, B->fx, 320, 240);  // only errorMultiplier2() = |fx| is read
  FramePtr frame(new plsvo::Frame(&cam, cv::Mat(), 0.0));
  frame->T_f_w_ = pose_from7(B->T_f_w + 7 * (size_t)b);
  frame->Cov_.setZero();

  std::vector<std::unique_ptr<plsvo::Point>> points;
  std::vector<std::unique_ptr<plsvo::LineSeg>> lines;
  std::vector<plsvo::PointFeat*> pts;
  std::vector<plsvo::LineFeat*> segs;
  int n_valid = 0;
  for (int i = 0; i < np; ++i) {
    const bool valid = !B->pt_valid || B->pt_valid[po + i];
    plsvo::Point* p3 = NULL;
    if (valid) {
      points.emplace_back(new plsvo::Point(v3(B->pt_pos + 3 * (po + i))));
      p3 = points.back().get();
      ++n_valid;
    }
    plsvo::PointFeat* f = new plsvo::PointFeat(frame.get(), p3, Vector2d(0, 0), v3(B->pt_f + 3 * (po + i)), B->pt_level[po + i]);
    frame->pt_fts_.push_back(f);
    pts.push_back(f);
  }
  for (int j = 0; j < ns; ++j) {
    const bool valid = !B->seg_valid || B->seg_valid[so + j];
    plsvo::LineSeg* l3 = NULL;
    if (valid) {
      lines.emplace_back(new plsvo::LineSeg(v3(B->seg_spos + 3 * (so + j)), v3(B->seg_epos + 3 * (so + j))));
      l3 = lines.back().get();
      ++n_valid;
    }
    // bearing vectors are placeholders: the optimiser reads only LineFeat::line, set below
    plsvo::LineFeat* f = new plsvo::LineFeat(frame.get(), l3, Vector2d(0, 0), Vector2d(1, 0), Vector3d(0, 0, 1), Vector3d(1, 0, 1),
                                             B->seg_level[so + j]);
    f->line = v3(B->seg_line + 3 * (so + j));
    frame->seg_fts_.push_back(f);
    segs.push_back(f);
  }

  double estimated_scale = 0, error_init = 0, error_final = 0;
  size_t num_obs_pt = 0, num_obs_ls = 0;
  // src/frame_handler_mono.cpp:327-329 (9-argument overload) / the exported 10-argument overload
  if (P->n_iter_ref < 0)
    plsvo::pose_optimizer::optimizeGaussNewton(P->reproj_thresh, (size_t)P->n_iter, false, frame, estimated_scale, error_init,
                                               error_final, num_obs_pt, num_obs_ls);
  else
    plsvo::pose_optimizer::optimizeGaussNewton(P->reproj_thresh, (size_t)P->n_iter, (size_t)P->n_iter_ref, false, frame,
                                               estimated_scale, error_init, error_final, num_obs_pt, num_obs_ls);

  const bool no_obs = (n_valid == 0);
  if (out->status) out->status[b] = no_obs ? 1 : 0;
  if (out->T_f_w) {
    if (no_obs)
      std::memcpy(out->T_f_w + 7 * (size_t)b, B->T_f_w + 7 * (size_t)b, 7 * sizeof(double));
    else
      pose_to7(frame->T_f_w_, out->T_f_w + 7 * (size_t)b);
  }
  if (out->cov)
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < 6; ++j) out->cov[36 * (size_t)b + 6 * i + j] = frame->Cov_(i, j);
  if (out->estimated_scale) out->estimated_scale[b] = estimated_scale;
  if (out->error_init) out->error_init[b] = error_init;
  if (out->error_final) out->error_final[b] = error_final;
  if (out->num_obs_pt) out->num_obs_pt[b] = (int64_t)num_obs_pt;
  if (out->num_obs_ls) out->num_obs_ls[b] = (int64_t)num_obs_ls;
  if (out->pt_outlier) {
    std::memset(out->pt_outlier + po, 0, B->n_pts);
    for (int i = 0; i < np; ++i) {
      const bool valid = !B->pt_valid || B->pt_valid[po + i];
      out->pt_outlier[po + i] = (valid && pts[i]->feat3D == NULL) ? 1 : 0;
    }
  }
  if (out->seg_outlier && B->n_segs) {
    std::memset(out->seg_outlier + so, 0, B->n_segs);
    for (int j = 0; j < ns; ++j) {
      const bool valid = !B->seg_valid || B->seg_valid[so + j];
      out->seg_outlier[so + j] = (valid && segs[j]->feat3D == NULL) ? 1 : 0;
    }
  }
  if (out->iters) out->iters[2 * (size_t)b] = out->iters[2 * (size_t)b + 1] = -1;  // not observable from outside
}

}  // namespace

extern "C" {

int plsvo_ref_align_batch(const plsvo_align_batch* batch, const plsvo_align_params* params, const plsvo_align_result* out,
                          int n_threads) {
  if (!batch || !params || !out) return PLSVO_ERR_INVALID;
  if (params->max_level < params->min_level || params->min_level < 0 || params->max_level >= PLSVO_MAX_LEVELS)
    return PLSVO_ERR_INVALID;
  parallel_for(batch->batch, n_threads, [&](int b) { align_one(batch, params, out, b); });
  return PLSVO_OK;
}

int plsvo_ref_poseopt_batch(const plsvo_poseopt_batch* batch, const plsvo_poseopt_params* params,
                            const plsvo_poseopt_result* out, int n_threads) {
  if (!batch || !params || !out) return PLSVO_ERR_INVALID;
  parallel_for(batch->batch, n_threads, [&](int b) { poseopt_one(batch, params, out, b); });
  return PLSVO_OK;
}

// feature_alignment::align2D / align1D (src/feature_alignment.cpp:160-290, :36-157) on one feature.
int plsvo_ref_align2d(const uint8_t* cur_img, int cols, int rows, size_t cur_step, const uint8_t* ref_patch_with_border,
                      const uint8_t* ref_patch, int n_iter, double* px) {
  cv::Mat img(rows, cols, CV_8U, const_cast<uint8_t*>(cur_img), cur_step);
  Vector2d est(px[0], px[1]);
  const bool ok = plsvo::feature_alignment::align2D(img, const_cast<uint8_t*>(ref_patch_with_border),
                                                    const_cast<uint8_t*>(ref_patch), n_iter, est);
  px[0] = est[0], px[1] = est[1];
  return ok ? 1 : 0;
}
int plsvo_ref_align1d(const uint8_t* cur_img, int cols, int rows, size_t cur_step, const float* dir,
                      const uint8_t* ref_patch_with_border, const uint8_t* ref_patch, int n_iter, double* px, double* h_inv) {
  cv::Mat img(rows, cols, CV_8U, const_cast<uint8_t*>(cur_img), cur_step);
  Vector2d est(px[0], px[1]);
  Eigen::Vector2f d(dir[0], dir[1]);
  const bool ok = plsvo::feature_alignment::align1D(img, d, const_cast<uint8_t*>(ref_patch_with_border),
                                                    const_cast<uint8_t*>(ref_patch), n_iter, est, *h_inv);
  px[0] = est[0], px[1] = est[1];
  return ok ? 1 : 0;
}

// Matcher::findMatchDirect(const Point&, const Frame&, Vector2d&) (src/matcher.cpp:159-211) per feature.
int plsvo_ref_match_direct_batch(const plsvo_match_batch* in, const plsvo_match_result* out) {
  if (!in || !out) return PLSVO_ERR_INVALID;
  plsvo::Config::nPyrLevels() = (size_t)in->n_pyr_levels;
  vk::PinholeCamera cam(in->cam.width, in->cam.height, in->cam.fx, in->cam.fy, in->cam.cx, in->cam.cy);
  auto make_frames = [&](int n, const uint8_t* const* img, const size_t* pitch, const size_t* stride, const double* T) {
    std::vector<FramePtr> frames;
    for (int r = 0; r < n; ++r) {
      FramePtr f(new plsvo::Frame(&cam, cv::Mat(), 0.0));
      f->img_pyr_.resize(PLSVO_MAX_LEVELS);
      for (int l = 0; l < PLSVO_MAX_LEVELS; ++l)
        if (img[l])
          f->img_pyr_[l] = cv::Mat(in->cam.height >> l, in->cam.width >> l, CV_8U, const_cast<uint8_t*>(img[l] + (size_t)r * stride[l]), pitch[l]);
      f->T_f_w_ = pose_from7(T + 7 * (size_t)r);
      frames.push_back(f);
    }
    return frames;
  };
  std::vector<FramePtr> refs = make_frames(in->n_ref_images, in->ref_img, in->ref_pitch, in->ref_stride, in->T_ref_w);
  std::vector<FramePtr> curs = make_frames(in->n_cur_images, in->cur_img, in->cur_pitch, in->cur_stride, in->T_cur_w);
  for (int i = 0; i < in->n_features; ++i) {
    plsvo::Frame* rf = refs[in->ref_index[i]].get();
    plsvo::Point pt(v3(in->pos + 3 * (size_t)i));
    plsvo::PointFeat ftr(rf, &pt, v2(in->ref_px + 2 * (size_t)i), v3(in->ref_f + 3 * (size_t)i), in->ref_level[i]);
    if (in->is_edgelet && in->is_edgelet[i]) {
      ftr.type = plsvo::PointFeat::EDGELET;
      ftr.grad = v2(in->ref_grad + 2 * (size_t)i);
    }
    pt.addFrameRef(&ftr);
    plsvo::Matcher matcher{};  // value-initialised: the patch buffers start zeroed (warpAffine may leave them untouched)
    matcher.search_level_ = -1;
    matcher.options_.align_max_iter = in->n_iter;
    Vector2d px(in->px_cur[2 * (size_t)i], in->px_cur[2 * (size_t)i + 1]);
    const bool ok = matcher.findMatchDirect(pt, *curs[in->cur_index[i]], px);
    out->px_cur[2 * (size_t)i] = px[0], out->px_cur[2 * (size_t)i + 1] = px[1];
    out->success[i] = ok ? 1 : 0;
    if (out->search_level) out->search_level[i] = matcher.search_level_;
    if (out->A_cur_ref && matcher.search_level_ >= 0)  // rows whose in-frame test failed stay as the caller passed them
      for (int r = 0; r < 2; ++r)
        for (int k = 0; k < 2; ++k) out->A_cur_ref[4 * (size_t)i + 2 * r + k] = matcher.A_cur_ref_(r, k);
  }
  return PLSVO_OK;
}

// Point::optimize / LineSeg::optimize (src/feature3D_impl.cpp:36-174) over CSR observation lists.
int plsvo_ref_structopt_batch(const plsvo_structopt_batch* in, const plsvo_structopt_result* out) {
  if (!in || !out) return PLSVO_ERR_INVALID;
  vk::PinholeCamera cam(640, 480, 300, 300, 320, 240);  // not read by optimize()
  std::vector<FramePtr> frames;
  for (int k = 0; k < in->n_frames; ++k) {
    FramePtr f(new plsvo::Frame(&cam, cv::Mat(), 0.0));
    f->T_f_w_ = pose_from7(in->T_f_w + 7 * (size_t)k);
    frames.push_back(f);
  }
  for (int i = 0; i < in->n_points; ++i) {
    plsvo::Point pt(v3(in->pt_pos + 3 * (size_t)i));
    std::vector<std::unique_ptr<plsvo::PointFeat>> obs;
    for (int o = in->pt_obs_begin[i + 1] - 1; o >= in->pt_obs_begin[i]; --o) {  // addFrameRef pushes to the front
      obs.emplace_back(new plsvo::PointFeat(frames[in->pt_obs_frame[o]].get(), &pt, Vector2d(0, 0), v3(in->pt_obs_f + 3 * (size_t)o), 0));
      pt.addFrameRef(obs.back().get());
    }
    pt.optimize((size_t)in->n_iter_pts);
    for (int k = 0; k < 3; ++k) out->pt_pos[3 * (size_t)i + k] = pt.pos_[k];
  }
  for (int i = 0; i < in->n_segs; ++i) {
    plsvo::LineSeg ls(v3(in->seg_spos + 3 * (size_t)i), v3(in->seg_epos + 3 * (size_t)i));
    std::vector<std::unique_ptr<plsvo::LineFeat>> obs;
    for (int o = in->seg_obs_begin[i + 1] - 1; o >= in->seg_obs_begin[i]; --o) {
      obs.emplace_back(new plsvo::LineFeat(frames[in->seg_obs_frame[o]].get(), &ls, Vector2d(0, 0), Vector2d(1, 0),
                                           v3(in->seg_obs_sf + 3 * (size_t)o), v3(in->seg_obs_ef + 3 * (size_t)o), 0));
      ls.addFrameRef(obs.back().get());
    }
    ls.optimize((size_t)in->n_iter_segs);
    for (int k = 0; k < 3; ++k) out->seg_spos[3 * (size_t)i + k] = ls.spos_[k], out->seg_epos[3 * (size_t)i + k] = ls.epos_[k];
  }
  return PLSVO_OK;
}


// Builds, from the flat CSR batch, the reference's own objects: keyframes, Points / LineSegs with their obs_ lists, and a
// current frame whose pt_fts_ / seg_fts_ reference every 3-D feature; `last` carries last_structure_optim_.
namespace {
struct StructScene {
  vk::PinholeCamera cam{640, 480, 300, 300, 320, 240};  // not read by optimize()
  std::vector<FramePtr> keyframes;
  FramePtr cur;
  std::vector<std::unique_ptr<plsvo::Point>> pts;
  std::vector<std::unique_ptr<plsvo::LineSeg>> segs;
  StructScene(const plsvo_structopt_batch* in, const int32_t* pt_last, const int32_t* seg_last, int frame_id) {
    for (int k = 0; k < in->n_frames; ++k) {
      FramePtr f(new plsvo::Frame(&cam, cv::Mat(), 0.0));
      f->T_f_w_ = pose_from7(in->T_f_w + 7 * (size_t)k);
      keyframes.push_back(f);
    }
    cur.reset(new plsvo::Frame(&cam, cv::Mat(), 1.0));
    cur->id_ = frame_id;
    for (int i = 0; i < in->n_points; ++i) {
      pts.emplace_back(new plsvo::Point(v3(in->pt_pos + 3 * (size_t)i)));
      plsvo::Point* pt = pts.back().get();
      pt->last_structure_optim_ = pt_last ? pt_last[i] : 0;
      for (int o = in->pt_obs_begin[i + 1] - 1; o >= in->pt_obs_begin[i]; --o) {  // addFrameRef pushes to the front
        plsvo::Frame* kf = keyframes[in->pt_obs_frame[o]].get();
        plsvo::PointFeat* ft = new plsvo::PointFeat(kf, pt, Vector2d(0, 0), v3(in->pt_obs_f + 3 * (size_t)o), 0);
        kf->pt_fts_.push_back(ft);  // owned by the keyframe
        pt->addFrameRef(ft);
      }
      cur->pt_fts_.push_back(new plsvo::PointFeat(cur.get(), pt, Vector2d(0, 0), Vector3d(0, 0, 1), 0));
    }
    for (int i = 0; i < in->n_segs; ++i) {
      segs.emplace_back(new plsvo::LineSeg(v3(in->seg_spos + 3 * (size_t)i), v3(in->seg_epos + 3 * (size_t)i)));
      plsvo::LineSeg* ls = segs.back().get();
      ls->last_structure_optim_ = seg_last ? seg_last[i] : 0;
      for (int o = in->seg_obs_begin[i + 1] - 1; o >= in->seg_obs_begin[i]; --o) {
        plsvo::Frame* kf = keyframes[in->seg_obs_frame[o]].get();
        plsvo::LineFeat* ft = new plsvo::LineFeat(kf, ls, Vector2d(0, 0), Vector2d(1, 0), v3(in->seg_obs_sf + 3 * (size_t)o),
                                                  v3(in->seg_obs_ef + 3 * (size_t)o), 0);
        kf->seg_fts_.push_back(ft);
        ls->addFrameRef(ft);
      }
      cur->seg_fts_.push_back(new plsvo::LineFeat(cur.get(), ls, Vector2d(0, 0), Vector2d(1, 0), Vector3d(0, 0, 1), Vector3d(0, 0, 1), 0));
    }
  }
  void read_back(const plsvo_structopt_batch* in, const plsvo_structopt_result* out, int32_t* pt_last, int32_t* seg_last) const {
    for (int i = 0; i < in->n_points; ++i) {
      for (int k = 0; k < 3; ++k) out->pt_pos[3 * (size_t)i + k] = pts[i]->pos_[k];
      if (pt_last) pt_last[i] = pts[i]->last_structure_optim_;
    }
    for (int i = 0; i < in->n_segs; ++i) {
      for (int k = 0; k < 3; ++k) out->seg_spos[3 * (size_t)i + k] = segs[i]->spos_[k], out->seg_epos[3 * (size_t)i + k] = segs[i]->epos_[k];
      if (seg_last) seg_last[i] = segs[i]->last_structure_optim_;
    }
  }
};
}  // namespace

// FrameHandlerBase::optimizeStructure (src/frame_handler_base.cpp:202-237) with the reference's own Point::optimize /
// LineSeg::optimize (src/feature3D_impl.cpp, compiled in place).  frame_handler_base.cpp itself needs the whole
// front end (detector, map, config), so its 30 lines of selection logic are restated here, line for line.
extern "C" int plsvo_ref_optimize_structure(const plsvo_structopt_batch* in, const plsvo_structopt_result* out, int32_t* pt_last,
                                            int32_t* seg_last, int max_n_pts_, int max_n_segs_, int frame_id) {
  if (!in || !out) return PLSVO_ERR_INVALID;
  StructScene sc(in, pt_last, seg_last, frame_id);
  FramePtr frame = sc.cur;
  size_t max_n_pts = (size_t)max_n_pts_, max_n_segs = (size_t)max_n_segs_;
  std::deque<plsvo::Point*> pts;  // :209-213
  for (auto it = frame->pt_fts_.begin(); it != frame->pt_fts_.end(); ++it)
    if ((*it)->feat3D != NULL) pts.push_back((*it)->feat3D);
  max_n_pts = std::min(max_n_pts, pts.size());  // :214
  std::nth_element(pts.begin(), pts.begin() + max_n_pts, pts.end(),
                   [](plsvo::Point* l, plsvo::Point* r) { return l->last_structure_optim_ < r->last_structure_optim_; });  // :215, :192-195
  for (auto it = pts.begin(); it != pts.begin() + max_n_pts; ++it) {  // :216-220
    (*it)->optimize(in->n_iter_pts);
    (*it)->last_structure_optim_ = frame->id_;
  }
  std::deque<plsvo::LineSeg*> segs;  // :222-226
  for (auto it = frame->seg_fts_.begin(); it != frame->seg_fts_.end(); ++it) {
    plsvo::LineFeat* f = static_cast<plsvo::LineFeat*>(*it);
    if (f->feat3D != NULL) segs.push_back(f->feat3D);
  }
  max_n_segs = std::min(max_n_segs, segs.size());  // :227
  std::nth_element(segs.begin(), segs.begin() + max_n_segs, segs.end(),
                   [](plsvo::LineSeg* l, plsvo::LineSeg* r) { return l->last_structure_optim_ < r->last_structure_optim_; });  // :228
  for (auto it = segs.begin(); it != segs.begin() + max_n_segs; ++it) {  // :229-233
    (*it)->optimize(in->n_iter_segs);
    (*it)->last_structure_optim_ = frame->id_;
  }
  sc.read_back(in, out, pt_last, seg_last);
  return PLSVO_OK;
}

// DepthFilter::updatePointSeeds (src/depth_filter.cpp:270-365) driven through the reference class itself: the seeds
// of one current frame are put into pt_seeds_, the protected update is called, and the mutated seeds are read back.
// Seed ageing and convergence never erase a seed here (same batch id; convergence threshold disabled) so that every
// seed's state can be read; the convergence flag is recomputed from the returned state by the caller.
namespace {
struct DepthFilterProbe : plsvo::DepthFilter {
  DepthFilterProbe(plsvo::feature_detection::DetectorPtr<plsvo::PointFeat> pd, plsvo::feature_detection::DetectorPtr<plsvo::LineFeat> ld)
      : plsvo::DepthFilter(pd, ld, [](plsvo::Point* p, double) { delete p; }, [](plsvo::LineSeg* l, double, double) { delete l; }) {}
  using plsvo::DepthFilter::pt_seeds_;
  using plsvo::DepthFilter::seg_seeds_;
  using plsvo::DepthFilter::matcher_;
  using plsvo::DepthFilter::matcherls_;
  void update(FramePtr f) { updatePointSeeds(f); }
  void update_lines(FramePtr f) { updateLineSeeds(f); }
};
}  // namespace

int plsvo_ref_seed_update_batch(const plsvo_seed_batch* in, const plsvo_seed_result* out) {
  if (!in || !out) return PLSVO_ERR_INVALID;
  plsvo::Config::nPyrLevels() = (size_t)in->n_pyr_levels;
  vk::PinholeCamera cam(in->cam.width, in->cam.height, in->cam.fx, in->cam.fy, in->cam.cx, in->cam.cy);
  auto make_frames = [&](int n, const uint8_t* const* img, const size_t* pitch, const size_t* stride, const double* T) {
    std::vector<FramePtr> frames;
    for (int r = 0; r < n; ++r) {
      FramePtr f(new plsvo::Frame(&cam, cv::Mat(), 0.0));
      f->img_pyr_.resize(PLSVO_MAX_LEVELS);
      for (int l = 0; l < PLSVO_MAX_LEVELS; ++l)
        if (img[l])
          f->img_pyr_[l] = cv::Mat(in->cam.height >> l, in->cam.width >> l, CV_8U, const_cast<uint8_t*>(img[l] + (size_t)r * stride[l]), pitch[l]);
      f->T_f_w_ = pose_from7(T + 7 * (size_t)r);
      frames.push_back(f);
    }
    return frames;
  };
  std::vector<FramePtr> refs = make_frames(in->n_ref_images, in->ref_img, in->ref_pitch, in->ref_stride, in->T_ref_w);
  std::vector<FramePtr> curs = make_frames(in->n_cur_images, in->cur_img, in->cur_pitch, in->cur_stride, in->T_cur_w);
  typedef plsvo::feature_detection::AbstractDetector<plsvo::PointFeat> VoidPt;
  typedef plsvo::feature_detection::AbstractDetector<plsvo::LineFeat> VoidLs;
  plsvo::feature_detection::DetectorPtr<plsvo::PointFeat> pd(new VoidPt(in->cam.width, in->cam.height, 25, in->n_pyr_levels));
  plsvo::feature_detection::DetectorPtr<plsvo::LineFeat> ld(new VoidLs(in->cam.width, in->cam.height, 25, in->n_pyr_levels));
  for (int c = 0; c < in->n_cur_images; ++c) {
    DepthFilterProbe df(pd, ld);
    df.options_.seed_convergence_sigma2_thresh = 1e300;  // never erase: every seed's state is read back below
    df.matcher_.options_.align_max_iter = in->n_iter;
    df.matcher_.options_.max_epi_search_steps = (size_t)in->max_epi_search_steps;
    df.matcher_.options_.align_1d = in->align_1d != 0;
    df.matcher_.options_.subpix_refinement = in->subpix_refinement != 0;
    df.matcher_.options_.epi_search_edgelet_filtering = in->epi_search_edgelet_filtering != 0;
    df.matcher_.options_.epi_search_edgelet_max_angle = in->epi_search_edgelet_max_angle;
    std::vector<std::unique_ptr<plsvo::PointFeat>> ftrs;
    std::vector<int> ids;
    for (int i = 0; i < in->n_seeds; ++i) {
      if (in->cur_index[i] != c) continue;
      ftrs.emplace_back(new plsvo::PointFeat(refs[in->ref_index[i]].get(), v2(in->ref_px + 2 * (size_t)i), v3(in->ref_f + 3 * (size_t)i),
                                             in->ref_level[i]));
      if (in->is_edgelet && in->is_edgelet[i]) {
        ftrs.back()->type = plsvo::PointFeat::EDGELET;
        ftrs.back()->grad = v2(in->ref_grad + 2 * (size_t)i);
      }
      plsvo::PointSeed seed(ftrs.back().get(), 1.0f, 1.0f);
      seed.batch_id = plsvo::Seed::batch_counter;
      seed.id = i;
      seed.a = in->a[i], seed.b = in->b[i], seed.mu = in->mu[i], seed.z_range = in->z_range[i], seed.sigma2 = in->sigma2[i];
      df.pt_seeds_.push_back(seed);
      out->status[i] = -1;  // stays -1 if the reference erased the seed (NaN search range, :355-359)
    }
    df.update(curs[c]);
    for (const plsvo::PointSeed& sd : df.pt_seeds_) {
      const int i = sd.id;
      out->a[i] = sd.a, out->b[i] = sd.b, out->mu[i] = sd.mu, out->sigma2[i] = sd.sigma2;
      out->status[i] = 0;  // state only: the reference does not report which branch a seed took
    }
  }
  return PLSVO_OK;
}

// DepthFilter::updateLineSeeds (src/depth_filter.cpp:367-471) through the reference class, as for the point seeds above.
int plsvo_ref_line_seed_update_batch(const plsvo_line_seed_batch* inl, const plsvo_line_seed_result* outl) {
  if (!inl || !outl) return PLSVO_ERR_INVALID;
  const plsvo_seed_batch* in = &inl->seeds;
  const plsvo_seed_result* out = &outl->seeds;
  plsvo::Config::nPyrLevels() = (size_t)in->n_pyr_levels;
  vk::PinholeCamera cam(in->cam.width, in->cam.height, in->cam.fx, in->cam.fy, in->cam.cx, in->cam.cy);
  auto make_frames = [&](int n, const uint8_t* const* img, const size_t* pitch, const size_t* stride, const double* T) {
    std::vector<FramePtr> frames;
    for (int r = 0; r < n; ++r) {
      FramePtr f(new plsvo::Frame(&cam, cv::Mat(), 0.0));
      f->img_pyr_.resize(PLSVO_MAX_LEVELS);
      for (int l = 0; l < PLSVO_MAX_LEVELS; ++l)
        if (img[l])
          f->img_pyr_[l] = cv::Mat(in->cam.height >> l, in->cam.width >> l, CV_8U, const_cast<uint8_t*>(img[l] + (size_t)r * stride[l]), pitch[l]);
      f->T_f_w_ = pose_from7(T + 7 * (size_t)r);
      frames.push_back(f);
    }
    return frames;
  };
  std::vector<FramePtr> refs = make_frames(in->n_ref_images, in->ref_img, in->ref_pitch, in->ref_stride, in->T_ref_w);
  std::vector<FramePtr> curs = make_frames(in->n_cur_images, in->cur_img, in->cur_pitch, in->cur_stride, in->T_cur_w);
  typedef plsvo::feature_detection::AbstractDetector<plsvo::PointFeat> VoidPt;
  typedef plsvo::feature_detection::AbstractDetector<plsvo::LineFeat> VoidLs;
  plsvo::feature_detection::DetectorPtr<plsvo::PointFeat> pd(new VoidPt(in->cam.width, in->cam.height, 25, in->n_pyr_levels));
  plsvo::feature_detection::DetectorPtr<plsvo::LineFeat> ld(new VoidLs(in->cam.width, in->cam.height, 25, in->n_pyr_levels));
  for (int c = 0; c < in->n_cur_images; ++c) {
    DepthFilterProbe df(pd, ld);
    df.options_.seed_convergence_sigma2_thresh = 1e300;
    df.matcherls_.options_.align_max_iter = in->n_iter;
    df.matcherls_.options_.max_epi_search_steps = (size_t)in->max_epi_search_steps;
    df.matcherls_.options_.align_1d = in->align_1d != 0;
    df.matcherls_.options_.subpix_refinement = in->subpix_refinement != 0;
    std::vector<std::unique_ptr<plsvo::LineFeat>> ftrs;
    for (int i = 0; i < in->n_seeds; ++i) {
      if (in->cur_index[i] != c) continue;
      const Vector2d px = v2(in->ref_px + 2 * (size_t)i);
      ftrs.emplace_back(new plsvo::LineFeat(refs[in->ref_index[i]].get(), px, px, v3(inl->ref_sf + 3 * (size_t)i),
                                            v3(inl->ref_ef + 3 * (size_t)i), in->ref_level[i]));
      ftrs.back()->px = px;                                 // base Feature fields the end-point search reads (matcher.cpp:440-447)
      ftrs.back()->f = v3(in->ref_f + 3 * (size_t)i);
      plsvo::LineSeed seed(ftrs.back().get(), 1.0f, 1.0f);
      seed.batch_id = plsvo::Seed::batch_counter;
      seed.id = i;
      seed.a = in->a[i], seed.b = in->b[i];
      seed.mu_s = in->mu[i], seed.z_range_s = in->z_range[i], seed.sigma2_s = in->sigma2[i];
      seed.mu_e = inl->mu_e[i], seed.z_range_e = inl->z_range_e[i], seed.sigma2_e = inl->sigma2_e[i];
      df.seg_seeds_.push_back(seed);
      out->status[i] = -1;
    }
    df.update_lines(curs[c]);
    for (const plsvo::LineSeed& sd : df.seg_seeds_) {
      const int i = sd.id;
      out->a[i] = sd.a, out->b[i] = sd.b, out->mu[i] = sd.mu_s, out->sigma2[i] = sd.sigma2_s;
      outl->mu_e[i] = sd.mu_e, outl->sigma2_e[i] = sd.sigma2_e;
      out->status[i] = 0;
    }
  }
  return PLSVO_OK;
}

// ---- the loops either side of the hot path on multi-observation scenes (oracle/next_scenes.h) ----
// Reprojector::refineBestCandidate -> refine (src/reprojector.cpp:236-387): one Matcher member answers every candidate of a
// frame in turn; what Reprojector::refine reads afterwards (return value, px, search_level_, A_cur_ref_, ref_ftr_) is recorded.
int plsvo_ref_match_scene(const plsvo_match_batch* in, int n_obs, const plsvo_scene_match_out* out) {
  if (!in || !out) return PLSVO_ERR_INVALID;
  plsvo::Config::nPyrLevels() = (size_t)in->n_pyr_levels;
  plsvo_scenes::MatchScene sc(in, n_obs);
  plsvo::Matcher m{};
  m.search_level_ = -1, m.ref_ftr_ = NULL;
  m.A_cur_ref_.setZero();
  m.options_.align_max_iter = in->n_iter;
  plsvo_scenes::last_loop_seconds() = 0.0;
  plsvo_scenes::LoopTimer timer;
  for (int c = 0; c < in->n_cur_images; ++c) {
    for (int i = 0; i < in->n_features; ++i) {
      if (in->cur_index[i] != c) continue;
      const size_t I = (size_t)i;
      Vector2d px(in->px_cur[2 * I], in->px_cur[2 * I + 1]);
      out->pt_found[i] = m.findMatchDirect(*sc.points[i], *sc.curs[c], px) ? 1 : 0;
      out->pt_px[2 * I] = px[0], out->pt_px[2 * I + 1] = px[1];
      out->pt_level[i] = m.search_level_;
      for (int r = 0; r < 2; ++r)
        for (int k = 0; k < 2; ++k) out->pt_A[4 * I + 2 * r + k] = m.A_cur_ref_(r, k);
      out->pt_ref[i] = m.ref_ftr_ ? plsvo_scenes::frame_slot(sc.refs, m.ref_ftr_->frame) : -1;
    }
    for (size_t j = 0; j < sc.segs.size(); ++j) {
      if (!sc.segs[j] || in->cur_index[2 * j] != c) continue;
      Vector2d spx(in->px_cur[4 * j], in->px_cur[4 * j + 1]), epx(in->px_cur[4 * j + 2], in->px_cur[4 * j + 3]);
      out->seg_found[j] = m.findMatchDirect(*sc.segs[j], *sc.curs[c], spx, epx) ? 1 : 0;
      out->seg_spx[2 * j] = spx[0], out->seg_spx[2 * j + 1] = spx[1], out->seg_epx[2 * j] = epx[0], out->seg_epx[2 * j + 1] = epx[1];
      out->seg_level[j] = m.search_level_;
      for (int r = 0; r < 2; ++r)
        for (int k = 0; k < 2; ++k) out->seg_A[4 * j + 2 * r + k] = m.A_cur_ref_(r, k);
      out->seg_ref[j] = m.ref_ftr_ ? plsvo_scenes::frame_slot(sc.refs, m.ref_ftr_->frame) : -1;
    }
  }
  return PLSVO_OK;
}

// DepthFilter::updateSeeds (src/depth_filter.cpp:262-471) — the reference class itself with its real convergence threshold,
// seed ageing, callbacks and detector marks.
namespace {
struct DepthFilterSceneProbe : plsvo::DepthFilter {
  using plsvo::DepthFilter::DepthFilter;
  using plsvo::DepthFilter::pt_seeds_;
  using plsvo::DepthFilter::seg_seeds_;
  using plsvo::DepthFilter::matcher_;
  using plsvo::DepthFilter::matcherls_;
  int update(FramePtr f) {
    updateSeeds(f);
    return PLSVO_OK;
  }
};
}  // namespace
int plsvo_ref_seed_scene(const plsvo_seed_batch* in, const plsvo_line_seed_batch* lin, const int32_t* pt_age, const int32_t* seg_age,
                         int is_keyframe, const plsvo_scene_seed_out* out) {
  return plsvo_scenes::run_seed_scene<DepthFilterSceneProbe>(in, lin, pt_age, seg_age, is_keyframe, out);
}

double plsvo_ref_last_loop_seconds(void) { return plsvo_scenes::last_loop_seconds(); }

const char* plsvo_ref_describe(void) {
  return "rubengooj/pl-svo src/{sparse_img_align,pose_optimizer,feature,feature_alignment,matcher,config,feature3D_impl,depth_filter}.cpp compiled unmodified against stand-in "
         "Eigen/Sophus/vikit/OpenCV/boost headers (oracle/refdeps)";
}
}
