// shimref_harness.cpp — the reference's OWN object model driven through the B200 shim.
//
// TEST INFRASTRUCTURE (same rule as the other files in oracle/).  `make -C oracle shimref` compiles, where they lie,
//     /root/reference/src/feature.cpp                 (PointFeat / LineFeat constructors of the reference)
// together with pl-svo_b200/host/plsvo_shim.cpp in -DPLSVO_SHIM_WITH_REFERENCE_HEADERS mode — i.e. against the reference's
// real plsvo::Frame / PointFeat / LineFeat / Point / LineSeg / Sophus::SE3 definitions, with plsvo/sparse_img_align.h and
// plsvo/pose_optimizer.h resolved to the shim through pl-svo_b200/host/overlay/ — and links the CUDA C-ABI library.
// This file builds reference-typed frames from a flat batch (exactly as oracle/ref_harness.cpp does for the CPU
// reference) and makes the two calls of src/frame_handler_mono.cpp:272-274 and :327-329; what answers them here is the
// GPU.  tests/test_gpu_shim.py compares the poses / killed segments / outliers / covariance that come back in the
// reference's own objects with the C ABI called directly and with the oracle.
#include <plsvo/feature.h>
#include <plsvo/feature3D.h>
#include <plsvo/frame.h>
#include <plsvo/pose_optimizer.h>    // -> plsvo_shim.h through the overlay
#include <plsvo/sparse_img_align.h>  // -> plsvo_shim.h through the overlay
#include <vikit/pinhole_camera.h>

#include <cstring>
#include <chrono>
#include <memory>
#include <vector>

#include "../include/plsvo_b200.h"
#include "next_scenes.h"
#include "plsvo_shim_next.h"

namespace plsvo {
int Frame::frame_counter_ = 0;
Frame::Frame(vk::AbstractCamera* cam, const cv::Mat&, double timestamp)
    : id_(0), timestamp_(timestamp), cam_(cam), key_pts_(5), is_keyframe_(false), v_kf_(NULL) {}
Frame::~Frame() {
  for (PointFeat* f : pt_fts_) delete f;
  for (LineFeat* f : seg_fts_) delete f;
}
// Point / LineSeg (constructors, getCloseViewObs, optimize) are the reference's own: src/feature3D.cpp and
// src/feature3D_impl.cpp are compiled in (oracle/Makefile, SHIMREF_REF_SRCS).
}  // namespace plsvo

namespace {
using Eigen::Quaterniond;
using Eigen::Vector2d;
using Eigen::Vector3d;
using plsvo::FramePtr;
using Sophus::SE3;
SE3 pose_from7(const double* p) { return SE3(Quaterniond(p[3], p[0], p[1], p[2]), Vector3d(p[4], p[5], p[6])); }
void pose_to7(const SE3& T, double* p) {
  const Quaterniond& q = T.unit_quaternion();
  p[0] = q.x(), p[1] = q.y(), p[2] = q.z(), p[3] = q.w();
  p[4] = T.translation()[0], p[5] = T.translation()[1], p[6] = T.translation()[2];
}
Vector3d v3(const double* p) { return Vector3d(p[0], p[1], p[2]); }
Vector2d v2(const double* p) { return Vector2d(p[0], p[1]); }
}  // namespace


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

// wall time of every SparseImgAlign::run call of the last plsvo_shimref_align_batch (B = 1 latency, tools/b1_latency.py)
static std::vector<double> g_run_seconds;
static std::vector<double> g_po_seconds;  // per frame: the optimizeGaussNewton call alone

extern "C" {

int plsvo_shimref_set_device(int device) { return plsvo::shim_set_device(device); }
int plsvo_shimref_run_seconds(double* out, int n) {
  const int m = std::min<int>(n, (int)g_run_seconds.size());
  for (int i = 0; i < m; ++i) out[i] = g_run_seconds[i];
  return m;
}

// SparseImgAlign(max, min, n_iter, GaussNewton, false, false).run(ref, cur) on reference-typed frames, pair by pair
int plsvo_shimref_align_batch(const plsvo_align_batch* B, const plsvo_align_params* P, const plsvo_align_result* out) {
  if (!B || !P || !out) return PLSVO_ERR_INVALID;
  g_run_seconds.clear();
  for (int b = 0; b < B->batch; ++b) {
    const int np = B->pt_count ? B->pt_count[b] : B->n_pts;
    const int ns = B->seg_count ? B->seg_count[b] : B->n_segs;
    const size_t po = (size_t)b * B->n_pts, so = (size_t)b * B->n_segs;
    vk::PinholeCamera cam(B->cam.width, B->cam.height, B->cam.fx, B->cam.fy, B->cam.cx, B->cam.cy);
    FramePtr ref(new plsvo::Frame(&cam, cv::Mat(), 0.0)), cur(new plsvo::Frame(&cam, cv::Mat(), 1.0));
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
    std::vector<plsvo::LineFeat*> segs;
    for (int i = 0; i < np; ++i) {
      plsvo::Point* p3 = NULL;
      if (!B->pt_valid || B->pt_valid[po + i]) {
        points.emplace_back(new plsvo::Point(v3(B->pt_pos + 3 * (po + i))));
        p3 = points.back().get();
      }
      ref->pt_fts_.push_back(new plsvo::PointFeat(ref.get(), p3, v2(B->pt_px + 2 * (po + i)), v3(B->pt_f + 3 * (po + i)), 0));
    }
    for (int j = 0; j < ns; ++j) {
      plsvo::LineSeg* l3 = NULL;
      if (!B->seg_valid || B->seg_valid[so + j]) {
        lines.emplace_back(new plsvo::LineSeg(v3(B->seg_spos + 3 * (so + j)), v3(B->seg_epos + 3 * (so + j))));
        l3 = lines.back().get();
      }
      plsvo::LineFeat* f = new plsvo::LineFeat(ref.get(), l3, v2(B->seg_spx + 2 * (so + j)), v2(B->seg_epx + 2 * (so + j)),
                                               v3(B->seg_sf + 3 * (so + j)), v3(B->seg_ef + 3 * (so + j)), 0);
      f->length = B->seg_length[so + j];
      ref->seg_fts_.push_back(f);
      segs.push_back(f);
    }
    // src/frame_handler_mono.cpp:272-274, verbatim but for the Config:: constants
    plsvo::SparseImgAlign img_align(P->max_level, P->min_level, P->n_iter, plsvo::SparseImgAlign::GaussNewton, false, false);
    const auto t_run0 = std::chrono::steady_clock::now();
    const size_t img_align_n_tracked = img_align.run(ref, cur);
    g_run_seconds.push_back(std::chrono::duration<double>(std::chrono::steady_clock::now() - t_run0).count());
    pose_to7(cur->T_f_w_, out->T_cur_w + 7 * (size_t)b);
    out->n_tracked[b] = (int64_t)img_align_n_tracked;
    if (out->H) img_align.getFisherInformation(out->H + 36 * (size_t)b);  // H / (5e-4 * 255^2)
    if (out->seg_killed) {
      for (int j = 0; j < B->n_segs; ++j) out->seg_killed[so + j] = 0;
      for (int j = 0; j < ns; ++j)
        out->seg_killed[so + j] = ((!B->seg_valid || B->seg_valid[so + j]) && segs[j]->feat3D == NULL) ? 1 : 0;
    }
  }
  return PLSVO_OK;
}

// FrameHandlerBase::optimizeStructure(frame, max_n_pts, max_iter, max_n_segs, max_iter_segs) (frame_handler_base.cpp:
// 202-237) on reference-typed objects through plsvo::b200::optimizeStructure (the shim's drop-in body): pt_last / seg_last
// = last_structure_optim_ on entry and on return, out = every feature's position afterwards (optimised or not).
int plsvo_shimref_optimize_structure(const plsvo_structopt_batch* in, const plsvo_structopt_result* out, int32_t* pt_last,
                                     int32_t* seg_last, int max_n_pts, int max_n_segs, int frame_id) {
  if (!in || !out) return PLSVO_ERR_INVALID;
  StructScene sc(in, pt_last, seg_last, frame_id);
  const int rc = plsvo::b200::optimizeStructure(sc.cur, (size_t)max_n_pts, in->n_iter_pts, (size_t)max_n_segs, in->n_iter_segs);
  if (rc != PLSVO_OK) return rc;
  sc.read_back(in, out, pt_last, seg_last);
  return PLSVO_OK;
}

// pose_optimizer::optimizeGaussNewton(thresh, n_iter[, n_iter_ref], false, frame, ...) on reference-typed frames
int plsvo_shimref_poseopt_batch(const plsvo_poseopt_batch* B, const plsvo_poseopt_params* P, const plsvo_poseopt_result* out) {
  if (!B || !P || !out) return PLSVO_ERR_INVALID;
  for (int b = 0; b < B->batch; ++b) {
    const int np = B->pt_count ? B->pt_count[b] : B->n_pts;
    const int ns = B->seg_count ? B->seg_count[b] : B->n_segs;
    const size_t po = (size_t)b * B->n_pts, so = (size_t)b * B->n_segs;
    vk::PinholeCamera cam(640, 480, B->fx, B->fx, 320, 240);
    FramePtr frame(new plsvo::Frame(&cam, cv::Mat(), 0.0));
    frame->T_f_w_ = pose_from7(B->T_f_w + 7 * (size_t)b);
    frame->Cov_.setZero();
    std::vector<std::unique_ptr<plsvo::Point>> points;
    std::vector<std::unique_ptr<plsvo::LineSeg>> lines;
    std::vector<plsvo::PointFeat*> pts;
    std::vector<plsvo::LineFeat*> segs;
    for (int i = 0; i < np; ++i) {
      plsvo::Point* p3 = NULL;
      if (!B->pt_valid || B->pt_valid[po + i]) {
        points.emplace_back(new plsvo::Point(v3(B->pt_pos + 3 * (po + i))));
        p3 = points.back().get();
      }
      pts.push_back(new plsvo::PointFeat(frame.get(), p3, Vector2d(0, 0), v3(B->pt_f + 3 * (po + i)), B->pt_level[po + i]));
      frame->pt_fts_.push_back(pts.back());
    }
    for (int j = 0; j < ns; ++j) {
      plsvo::LineSeg* l3 = NULL;
      if (!B->seg_valid || B->seg_valid[so + j]) {
        lines.emplace_back(new plsvo::LineSeg(v3(B->seg_spos + 3 * (so + j)), v3(B->seg_epos + 3 * (so + j))));
        l3 = lines.back().get();
      }
      plsvo::LineFeat* f = new plsvo::LineFeat(frame.get(), l3, Vector2d(0, 0), Vector2d(1, 0), Vector3d(0, 0, 1), Vector3d(1, 0, 1),
                                               B->seg_level[so + j]);
      f->line = v3(B->seg_line + 3 * (so + j));
      frame->seg_fts_.push_back(f);
      segs.push_back(f);
    }
    double estimated_scale = 0, error_init = 0, error_final = 0;
    size_t num_obs_pt = 0, num_obs_ls = 0;
    if (b == 0) g_po_seconds.clear();
    const auto t_po0 = std::chrono::steady_clock::now();
    if (P->n_iter_ref < 0)  // src/frame_handler_mono.cpp:327-329
      plsvo::pose_optimizer::optimizeGaussNewton(P->reproj_thresh, (size_t)P->n_iter, false, frame, estimated_scale, error_init,
                                                 error_final, num_obs_pt, num_obs_ls);
    else
      plsvo::pose_optimizer::optimizeGaussNewton(P->reproj_thresh, (size_t)P->n_iter, (size_t)P->n_iter_ref, false, frame,
                                                 estimated_scale, error_init, error_final, num_obs_pt, num_obs_ls);
    g_po_seconds.push_back(std::chrono::duration<double>(std::chrono::steady_clock::now() - t_po0).count());
    pose_to7(frame->T_f_w_, out->T_f_w + 7 * (size_t)b);
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
      for (int i = 0; i < np; ++i) out->pt_outlier[po + i] = ((!B->pt_valid || B->pt_valid[po + i]) && pts[i]->feat3D == NULL) ? 1 : 0;
    }
    if (out->seg_outlier && B->n_segs) {
      std::memset(out->seg_outlier + so, 0, B->n_segs);
      for (int j = 0; j < ns; ++j) out->seg_outlier[so + j] = ((!B->seg_valid || B->seg_valid[so + j]) && segs[j]->feat3D == NULL) ? 1 : 0;
    }
  }
  return PLSVO_OK;
}
// Reprojector-style pass (see plsvo_ref_match_scene in ref_harness.cpp) answered by plsvo::b200::DirectMatcher: every
// candidate of a frame is enqueued, ONE device call, then the same per-candidate reads in the same order.
int plsvo_shimref_match_scene(const plsvo_match_batch* in, int n_obs, const plsvo_scene_match_out* out) {
  if (!in || !out) return PLSVO_ERR_INVALID;
  plsvo::Config::nPyrLevels() = (size_t)in->n_pyr_levels;
  plsvo_scenes::MatchScene sc(in, n_obs);
  plsvo::b200::DirectMatcher m(in->n_iter);
  m.search_level_ = -1, m.ref_ftr_ = NULL;
  m.A_cur_ref_.setZero();
  plsvo_scenes::last_loop_seconds() = 0.0;
  plsvo_scenes::LoopTimer timer;
  for (int c = 0; c < in->n_cur_images; ++c) {
    m.reset(*sc.curs[c]);
    std::vector<size_t> kp(in->n_features, 0), ks(sc.segs.size(), 0);
    for (int i = 0; i < in->n_features; ++i)
      if (in->cur_index[i] == c) kp[i] = m.enqueue(sc.points[i].get(), Vector2d(in->px_cur[2 * (size_t)i], in->px_cur[2 * (size_t)i + 1]));
    for (size_t j = 0; j < sc.segs.size(); ++j)
      if (sc.segs[j] && in->cur_index[2 * j] == c)
        ks[j] = m.enqueue(sc.segs[j].get(), Vector2d(in->px_cur[4 * j], in->px_cur[4 * j + 1]), Vector2d(in->px_cur[4 * j + 2], in->px_cur[4 * j + 3]));
    const int rc = m.run();
    if (rc != PLSVO_OK) return rc;
    for (int i = 0; i < in->n_features; ++i) {
      if (in->cur_index[i] != c) continue;
      const size_t I = (size_t)i;
      Vector2d px(in->px_cur[2 * I], in->px_cur[2 * I + 1]);
      out->pt_found[i] = m.findMatchDirect(kp[i], px) ? 1 : 0;
      out->pt_px[2 * I] = px[0], out->pt_px[2 * I + 1] = px[1];
      out->pt_level[i] = m.search_level_;
      for (int r = 0; r < 2; ++r)
        for (int k = 0; k < 2; ++k) out->pt_A[4 * I + 2 * r + k] = m.A_cur_ref_(r, k);
      out->pt_ref[i] = m.ref_ftr_ ? plsvo_scenes::frame_slot(sc.refs, m.ref_ftr_->frame) : -1;
    }
    for (size_t j = 0; j < sc.segs.size(); ++j) {
      if (!sc.segs[j] || in->cur_index[2 * j] != c) continue;
      Vector2d spx(in->px_cur[4 * j], in->px_cur[4 * j + 1]), epx(in->px_cur[4 * j + 2], in->px_cur[4 * j + 3]);
      out->seg_found[j] = m.findMatchDirect(ks[j], spx, epx) ? 1 : 0;
      out->seg_spx[2 * j] = spx[0], out->seg_spx[2 * j + 1] = spx[1], out->seg_epx[2 * j] = epx[0], out->seg_epx[2 * j + 1] = epx[1];
      out->seg_level[j] = m.search_level_;
      for (int r = 0; r < 2; ++r)
        for (int k = 0; k < 2; ++k) out->seg_A[4 * j + 2 * r + k] = m.A_cur_ref_(r, k);
      out->seg_ref[j] = m.ref_ftr_ ? plsvo_scenes::frame_slot(sc.refs, m.ref_ftr_->frame) : -1;
    }
  }
  return PLSVO_OK;
}

// DepthFilter::updateSeeds answered by plsvo::b200::DepthFilterB200 (see plsvo_ref_seed_scene in ref_harness.cpp)
namespace {
struct DepthFilterB200SceneProbe : plsvo::b200::DepthFilterB200 {
  using plsvo::b200::DepthFilterB200::DepthFilterB200;
  using plsvo::DepthFilter::pt_seeds_;
  using plsvo::DepthFilter::seg_seeds_;
  using plsvo::DepthFilter::matcher_;
  using plsvo::DepthFilter::matcherls_;
  int update(FramePtr f) {
    updateSeeds(f);
    return last_status();
  }
};
}  // namespace
int plsvo_shimref_seed_scene(const plsvo_seed_batch* in, const plsvo_line_seed_batch* lin, const int32_t* pt_age, const int32_t* seg_age,
                             int is_keyframe, const plsvo_scene_seed_out* out) {
  return plsvo_scenes::run_seed_scene<DepthFilterB200SceneProbe>(in, lin, pt_age, seg_age, is_keyframe, out);
}
double plsvo_shimref_last_loop_seconds(void) { return plsvo_scenes::last_loop_seconds(); }
int plsvo_shimref_poseopt_seconds(double* out, int n) {
  const int m = std::min<int>(n, (int)g_po_seconds.size());
  for (int i = 0; i < m; ++i) out[i] = g_po_seconds[i];
  return m;
}
}
