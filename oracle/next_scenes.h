// next_scenes.h — TEST INFRASTRUCTURE (same rule as the other files in oracle/): reference-typed scenes for the loops
// either side of the hot path, shared by oracle/ref_harness.cpp (the reference's own Matcher / DepthFilter answer) and
// oracle/shimref_harness.cpp (plsvo::b200::DirectMatcher / DepthFilterB200 answer, i.e. the device or, in the
// oracle-backed build, the CPU oracle).  Both sides build bit-identical objects from the same flat batch, so whatever
// differs afterwards is the binding under test.
#pragma once
#include <plsvo/config.h>
#include <plsvo/depth_filter.h>
#include <plsvo/feature.h>
#include <plsvo/feature3D.h>
#include <plsvo/feature_detection.h>
#include <plsvo/frame.h>
#include <vikit/pinhole_camera.h>

#include <chrono>
#include <memory>
#include <utility>
#include <vector>

#include "../include/plsvo_b200.h"

extern "C" {
// per-candidate record of a Reprojector-style pass over n point candidates (rows of a plsvo_match_batch) and n/2 segment
// candidates (rows 2j, 2j+1 as start / end point), ONE matcher object used for all of them in order
typedef struct plsvo_scene_match_out {
  uint8_t* pt_found;   // [n]    return value of findMatchDirect
  double* pt_px;       // [n][2] px_est afterwards
  int32_t* pt_level;   // [n]    matcher.search_level_ afterwards (persists from earlier candidates when the call exits early)
  double* pt_A;        // [n][4] matcher.A_cur_ref_ afterwards
  int32_t* pt_ref;     // [n]    keyframe index of matcher.ref_ftr_ afterwards (-1: NULL)
  uint8_t* seg_found;  // [n/2]
  double* seg_spx;     // [n/2][2]
  double* seg_epx;     // [n/2][2]
  int32_t* seg_level;  // [n/2]
  double* seg_A;       // [n/2][4]
  int32_t* seg_ref;    // [n/2]
} plsvo_scene_match_out;

// what DepthFilter::updateSeeds left behind, per seed of the batch (id = row)
typedef struct plsvo_scene_seed_out {
  int32_t* pt_fate;    // [n] 0 alive, 1 converged (Point created, callback called, seed erased), 2 erased otherwise (age / NaN range)
  float* pt_state;     // [n][4] a b mu sigma2 of a live seed
  double* pt_xyz;      // [n][3] Point::pos_ handed to the callback
  double* pt_cb_sigma2;  // [n]
  int32_t* n_pt_marks;   // [1] setGridOccpuancy calls on the point detector, then their positions in call order
  double* pt_marks;      // [n][2]
  int32_t* seg_fate;   // [m]
  float* seg_state;    // [m][6] a b mu_s sigma2_s mu_e sigma2_e
  double* seg_xyz;     // [m][6] LineSeg::spos_, epos_
  double* seg_cb_sigma2;  // [m][2]
  int32_t* n_seg_marks;
  double* seg_marks;   // [m][4] spx, epx of the LineFeat handed to setGridOccpuancy
} plsvo_scene_seed_out;
}

namespace plsvo_scenes {
// wall-clock seconds the last scene driver spent in the loop under test (scene construction and read-back excluded)
inline double& last_loop_seconds() {
  static double s = 0.0;
  return s;
}
struct LoopTimer {
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  ~LoopTimer() { last_loop_seconds() += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};
using Eigen::Quaterniond;
using Eigen::Vector2d;
using Eigen::Vector3d;
using plsvo::FramePtr;

inline Sophus::SE3 pose_of7(const double* p) { return Sophus::SE3(Quaterniond(p[3], p[0], p[1], p[2]), Vector3d(p[4], p[5], p[6])); }
inline Vector3d vec3(const double* p) { return Vector3d(p[0], p[1], p[2]); }
inline Vector2d vec2(const double* p) { return Vector2d(p[0], p[1]); }

inline std::vector<FramePtr> make_frames(vk::PinholeCamera* cam, int n, const uint8_t* const* img, const size_t* pitch, const size_t* stride,
                                         const double* T, int id0) {
  std::vector<FramePtr> frames;
  for (int r = 0; r < n; ++r) {
    FramePtr f(new plsvo::Frame(cam, cv::Mat(), 0.0));
    f->id_ = id0 + r;
    f->img_pyr_.resize(PLSVO_MAX_LEVELS);
    for (int l = 0; l < PLSVO_MAX_LEVELS; ++l)
      if (img[l])
        f->img_pyr_[l] = cv::Mat(cam->height() >> l, cam->width() >> l, CV_8U, const_cast<uint8_t*>(img[l] + (size_t)r * stride[l]), pitch[l]);
    f->T_f_w_ = pose_of7(T + 7 * (size_t)r);
    frames.push_back(f);
  }
  return frames;
}

inline int frame_slot(const std::vector<FramePtr>& frames, const plsvo::Frame* f) {
  for (size_t k = 0; k < frames.size(); ++k)
    if (frames[k].get() == f) return (int)k;
  return -1;
}

// n map points, each observed in n_obs keyframes (the batch's own observation + projections into the next keyframes), and
// n/2 map segments made of consecutive rows; candidates carry the batch's px_cur estimates.
struct MatchScene {
  vk::PinholeCamera cam;
  std::vector<FramePtr> refs, curs;
  std::vector<std::unique_ptr<plsvo::Point>> points;
  std::vector<std::unique_ptr<plsvo::LineSeg>> segs;  // [n/2]; NULL where rows 2j, 2j+1 belong to different current frames
  std::vector<std::unique_ptr<plsvo::PointFeat>> pt_obs;
  std::vector<std::unique_ptr<plsvo::LineFeat>> seg_obs;
  MatchScene(const plsvo_match_batch* in, int n_obs)
      : cam(in->cam.width, in->cam.height, in->cam.fx, in->cam.fy, in->cam.cx, in->cam.cy) {
    refs = make_frames(&cam, in->n_ref_images, in->ref_img, in->ref_pitch, in->ref_stride, in->T_ref_w, 0);
    curs = make_frames(&cam, in->n_cur_images, in->cur_img, in->cur_pitch, in->cur_stride, in->T_cur_w, 1000);
    n_obs = std::max(1, std::min(n_obs, in->n_ref_images));
    for (int i = 0; i < in->n_features; ++i) {
      const size_t I = (size_t)i;
      points.emplace_back(new plsvo::Point(vec3(in->pos + 3 * I)));
      plsvo::Point* pt = points.back().get();
      std::vector<plsvo::PointFeat*> obs;
      for (int k = 0; k < n_obs; ++k) {
        plsvo::Frame* kf = refs[(in->ref_index[i] + k) % in->n_ref_images].get();
        plsvo::PointFeat* f;
        if (k == 0) {
          f = new plsvo::PointFeat(kf, pt, vec2(in->ref_px + 2 * I), vec3(in->ref_f + 3 * I), in->ref_level[i]);
          if (in->is_edgelet && in->is_edgelet[i]) f->type = plsvo::PointFeat::EDGELET, f->grad = vec2(in->ref_grad + 2 * I);
        } else {
          const Vector2d px = kf->w2c(pt->pos_);
          f = new plsvo::PointFeat(kf, pt, px, kf->c2f(px), in->ref_level[i]);
        }
        pt_obs.emplace_back(f);
        obs.push_back(f);
      }
      if (i & 1) std::reverse(obs.begin(), obs.end());  // the batch's own observation is not always the first of the list
      for (plsvo::PointFeat* f : obs) pt->addFrameRef(f);
    }
    for (int j = 0; j + 1 < in->n_features; j += 2) {
      segs.emplace_back();
      if (in->cur_index[j] != in->cur_index[j + 1]) continue;
      const size_t S = (size_t)j, E = (size_t)j + 1;
      segs.back().reset(new plsvo::LineSeg(vec3(in->pos + 3 * S), vec3(in->pos + 3 * E)));
      plsvo::LineSeg* ls = segs.back().get();
      for (int k = 0; k < n_obs; ++k) {
        plsvo::Frame* kf = refs[(in->ref_index[j] + k) % in->n_ref_images].get();
        const Vector2d spx = k == 0 ? vec2(in->ref_px + 2 * S) : Vector2d(kf->w2c(ls->spos_));
        const Vector3d sf = k == 0 ? vec3(in->ref_f + 3 * S) : Vector3d(kf->c2f(spx));
        const Vector2d epx = kf->w2c(ls->epos_);
        plsvo::LineFeat* f = new plsvo::LineFeat(kf, ls, spx, epx, sf, kf->c2f(epx), in->ref_level[j]);
        seg_obs.emplace_back(f);
        ls->addFrameRef(f);
      }
    }
  }
};

// detectors that record what DepthFilter marks on keyframes
struct MarkingPointDetector : plsvo::feature_detection::AbstractDetector<plsvo::PointFeat> {
  std::vector<Vector2d> marks;
  MarkingPointDetector(int w, int h, int levels) : AbstractDetector(w, h, 25, levels) {}
  void setGridOccpuancy(const plsvo::PointFeat& ft) override { marks.push_back(ft.px); }
};
struct MarkingLineDetector : plsvo::feature_detection::AbstractDetector<plsvo::LineFeat> {
  std::vector<std::pair<Vector2d, Vector2d>> marks;
  MarkingLineDetector(int w, int h, int levels) : AbstractDetector(w, h, 25, levels) {}
  void setGridOccpuancy(const plsvo::LineFeat& ft) override { marks.push_back(std::make_pair(ft.spx, ft.epx)); }
};

// The seeds of ONE current frame `c` of a point-seed batch and a line-seed batch, as DepthFilter keeps them.
// DF is a probe subclass of plsvo::DepthFilter or of plsvo::b200::DepthFilterB200 that exposes pt_seeds_, seg_seeds_,
// matcher_, matcherls_ and `int update(FramePtr)` = updateSeeds(frame) + status.
template <class DF>
struct SeedScene {
  vk::PinholeCamera cam;
  std::vector<FramePtr> refs, curs, lrefs;
  std::vector<std::unique_ptr<plsvo::PointFeat>> pt_ftrs;
  std::vector<std::unique_ptr<plsvo::LineFeat>> seg_ftrs;
  std::vector<std::unique_ptr<plsvo::Point>> made_points;
  std::vector<std::unique_ptr<plsvo::LineSeg>> made_segs;
  boost::shared_ptr<MarkingPointDetector> pdet;
  boost::shared_ptr<MarkingLineDetector> ldet;
  std::unique_ptr<DF> df;
  const plsvo_scene_seed_out* out;

  SeedScene(const plsvo_seed_batch* in, const plsvo_line_seed_batch* lin, const plsvo_scene_seed_out* o)
      : cam(in->cam.width, in->cam.height, in->cam.fx, in->cam.fy, in->cam.cx, in->cam.cy), out(o) {
    refs = make_frames(&cam, in->n_ref_images, in->ref_img, in->ref_pitch, in->ref_stride, in->T_ref_w, 0);
    curs = make_frames(&cam, in->n_cur_images, in->cur_img, in->cur_pitch, in->cur_stride, in->T_cur_w, 1000);
    if (lin) {
      const plsvo_seed_batch* ls = &lin->seeds;
      lrefs = make_frames(&cam, ls->n_ref_images, ls->ref_img, ls->ref_pitch, ls->ref_stride, ls->T_ref_w, 2000);
    }
  }

  // fresh filter holding the seeds whose cur_index is c; age[i] keyframes old
  void load(const plsvo_seed_batch* in, const plsvo_line_seed_batch* lin, int c, const int32_t* pt_age, const int32_t* seg_age) {
    pdet.reset(new MarkingPointDetector(in->cam.width, in->cam.height, in->n_pyr_levels));
    ldet.reset(new MarkingLineDetector(in->cam.width, in->cam.height, in->n_pyr_levels));
    SeedScene* self = this;
    df.reset(new DF(pdet, ldet,
                    [self](plsvo::Point* p, double sigma2) {
                      const int i = self->pt_id(p->obs_.front());
                      for (int k = 0; k < 3; ++k) self->out->pt_xyz[3 * (size_t)i + k] = p->pos_[k];
                      self->out->pt_cb_sigma2[i] = sigma2;
                      self->out->pt_fate[i] = 1;
                      self->made_points.emplace_back(p);
                    },
                    [self](plsvo::LineSeg* l, double s2s, double s2e) {
                      const int i = self->seg_id(l->obs_.front());
                      for (int k = 0; k < 3; ++k) self->out->seg_xyz[6 * (size_t)i + k] = l->spos_[k], self->out->seg_xyz[6 * (size_t)i + 3 + k] = l->epos_[k];
                      self->out->seg_cb_sigma2[2 * (size_t)i] = s2s, self->out->seg_cb_sigma2[2 * (size_t)i + 1] = s2e;
                      self->out->seg_fate[i] = 1;
                      self->made_segs.emplace_back(l);
                    }));
    auto set_options = [&](plsvo::Matcher::Options& mo, const plsvo_seed_batch* b) {
      mo.align_max_iter = b->n_iter, mo.max_epi_search_steps = (size_t)b->max_epi_search_steps;
      mo.align_1d = b->align_1d != 0, mo.subpix_refinement = b->subpix_refinement != 0;
      mo.epi_search_edgelet_filtering = b->epi_search_edgelet_filtering != 0;
      mo.epi_search_edgelet_max_angle = b->epi_search_edgelet_max_angle;
    };
    df->options_.seed_convergence_sigma2_thresh = in->seed_convergence_sigma2_thresh;
    set_options(df->matcher_.options_, in);
    df->matcher_.px_cur_ = Vector2d(-1.0, -1.0);  // Matcher() leaves it uninitialised; both sides start from the same value
    df->matcherls_.px_cur_ = Vector2d(-1.0, -1.0);
    pt_ftrs.clear(), seg_ftrs.clear();
    pt_ids.clear(), seg_ids.clear();
    for (int i = 0; i < in->n_seeds; ++i) {
      if (in->cur_index[i] != c) continue;
      const size_t I = (size_t)i;
      pt_ftrs.emplace_back(new plsvo::PointFeat(refs[in->ref_index[i]].get(), vec2(in->ref_px + 2 * I), vec3(in->ref_f + 3 * I), in->ref_level[i]));
      if (in->is_edgelet && in->is_edgelet[i]) pt_ftrs.back()->type = plsvo::PointFeat::EDGELET, pt_ftrs.back()->grad = vec2(in->ref_grad + 2 * I);
      plsvo::PointSeed seed(pt_ftrs.back().get(), 1.0f, 1.0f);
      seed.batch_id = plsvo::Seed::batch_counter - (pt_age ? pt_age[i] : 0);
      seed.id = i;
      seed.a = in->a[i], seed.b = in->b[i], seed.mu = in->mu[i], seed.z_range = in->z_range[i], seed.sigma2 = in->sigma2[i];
      df->pt_seeds_.push_back(seed);
      pt_ids.push_back(std::make_pair((const plsvo::Feature*)pt_ftrs.back().get(), i));
      out->pt_fate[i] = 2;  // until found alive below or reported by the callback
    }
    if (lin) {
      const plsvo_seed_batch* lb = &lin->seeds;
      set_options(df->matcherls_.options_, lb);
      for (int i = 0; i < lb->n_seeds; ++i) {
        if (lb->cur_index[i] != c) continue;
        const size_t I = (size_t)i;
        const Vector2d px = vec2(lb->ref_px + 2 * I);
        seg_ftrs.emplace_back(new plsvo::LineFeat(lrefs[lb->ref_index[i]].get(), px, px, vec3(lin->ref_sf + 3 * I), vec3(lin->ref_ef + 3 * I), lb->ref_level[i]));
        seg_ftrs.back()->px = px;  // base Feature fields the end-point search reads (matcher.cpp:440-447)
        seg_ftrs.back()->f = vec3(lb->ref_f + 3 * I);
        plsvo::LineSeed seed(seg_ftrs.back().get(), 1.0f, 1.0f);
        seed.batch_id = plsvo::Seed::batch_counter - (seg_age ? seg_age[i] : 0);
        seed.id = i;
        seed.a = lb->a[i], seed.b = lb->b[i];
        seed.mu_s = lb->mu[i], seed.z_range_s = lb->z_range[i], seed.sigma2_s = lb->sigma2[i];
        seed.mu_e = lin->mu_e[i], seed.z_range_e = lin->z_range_e[i], seed.sigma2_e = lin->sigma2_e[i];
        df->seg_seeds_.push_back(seed);
        seg_ids.push_back(std::make_pair((const plsvo::Feature*)seg_ftrs.back().get(), i));
        out->seg_fate[i] = 2;
      }
    }
  }

  // after df->update(frame): live seeds, detector marks (appended)
  void read_back() {
    for (const plsvo::PointSeed& sd : df->pt_seeds_) {
      const size_t i = (size_t)sd.id;
      out->pt_fate[i] = 0;
      out->pt_state[4 * i] = sd.a, out->pt_state[4 * i + 1] = sd.b, out->pt_state[4 * i + 2] = sd.mu, out->pt_state[4 * i + 3] = sd.sigma2;
    }
    for (const plsvo::LineSeed& sd : df->seg_seeds_) {
      const size_t i = (size_t)sd.id;
      out->seg_fate[i] = 0;
      float* s = out->seg_state + 6 * i;
      s[0] = sd.a, s[1] = sd.b, s[2] = sd.mu_s, s[3] = sd.sigma2_s, s[4] = sd.mu_e, s[5] = sd.sigma2_e;
    }
    for (const Vector2d& m : pdet->marks) {
      const size_t k = (size_t)(*out->n_pt_marks)++;
      out->pt_marks[2 * k] = m[0], out->pt_marks[2 * k + 1] = m[1];
    }
    for (const auto& m : ldet->marks) {
      const size_t k = (size_t)(*out->n_seg_marks)++;
      out->seg_marks[4 * k] = m.first[0], out->seg_marks[4 * k + 1] = m.first[1], out->seg_marks[4 * k + 2] = m.second[0],
                         out->seg_marks[4 * k + 3] = m.second[1];
    }
    // converged seeds: the created Point / LineSeg reference the seed's feature; undo that before the features go away
    for (auto& f : pt_ftrs) f->feat3D = NULL;
    for (auto& f : seg_ftrs) f->feat3D = NULL;
    made_points.clear(), made_segs.clear();
  }

  std::vector<std::pair<const plsvo::Feature*, int>> pt_ids, seg_ids;
  int pt_id(const plsvo::Feature* f) const {
    for (const auto& p : pt_ids)
      if (p.first == f) return p.second;
    return 0;
  }
  int seg_id(const plsvo::Feature* f) const {
    for (const auto& p : seg_ids)
      if (p.first == f) return p.second;
    return 0;
  }
};

// the whole scenario for either filter type: every current frame of the batch in turn, marked as keyframe or not
template <class DF>
int run_seed_scene(const plsvo_seed_batch* in, const plsvo_line_seed_batch* lin, const int32_t* pt_age, const int32_t* seg_age,
                   int is_keyframe, const plsvo_scene_seed_out* out) {
  if (!in || !out) return PLSVO_ERR_INVALID;
  if (lin && lin->seeds.n_cur_images != in->n_cur_images) return PLSVO_ERR_INVALID;
  plsvo::Config::nPyrLevels() = (size_t)in->n_pyr_levels;
  *out->n_pt_marks = 0, *out->n_seg_marks = 0;
  last_loop_seconds() = 0.0;
  SeedScene<DF> sc(in, lin, out);
  for (int c = 0; c < in->n_cur_images; ++c) {
    sc.load(in, lin, c, pt_age, seg_age);
    FramePtr frame = sc.curs[c];  // one frame serves both seed kinds: the two batches must describe the same current views
    frame->is_keyframe_ = is_keyframe != 0;
    int rc;
    {
      LoopTimer timer;
      rc = sc.df->update(frame);  // DepthFilter::updateSeeds(frame)
    }
    if (rc != PLSVO_OK) return rc;
    sc.read_back();
  }
  return PLSVO_OK;
}

}  // namespace plsvo_scenes
