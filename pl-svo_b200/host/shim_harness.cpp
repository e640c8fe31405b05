// shim_harness.cpp — test-only C entry points that build Frame objects (compat types) from flat arrays
// and call the shim exactly as src/frame_handler_mono.cpp does (:272-274 and :327-329), so the Python
// tests can drive the signature-preserving path end to end.
#include <cstring>
#include <vector>

#include "plsvo_shim.h"

using namespace plsvo;

namespace {
struct Owned {
  std::vector<PointFeat> pf;
  std::vector<LineFeat> lf;
  std::vector<Point> pts;
  std::vector<LineSeg> segs;
};
template <int N, class V>
void set(V& v, const double* s) {
  for (int i = 0; i < N; ++i) v[i] = s[i];
}
Sophus::SE3 se3(const double* p) {
  Eigen::Vector3d t;
  t[0] = p[4], t[1] = p[5], t[2] = p[6];
  return Sophus::SE3(Eigen::Quaterniond(p[3], p[0], p[1], p[2]), t);
}
}  // namespace

extern "C" {

// one frame pair through plsvo::SparseImgAlign(...).run(last_frame, new_frame)
long long plsvo_shim_test_align(int width, int height, double fx, double fy, double cx, double cy, int n_levels,
                                unsigned char** ref_levels, unsigned char** cur_levels, const double* T_ref_w,
                                const double* T_cur_w, int n_pts, const double* pt_px, const double* pt_f,
                                const double* pt_pos, const unsigned char* pt_valid, int n_segs, const double* seg_spx,
                                const double* seg_epx, const double* seg_sf, const double* seg_ef, const double* seg_spos,
                                const double* seg_epos, const double* seg_length, const unsigned char* seg_valid,
                                int max_level, int min_level, int n_iter, double* T_out, unsigned char* seg_killed,
                                double* fisher36) {
  vk::PinholeCamera cam(width, height, fx, fy, cx, cy);
  FramePtr last_frame(new Frame()), new_frame(new Frame());
  last_frame->cam_ = new_frame->cam_ = &cam;
  last_frame->T_f_w_ = se3(T_ref_w);
  new_frame->T_f_w_ = se3(T_cur_w);
  for (int l = 0; l < n_levels; ++l) {
    cv::Mat r, c;
    r.cols = c.cols = width >> l, r.rows = c.rows = height >> l;
    r.step[0] = c.step[0] = (size_t)(width >> l);
    r.data = ref_levels[l], c.data = cur_levels[l];
    last_frame->img_pyr_.push_back(r);
    new_frame->img_pyr_.push_back(c);
  }
  Owned o;
  o.pf.resize(n_pts), o.pts.resize(n_pts), o.lf.resize(n_segs), o.segs.resize(n_segs);
  for (int i = 0; i < n_pts; ++i) {
    set<2>(o.pf[i].px, pt_px + 2 * i), set<3>(o.pf[i].f, pt_f + 3 * i), set<3>(o.pts[i].pos_, pt_pos + 3 * i);
    o.pf[i].feat3D = (!pt_valid || pt_valid[i]) ? &o.pts[i] : NULL;
    last_frame->pt_fts_.push_back(&o.pf[i]);
  }
  for (int j = 0; j < n_segs; ++j) {
    set<2>(o.lf[j].spx, seg_spx + 2 * j), set<2>(o.lf[j].epx, seg_epx + 2 * j);
    set<3>(o.lf[j].sf, seg_sf + 3 * j), set<3>(o.lf[j].ef, seg_ef + 3 * j);
    set<3>(o.segs[j].spos_, seg_spos + 3 * j), set<3>(o.segs[j].epos_, seg_epos + 3 * j);
    o.lf[j].length = seg_length[j];
    o.lf[j].feat3D = (!seg_valid || seg_valid[j]) ? &o.segs[j] : NULL;
    last_frame->seg_fts_.push_back(&o.lf[j]);
  }
  // ---- verbatim call pattern of FrameHandlerMono::processFrame ----
  bool display = false;
  bool verbose = false;
  SparseImgAlign img_align(max_level, min_level, n_iter, SparseImgAlign::GaussNewton, display, verbose);
  size_t img_align_n_tracked = img_align.run(last_frame, new_frame);
  // -----------------------------------------------------------------
  const auto& q = new_frame->T_f_w_.unit_quaternion();
  const auto& t = new_frame->T_f_w_.translation();
  T_out[0] = q.x(), T_out[1] = q.y(), T_out[2] = q.z(), T_out[3] = q.w(), T_out[4] = t[0], T_out[5] = t[1], T_out[6] = t[2];
  for (int j = 0; j < n_segs; ++j)
    seg_killed[j] = ((!seg_valid || seg_valid[j]) && o.lf[j].feat3D == NULL) ? 1 : 0;
  img_align.getFisherInformation(fisher36);
  return (long long)img_align_n_tracked;
}

// one frame through pose_optimizer::optimizeGaussNewton (9-arg when n_iter_ref < 0)
int plsvo_shim_test_poseopt(double fx, const double* T_f_w, int n_pts, const double* pt_f, const double* pt_pos,
                            const int* pt_level, const unsigned char* pt_valid, int n_segs, const double* seg_line,
                            const double* seg_spos, const double* seg_epos, const int* seg_level,
                            const unsigned char* seg_valid, double reproj_thresh, int n_iter, int n_iter_ref,
                            double* T_out, double* cov36, double* scalars5, unsigned char* pt_outlier,
                            unsigned char* seg_outlier) {
  vk::PinholeCamera cam(640, 480, fx, fx, 319.5, 239.5);
  FramePtr new_frame(new Frame());
  new_frame->cam_ = &cam;
  new_frame->T_f_w_ = se3(T_f_w);
  Owned o;
  o.pf.resize(n_pts), o.pts.resize(n_pts), o.lf.resize(n_segs), o.segs.resize(n_segs);
  for (int i = 0; i < n_pts; ++i) {
    set<3>(o.pf[i].f, pt_f + 3 * i), set<3>(o.pts[i].pos_, pt_pos + 3 * i);
    o.pf[i].level = pt_level[i];
    o.pf[i].feat3D = (!pt_valid || pt_valid[i]) ? &o.pts[i] : NULL;
    new_frame->pt_fts_.push_back(&o.pf[i]);
  }
  for (int j = 0; j < n_segs; ++j) {
    set<3>(o.lf[j].line, seg_line + 3 * j);
    set<3>(o.segs[j].spos_, seg_spos + 3 * j), set<3>(o.segs[j].epos_, seg_epos + 3 * j);
    o.lf[j].level = seg_level[j];
    o.lf[j].feat3D = (!seg_valid || seg_valid[j]) ? &o.segs[j] : NULL;
    new_frame->seg_fts_.push_back(&o.lf[j]);
  }
  // ---- verbatim call pattern of FrameHandlerMono::processFrame ----
  size_t sfba_n_edges_final_pt = 0, sfba_n_edges_final_ls = 0;
  double sfba_thresh = -1, sfba_error_init = -1, sfba_error_final = -1;
  if (n_iter_ref < 0)
    pose_optimizer::optimizeGaussNewton(reproj_thresh, (size_t)n_iter, false, new_frame, sfba_thresh, sfba_error_init,
                                        sfba_error_final, sfba_n_edges_final_pt, sfba_n_edges_final_ls);
  else
    pose_optimizer::optimizeGaussNewton(reproj_thresh, (size_t)n_iter, (size_t)n_iter_ref, false, new_frame, sfba_thresh,
                                        sfba_error_init, sfba_error_final, sfba_n_edges_final_pt, sfba_n_edges_final_ls);
  // -----------------------------------------------------------------
  const auto& q = new_frame->T_f_w_.unit_quaternion();
  const auto& t = new_frame->T_f_w_.translation();
  T_out[0] = q.x(), T_out[1] = q.y(), T_out[2] = q.z(), T_out[3] = q.w(), T_out[4] = t[0], T_out[5] = t[1], T_out[6] = t[2];
  for (int i = 0; i < 36; ++i) cov36[i] = new_frame->Cov_.m[i];
  scalars5[0] = sfba_thresh, scalars5[1] = sfba_error_init, scalars5[2] = sfba_error_final;
  scalars5[3] = (double)sfba_n_edges_final_pt, scalars5[4] = (double)sfba_n_edges_final_ls;
  for (int i = 0; i < n_pts; ++i) pt_outlier[i] = ((!pt_valid || pt_valid[i]) && o.pf[i].feat3D == NULL) ? 1 : 0;
  for (int j = 0; j < n_segs; ++j) seg_outlier[j] = ((!seg_valid || seg_valid[j]) && o.lf[j].feat3D == NULL) ? 1 : 0;
  return 0;
}
}
