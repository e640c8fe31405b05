/*
 * plsvo_b200.h — C ABI of the B200-native PL-SVO per-frame optimisation path.
 *
 * Two entry-point families, one per reference symbol they replace:
 *
 *   plsvo_align_*    replaces  plsvo::SparseImgAlign::run()
 *                    (reference: include/plsvo/sparse_img_align.h:56-70,
 *                     src/sparse_img_align.cpp:54-95; call sites
 *                     src/frame_handler_mono.cpp:272-274 and :418-420)
 *   plsvo_poseopt_*  replaces  plsvo::pose_optimizer::optimizeGaussNewton()
 *                    (reference: include/plsvo/pose_optimizer.h:47-64,
 *                     src/pose_optimizer.cpp:38-260 (9-arg) and :262-582 (10-arg);
 *                     call site src/frame_handler_mono.cpp:327-329)
 *
 * The reference has no FFI layer: its boundary is two C++ link-time symbols that take
 * boost::shared_ptr<Frame>.  The C++ shim in pl-svo_b200/host/ keeps those two signatures
 * and packs Frame / Feature lists into the flat arrays declared here.
 *
 * Conventions
 *   - plain C types only, caller-allocated outputs, int return codes, no exceptions.
 *   - a batch is B independent frame pairs (align) or B independent frames (pose-opt).
 *   - SE3 poses are 7 doubles {qx,qy,qz,qw,tx,ty,tz}: the unit quaternion (Eigen coeffs()
 *     order) and translation that the reference's Sophus::SE3 stores (include/plsvo/frame.h:62).
 *   - 6x6 matrices are 36 doubles (symmetric, so row/column order is immaterial).
 *   - every call is stream-ordered on the context's stream; *_download and *_batch
 *     synchronise before returning.
 *   - there is NO CPU fallback: without a CUDA device every call returns PLSVO_ERR_NO_DEVICE.
 */
#ifndef PLSVO_B200_H_
#define PLSVO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PLSVO_MAX_LEVELS 8   /* pyramid levels addressable through the ABI (reference uses 5) */
#define PLSVO_PATCH_AREA 16  /* 4x4 patch, include/plsvo/sparse_img_align.h:48-50 */

/* return codes */
#define PLSVO_OK 0
#define PLSVO_ERR_INVALID (-1)    /* bad argument / inconsistent batch description */
#define PLSVO_ERR_CUDA (-2)       /* CUDA runtime error, see plsvo_last_error() */
#define PLSVO_ERR_NO_DEVICE (-3)  /* no usable CUDA device: there is no CPU path */
#define PLSVO_ERR_STATE (-4)      /* launch/download without a prior upload */

typedef struct plsvo_ctx plsvo_ctx;

/* Undistorted pinhole camera: what vk::PinholeCamera::world2cam / errorMultiplier2 /
 * isInFrame use when the handler is given the undistorted model (app/run_pipeline.cpp:786-795). */
typedef struct plsvo_camera {
  int32_t width, height; /* level-0 image size */
  int32_t reserved0, reserved1;
  double fx, fy, cx, cy;
} plsvo_camera;

/* ------------------------------------------------------------------------------------------
 * Sparse image alignment
 * ---------------------------------------------------------------------------------------- */

/* SparseImgAlign constructor arguments (src/sparse_img_align.cpp:40-52); defaults at the call
 * site are max_level=4, min_level=2, n_iter=30 (src/frame_handler_mono.cpp:272-273,
 * src/config.cpp:98-99); eps is hard-coded 1e-6 in the reference (:51). */
typedef struct plsvo_align_params {
  int32_t max_level;
  int32_t min_level;
  int32_t n_iter;
  int32_t reserved;
  double eps;
} plsvo_align_params;

/* One batch of B frame pairs.  All pairs share the camera and the array strides n_pts/n_segs;
 * per-pair feature counts may be smaller (pt_count/seg_count) and individual features may be
 * flagged invalid (feat3D == NULL in the reference).
 *
 * Images: for pyramid level l in [min_level,max_level], image of pair b starts at
 * ref_img[l] + b*img_stride[l] with row pitch img_pitch[l] bytes and (width>>l) x (height>>l)
 * u8 pixels (Frame::img_pyr_, include/plsvo/frame.h:64).  Levels outside the range may be NULL.
 * Levels ABOVE the lowest provided one may also be NULL inside the range: they are then derived on the device by
 * repeated vk::halfSample (the truncating 2x2 mean of frame_utils::createImgPyramid, src/frame.cpp:171-180 —
 * bit-identical to the host pyramid), which saves their host->device copy; this needs 16-byte aligned, 16-byte
 * pitched rows of the source level.
 *
 * The reference only ever uses the distance of a 3-D feature from the reference camera centre
 * (`(pos_ - ref_pos).norm()`, sparse_img_align.cpp:229,337-340).  A caller that already holds these distances may pass
 * them in pt_depth / seg_sdepth / seg_edepth and leave pt_pos / seg_spos / seg_epos NULL (16 bytes less per point,
 * 32 per segment over PCIe).
 *
 * pt_f / seg_sf / seg_ef may be NULL when the camera is the undistorted pinhole `cam`: the bearing vectors are then
 * formed on the device exactly as the reference's feature constructors form them, `cam_->cam2world(px)`
 * (src/feature.cpp:42,98-99; vk::PinholeCamera::cam2world = ((u-cx)/fx, (v-cy)/fy, 1).normalized()).  24 bytes less
 * per point, 48 per segment.
 *
 * Frame chains (flags & PLSVO_ALIGN_FRAME_CHAIN).  FrameHandlerMono aligns consecutive frames: the current frame of
 * one call is the reference frame of the next (src/frame_handler_mono.cpp:272 `run(last_frame_, new_frame_)`, :176
 * `last_frame_ = new_frame_`).  A batch that replays such a sequence — pair b = (frame b, frame b+1) — gives
 * ref_img[l] as a stack of B+1 frames (frame k at ref_img[l] + k*img_stride[l]) and leaves cur_img[l] NULL: every
 * frame crosses the host link and sits in device memory once instead of twice (B+1 frames instead of 2B).  Everything
 * else (poses, the features of each pair's reference frame, outputs) is per pair as before; results are bit-identical
 * to the same batch given as two stacks.
 */
#define PLSVO_ALIGN_FRAME_CHAIN 1 /* plsvo_align_batch.flags: ref_img holds B+1 chained frames, cur_img is ignored */

typedef struct plsvo_align_batch {
  int32_t batch;   /* B */
  int32_t n_pts;   /* array stride: points per pair   (Frame::pt_fts_,  frame.h:65) */
  int32_t n_segs;  /* array stride: segments per pair (Frame::seg_fts_, frame.h:66) */
  int32_t flags;   /* 0 or PLSVO_ALIGN_FRAME_CHAIN (was `reserved`: zero keeps the two-stack layout) */
  plsvo_camera cam;

  const uint8_t* ref_img[PLSVO_MAX_LEVELS];
  const uint8_t* cur_img[PLSVO_MAX_LEVELS];
  size_t img_pitch[PLSVO_MAX_LEVELS];
  size_t img_stride[PLSVO_MAX_LEVELS];

  const double* T_ref_w; /* [B][7]  ref_frame->T_f_w_ */
  const double* T_cur_w; /* [B][7]  cur_frame->T_f_w_ on entry (initial guess) */

  const int32_t* pt_count;  /* [B] or NULL (= n_pts)  : pt_fts_.size() */
  const double* pt_px;      /* [B][n_pts][2]  PointFeat::px  (level-0 pixels) */
  const double* pt_f;       /* [B][n_pts][3]  PointFeat::f   (unit bearing), or NULL = cam2world(px) */
  const double* pt_pos;     /* [B][n_pts][3]  PointFeat::feat3D->pos_ (world) */
  const uint8_t* pt_valid;  /* [B][n_pts] or NULL (= all valid): feat3D != NULL */

  const int32_t* seg_count; /* [B] or NULL (= n_segs) : seg_fts_.size() */
  const double* seg_spx;    /* [B][n_segs][2] LineFeat::spx */
  const double* seg_epx;    /* [B][n_segs][2] LineFeat::epx */
  const double* seg_sf;     /* [B][n_segs][3] LineFeat::sf, or NULL = cam2world(spx) */
  const double* seg_ef;     /* [B][n_segs][3] LineFeat::ef, or NULL = cam2world(epx) */
  const double* seg_spos;   /* [B][n_segs][3] LineFeat::feat3D->spos_ */
  const double* seg_epos;   /* [B][n_segs][3] LineFeat::feat3D->epos_ */
  const double* seg_length; /* [B][n_segs]    LineFeat::length */
  const uint8_t* seg_valid; /* [B][n_segs] or NULL: feat3D != NULL */

  const double* pt_depth;   /* [B][n_pts]  or NULL: |pos_ - ref_frame->pos()| (then pt_pos may be NULL) */
  const double* seg_sdepth; /* [B][n_segs] or NULL: |spos_ - ref_frame->pos()| (then seg_spos may be NULL) */
  const double* seg_edepth; /* [B][n_segs] or NULL: |epos_ - ref_frame->pos()| (then seg_epos may be NULL) */
} plsvo_align_batch;

/* Caller-allocated outputs; any pointer may be NULL to skip that output. */
typedef struct plsvo_align_result {
  double* T_cur_w;        /* [B][7]  cur_frame->T_f_w_ on return (sparse_img_align.cpp:92) */
  int64_t* n_tracked;     /* [B]     return value of run(): n_meas_/16 (:94) */
  double* H;              /* [B][36] H_ of the last evaluated iteration (getFisherInformation, :97-102) */
  uint8_t* seg_killed;    /* [B][n_segs] 1 where the reference sets ref seg feat3D = NULL (:687-688) */
  int32_t* iters;         /* [B][PLSVO_MAX_LEVELS] residual passes executed at each level (index = level) */
  int32_t* status;        /* [B] bit0: early-out "no features" (:58-62); bit1: solver stop_ was raised */
  uint32_t* patch_iters;  /* [B] sum over executed passes of patches evaluated (roofline accounting) */
  uint32_t* patch_levels; /* [B] sum over levels of patches precomputed (roofline accounting) */
} plsvo_align_result;

/* ------------------------------------------------------------------------------------------
 * Pose optimiser
 * ---------------------------------------------------------------------------------------- */

/* Arguments of pose_optimizer::optimizeGaussNewton (pose_optimizer.h:47-64); defaults
 * reproj_thresh=2.0, n_iter=10, n_iter_ref=3 (src/config.cpp:102-104).  n_iter_ref < 0
 * selects the 9-argument overload (no refinement loop). */
typedef struct plsvo_poseopt_params {
  double reproj_thresh;
  int32_t n_iter;
  int32_t n_iter_ref;
} plsvo_poseopt_params;

typedef struct plsvo_poseopt_batch {
  int32_t batch;   /* B frames */
  int32_t n_pts;   /* array stride */
  int32_t n_segs;  /* array stride */
  int32_t reserved;
  double fx;       /* frame->cam_->errorMultiplier2() */

  const double* T_f_w;      /* [B][7] frame->T_f_w_ on entry */

  const int32_t* pt_count;  /* [B] or NULL */
  const double* pt_f;       /* [B][n_pts][3]  PointFeat::f */
  const double* pt_pos;     /* [B][n_pts][3]  feat3D->pos_ */
  const int32_t* pt_level;  /* [B][n_pts]     Feature::level */
  const uint8_t* pt_valid;  /* [B][n_pts] or NULL */

  const int32_t* seg_count; /* [B] or NULL */
  const double* seg_line;   /* [B][n_segs][3] LineFeat::line */
  const double* seg_spos;   /* [B][n_segs][3] feat3D->spos_ */
  const double* seg_epos;   /* [B][n_segs][3] feat3D->epos_ */
  const int32_t* seg_level; /* [B][n_segs] */
  const uint8_t* seg_valid; /* [B][n_segs] or NULL */
} plsvo_poseopt_batch;

typedef struct plsvo_poseopt_result {
  double* T_f_w;            /* [B][7]  frame->T_f_w_ on return */
  double* cov;              /* [B][36] frame->Cov_ (pose_optimizer.cpp:199) */
  double* estimated_scale;  /* [B] */
  double* error_init;       /* [B] */
  double* error_final;      /* [B] */
  int64_t* num_obs_pt;      /* [B] */
  int64_t* num_obs_ls;      /* [B] */
  uint8_t* pt_outlier;      /* [B][n_pts]  1 where the reference sets feat3D = NULL (:218) */
  uint8_t* seg_outlier;     /* [B][n_segs] (:239) */
  int32_t* iters;           /* [B][2] GN passes executed in the main / refinement loop */
  int32_t* status;          /* [B] bit0: early return "no observations" (:88-89), outputs untouched */
} plsvo_poseopt_result;

/* ------------------------------------------------------------------------------------------
 * Context, memory, execution
 * ---------------------------------------------------------------------------------------- */

/* device: CUDA ordinal.  stream: a cudaStream_t to run on, or NULL to create a private one. */
int plsvo_ctx_create(int device, void* stream, plsvo_ctx** out);
void plsvo_ctx_destroy(plsvo_ctx* ctx);
/* last error text of this context (or of ctx creation when ctx == NULL) */
const char* plsvo_last_error(const plsvo_ctx* ctx);
/* the cudaStream_t all work of this context is ordered on */
void* plsvo_ctx_stream(plsvo_ctx* ctx);
int plsvo_sync(plsvo_ctx* ctx);

/* page-locked host memory for batch arrays (keeps the H2D/D2H legs at PCIe speed) */
int plsvo_host_alloc(void** ptr, size_t bytes);
int plsvo_host_free(void* ptr);

/* Alignment.  upload: host arrays -> device layout (async).  launch: the whole coarse-to-fine
 * optimisation of every pair, device-resident in and out (async).  download: device -> host
 * outputs, then synchronise.  plsvo_align_batch_run = upload + launch + download.
 * Every one-call form (plsvo_*_batch_run) has finished with the caller's arrays when it returns, whatever it returns:
 * an error found after copies had been queued first drains every stream of the context. */
int plsvo_align_upload(plsvo_ctx* ctx, const plsvo_align_batch* batch);
int plsvo_align_launch(plsvo_ctx* ctx, const plsvo_align_params* params);
int plsvo_align_download(plsvo_ctx* ctx, const plsvo_align_result* out);
int plsvo_align_batch_run(plsvo_ctx* ctx, const plsvo_align_batch* batch,
                          const plsvo_align_params* params, const plsvo_align_result* out);

/* Pose optimiser, same three legs. */
int plsvo_poseopt_upload(plsvo_ctx* ctx, const plsvo_poseopt_batch* batch);
int plsvo_poseopt_launch(plsvo_ctx* ctx, const plsvo_poseopt_params* params);
int plsvo_poseopt_download(plsvo_ctx* ctx, const plsvo_poseopt_result* out);
int plsvo_poseopt_batch_run(plsvo_ctx* ctx, const plsvo_poseopt_batch* batch,
                            const plsvo_poseopt_params* params, const plsvo_poseopt_result* out);

/* The two hot-path calls of FrameHandlerMono::processFrame back to back (src/frame_handler_mono.cpp:272-274 sparse
 * image alignment, :327-329 pose optimisation) for a batch of frames, with the pose staying on the device in between:
 * frame b of `po_batch` starts from the alignment result of pair b of `al_batch` when po_batch->T_f_w is NULL
 * (otherwise from the poses given).  `al_out` may be NULL.  Equivalent to plsvo_align_batch_run followed by
 * plsvo_poseopt_batch_run with T_f_w = the aligned poses, minus one device->host->device trip and one sync.
 * plsvo_track_upload / plsvo_track_launch are its two device-side legs (results: plsvo_align_download and
 * plsvo_poseopt_download). */
int plsvo_track_upload(plsvo_ctx* ctx, const plsvo_align_batch* al_batch, const plsvo_poseopt_batch* po_batch);
int plsvo_track_launch(plsvo_ctx* ctx, const plsvo_align_params* al_params, const plsvo_poseopt_params* po_params);
int plsvo_track_batch_run(plsvo_ctx* ctx, const plsvo_align_batch* al_batch, const plsvo_align_params* al_params,
                          const plsvo_poseopt_batch* po_batch, const plsvo_poseopt_params* po_params,
                          const plsvo_align_result* al_out, const plsvo_poseopt_result* po_out);

/* ------------------------------------------------------------------------------------------
 * Image pyramid (SURVEY.md §8f "next", rank 2): frame_utils::createImgPyramid, src/frame.cpp:171-180,
 * i.e. repeated vk::halfSample (truncating 2x2 mean), for B frames.  Host in, host out.
 * level[0] of the result may be NULL (level 0 is the input); n_levels <= 7.
 * ---------------------------------------------------------------------------------------- */
typedef struct plsvo_pyramid_batch {
  int32_t batch, width, height, n_levels;
  const uint8_t* img0; /* [B] level-0 images: image b at img0 + b*stride0, rows pitch0 bytes */
  size_t pitch0, stride0;
} plsvo_pyramid_batch;

typedef struct plsvo_pyramid_result {
  uint8_t* level[PLSVO_MAX_LEVELS]; /* caller-allocated; level l holds (width>>l) x (height>>l) pixels */
  size_t pitch[PLSVO_MAX_LEVELS];
  size_t stride[PLSVO_MAX_LEVELS];
} plsvo_pyramid_result;

int plsvo_pyramid_batch_run(plsvo_ctx* ctx, const plsvo_pyramid_batch* in, const plsvo_pyramid_result* out);

/* ------------------------------------------------------------------------------------------
 * Feature alignment (SURVEY.md §8f "next", rank 1): feature_alignment::align2D,
 * include/plsvo/feature_alignment.h:49-55, src/feature_alignment.cpp:160-290 (scalar path) — the 8x8
 * inverse-compositional refinement Matcher::findMatchDirect runs per feature (src/matcher.cpp:201).
 * n features; feature i searches pyramid level level[i] of frame image_index[i].
 * ---------------------------------------------------------------------------------------- */
typedef struct plsvo_align2d_batch {
  int32_t n_features, n_images, width, height; /* width/height = level-0 size of the frames */
  int32_t n_iter, reserved;
  const uint8_t* img[PLSVO_MAX_LEVELS]; /* cur_img per level: frame b at img[l] + b*img_stride[l] */
  size_t img_pitch[PLSVO_MAX_LEVELS];
  size_t img_stride[PLSVO_MAX_LEVELS];
  const int32_t* image_index;           /* [n] */
  const int32_t* level;                 /* [n] */
  const uint8_t* ref_patch_with_border; /* [n][10*10] */
  const uint8_t* ref_patch;             /* [n][8*8]   */
  const double* px;                     /* [n][2] cur_px_estimate on entry (pixels of the search level) */
} plsvo_align2d_batch;

typedef struct plsvo_align2d_result {
  double* px;         /* [n][2] cur_px_estimate on return */
  uint8_t* converged; /* [n]    return value of align2D */
} plsvo_align2d_result;

int plsvo_align2d_batch_run(plsvo_ctx* ctx, const plsvo_align2d_batch* in, const plsvo_align2d_result* out);

/* ---- feature_alignment::align1D ("next") --------------------------------
 * Replaces, for n features at once, include/plsvo/feature_alignment.h:39-46 /
 * src/feature_alignment.cpp:40-157:
 *     bool align1D(const cv::Mat& cur_img, const Vector2f& dir, uint8_t* ref_patch_with_border,
 *                  uint8_t* ref_patch, const int n_iter, Vector2d& cur_px_estimate, double& h_inv);
 * the 1-DoF variant Matcher::findMatchDirect / findEpipolarMatchDirect use for edgelets
 * (src/matcher.cpp:195,331,400,497): the patch may only move along `dir`.  Same feature arrays as
 * plsvo_align2d_batch plus one direction per feature. */
typedef struct plsvo_align1d_batch {
  plsvo_align2d_batch features; /* images, image_index, level, patches, px — as for align2D */
  const float* dir;             /* [n][2] direction in which the patch is allowed to move */
} plsvo_align1d_batch;

typedef struct plsvo_align1d_result {
  double* px;         /* [n][2] cur_px_estimate on return */
  uint8_t* converged; /* [n]    return value of align1D */
  double* h_inv;      /* [n]    h_inv on return (:75) */
} plsvo_align1d_result;

int plsvo_align1d_batch_run(plsvo_ctx* ctx, const plsvo_align1d_batch* in, const plsvo_align1d_result* out);

/* ---- Matcher::findMatchDirect ("next") ----------------------------------
 * Replaces, for n (3D feature, current frame) candidates at once, include/plsvo/matcher.h:104-107 /
 * src/matcher.cpp:159-211:
 *     bool Matcher::findMatchDirect(const Point& pt, const Frame& cur_frame, Vector2d& px_cur);
 * i.e. everything after pt.getCloseViewObs() has picked the reference observation ref_ftr_ (list
 * logic, stays on the host): the in-frame test of the reference patch (:169-171), the affine warp
 * matrix (warp::getWarpMatrixAffine, :42-71), the search level (warp::getBestSearchLevel, :73-87), the
 * warped 10x10 reference patch (warp::warpAffine, :89-133, bilinear vk::interpolateMat_8u), the 8x8
 * patch cut out of it (:148-157) and the align2D / align1D refinement at the search level (:189-208).
 * The line-segment overload (:241-275) is the same computation on the two end points
 * (precomputeRefPatch, :213-232, + align2D): pass each end point as one feature and AND the flags.
 * Both frames use the same undistorted pinhole camera (app/run_pipeline.cpp:786-795). */
typedef struct plsvo_match_batch {
  int32_t n_features;
  int32_t n_ref_images;   /* keyframes holding the reference observations */
  int32_t n_cur_images;   /* current frames */
  int32_t n_pyr_levels;   /* Config::nPyrLevels(): the search level is < n_pyr_levels (:180) */
  int32_t n_iter;         /* Matcher::Options::align_max_iter (10, matcher.h:85) */
  int32_t reserved;
  plsvo_camera cam;
  const uint8_t* ref_img[PLSVO_MAX_LEVELS]; /* ref_ftr_->frame->img_pyr_[l]: frame r at ref_img[l] + r*ref_stride[l] */
  size_t ref_pitch[PLSVO_MAX_LEVELS];
  size_t ref_stride[PLSVO_MAX_LEVELS];
  const uint8_t* cur_img[PLSVO_MAX_LEVELS]; /* cur_frame.img_pyr_[l] */
  size_t cur_pitch[PLSVO_MAX_LEVELS];
  size_t cur_stride[PLSVO_MAX_LEVELS];
  const double* T_ref_w;      /* [n_ref_images][7] ref_ftr_->frame->T_f_w_ */
  const double* T_cur_w;      /* [n_cur_images][7] cur_frame.T_f_w_ */
  const int32_t* ref_index;   /* [n] keyframe of the reference observation */
  const int32_t* cur_index;   /* [n] current frame */
  const double* ref_px;       /* [n][2] ref_ftr_->px (level-0 pixels) */
  const double* ref_f;        /* [n][3] ref_ftr_->f */
  const int32_t* ref_level;   /* [n]    ref_ftr_->level */
  const uint8_t* is_edgelet;  /* [n] or NULL: PointFeat::type == EDGELET -> align1D along A_cur_ref*grad */
  const double* ref_grad;     /* [n][2] or NULL: PointFeat::grad */
  const double* pos;          /* [n][3] pt.pos_ (world) */
  const double* px_cur;       /* [n][2] px_cur on entry: the projection estimate (level-0 pixels) */
} plsvo_match_batch;

typedef struct plsvo_match_result {
  double* px_cur;        /* [n][2] px_cur on return (untouched where the in-frame test fails) */
  uint8_t* success;      /* [n]    return value of findMatchDirect */
  int32_t* search_level; /* [n]    Matcher::search_level_ (-1 where the in-frame test fails) */
  double* A_cur_ref;     /* [n][4] or NULL: Matcher::A_cur_ref_ row-major (A00 A01 A10 A11), which Reprojector::refine reads for
                          *        edgelets (src/reprojector.cpp:318-320); untouched where the in-frame test fails */
} plsvo_match_result;

int plsvo_match_direct_batch_run(plsvo_ctx* ctx, const plsvo_match_batch* in, const plsvo_match_result* out);

/* ---- Structure optimisation: Point::optimize / LineSeg::optimize ("next") ---
 * Replaces, for a batch of 3D features, include/plsvo/feature3D.h:120,157 / src/feature3D_impl.cpp:36-95,
 * 97-174 as driven by FrameHandlerBase::optimizeStructure (src/frame_handler_base.cpp:202-237):
 *     void Point::optimize(const size_t n_iter);     void LineSeg::optimize(const size_t n_iter);
 * 3x3 Gauss-Newton on the reprojection error (unit plane) of a 3D point over its observations obs_
 * (a LineSeg optimises its two end points with a coupled accept/roll-back/convergence test).
 * Observations are given in CSR form, in obs_ list order (the summation order of the reference). */
typedef struct plsvo_structopt_batch {
  int32_t n_points, n_segs, n_frames;
  int32_t n_iter_pts;  /* Config::structureOptimNumIter() */
  int32_t n_iter_segs; /* Config::structureOptimNumIterSegs() */
  int32_t reserved;
  const double* T_f_w;           /* [n_frames][7] (*it)->frame->T_f_w_ of the observing keyframes */
  const int32_t* pt_obs_begin;   /* [n_points+1] offsets into pt_obs_* */
  const int32_t* pt_obs_frame;   /* [n_pt_obs]   index into T_f_w */
  const double* pt_obs_f;        /* [n_pt_obs][3] PointFeat::f */
  const double* pt_pos;          /* [n_points][3] Point::pos_ on entry */
  const int32_t* seg_obs_begin;  /* [n_segs+1] */
  const int32_t* seg_obs_frame;  /* [n_seg_obs] */
  const double* seg_obs_sf;      /* [n_seg_obs][3] LineFeat::sf */
  const double* seg_obs_ef;      /* [n_seg_obs][3] LineFeat::ef */
  const double* seg_spos;        /* [n_segs][3] LineSeg::spos_ on entry */
  const double* seg_epos;        /* [n_segs][3] LineSeg::epos_ on entry */
} plsvo_structopt_batch;

typedef struct plsvo_structopt_result {
  double* pt_pos;   /* [n_points][3] Point::pos_ on return */
  double* seg_spos; /* [n_segs][3] */
  double* seg_epos; /* [n_segs][3] */
  int32_t* pt_iters;  /* [n_points] or NULL: GN iterations executed (diagnostic) */
  int32_t* seg_iters; /* [n_segs] or NULL */
} plsvo_structopt_result;

int plsvo_structopt_batch_run(plsvo_ctx* ctx, const plsvo_structopt_batch* in, const plsvo_structopt_result* out);

/* ---- Depth-filter point-seed update ("next") ------------------------------
 * Replaces, for n point seeds at once, the body of DepthFilter::updatePointSeeds (src/depth_filter.cpp:270-365):
 * visibility test of the seed in the current frame (:291-304), inverse-depth search range (:307-308),
 *     bool Matcher::findEpipolarMatchDirect(ref_frame, cur_frame, ref_ftr, d_estimate, d_min, d_max, depth)
 * (include/plsvo/matcher.h, src/matcher.cpp:277-420: epipolar segment, affine warp, edgelet pre-selection, ZMSSD
 * search along the epipolar line or direct alignment when it is shorter than 2 px, sub-pixel align2D/align1D,
 * depthFromTriangulation :135-146), DepthFilter::computeTau (:568-584) and the Gaussian x Beta update
 * DepthFilter::updatePointSeed (:489-512, Vogiatzis & Hernandez 2011).  The list logic around it (seed ageing,
 * creating a Point from a converged seed, the detector's occupancy grid) stays on the host; `converged` reports
 * the reference's test sqrt(sigma2) < z_range / seed_convergence_sigma2_thresh (:334).
 * Images must be dense (pitch == width of the level): the reference indexes the ZMSSD patch with Mat::cols (:380-382). */
typedef struct plsvo_seed_batch {
  int32_t n_seeds;
  int32_t n_ref_images;
  int32_t n_cur_images;
  int32_t n_pyr_levels;          /* Config::nPyrLevels() */
  int32_t n_iter;                /* Matcher::Options::align_max_iter (10) */
  int32_t max_epi_search_steps;  /* Matcher::Options::max_epi_search_steps (1000) */
  uint8_t align_1d;              /* Matcher::Options::align_1d (false) */
  uint8_t subpix_refinement;     /* Matcher::Options::subpix_refinement (true) */
  uint8_t epi_search_edgelet_filtering; /* (true) */
  uint8_t reserved0[5];
  double epi_search_edgelet_max_angle;   /* (0.7) */
  double seed_convergence_sigma2_thresh; /* DepthFilter::Options (200.0) */
  plsvo_camera cam;
  const uint8_t* ref_img[PLSVO_MAX_LEVELS];
  size_t ref_pitch[PLSVO_MAX_LEVELS];
  size_t ref_stride[PLSVO_MAX_LEVELS];
  const uint8_t* cur_img[PLSVO_MAX_LEVELS];
  size_t cur_pitch[PLSVO_MAX_LEVELS];
  size_t cur_stride[PLSVO_MAX_LEVELS];
  const double* T_ref_w;     /* [n_ref_images][7] it->ftr->frame->T_f_w_ */
  const double* T_cur_w;     /* [n_cur_images][7] frame->T_f_w_ */
  const int32_t* ref_index;  /* [n] */
  const int32_t* cur_index;  /* [n] */
  const double* ref_px;      /* [n][2] it->ftr->px */
  const double* ref_f;       /* [n][3] it->ftr->f */
  const int32_t* ref_level;  /* [n]    it->ftr->level */
  const uint8_t* is_edgelet; /* [n] or NULL */
  const double* ref_grad;    /* [n][2] or NULL */
  const float* a;            /* [n] PointSeed::a on entry */
  const float* b;            /* [n] */
  const float* mu;           /* [n] inverse depth mean */
  const float* z_range;      /* [n] */
  const float* sigma2;       /* [n] */
} plsvo_seed_batch;

#define PLSVO_SEED_NOT_VISIBLE 0 /* behind the camera / outside the image: seed untouched (:296-304) */
#define PLSVO_SEED_NO_MATCH 1    /* findEpipolarMatchDirect failed: b += 1 (:312-317) */
#define PLSVO_SEED_UPDATED 2     /* Bayesian update applied (:320-325) */

typedef struct plsvo_seed_result {
  float* a;           /* [n] PointSeed state on return */
  float* b;
  float* mu;
  float* sigma2;
  int32_t* status;    /* [n] PLSVO_SEED_* */
  uint8_t* converged; /* [n] */
  double* depth;      /* [n] z of findEpipolarMatchDirect (NaN unless status == UPDATED) */
  double* px_cur;     /* [n][2] Matcher::px_cur_ (diagnostic; NaN where no position was computed) */
} plsvo_seed_result;

int plsvo_seed_update_batch_run(plsvo_ctx* ctx, const plsvo_seed_batch* in, const plsvo_seed_result* out);

/* Line-seed variant: the body of DepthFilter::updateLineSeeds (src/depth_filter.cpp:367-471).  A LineSeed carries one
 * inverse-depth Gaussian per end point and a shared Beta (a, b).  Both end points are searched with
 * Matcher::findEpipolarMatchDirectSegmentEndpoint (src/matcher.cpp:420-588) — which, as in the reference, warps and
 * searches around the segment feature's own px / f (its mid point, base Feature fields) for BOTH depth hypotheses and
 * has no edgelet pre-selection — then computeTau uses the end-point bearings sf / ef (:415-418) and
 * DepthFilter::updateLineSeed (:514-565) updates both Gaussians and takes a = max(a_s, a_e), b = min(b_s, b_e).
 * `seeds` describes the start point: ref_px / ref_f = LineFeat::px / f, mu / z_range / sigma2 = mu_s / z_range_s / sigma2_s;
 * is_edgelet / ref_grad are ignored. */
typedef struct plsvo_line_seed_batch {
  plsvo_seed_batch seeds;
  const double* ref_sf;    /* [n][3] LineFeat::sf */
  const double* ref_ef;    /* [n][3] LineFeat::ef */
  const float* mu_e;       /* [n] */
  const float* z_range_e;  /* [n] */
  const float* sigma2_e;   /* [n] */
} plsvo_line_seed_batch;

typedef struct plsvo_line_seed_result {
  plsvo_seed_result seeds; /* a, b, mu_s, sigma2_s, status, converged (both end points), depth = z_s, px_cur of the start search */
  float* mu_e;             /* [n] */
  float* sigma2_e;         /* [n] */
  double* depth_e;         /* [n] z_e (NaN unless status == UPDATED) */
  double* px_cur_e;        /* [n][2] or NULL: Matcher::px_cur_ of the end-point search — what matcherls_.px_cur_ holds when
                            *        DepthFilter::updateLineSeeds marks the detector grid on keyframes (:426-430); NaN where the
                            *        end-point search did not run or computed no position */
} plsvo_line_seed_result;

int plsvo_line_seed_update_batch_run(plsvo_ctx* ctx, const plsvo_line_seed_batch* in, const plsvo_line_seed_result* out);

/* device time (CUDA events on the context's stream) of the kernel launched by the last
 * plsvo_pyramid / align2d / align1d / match_direct / seed_update / structopt _batch_run call: the kernel alone,
 * without the host<->device copies those calls also make.  Measurement aid, no reference counterpart. */
int plsvo_last_kernel_ms(plsvo_ctx* ctx, float* ms);

/* number of kernels this context has launched since creation (bench "gpu_launches") */
int64_t plsvo_launch_count(const plsvo_ctx* ctx);

/* Device self-test of the fp32 weight kernel: evaluates w = 1/(1+a) with the product's fp32
 * sequence and with the reference's double-then-narrow expression (sparse_img_align.cpp:479) for
 * n pseudo-random a in [0,256) plus all n_exhaustive first float bit patterns of [0,256), and
 * returns the number of bitwise mismatches. */
int plsvo_selftest_weight(plsvo_ctx* ctx, uint32_t n, uint32_t seed, uint64_t* mismatches);

/* library build info: "plsvo_b200 <version> sm_100a" */
const char* plsvo_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PLSVO_B200_H_ */
