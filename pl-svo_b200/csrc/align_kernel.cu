// align_kernel.cu — sparse image alignment (plsvo::SparseImgAlign::run, src/sparse_img_align.cpp:54-95)
// as ONE persistent sm_100a kernel: a CTA owns a frame pair for its whole coarse-to-fine
// Gauss-Newton optimisation, so a pair costs no host round trips and no re-launches.
//
// Mapping:
//   * work queue: CTAs pull pair indices from an atomic counter (iteration counts vary per pair).
//   * per level: one thread issues a bulk async copy (TMA engine, cp.async.bulk -> UBLKCP) of the
//     current image level into shared memory while all threads precompute the reference-patch
//     cache (4x4 bilinear intensities + central-difference gradients: sparse_img_align.cpp:195-378)
//     from the reference image in global memory.
//   * per GN iteration: thread-per-patch residual pass (:380-502 points, :504-695 segment samples):
//     warp the patch centre with T_cur_from_ref (double), bilinear 4x4 residuals in float with the
//     reference's exact operation order, five per-patch sums (w*dx*dx, w*dx*dy, w*dy*dy, w*dx*r,
//     w*dy*r) in double, then a rank-2 update of the thread's 21+6 accumulators using the two
//     projection-Jacobian rows of the patch (J_px = (dx*row0 + dy*row1)*fx/2^l, :261-262) — this
//     factorisation replaces the reference's 6x(N*16) double Jacobian cache (768 B/patch) by
//     128 B/patch of float gradients.  Accumulators are reduced with a register-halving warp
//     shuffle tree, then across warps through shared memory in fixed order (deterministic).
//   * thread 0 solves the 6x6 system (pivoted LDLT), applies T <- T*exp(-x) and the vikit
//     NLLSSolver accept / rollback / convergence logic on chip.
#include <cuda_runtime.h>
#include <stdint.h>

#include "device_math.cuh"
#include "internal.h"

namespace plsvo {

namespace {

struct PairCtl {
  double R[9];
  double t[3];
  double model[7];      // T_cur_from_ref (q, t)
  double old_model[7];
  double T_ref[7];
  double ref_pos[3];
  double chi2_prev;
  double H_last[36];
  double scratch[36];
  double g[6];
  double x[6];
  unsigned long long mbar;
  long long n_meas_last;
  int pair;
  int flag;
  int stop;
  int iter;
  int n_seg_patches;
  int n_seg_slots;  // lane slots taken by the segment groups at the current level
  unsigned int patch_iters;
  unsigned int patch_levels;
  int iters_level[PLSVO_MAX_LEVELS];
};

struct PatchSums {  // five fp32 in-patch sums of one pass + how to apply them (0 skip, 1 point, 2+j segment j)
  float S[5];
  int kind;
};

struct Layout {
  uint32_t ctl, red, tot, seg_alive, seg_N, seg_off, seg_slot, slot_seg, seg_px, seg_scale, prec, pt_vis, xyz, cache, img, total;
};

__host__ __device__ inline uint32_t align_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

__host__ __device__ inline Layout make_layout(int n_pts, int n_segs, int max_patches, int max_seg_patches,
                                              int img_bytes, bool cache_in_smem) {
  Layout L;
  uint32_t o = 0;
  L.ctl = o;
  o = align_up(o + (uint32_t)sizeof(PairCtl), 16);
  L.red = o;
  o += 8 * 32 * 8;  // cross-warp partials (up to 8 warps)
  L.tot = o;
  o += 32 * 8;
  L.seg_N = o;
  o += 4u * (uint32_t)n_segs;
  L.seg_off = o;
  o += 4u * (uint32_t)n_segs;
  L.seg_slot = o;
  o += 4u * (uint32_t)n_segs;
  L.slot_seg = o;
  o += 2u * (2u * (uint32_t)max_seg_patches + 64u);  // lane slot -> segment (groups of 2^k lanes, k per segment)
  L.seg_px = align_up(o, 16);
  o = L.seg_px + 16u * (uint32_t)max_seg_patches;  // 2D centre of every segment sample (precompute only)
  L.seg_scale = o;
  o += 16u * (uint32_t)n_segs;  // per-segment (weight/res_, weight) of the current pass
  L.prec = o;
  o += (uint32_t)sizeof(PatchSums) * (uint32_t)max_patches;  // per-patch sums of the current pass
  L.seg_alive = o;
  o += (uint32_t)n_segs;
  L.pt_vis = o;
  o += (uint32_t)n_pts;
  o = align_up(o, 16);
  L.xyz = o;
  o += 4u * 8u * (uint32_t)max_patches;  // X, Y, Z, 1/Z of every patch's 3D point in the ref frame
  o = align_up(o, 16);
  L.cache = o;
  if (cache_in_smem) o += (uint32_t)kCacheRows * 16u * (uint32_t)max_patches;
  o = align_up(o, 128);
  L.img = o;
  o += (uint32_t)img_bytes + 16u;  // slack: the 5-byte row reads fetch whole aligned words
  L.total = o;
  return L;
}

__device__ __forceinline__ float u8f(uint8_t v) { return (float)v; }

// bilinear sample with the reference's operation order: ((wTL*a + wTR*b) + wBL*c) + wBR*d,
// every product and sum rounded separately (no FMA contraction) — sparse_img_align.cpp:458
__device__ __forceinline__ float bilin(float wTL, float wTR, float wBL, float wBR, float a, float b, float c, float d) {
  return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(wTL, a), __fmul_rn(wTR, b)), __fmul_rn(wBL, c)), __fmul_rn(wBR, d));
}

// Patch::setPosition + isInFrame + computeInterpWeights (src/feature.cpp:189-208, feature.h:139-144).
// The weights are computed in float: for accepted patches (floor >= boundary >= 2) 1-subpix is
// exact in float and each product is rounded once, which equals the reference's double-then-narrow.
__device__ __forceinline__ bool patch_setup(double u, double v, int cols, int rows, int boundary, int& ui, int& vi,
                                            float& wTL, float& wTR, float& wBL, float& wBR) {
  const float uf = (float)u, vf = (float)v;
  const float fu = floorf(uf), fv = floorf(vf);
  ui = (int)fu;
  vi = (int)fv;
  if (ui < boundary || vi < boundary || ui >= cols - boundary || vi >= rows - boundary) return false;
  const float su = __fsub_rn(uf, fu), sv = __fsub_rn(vf, fv);
  const float omu = __fsub_rn(1.0f, su), omv = __fsub_rn(1.0f, sv);
  wTL = __fmul_rn(omu, omv);
  wTR = __fmul_rn(su, omv);
  wBL = __fmul_rn(omu, sv);
  wBR = __fmul_rn(su, sv);
  return true;
}

// LineFeat::setupSampling (src/feature.cpp:160-173) followed by the per-level decimation (:320)
__device__ __forceinline__ int seg_num_samples(const double* spx, const double* epx, double length, int level,
                                               double* dif) {
  dif[0] = epx[0] - spx[0];
  dif[1] = epx[1] - spx[1];
  const double a0 = fabs(dif[0]), a1 = fabs(dif[1]);
  const double tan_dir = fmin(a0, a1) / fmax(a0, a1);
  const double sin_dir = tan_dir / sqrt(1.0 + tan_dir * tan_dir);
  const double correction = 2.0 * sqrt(1.0 + sin_dir * sin_dir);
  const double nd = fmax(1.0, length / (2.0 * 4 * correction));
  const unsigned long long n0 = (unsigned long long)nd;
  return (int)(1 + (n0 - 1) / (unsigned long long)(1 << level));
}

__device__ __forceinline__ bool cam_in_frame(int ox, int oy, int boundary, int level, int width, int height) {
  return ox >= boundary && ox < width / (1 << level) - boundary && oy >= boundary &&
         oy < height / (1 << level) - boundary;
}

// rank-2 update of the 21 (upper-triangular H) + 6 (Jres) accumulators of one thread:
//   H += Sxx r0 r0^T + Sxy (r0 r1^T + r1 r0^T) + Syy r1 r1^T ,  Jres -= Sxr r0 + Syr r1
__device__ __forceinline__ void rank2_update(double* acc, double X, double Y, double z_inv, double Sxx, double Sxy,
                                             double Syy, double Sxr, double Syr) {
  double r0[6], r1[6];
  jacobian_rows_zinv(X, Y, z_inv, r0, r1);
  double p[6], q[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    p[i] = Sxx * r0[i] + Sxy * r1[i];
    q[i] = Sxy * r0[i] + Syy * r1[i];
  }
  int idx = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = i; j < 6; ++j) acc[idx++] += p[i] * r0[j] + q[i] * r1[j];
#pragma unroll
  for (int i = 0; i < 6; ++i) acc[21 + i] -= Sxr * r0[i] + Syr * r1[i];
}

// One patch of the residual pass.  WEIGHTED = point patch (:450-500: w = 1/(1+|r|)), otherwise a
// segment sample (:612-637: unweighted sums, |r| collected).  Returns false if the warped patch
// is not fully inside the current image (isInFrame(halfsize)).
// Per-pixel values (bilinear intensity, residual, weight, chi2 term) are bit-identical to the
// reference's float arithmetic; the five in-patch sums run in fp32 FMAs over the 16 pixels and
// are widened to double per patch (the reference accumulates every pixel in double: the
// difference is ~1e-7 relative on one patch's contribution and does not move the fixed point).
// five consecutive image bytes starting at byte offset (sh/8) of the aligned word pair at `row`
__device__ __forceinline__ void load_row5(const uint8_t* row, int sh, float* f) {
  const uint32_t w0 = *reinterpret_cast<const uint32_t*>(row);
  const uint32_t w1 = *reinterpret_cast<const uint32_t*>(row + 4);
  const uint32_t lo = __funnelshift_r(w0, w1, sh);
  const uint32_t hi = w1 >> sh;
  f[0] = byte_to_float(lo, 0);
  f[1] = byte_to_float(lo, 1);
  f[2] = byte_to_float(lo, 2);
  f[3] = byte_to_float(lo, 3);
  f[4] = byte_to_float(hi, 0);
}
// seven consecutive bytes (reference image, global memory, read-only path)
__device__ __forceinline__ void load_row7(const uint8_t* row, int sh, float* g) {
  const uint32_t* wp = reinterpret_cast<const uint32_t*>(row);
  const uint32_t w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2);
  const uint32_t lo = __funnelshift_r(w0, w1, sh);
  const uint32_t hi = __funnelshift_r(w1, w2, sh);
  g[0] = byte_to_float(lo, 0);
  g[1] = byte_to_float(lo, 1);
  g[2] = byte_to_float(lo, 2);
  g[3] = byte_to_float(lo, 3);
  g[4] = byte_to_float(hi, 0);
  g[5] = byte_to_float(hi, 1);
  g[6] = byte_to_float(hi, 2);
}

// One patch of the residual pass.  weighted = point patch (:450-500: w = 1/(1+|r|)); otherwise a segment
// sample (:612-637: unweighted sums, |r| collected).  One body serves both (w = 1 is exact for the
// unweighted sums) so that the hot loop stays small enough for the instruction cache.  Returns false
// if the warped patch is not fully inside the current image (isInFrame(halfsize)).
// Per-pixel values (bilinear intensity, residual, weight, chi2 term) are bit-identical to the
// reference's float arithmetic; the five in-patch sums run in fp32 FMAs over the 16 pixels and are
// widened to double per patch (the reference accumulates every pixel in double: the difference is
// ~1e-7 relative on one patch's contribution and does not move the fixed point).
template <bool weighted>
__device__ __forceinline__ bool eval_patch(const uint8_t* __restrict__ img, int pitch, int cols, int rows,
                                           const float4* cache, int MP, int p, double u, double v,
                                           float* S /*[5]*/, float& acc_out) {
  int ui, vi;
  float wTL, wTR, wBL, wBR;
  if (!patch_setup(u, v, cols, rows, 2, ui, vi, wTL, wTR, wBL, wBR)) return false;
  // 5x5 footprint, streamed row by row (two aligned 32-bit loads + funnel shift per row; rows are
  // 16B-pitched).  The row loop is kept rolled.
  const int c0 = ui - 2;
  const int sh = (c0 & 3) * 8;
  const uint8_t* rowp = img + (size_t)(vi - 2) * pitch + (c0 & ~3);
  float ra[5], rb[5];
  load_row5(rowp, sh, ra);
#ifdef PLSVO_FP64_SUMS
  double Sxx = 0, Sxy = 0, Syy = 0, Sxr = 0, Syr = 0;
#else
  float Sxx = 0.f, Sxy = 0.f, Syy = 0.f, Sxr = 0.f, Syr = 0.f;
#endif
  float acc_f = 0.f;
  const float4* cp = cache + p;
#pragma unroll 1
  for (int y = 0; y < 4; ++y) {
    rowp += pitch;
    load_row5(rowp, sh, rb);
    const float4 ref4 = cp[0];
    const float4 dx4 = cp[4 * MP];
    const float4 dy4 = cp[8 * MP];
    cp += MP;
    const float refv[4] = {ref4.x, ref4.y, ref4.z, ref4.w};
    const float dxv[4] = {dx4.x, dx4.y, dx4.z, dx4.w};
    const float dyv[4] = {dy4.x, dy4.y, dy4.z, dy4.w};
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      const float cur = bilin(wTL, wTR, wBL, wBR, ra[x], ra[x + 1], rb[x], rb[x + 1]);
      const float res = __fsub_rn(cur, refv[x]);
      const float dx = dxv[x], dy = dyv[x];
      const float ares = fabsf(res);
      const float w = weighted ? weight_rcp(ares) : 1.0f;                        // :479
      const float term = weighted ? __fmul_rn(__fmul_rn(res, res), w) : ares;   // :484 / :643
      acc_f = __fadd_rn(acc_f, term);
#ifdef PLSVO_FP64_SUMS
      const double wdx = (double)w * (double)dx, wdy = (double)w * (double)dy;
      Sxx += wdx * (double)dx;
      Sxy += wdx * (double)dy;
      Syy += wdy * (double)dy;
      Sxr += wdx * (double)res;
      Syr += wdy * (double)res;
#else
      const float wdx = __fmul_rn(w, dx), wdy = __fmul_rn(w, dy);
      Sxx = fmaf(wdx, dx, Sxx);
      Sxy = fmaf(wdx, dy, Sxy);
      Syy = fmaf(wdy, dy, Syy);
      Sxr = fmaf(wdx, res, Sxr);
      Syr = fmaf(wdy, res, Syr);
#endif
    }
#pragma unroll
    for (int c = 0; c < 5; ++c) ra[c] = rb[c];
  }
  S[0] = (float)Sxx, S[1] = (float)Sxy, S[2] = (float)Syy, S[3] = (float)Sxr, S[4] = (float)Syr;
  acc_out = acc_f;  // chi2 of this patch (points) / sum of |res| (segment sample)
  return true;
}

// Reference-patch precompute for one patch (:243-264 / :354-375): 16 interpolated intensities and
// central-difference gradients of the interpolated image, written as 12 float4 rows.
// V[a][b] = interpolated intensity at integer offset (a-1, b-1) from the patch origin; the reference
// evaluates the same bilinear expression for ref / dx / dy of neighbouring pixels (:251-258), so each
// value is computed once, in a rolled sliding window over the rows (V rows y, y+1, y+2 for pixel row y).
__device__ __forceinline__ void precompute_patch(const uint8_t* __restrict__ img, int pitch, int ui, int vi, float wTL,
                                                 float wTR, float wBL, float wBR, float4* __restrict__ cache, int MP,
                                                 int p) {
  const int c0 = ui - 3;
  const int sh = (c0 & 3) * 8;
  const uint8_t* rowp = img + (size_t)(vi - 3) * pitch + (c0 & ~3);
  float g0[7], g1[7], Va[6], Vb[6], Vc[6];
  load_row7(rowp, sh, g0);
  load_row7(rowp + pitch, sh, g1);
#pragma unroll
  for (int c = 0; c < 6; ++c) Va[c] = bilin(wTL, wTR, wBL, wBR, g0[c], g0[c + 1], g1[c], g1[c + 1]);
  rowp += 2 * pitch;
  load_row7(rowp, sh, g0);
#pragma unroll
  for (int c = 0; c < 6; ++c) Vb[c] = bilin(wTL, wTR, wBL, wBR, g1[c], g1[c + 1], g0[c], g0[c + 1]);
  // here g0 holds block row 2; loop invariant: g0 = block row y+2
  float4* cp = cache + p;
#pragma unroll 1
  for (int y = 0; y < 4; ++y) {
    rowp += pitch;
    load_row7(rowp, sh, g1);  // block row y+3
#pragma unroll
    for (int c = 0; c < 6; ++c) Vc[c] = bilin(wTL, wTR, wBL, wBR, g0[c], g0[c + 1], g1[c], g1[c + 1]);
    float refv[4], dxv[4], dyv[4];
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      refv[x] = Vb[x + 1];
      dxv[x] = __fmul_rn(0.5f, __fsub_rn(Vb[x + 2], Vb[x]));
      dyv[x] = __fmul_rn(0.5f, __fsub_rn(Vc[x + 1], Va[x + 1]));
    }
    cp[0] = make_float4(refv[0], refv[1], refv[2], refv[3]);
    cp[4 * MP] = make_float4(dxv[0], dxv[1], dxv[2], dxv[3]);
    cp[8 * MP] = make_float4(dyv[0], dyv[1], dyv[2], dyv[3]);
    cp += MP;
#pragma unroll
    for (int c = 0; c < 6; ++c) Va[c] = Vb[c], Vb[c] = Vc[c];
#pragma unroll
    for (int c = 0; c < 7; ++c) g0[c] = g1[c];
  }
}

__device__ __forceinline__ void zero_gradients(float4* cache, int MP, int p) {
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int y = 0; y < 8; ++y) cache[(4 + y) * MP + p] = z;
}

// Thread 0: one Gauss-Newton step of vk::NLLSSolver::optimizeGaussNewton with SparseImgAlign's
// solve()/update() (:697-710).  tot = block totals [0..20]=H upper, [21..26]=Jres, [27]=chi2,
// [28]=n_meas, [29]=patches evaluated.
__device__ __noinline__ void gn_step(PairCtl* ctl, const double* tot, int level, int n_iter, double eps) {
  const long long n_meas = (long long)tot[28];
  ctl->n_meas_last = n_meas;
  ctl->patch_iters += (unsigned int)tot[29];
  ctl->iters_level[level] += 1;
  // chi2/n_meas_ : float / size_t -> float (:192)
  const double new_chi2 = (double)((float)tot[27] / (float)(unsigned long long)n_meas);
  {
    double Hu[21], gg[6], xx[6];
#pragma unroll
    for (int i = 0; i < 21; ++i) Hu[i] = tot[i];
#pragma unroll
    for (int i = 0; i < 6; ++i) gg[i] = tot[21 + i];
#ifdef PLSVO_PIVOT_ALWAYS
    if (false) {
#else
    if (ldlt6_reg(Hu, gg, xx)) {
#endif
#pragma unroll
      for (int i = 0; i < 6; ++i) ctl->x[i] = xx[i];
    } else {
      // degenerate system: pivoted Eigen-style routine on the full symmetric matrix
      double* H = ctl->H_last;
      int idx = 0;
      for (int i = 0; i < 6; ++i)
        for (int j = i; j < 6; ++j) {
          H[i * 6 + j] = tot[idx];
          H[j * 6 + i] = tot[idx];
          ++idx;
        }
      for (int i = 0; i < 6; ++i) ctl->g[i] = tot[21 + i];
      ldlt6_solve(H, ctl->g, ctl->x, ctl->scratch);
    }
  }
  if (isnan(ctl->x[0])) ctl->stop = 1;
  const bool reject = (ctl->iter > 0 && new_chi2 > ctl->chi2_prev) || ctl->stop;
  int flag;
  if (reject) {
    for (int i = 0; i < 7; ++i) ctl->model[i] = ctl->old_model[i];
    flag = 1;
  } else {
    double mx[6];
    double nm = 0.0;
    for (int i = 0; i < 6; ++i) {
      mx[i] = -ctl->x[i];
      nm = fmax(nm, fabs(ctl->x[i]));
    }
    SE3q model;
    model.q.x = ctl->model[0], model.q.y = ctl->model[1], model.q.z = ctl->model[2], model.q.w = ctl->model[3];
    model.t = v3(ctl->model[4], ctl->model[5], ctl->model[6]);
    const SE3q nm_model = se3_mul(model, se3_exp(mx));
    for (int i = 0; i < 7; ++i) ctl->old_model[i] = ctl->model[i];
    se3_store(nm_model, ctl->model);
    ctl->chi2_prev = new_chi2;
    flag = (nm <= eps) ? 1 : 0;
  }
  ctl->iter += 1;
  if (ctl->iter >= n_iter) flag = 1;
  Quat q;
  q.x = ctl->model[0], q.y = ctl->model[1], q.z = ctl->model[2], q.w = ctl->model[3];
  quat_to_R(q, ctl->R);
  ctl->t[0] = ctl->model[4], ctl->t[1] = ctl->model[5], ctl->t[2] = ctl->model[6];
  if (flag) {  // last evaluated pass of this level: keep H_ (getFisherInformation, :97-102)
    int idx = 0;
    for (int i = 0; i < 6; ++i)
      for (int j = i; j < 6; ++j) {
        ctl->H_last[i * 6 + j] = tot[idx];
        ctl->H_last[j * 6 + i] = tot[idx];
        ++idx;
      }
  }
  ctl->flag = flag;
}

template <bool CACHE_SMEM, int NT>
__global__ void __launch_bounds__(NT, 512 / NT) sparse_img_align_kernel(const AlignArgs a) {
  constexpr int kAlignThreads = NT;
  constexpr int kAlignWarps = NT / 32;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int MP = a.max_patches;
  const Layout L = make_layout(a.n_pts, a.n_segs, MP, a.max_seg_patches, a.smem_img_bytes, CACHE_SMEM);
  PairCtl* ctl = reinterpret_cast<PairCtl*>(smem + L.ctl);
  double* red = reinterpret_cast<double*>(smem + L.red);
  double* tot = reinterpret_cast<double*>(smem + L.tot);
  uint8_t* seg_alive = smem + L.seg_alive;
  int* seg_N = reinterpret_cast<int*>(smem + L.seg_N);
  int* seg_off = reinterpret_cast<int*>(smem + L.seg_off);
  int* seg_slot = reinterpret_cast<int*>(smem + L.seg_slot);
  uint16_t* slot_seg = reinterpret_cast<uint16_t*>(smem + L.slot_seg);
  const int max_slots = 2 * a.max_seg_patches + 64;
  double* seg_px = reinterpret_cast<double*>(smem + L.seg_px);
  double* seg_scale = reinterpret_cast<double*>(smem + L.seg_scale);
  PatchSums* prec = reinterpret_cast<PatchSums*>(smem + L.prec);
  uint8_t* pt_vis = smem + L.pt_vis;
  uint8_t* img_s = smem + L.img;
  double* xyz = reinterpret_cast<double*>(smem + L.xyz);
  float4* cache;
  if (CACHE_SMEM) {
    cache = reinterpret_cast<float4*>(smem + L.cache);
  } else {
    cache = a.ws_cache + (size_t)blockIdx.x * kCacheRows * MP;
  }
  uint64_t* bar = reinterpret_cast<uint64_t*>(&ctl->mbar);
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbarrier_init();
    fence_proxy_async();
  }
  __syncthreads();
  uint32_t bar_parity = 0;

  for (;;) {
    __syncthreads();  // everyone is done with ctl of the previous pair
    if (tid == 0) {
      const int nb = (int)atomicAdd(a.work_counter, 1u);
      if (a.gate_chunk > 0 && nb < a.B) {
        // host-buffer pipeline: this pair's inputs are still in flight over PCIe until the copy stream
        // has bumped the arrival counter past its chunk
        const unsigned need = (unsigned)(nb / a.gate_chunk) + 1u;
        while (*reinterpret_cast<const volatile unsigned int*>(a.arrived) < need) __nanosleep(1000);
        __threadfence_system();
      }
      ctl->pair = nb;
    }
    __syncthreads();
    const int b = ctl->pair;
    if (b >= a.B) break;

    const int np = a.pt_count ? a.pt_count[b] : a.n_pts;
    const int ns = a.seg_count ? a.seg_count[b] : a.n_segs;
    const size_t po = (size_t)b * a.n_pts, so = (size_t)b * a.n_segs;

    if (np == 0 && ns == 0) {  // :58-62 early-out: return 0, cur pose untouched
      if (tid == 0) {
        for (int i = 0; i < 7; ++i) a.out_T[(size_t)b * 7 + i] = a.T_cur_w[(size_t)b * 7 + i];
        a.out_n_tracked[b] = 0;
        for (int i = 0; i < 36; ++i) a.out_H[(size_t)b * 36 + i] = 0.0;
        for (int l = 0; l < PLSVO_MAX_LEVELS; ++l) a.out_iters[(size_t)b * PLSVO_MAX_LEVELS + l] = 0;
        a.out_status[b] = 1;
        a.out_patch_iters[b] = 0;
        a.out_patch_levels[b] = 0;
      }
      for (int j = tid; j < a.n_segs; j += kAlignThreads) a.out_seg_killed[so + j] = 0;
      continue;
    }

    if (tid == 0) {
      const SE3q T_ref = se3_load(a.T_ref_w + (size_t)b * 7);
      const SE3q T_cur = se3_load(a.T_cur_w + (size_t)b * 7);
      const SE3q T_ref_inv = se3_inverse(T_ref);
      const SE3q model = se3_mul(T_cur, T_ref_inv);  // :80
      se3_store(T_ref, ctl->T_ref);
      se3_store(model, ctl->model);
      ctl->ref_pos[0] = T_ref_inv.t.x, ctl->ref_pos[1] = T_ref_inv.t.y, ctl->ref_pos[2] = T_ref_inv.t.z;
      quat_to_R(model.q, ctl->R);
      ctl->t[0] = model.t.x, ctl->t[1] = model.t.y, ctl->t[2] = model.t.z;
      ctl->chi2_prev = 1e10;
      ctl->stop = 0;
      ctl->n_meas_last = 0;
      ctl->patch_iters = 0;
      ctl->patch_levels = 0;
      for (int i = 0; i < 36; ++i) ctl->H_last[i] = 0.0;
      for (int l = 0; l < PLSVO_MAX_LEVELS; ++l) ctl->iters_level[l] = 0;
    }
    __syncthreads();
    const double rpx = ctl->ref_pos[0], rpy = ctl->ref_pos[1], rpz = ctl->ref_pos[2];

    // per-pair point setup: xyz_ref = f * |pos - ref_pos|  (:229-230), visibility cleared
    for (int i = tid; i < np; i += kAlignThreads) {
      pt_vis[i] = 0;
      const double* pos = a.pt_pos + (po + i) * 3;
      const double* f = a.pt_f + (po + i) * 3;
      const double dx = pos[0] - rpx, dy = pos[1] - rpy, dz = pos[2] - rpz;
      const double depth = sqrt(dx * dx + dy * dy + dz * dz);
      const double Zr = f[2] * depth;
      xyz[0 * MP + i] = f[0] * depth;
      xyz[1 * MP + i] = f[1] * depth;
      xyz[2 * MP + i] = Zr;
      xyz[3 * MP + i] = 1. / Zr;  // z_inv of Frame::jacobian_xyz2uv (frame.h:144), constant per pair
    }
    for (int j = tid; j < ns; j += kAlignThreads) seg_alive[j] = a.seg_valid ? (a.seg_valid[so + j] ? 1 : 0) : 1;
    unsigned int my_patch_levels = 0;

    for (int level = a.max_level; level >= a.min_level; --level) {
      const int cols = a.width >> level, rows = a.height >> level;
      const int pitch = (int)a.pitch[level];
      const float scale = 1.0f / (float)(1 << level);
      const double dscale = (double)scale;
      const uint8_t* ref_img = a.ref_img[level] + (size_t)b * a.stride[level];
      const uint8_t* cur_img_g = a.cur_img[level] + (size_t)b * a.stride[level];
      const bool stage = a.img_in_smem[level] != 0;
      const uint8_t* cur_img = stage ? img_s : cur_img_g;
      __syncthreads();  // previous level's readers of img_s are done
      if (tid == 0) {
        if (stage) {
          const uint32_t bytes = (uint32_t)rows * (uint32_t)pitch;
          fence_proxy_async();
          mbar_expect_tx(bar, bytes);
          bulk_g2s(img_s, cur_img_g, bytes, bar);
        }
        ctl->iter = 0;
        for (int i = 0; i < 7; ++i) ctl->old_model[i] = ctl->model[i];
      }
      // ---- segment sampling at this level (:285-332) ----
      for (int j = tid; j < ns; j += kAlignThreads) {
        int N = 0;
        if (seg_alive[j]) {
          const double* spx = a.seg_spx + (so + j) * 2;
          const double* epx = a.seg_epx + (so + j) * 2;
          const int sx = (int)(spx[0] * dscale), sy = (int)(spx[1] * dscale);
          const int ex = (int)(epx[0] * dscale), ey = (int)(epx[1] * dscale);
          if (cam_in_frame(sx, sy, 3, level, a.width, a.height) && cam_in_frame(ex, ey, 3, level, a.width, a.height)) {
            double dif[2];
            N = seg_num_samples(spx, epx, a.seg_length[so + j], level, dif);
          }
        }
        seg_N[j] = N;
      }
      for (int q = tid; q < max_slots; q += kAlignThreads) slot_seg[q] = 0xffffu;
      __syncthreads();
      if (warp == 0) {  // exclusive scan of seg_N -> seg_off (cache offsets in patches, :282-292) + lane groups
        int carry = 0;
        for (int base = 0; base < ns; base += 32) {
          const int j = base + lane;
          const int v = (j < ns) ? seg_N[j] : 0;
          int incl = v;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += n;
          }
          if (j < ns) seg_off[j] = carry + incl - v;
          carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        // Lane groups: segment j gets 2^k consecutive lanes, 2^k = smallest power of two >= min(N_j, 32).
        // Groups are laid out class by class, largest first, so every group is aligned to its own size
        // and never straddles a warp.
        int slot_base = 0;
        for (int cls = 5; cls >= 0; --cls) {
          int cnt = 0;
          for (int base = 0; base < ns; base += 32) {
            const int j = base + lane;
            const int N = (j < ns) ? seg_N[j] : 0;
            int k = -1;
            if (N > 0) {
              k = 0;
              while ((1 << k) < N && k < 5) ++k;
            }
            const unsigned m = __ballot_sync(0xffffffffu, k == cls);
            if (k == cls) seg_slot[j] = slot_base + ((cnt + __popc(m & ((1u << lane) - 1u))) << cls);
            cnt += __popc(m);
          }
          slot_base += cnt << cls;
        }
        if (lane == 0) {
          ctl->n_seg_patches = carry;
          ctl->n_seg_slots = slot_base;
        }
      }
      __syncthreads();
      const int n_sp = min(ctl->n_seg_patches, a.max_seg_patches);
      const int n_patches = np + n_sp;
      const int n_seg_slots = min((ctl->n_seg_slots + 31) & ~31, max_slots & ~31);
      // ---- expand segments into sample patches: 2D centre and 3D point by repeated addition (:323-335) ----
      for (int j = tid; j < ns; j += kAlignThreads) {
        const int N = seg_N[j];
        if (N == 0) continue;
        const double* spx = a.seg_spx + (so + j) * 2;
        const double* epx = a.seg_epx + (so + j) * 2;
        double dif[2];
        seg_num_samples(spx, epx, a.seg_length[so + j], level, dif);
        const double nm1 = (double)(unsigned long long)(N - 1);
        const double inc2d0 = dif[0] * dscale / nm1, inc2d1 = dif[1] * dscale / nm1;
        double px0 = spx[0] * dscale, px1 = spx[1] * dscale;
        const double* sp = a.seg_spos + (so + j) * 3;
        const double* ep = a.seg_epos + (so + j) * 3;
        const double* sf = a.seg_sf + (so + j) * 3;
        const double* ef = a.seg_ef + (so + j) * 3;
        double d0 = sp[0] - rpx, d1 = sp[1] - rpy, d2 = sp[2] - rpz;
        const double p_depth = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
        d0 = ep[0] - rpx, d1 = ep[1] - rpy, d2 = ep[2] - rpz;
        const double q_depth = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
        const double P0 = sf[0] * p_depth, P1 = sf[1] * p_depth, P2 = sf[2] * p_depth;
        const double Q0 = ef[0] * q_depth, Q1 = ef[1] * q_depth, Q2 = ef[2] * q_depth;
        const double i0 = (Q0 - P0) / nm1, i1 = (Q1 - P1) / nm1, i2 = (Q2 - P2) / nm1;
        double X = P0, Y = P1, Z = P2;
        const int off = seg_off[j];
        {
          int g = 1;
          while (g < N && g < 32) g <<= 1;
          const int s0 = seg_slot[j];
          for (int n = 0; n < g; ++n)
            if (s0 + n < max_slots) slot_seg[s0 + n] = (uint16_t)j;
        }
        for (int n = 0; n < N; ++n) {
          const int sp_idx = off + n;
          if (sp_idx < n_sp) {
            seg_px[2 * sp_idx] = px0;
            seg_px[2 * sp_idx + 1] = px1;
            xyz[0 * MP + np + sp_idx] = X;
            xyz[1 * MP + np + sp_idx] = Y;
            xyz[2 * MP + np + sp_idx] = Z;
            xyz[3 * MP + np + sp_idx] = 1. / Z;
          }
          px0 += inc2d0, px1 += inc2d1;
          X += i0, Y += i1, Z += i2;
        }
      }
      __syncthreads();
      // ---- reference patch cache (:195-378) ----
      for (int p = tid; p < n_patches; p += kAlignThreads) {
        double u, v;
        const bool is_pt = p < np;
        if (is_pt) {
          if (a.pt_valid && !a.pt_valid[po + p]) continue;
          const double* px = a.pt_px + (po + p) * 2;
          u = px[0] * dscale, v = px[1] * dscale;
        } else {
          u = seg_px[2 * (p - np)], v = seg_px[2 * (p - np) + 1];
        }
        int ui, vi;
        float wTL, wTR, wBL, wBR;
        const bool in = patch_setup(u, v, cols, rows, 3, ui, vi, wTL, wTR, wBL, wBR);
        if (!in) {
          // points: skipped at this level (:218-219); their Jacobian columns were zeroed (:85).
          // segment samples are inside by construction; guard only protects against malformed input.
          if (!is_pt || pt_vis[p]) zero_gradients(cache, MP, p);
          continue;
        }
        if (is_pt) pt_vis[p] = 1;
        precompute_patch(ref_img, pitch, ui, vi, wTL, wTR, wBL, wBR, cache, MP, p);
        ++my_patch_levels;
      }
      if (stage) {
        mbar_wait(bar, bar_parity);
        bar_parity ^= 1u;
      }
      __syncthreads();

      // ---- Gauss-Newton iterations at this level (vk::NLLSSolver::optimizeGaussNewton) ----
      const double cJ = fabs(a.fx) / (double)(1 << level);  // focal_length / 2^level (:262)
      const double cJ2 = cJ * cJ;
      for (;;) {
        const double R0 = ctl->R[0], R1 = ctl->R[1], R2 = ctl->R[2], R3 = ctl->R[3], R4 = ctl->R[4], R5 = ctl->R[5],
                     R6 = ctl->R[6], R7 = ctl->R[7], R8 = ctl->R[8];
        const double t0 = ctl->t[0], t1 = ctl->t[1], t2 = ctl->t[2];
        double chi2_acc = 0.0;
        int n_meas_acc = 0, n_patch_acc = 0;
        // ======== phase 1: residuals.  Each thread evaluates its patches and leaves five fp32 sums per
        // patch in shared memory; the 27 double accumulators are not live here, which keeps the pixel
        // loop free of register spills. ========
        // ---- point patches (:380-502): thread per patch ----
        for (int p = tid; p < np; p += kAlignThreads) {
          int kind = 0;
          if (pt_vis[p]) {
            const double X = xyz[0 * MP + p], Y = xyz[1 * MP + p], Z = xyz[2 * MP + p];
            const double xc = R0 * X + R1 * Y + R2 * Z + t0;
            const double yc = R3 * X + R4 * Y + R5 * Z + t1;
            const double zc = R6 * X + R7 * Y + R8 * Z + t2;
            const double izc = 1.0 / zc;
            const double u = (a.fx * (xc * izc) + a.cx) * dscale;  // world2cam(xyz)*scale (:425)
            const double v = (a.fy * (yc * izc) + a.cy) * dscale;
            float S[5], aux;
            if (eval_patch<true>(cur_img, pitch, cols, rows, cache, MP, p, u, v, S, aux)) {
              kind = 1;
#pragma unroll
              for (int k = 0; k < 5; ++k) prec[p].S[k] = S[k];
              chi2_acc += (double)aux;
              n_meas_acc += 16;
              n_patch_acc += 1;
            }
          }
          prec[p].kind = kind;
        }
        // ---- segment samples (:504-695): every segment owns a group of G = 2^gshift consecutive lanes
        // of one warp (G >= its sample count, or the whole warp looping over samples), so the
        // per-segment gate/weight (:640-688) is a few shuffles: no block barrier.
        // Warps take segment rounds from the top so they interleave with the point rounds.
        for (int base = (kAlignWarps - 1 - warp) * 32; base < n_seg_slots; base += kAlignThreads) {
          const int q = base + lane;
          const int j = slot_seg[q];
          const bool seg_ok = (j < ns) && seg_alive[j];
          const int N = seg_ok ? seg_N[j] : 0;
          const int off = seg_ok ? seg_off[j] : 0;
          const int n0 = seg_ok ? q - seg_slot[j] : 0;
          int G = 1;
          while (G < N && G < 32) G <<= 1;
          float my_abs = 0.f;
          int first_bad = 0x7fffffff;
          for (int n = n0; n < N; n += G) {  // one trip unless a segment has more samples than a warp
            const int p = np + off + n;
            const double X = xyz[0 * MP + p], Y = xyz[1 * MP + p], Z = xyz[2 * MP + p];
            const double xc = R0 * X + R1 * Y + R2 * Z + t0;
            const double yc = R3 * X + R4 * Y + R5 * Z + t1;
            const double zc = R6 * X + R7 * Y + R8 * Z + t2;
            const double izc = 1.0 / zc;
            const double u = (a.fx * (xc * izc) + a.cx) * dscale;
            const double v = (a.fy * (yc * izc) + a.cy) * dscale;
            float S[5], aux;
            if (eval_patch<false>(cur_img, pitch, cols, rows, cache, MP, p, u, v, S, aux)) {
              my_abs = __fadd_rn(my_abs, aux);
#pragma unroll
              for (int k = 0; k < 5; ++k) prec[p].S[k] = S[k];
              prec[p].kind = 2 + j;
            } else {
              first_bad = min(first_bad, n);
              prec[p].kind = 0;
            }
          }
          // group reductions (xor tree inside the group: partners at distance d < G stay in the group)
          float res_ = my_abs;
#pragma unroll
          for (int d = 16; d >= 1; d >>= 1) {
            const float o = __shfl_xor_sync(0xffffffffu, res_, d);
            const int fb = __shfl_xor_sync(0xffffffffu, first_bad, d);
            if (d < G) {
              res_ = __fadd_rn(res_, o);
              first_bad = min(first_bad, fb);
            }
          }
          if (N == 0 || n0 != 0) continue;  // the group's first lane settles the segment
          const bool good = first_bad >= N;
          n_patch_acc += good ? N : first_bad;  // samples evaluated before the loop stops (:588-594)
          res_ = (float)((double)res_ / (double)(unsigned long long)N);  // :647
          if (good && (double)res_ < 200.0) {
            const float w = (float)(1.0 / (1.0 + (double)res_));  // :675
            seg_scale[2 * j] = (double)w / (double)res_ * cJ2;    // H += H_*weight/res_ (:681)
            seg_scale[2 * j + 1] = (double)w * cJ;                // Jres += Jres_*weight (:682)
            chi2_acc += (double)__fmul_rn(__fmul_rn(res_, res_), w);  // :683
            n_meas_acc += 1;                                         // :684
          } else {
            seg_scale[2 * j] = 0.0;  // rejected: its samples are skipped in phase 2
            seg_scale[2 * j + 1] = 0.0;
            seg_alive[j] = 0;  // it->feat3D = NULL (:688); the group's lanes have all read it already
          }
        }
        __syncwarp();  // phase 2 reads back what lanes of this warp wrote (records of own patches, seg_scale)
        // ======== phase 2: normal equations.  Rank-2 update of this thread's 21+6 accumulators per patch. ========
        double acc[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = 0.0;
        for (int p = tid; p < np; p += kAlignThreads) {
          if (prec[p].kind != 1) continue;
          const PatchSums ps = prec[p];
          rank2_update(acc, xyz[0 * MP + p], xyz[1 * MP + p], xyz[3 * MP + p], (double)ps.S[0] * cJ2,
                       (double)ps.S[1] * cJ2, (double)ps.S[2] * cJ2, (double)ps.S[3] * cJ, (double)ps.S[4] * cJ);
        }
        for (int base = (kAlignWarps - 1 - warp) * 32; base < n_seg_slots; base += kAlignThreads) {
          const int q = base + lane;
          const int j = slot_seg[q];
          if (j >= ns) continue;
          const int N = seg_N[j];
          const double sH = seg_scale[2 * j], sJ = seg_scale[2 * j + 1];
          if (sH == 0.0 && sJ == 0.0) continue;
          int G = 1;
          while (G < N && G < 32) G <<= 1;
          for (int n = q - seg_slot[j]; n < N; n += G) {
            const int p = np + seg_off[j] + n;
            if (prec[p].kind != 2 + j) continue;
            const PatchSums ps = prec[p];
            rank2_update(acc, xyz[0 * MP + p], xyz[1 * MP + p], xyz[3 * MP + p], (double)ps.S[0] * sH, (double)ps.S[1] * sH,
                         (double)ps.S[2] * sH, (double)ps.S[3] * sJ, (double)ps.S[4] * sJ);
          }
        }
        acc[27] = chi2_acc;
        acc[28] = (double)n_meas_acc;
        acc[29] = (double)n_patch_acc;
        // ---- block reduction (deterministic order) ----
        const double mine = warp_reduce32(acc, lane);
        red[warp * 32 + lane] = mine;
        __syncthreads();
        if (warp == 0) {
          double s = 0.0;
#pragma unroll
          for (int w = 0; w < kAlignWarps; ++w) s += red[w * 32 + lane];
          tot[lane] = s;
          __syncwarp();
          if (lane == 0) gn_step(ctl, tot, level, a.n_iter, a.eps);
        }
        __syncthreads();
        if (ctl->flag) break;
      }
    }  // levels

    // ---- results ----
    {
      unsigned int v = my_patch_levels;
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
      if (lane == 0 && v) atomicAdd(&ctl->patch_levels, v);
    }
    for (int j = tid; j < a.n_segs; j += kAlignThreads) {
      const bool valid0 = (j < ns) && (a.seg_valid ? a.seg_valid[so + j] != 0 : true);
      a.out_seg_killed[so + j] = (valid0 && !seg_alive[j]) ? 1 : 0;
    }
    __syncthreads();
    if (tid == 0) {
      SE3q model, T_ref;
      model.q.x = ctl->model[0], model.q.y = ctl->model[1], model.q.z = ctl->model[2], model.q.w = ctl->model[3];
      model.t = v3(ctl->model[4], ctl->model[5], ctl->model[6]);
      T_ref.q.x = ctl->T_ref[0], T_ref.q.y = ctl->T_ref[1], T_ref.q.z = ctl->T_ref[2], T_ref.q.w = ctl->T_ref[3];
      T_ref.t = v3(ctl->T_ref[4], ctl->T_ref[5], ctl->T_ref[6]);
      const SE3q T_cur = se3_mul(model, T_ref);  // :92
      se3_store(T_cur, a.out_T + (size_t)b * 7);
      a.out_n_tracked[b] = ctl->n_meas_last / 16;  // :94
      for (int i = 0; i < 36; ++i) a.out_H[(size_t)b * 36 + i] = ctl->H_last[i];
      for (int l = 0; l < PLSVO_MAX_LEVELS; ++l) a.out_iters[(size_t)b * PLSVO_MAX_LEVELS + l] = ctl->iters_level[l];
      a.out_status[b] = ctl->stop ? 2 : 0;
      a.out_patch_iters[b] = ctl->patch_iters;
      a.out_patch_levels[b] = ctl->patch_levels;
    }
  }
}

}  // namespace

namespace {
__global__ void weight_selftest_kernel(uint32_t n, uint32_t seed, unsigned long long* mismatch) {
  unsigned long long bad = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    // half of the samples: random bit patterns in [0,256); other half: residual-like values k/2^m
    uint32_t h = (i ^ seed) * 2654435761u;
    h ^= h >> 15;
    h *= 2246822519u;
    h ^= h >> 13;
    float a;
    if (i & 1) {
      a = __uint_as_float(h % 0x43800000u);  // all floats in [0,256)
    } else {
      a = (float)(h & 0xffffff) * (1.0f / 65536.0f);  // multiples of 2^-16 below 256
    }
    const float fast = weight_rcp(a);
    const float ref = (float)(1.0 / (1.0 + (double)a));
    if (__float_as_uint(fast) != __float_as_uint(ref)) ++bad;
  }
  if (bad) atomicAdd(mismatch, bad);
}
}  // namespace

cudaError_t weight_selftest_launch(uint32_t n, uint32_t seed, unsigned long long* d_mismatch, cudaStream_t s) {
  weight_selftest_kernel<<<592, 256, 0, s>>>(n, seed, d_mismatch);
  return cudaGetLastError();
}

size_t align_smem_bytes(int n_pts, int n_segs, int max_patches, int max_seg_patches, int img_bytes,
                        bool cache_in_smem) {
  return make_layout(n_pts, n_segs, max_patches, max_seg_patches, img_bytes, cache_in_smem).total;
}

namespace {
template <bool CS, int NT>
cudaError_t prepare_t(size_t smem_bytes, int* ctas_per_sm) {
  cudaError_t e = cudaFuncSetAttribute(sparse_img_align_kernel<CS, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem_bytes);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(sparse_img_align_kernel<CS, NT>, cudaFuncAttributePreferredSharedMemoryCarveout,
                           cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, sparse_img_align_kernel<CS, NT>, NT, smem_bytes);
}
}  // namespace

cudaError_t align_kernel_prepare(bool cache_in_smem, int threads, size_t smem_bytes, int* ctas_per_sm) {
  switch (threads) {
    case 64:
      return cache_in_smem ? prepare_t<true, 64>(smem_bytes, ctas_per_sm) : prepare_t<false, 64>(smem_bytes, ctas_per_sm);
    case 128:
      return cache_in_smem ? prepare_t<true, 128>(smem_bytes, ctas_per_sm) : prepare_t<false, 128>(smem_bytes, ctas_per_sm);
    case 256:
      return cache_in_smem ? prepare_t<true, 256>(smem_bytes, ctas_per_sm) : prepare_t<false, 256>(smem_bytes, ctas_per_sm);
    default:
      return cudaErrorInvalidValue;
  }
}

cudaError_t align_kernel_launch(const AlignArgs& a, int grid, int threads, size_t smem_bytes, bool cache_in_smem,
                                cudaStream_t s) {
#define PLSVO_LAUNCH(CS, NT) sparse_img_align_kernel<CS, NT><<<grid, NT, smem_bytes, s>>>(a)
  switch (threads) {
    case 64:
      if (cache_in_smem) PLSVO_LAUNCH(true, 64); else PLSVO_LAUNCH(false, 64);
      break;
    case 128:
      if (cache_in_smem) PLSVO_LAUNCH(true, 128); else PLSVO_LAUNCH(false, 128);
      break;
    case 256:
      if (cache_in_smem) PLSVO_LAUNCH(true, 256); else PLSVO_LAUNCH(false, 256);
      break;
    default:
      return cudaErrorInvalidValue;
  }
#undef PLSVO_LAUNCH
  return cudaGetLastError();
}

}  // namespace plsvo
