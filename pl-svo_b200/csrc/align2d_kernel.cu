// align2d_kernel.cu — feature_alignment::align2D (src/feature_alignment.cpp:160-290): 8x8 inverse-
// compositional Lucas-Kanade refinement of a feature position (2 DoF + mean intensity offset), the
// per-feature kernel of Matcher::findMatchDirect (src/matcher.cpp:201,257,268).  SURVEY.md §8f rank 1
// ("next"): the step between the two hot-path calls.
//
// One thread per feature keeps the reference's sequential fp32 accumulation order over the 64 pixels,
// so positions and convergence flags are bit-identical to the scalar reference code (the SSE2/NEON
// variants of the reference use fixed-point arithmetic and differ from its own scalar path).
// The 10x10 reference patch with border and the 8x8 reference patch of a feature live in shared
// memory; template gradients are recomputed from the border patch (two byte subtractions) instead
// of being cached.  Image bytes come straight from global memory / L2 (9x9 footprint per iteration).
#include <cuda_runtime.h>
#include <stdint.h>

#include "exact_math.cuh"
#include "internal.h"

namespace plsvo {
namespace {

constexpr int kA2Threads = 128;

// Exact u8 -> f32 and (u8 - u8) -> f32 without the conversion unit: 0x4B000000 | v is the float 2^23 + v, so one
// FADD gives v exactly; differences of two such values (|d| <= 255) and their halves are exact as well, i.e. the
// same numbers as the reference's int -> float / double conversions, produced on the FMA pipe instead of XU.
__device__ __forceinline__ float u8f(uint8_t v) { return __fsub_rn(__uint_as_float(0x4B000000u | (uint32_t)v), 8388608.0f); }
__device__ __forceinline__ float u8diff(uint8_t a, uint8_t b) {
  return __fsub_rn(__uint_as_float(0x4B000000u | (uint32_t)a), __uint_as_float(0x4B000000u | (uint32_t)b));
}

// Nine consecutive image bytes starting at an arbitrary address, as exact floats: three aligned 32-bit loads and
// two funnel shifts instead of nine byte loads (the image buffers are 256-byte aligned with 256 bytes of slack, so
// the aligned window never leaves the allocation).  Byte -> float through PRMT + FADD (see u8f).
__device__ __forceinline__ float bytef(uint32_t w, int k) {
  return __fsub_rn(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7440 | k)), 8388608.0f);
}
__device__ __forceinline__ void load_row9(const uint8_t* p, float f[9]) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~static_cast<uintptr_t>(3));
  const uint32_t sh = static_cast<uint32_t>(a & 3) * 8;
  const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
  const uint32_t lo = __funnelshift_r(w0, w1, sh), mid = __funnelshift_r(w1, w2, sh), hi = w2 >> sh;
#pragma unroll
  for (int k = 0; k < 4; ++k) f[k] = bytef(lo, k), f[4 + k] = bytef(mid, k);
  f[8] = bytef(hi, 0);
}

// align2D on one feature.  border = 10x10 reference patch with border, ref = 8x8 reference patch with row step
// ref_step (8 for a packed patch, 10 when it is the interior of `border`).  u,v in/out; returns `converged`.
__device__ __forceinline__ bool align2d_core(const uint8_t* border, const uint8_t* ref, const int ref_step, const uint8_t* img,
                                             const int cur_step, const int cols, const int rows, const int n_iter, float& u,
                                             float& v) {
  // ---- template Hessian (:183-201): J = (0.5*dx, 0.5*dy, 1), H = sum J J^T (exact in fp32) ----
  float H00 = 0, H01 = 0, H02 = 0, H11 = 0, H12 = 0, H22 = 0;
  for (int y = 0; y < 8; ++y) {
    const uint8_t* it = border + (y + 1) * 10 + 1;
    for (int x = 0; x < 8; ++x, ++it) {
      const float J0 = __fmul_rn(0.5f, u8diff(it[1], it[-1]));
      const float J1 = __fmul_rn(0.5f, u8diff(it[10], it[-10]));
      H00 = __fadd_rn(H00, __fmul_rn(J0, J0));
      H01 = __fadd_rn(H01, __fmul_rn(J0, J1));
      H02 = __fadd_rn(H02, J0);
      H11 = __fadd_rn(H11, __fmul_rn(J1, J1));
      H12 = __fadd_rn(H12, J1);
      H22 = __fadd_rn(H22, 1.0f);
    }
  }
  // ---- Hinv = H.inverse(): Eigen's fixed-size 3x3 path (cofactors of column 0, determinant, 1/det) ----
  const float m00 = H00, m01 = H01, m02 = H02, m10 = H01, m11 = H11, m12 = H12, m20 = H02, m21 = H12, m22 = H22;
#define COF(i1, j1, i2, j2, i3, j3, i4, j4) __fsub_rn(__fmul_rn(m##i1##j1, m##i2##j2), __fmul_rn(m##i3##j3, m##i4##j4))
  const float c00 = COF(1, 1, 2, 2, 1, 2, 2, 1);  // cofactor<0,0>
  const float c10 = COF(2, 1, 0, 2, 2, 2, 0, 1);  // cofactor<1,0>
  const float c20 = COF(0, 1, 1, 2, 0, 2, 1, 1);  // cofactor<2,0>
  const float det = __fadd_rn(__fadd_rn(__fmul_rn(c00, m00), __fmul_rn(c10, m10)), __fmul_rn(c20, m20));
  const float invdet = __fdiv_rn(1.0f, det);
  const float c01 = COF(1, 2, 2, 0, 1, 0, 2, 2);  // cofactor<0,1>
  const float c11 = COF(2, 2, 0, 0, 2, 0, 0, 2);  // cofactor<1,1>
  const float c21 = COF(0, 2, 1, 0, 0, 0, 1, 2);  // cofactor<2,1>
  const float c02 = COF(1, 0, 2, 1, 1, 1, 2, 0);  // cofactor<0,2>
  const float c12 = COF(2, 0, 0, 1, 2, 1, 0, 0);  // cofactor<1,2>
  const float c22 = COF(0, 0, 1, 1, 0, 1, 1, 0);  // cofactor<2,2>
#undef COF
  // result.row(0) = cofactors_col0 * invdet ; result(1,0)=c01*invdet ; (1,1)=c11 ; (1,2)=c21 ; (2,0)=c02 ; (2,1)=c12 ; (2,2)=c22
  const float I00 = __fmul_rn(c00, invdet), I01 = __fmul_rn(c10, invdet), I02 = __fmul_rn(c20, invdet);
  const float I10 = __fmul_rn(c01, invdet), I11 = __fmul_rn(c11, invdet), I12 = __fmul_rn(c21, invdet);
  const float I20 = __fmul_rn(c02, invdet), I21 = __fmul_rn(c12, invdet), I22 = __fmul_rn(c22, invdet);

  float mean_diff = 0.f;
  const float min_update_squared = (float)(0.03 * 0.03);
  bool converged = false;
  for (int iter = 0; iter < n_iter; ++iter) {
    // Patch::setPosition / isInFrame(halfsize=4) / computeInterpWeights (src/feature.cpp:189-208)
    const float fu = floorf(u), fv = floorf(v);
    const int ui = (int)fu, vi = (int)fv;
    if (ui < 4 || vi < 4 || ui >= cols - 4 || vi >= rows - 4) break;
    const float su = __fsub_rn(u, fu), sv = __fsub_rn(v, fv);
    const float wTL = (float)((1.0 - (double)su) * (1.0 - (double)sv));
    const float wTR = (float)((double)su * (1.0 - (double)sv));
    const float wBL = (float)((1.0 - (double)su) * (double)sv);
    const float wBR = __fmul_rn(su, sv);
    float J0 = 0.f, J1 = 0.f, J2 = 0.f;
    const uint8_t* row = img + (size_t)(vi - 4) * cur_step + (ui - 4);
    const uint8_t* ref_row = ref;
    float top[9], bot[9];
    load_row9(row, top);
#pragma unroll 1
    for (int y = 0; y < 8; ++y, ref_row += ref_step) {
      row += cur_step;
      load_row9(row, bot);
      const uint8_t* itb = border + (y + 1) * 10 + 1;
      const uint8_t* it_ref = ref_row;
#pragma unroll
      for (int x = 0; x < 8; ++x, ++it_ref, ++itb) {
        const float search_pixel =
            __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(wTL, top[x]), __fmul_rn(wTR, top[x + 1])), __fmul_rn(wBL, bot[x])),
                      __fmul_rn(wBR, bot[x + 1]));
        const float res = __fadd_rn(__fsub_rn(search_pixel, u8f(*it_ref)), mean_diff);
        const float dx = __fmul_rn(0.5f, u8diff(itb[1], itb[-1]));
        const float dy = __fmul_rn(0.5f, u8diff(itb[10], itb[-10]));
        J0 = __fsub_rn(J0, __fmul_rn(res, dx));
        J1 = __fsub_rn(J1, __fmul_rn(res, dy));
        J2 = __fsub_rn(J2, res);
      }
#pragma unroll
      for (int k = 0; k < 9; ++k) top[k] = bot[k];
    }
    // update = Hinv * Jres
    const float up0 = __fadd_rn(__fadd_rn(__fmul_rn(I00, J0), __fmul_rn(I01, J1)), __fmul_rn(I02, J2));
    const float up1 = __fadd_rn(__fadd_rn(__fmul_rn(I10, J0), __fmul_rn(I11, J1)), __fmul_rn(I12, J2));
    const float up2 = __fadd_rn(__fadd_rn(__fmul_rn(I20, J0), __fmul_rn(I21, J1)), __fmul_rn(I22, J2));
    u = __fadd_rn(u, up0);
    v = __fadd_rn(v, up1);
    mean_diff = __fadd_rn(mean_diff, up2);
    if (__fadd_rn(__fmul_rn(up0, up0), __fmul_rn(up1, up1)) < min_update_squared) {
      converged = true;
      break;
    }
  }
  return converged;
}

// align1D on one feature (same conventions as align2d_core); d0,d1 = direction, h_inv out.
__device__ __forceinline__ bool align1d_core(const uint8_t* border, const uint8_t* ref, const int ref_step, const uint8_t* img,
                                             const int cur_step, const int cols, const int rows, const int n_iter, const float d0,
                                             const float d1, float& u, float& v, double& h_inv) {
  // directional template derivative (:63-66): J0 = 0.5*(dir0*(I[x+1]-I[x-1]) + dir1*(I[y+1]-I[y-1])) in float, J1 = 1
  auto dv_at = [&](const uint8_t* it) {
    const float gx = __fmul_rn(d0, u8diff(it[1], it[-1]));
    const float gy = __fmul_rn(d1, u8diff(it[10], it[-10]));
    return __fmul_rn(0.5f, __fadd_rn(gx, gy));
  };
  float H00 = 0, H01 = 0, H11 = 0;
  for (int y = 0; y < 8; ++y) {
    const uint8_t* it = border + (y + 1) * 10 + 1;
    for (int x = 0; x < 8; ++x, ++it) {
      const float J0 = dv_at(it);
      H00 = __fadd_rn(H00, __fmul_rn(J0, J0));
      H01 = __fadd_rn(H01, J0);
      H11 = __fadd_rn(H11, 1.0f);
    }
  }
  h_inv = 1.0 / (double)H00 * 8 * 8;  // :75
  // Matrix2f::inverse(): 1/det, (d, -c; -b, a) * invdet
  const float det = __fsub_rn(__fmul_rn(H00, H11), __fmul_rn(H01, H01));
  const float invdet = __fdiv_rn(1.0f, det);
  const float I00 = __fmul_rn(H11, invdet), I10 = __fmul_rn(-H01, invdet);
  const float I01 = __fmul_rn(-H01, invdet), I11 = __fmul_rn(H00, invdet);

  float mean_diff = 0.f;
  const float min_update_squared = (float)(0.03 * 0.03);
  float chi2 = 0.f, up0 = 0.f, up1 = 0.f;
  bool converged = false;
  for (int iter = 0; iter < n_iter; ++iter) {
    const float fu = floorf(u), fv = floorf(v);
    const int ui = (int)fu, vi = (int)fv;
    if (ui < 4 || vi < 4 || ui >= cols - 4 || vi >= rows - 4) break;  // NaN never passes this test (:87-92)
    const float su = __fsub_rn(u, fu), sv = __fsub_rn(v, fv);
    const float wTL = (float)((1.0 - (double)su) * (1.0 - (double)sv));
    const float wTR = (float)((double)su * (1.0 - (double)sv));
    const float wBL = (float)((1.0 - (double)su) * (double)sv);
    const float wBR = __fmul_rn(su, sv);
    float J0 = 0.f, J1 = 0.f, new_chi2 = 0.f;
    const uint8_t* row = img + (size_t)(vi - 4) * cur_step + (ui - 4);
    const uint8_t* ref_row = ref;
    float top[9], bot[9];
    load_row9(row, top);
#pragma unroll 1
    for (int y = 0; y < 8; ++y, ref_row += ref_step) {
      row += cur_step;
      load_row9(row, bot);
      const uint8_t* itb = border + (y + 1) * 10 + 1;
      const uint8_t* it_ref = ref_row;
#pragma unroll
      for (int x = 0; x < 8; ++x, ++it_ref, ++itb) {
        const float search_pixel =
            __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(wTL, top[x]), __fmul_rn(wTR, top[x + 1])), __fmul_rn(wBL, bot[x])),
                      __fmul_rn(wBR, bot[x + 1]));
        const float res = __fadd_rn(__fsub_rn(search_pixel, u8f(*it_ref)), mean_diff);
        J0 = __fsub_rn(J0, __fmul_rn(res, dv_at(itb)));
        J1 = __fsub_rn(J1, res);
        new_chi2 = __fadd_rn(new_chi2, __fmul_rn(res, res));
      }
#pragma unroll
      for (int k = 0; k < 9; ++k) top[k] = bot[k];
    }
    if (iter > 0 && new_chi2 > chi2) {  // :124-132 (the back-off subtracts the raw update, as the reference does)
      u = __fsub_rn(u, up0);
      v = __fsub_rn(v, up1);
      break;
    }
    chi2 = new_chi2;
    up0 = __fadd_rn(__fmul_rn(I00, J0), __fmul_rn(I01, J1));
    up1 = __fadd_rn(__fmul_rn(I10, J0), __fmul_rn(I11, J1));
    u = __fadd_rn(u, __fmul_rn(up0, d0));
    v = __fadd_rn(v, __fmul_rn(up0, d1));
    mean_diff = __fadd_rn(mean_diff, up1);
    if (__fadd_rn(__fmul_rn(up0, up0), __fmul_rn(up1, up1)) < min_update_squared) {
      converged = true;
      break;
    }
  }
  return converged;
}

__global__ void __launch_bounds__(kA2Threads) align2d_kernel(const Align2DArgs a) {
  __shared__ __align__(4) uint8_t s_border[kA2Threads][108];  // 100 used; 27-word pitch (odd) keeps the threads of a warp on distinct banks
  __shared__ __align__(4) uint8_t s_ref[kA2Threads][68];  // 64 used; 17-word pitch (odd): no bank conflicts
  const int tid = threadIdx.x;
  const int i = blockIdx.x * kA2Threads + tid;
  const bool active = i < a.n;
  if (active) {
    const uint32_t* gb = reinterpret_cast<const uint32_t*>(a.ref_patch_with_border + (size_t)i * 100);
    const uint32_t* gr = reinterpret_cast<const uint32_t*>(a.ref_patch + (size_t)i * 64);
    uint32_t* sb = reinterpret_cast<uint32_t*>(s_border[tid]);
    uint32_t* sr = reinterpret_cast<uint32_t*>(s_ref[tid]);
#pragma unroll
    for (int k = 0; k < 25; ++k) sb[k] = gb[k];
#pragma unroll
    for (int k = 0; k < 16; ++k) sr[k] = gr[k];
  }
  if (!active) return;
  const uint8_t* border = s_border[tid];
  const uint8_t* ref = s_ref[tid];
  const int level = a.level[i];
  const int cols = a.width >> level, rows = a.height >> level;
  const int cur_step = (int)a.pitch[level];
  const uint8_t* img = a.img[level] + (size_t)a.image_index[i] * a.stride[level];

  float u = (float)a.px[2 * (size_t)i], v = (float)a.px[2 * (size_t)i + 1];
  const bool converged = align2d_core(border, ref, 8, img, cur_step, cols, rows, a.n_iter, u, v);
  a.out_px[2 * (size_t)i] = (double)u;
  a.out_px[2 * (size_t)i + 1] = (double)v;
  a.out_converged[i] = converged ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
// feature_alignment::align1D (src/feature_alignment.cpp:40-157): the patch moves along `dir` only
// (edgelets), 1 DoF + mean intensity offset, with the reference's chi2 back-off.  Same layout and
// the same bit-exact fp32 sequencing as align2d_kernel.
__global__ void __launch_bounds__(kA2Threads) align1d_kernel(const Align2DArgs a) {
  __shared__ __align__(4) uint8_t s_border[kA2Threads][108];  // 100 used; 27-word pitch (odd) keeps the threads of a warp on distinct banks
  __shared__ __align__(4) uint8_t s_ref[kA2Threads][68];  // 64 used; 17-word pitch (odd): no bank conflicts
  const int tid = threadIdx.x;
  const int i = blockIdx.x * kA2Threads + tid;
  if (i >= a.n) return;
  {
    const uint32_t* gb = reinterpret_cast<const uint32_t*>(a.ref_patch_with_border + (size_t)i * 100);
    const uint32_t* gr = reinterpret_cast<const uint32_t*>(a.ref_patch + (size_t)i * 64);
    uint32_t* sb = reinterpret_cast<uint32_t*>(s_border[tid]);
    uint32_t* sr = reinterpret_cast<uint32_t*>(s_ref[tid]);
#pragma unroll
    for (int k = 0; k < 25; ++k) sb[k] = gb[k];
#pragma unroll
    for (int k = 0; k < 16; ++k) sr[k] = gr[k];
  }
  const uint8_t* border = s_border[tid];
  const uint8_t* ref = s_ref[tid];
  const int level = a.level[i];
  const int cols = a.width >> level, rows = a.height >> level;
  const int cur_step = (int)a.pitch[level];
  const uint8_t* img = a.img[level] + (size_t)a.image_index[i] * a.stride[level];
  const float d0 = a.dir[2 * (size_t)i], d1 = a.dir[2 * (size_t)i + 1];

  float u = (float)a.px[2 * (size_t)i], v = (float)a.px[2 * (size_t)i + 1];
  double h_inv;
  const bool converged = align1d_core(border, ref, 8, img, cur_step, cols, rows, a.n_iter, d0, d1, u, v, h_inv);
  a.out_h_inv[i] = h_inv;
  a.out_px[2 * (size_t)i] = (double)u;
  a.out_px[2 * (size_t)i + 1] = (double)v;
  a.out_converged[i] = converged ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
// Matcher::findMatchDirect(const Point&, const Frame&, Vector2d&) after getCloseViewObs (src/matcher.cpp:159-211):
// in-frame test, warp::getWarpMatrixAffine (:42-71), getBestSearchLevel (:73-87), warp::warpAffine (:89-133,
// vk::interpolateMat_8u), createPatchFromPatchWithBorder (:148-157), align2D / align1D at the search level.
// One thread per candidate.  The double-precision geometry is written with explicit round-to-nearest intrinsics
// so that the compiler cannot contract a*b+c into an FMA: A_cur_ref, the search level, every byte of the warped
// patch and therefore the refined position are bit-identical to the reference's scalar code.
__global__ void __launch_bounds__(kA2Threads) match_direct_kernel(const MatchArgs a) {
  __shared__ __align__(4) uint8_t s_border[kA2Threads][108];  // 100 used; 27-word pitch (odd) keeps the threads of a warp on distinct banks
  const int tid = threadIdx.x;
  const int i = blockIdx.x * kA2Threads + tid;
  if (i >= a.n) return;
  const size_t I = (size_t)i;
  const double px_ref0 = a.ref_px[2 * I], px_ref1 = a.ref_px[2 * I + 1];
  const int level_ref = a.ref_level[i];
  const int r = a.ref_index[i], c = a.cur_index[i];
  a.out_px[2 * I] = a.px_cur[2 * I];
  a.out_px[2 * I + 1] = a.px_cur[2 * I + 1];
  a.out_success[i] = 0;
  a.out_level[i] = -1;
  // :169-171  cam.isInFrame(px.cast<int>() / (1 << level), halfpatch_size_ + 2, level)
  {
    const int ox = (int)px_ref0 / (1 << level_ref), oy = (int)px_ref1 / (1 << level_ref);
    const int b = 6;
    if (!(ox >= b && ox < a.width / (1 << level_ref) - b && oy >= b && oy < a.height / (1 << level_ref) - b)) return;
  }
  const Pose T_w_ref = pose_inverse(pose_load(a.T_ref_w + 7 * (size_t)r));
  const Pose T_cur_ref = pose_mul(pose_load(a.T_cur_w + 7 * (size_t)c), T_w_ref);
  const V3 pos{a.pos[3 * I], a.pos[3 * I + 1], a.pos[3 * I + 2]};
  const V3 f_ref{a.ref_f[3 * I], a.ref_f[3 * I + 1], a.ref_f[3 * I + 2]};
  const double depth_ref = v_norm(v_sub(T_w_ref.t, pos));
  auto cam2world = [&](double u, double v) {  // PinholeCamera::cam2world, undistorted, then normalized()
    const V3 xyz{DD(DS(u, a.cx), a.fx), DD(DS(v, a.cy), a.fy), 1.0};
    const double n = v_norm(xyz);
    return V3{DD(xyz.x, n), DD(xyz.y, n), DD(xyz.z, n)};
  };
  auto world2cam = [&](V3 p, double& u, double& v) {  // world2cam(project2d(xyz))
    u = DA(DM(a.fx, DD(p.x, p.z)), a.cx);
    v = DA(DM(a.fy, DD(p.y, p.z)), a.cy);
  };
  // ---- warp::getWarpMatrixAffine ----
  double A00, A01, A10, A11;
  {
    const V3 xyz_ref = v_scale(f_ref, depth_ref);
    const double scale_ref = (double)(1 << level_ref);
    const double step = DM(5.0, scale_ref), zero = DM(0.0, scale_ref);
    V3 xyz_du = cam2world(DA(px_ref0, step), DA(px_ref1, zero));
    V3 xyz_dv = cam2world(DA(px_ref0, zero), DA(px_ref1, step));
    xyz_du = v_scale(xyz_du, DD(xyz_ref.z, xyz_du.z));
    xyz_dv = v_scale(xyz_dv, DD(xyz_ref.z, xyz_dv.z));
    double pc0, pc1, pu0, pu1, pv0, pv1;
    world2cam(pose_act(T_cur_ref, xyz_ref), pc0, pc1);
    world2cam(pose_act(T_cur_ref, xyz_du), pu0, pu1);
    world2cam(pose_act(T_cur_ref, xyz_dv), pv0, pv1);
    A00 = DD(DS(pu0, pc0), 5.0), A10 = DD(DS(pu1, pc1), 5.0);
    A01 = DD(DS(pv0, pc0), 5.0), A11 = DD(DS(pv1, pc1), 5.0);
  }
  // ---- warp::getBestSearchLevel ----
  const double det = DS(DM(A00, A11), DM(A10, A01));
  int search_level = 0;
  {
    double D = det;
    const int max_level = a.n_pyr_levels - 1;
    while (D > 3.0 && search_level < max_level) {
      search_level += 1;
      D = DM(D, 0.25);
    }
  }
  a.out_level[i] = search_level;
  // ---- warp::warpAffine into the 10x10 patch with border ----
  uint8_t* border = s_border[tid];
  {
    const double invdet = DD(1.0, det);
    const float R00 = (float)DM(A11, invdet), R10 = (float)DM(-A10, invdet);
    const float R01 = (float)DM(-A01, invdet), R11 = (float)DM(A00, invdet);
    const bool bad = isnan(R00);  // "Affine warp is NaN": the reference leaves the (zeroed) patch untouched
    const int cols = a.width >> level_ref, rows = a.height >> level_ref;
    const uint8_t* img = a.ref_img[level_ref] + (size_t)r * a.ref_stride[level_ref];
    const int stride = (int)a.ref_pitch[level_ref];
    const float fs = (float)(1 << level_ref);
    const float pr0 = __fdiv_rn((float)px_ref0, fs), pr1 = __fdiv_rn((float)px_ref1, fs);
    const float ss = (float)(1 << search_level);
    const float xmax = (float)(cols - 1), ymax = (float)(rows - 1);
    for (int y = 0; y < 10; ++y) {
      const float p1 = __fmul_rn((float)(y - 5), ss);
      for (int x = 0; x < 10; ++x) {
        uint8_t val = 0;
        if (!bad) {
          const float p0 = __fmul_rn((float)(x - 5), ss);
          const float q0 = __fadd_rn(__fadd_rn(__fmul_rn(R00, p0), __fmul_rn(R01, p1)), pr0);
          const float q1 = __fadd_rn(__fadd_rn(__fmul_rn(R10, p0), __fmul_rn(R11, p1)), pr1);
          if (!(q0 < 0 || q1 < 0 || q0 >= xmax || q1 >= ymax)) {
            // vk::interpolateMat_8u
            const float fx = floorf(q0), fy = floorf(q1);
            const int ix = (int)fx, iy = (int)fy;
            const float sx = __fsub_rn(q0, fx), sy = __fsub_rn(q1, fy);
            const float w00 = __fmul_rn(__fsub_rn(1.0f, sx), __fsub_rn(1.0f, sy));
            const float w01 = __fmul_rn(__fsub_rn(1.0f, sx), sy);
            const float w10 = __fmul_rn(sx, __fsub_rn(1.0f, sy));
            const float w11 = __fsub_rn(__fsub_rn(__fsub_rn(1.0f, w00), w01), w10);
            const uint8_t* ptr = img + (size_t)iy * stride + ix;
            const float I = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w00, u8f(ptr[0])), __fmul_rn(w01, u8f(ptr[stride]))),
                                                __fmul_rn(w10, u8f(ptr[1]))),
                                      __fmul_rn(w11, u8f(ptr[stride + 1])));
            val = (uint8_t)I;
          }
        }
        border[y * 10 + x] = val;
      }
    }
  }
  // ---- align at the search level; the 8x8 reference patch is the interior of the border patch ----
  const double scale = (double)(1 << search_level);
  float u = (float)DD(a.px_cur[2 * I], scale), v = (float)DD(a.px_cur[2 * I + 1], scale);
  const uint8_t* cur = a.cur_img[search_level] + (size_t)c * a.cur_stride[search_level];
  const int ccols = a.width >> search_level, crows = a.height >> search_level;
  const int cur_step = (int)a.cur_pitch[search_level];
  const uint8_t* ref = border + 11;
  bool ok;
  if (a.is_edgelet && a.is_edgelet[i]) {
    const double g0 = a.ref_grad[2 * I], g1 = a.ref_grad[2 * I + 1];
    double d0 = DA(DM(A00, g0), DM(A01, g1)), d1 = DA(DM(A10, g0), DM(A11, g1));
    const double n = __dsqrt_rn(DA(DM(d0, d0), DM(d1, d1)));
    d0 = DD(d0, n), d1 = DD(d1, n);
    double h_inv;
    ok = align1d_core(border, ref, 10, cur, cur_step, ccols, crows, a.n_iter, (float)d0, (float)d1, u, v, h_inv);
  } else {
    ok = align2d_core(border, ref, 10, cur, cur_step, ccols, crows, a.n_iter, u, v);
  }
  a.out_px[2 * I] = DM((double)u, scale);
  a.out_px[2 * I + 1] = DM((double)v, scale);
  a.out_success[i] = ok ? 1 : 0;
}

}  // namespace

cudaError_t match_direct_kernel_launch(const MatchArgs& a, cudaStream_t s) {
  if (a.n <= 0) return cudaSuccess;
  match_direct_kernel<<<(a.n + kA2Threads - 1) / kA2Threads, kA2Threads, 0, s>>>(a);
  return cudaGetLastError();
}

cudaError_t align1d_kernel_launch(const Align2DArgs& a, cudaStream_t s) {
  if (a.n <= 0) return cudaSuccess;
  align1d_kernel<<<(a.n + kA2Threads - 1) / kA2Threads, kA2Threads, 0, s>>>(a);
  return cudaGetLastError();
}

cudaError_t align2d_kernel_launch(const Align2DArgs& a, cudaStream_t s) {
  if (a.n <= 0) return cudaSuccess;
  align2d_kernel<<<(a.n + kA2Threads - 1) / kA2Threads, kA2Threads, 0, s>>>(a);
  return cudaGetLastError();
}

}  // namespace plsvo
