// structopt_kernel.cu — Point::optimize (src/feature3D_impl.cpp:36-95) and LineSeg::optimize (:97-174) as
// driven by FrameHandlerBase::optimizeStructure (src/frame_handler_base.cpp:202-237).  SURVEY.md §8f rank 3
// ("next"): the step after the pose optimiser.
//
// 3x3 Gauss-Newton on the unit-plane reprojection error of a 3D point over its observations.  One thread
// per 3D feature walks its observation list in obs_ order (the reference's summation order); a line
// segment's thread optimises both end points with the reference's coupled accept / roll-back /
// convergence test.  All arithmetic goes through exact_math.cuh (no FMA contraction) and the 3x3 solve is
// Eigen's pivoted LDLT (ldlt_inplace<Lower>::unblocked + solve), so positions are bit-identical to the
// reference's scalar code.  Tiny, latency-bound work: useful only batched over many features/frames.
#include <cuda_runtime.h>
#include <stdint.h>

#include "exact_math.cuh"
#include "internal.h"

namespace plsvo {
namespace {

struct Normal3 {
  double A[3][3], b[3], chi2;
};
__device__ __forceinline__ void normal_clear(Normal3& n) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    n.b[r] = 0.0;
#pragma unroll
    for (int c = 0; c < 3; ++c) n.A[r][c] = 0.0;
  }
  n.chi2 = 0.0;
}

// one observation: A += J^T J, b -= J^T e, chi2 += |e|^2   (feature3D_impl.cpp:49-59, feature3D.h:126-140)
__device__ __forceinline__ void accumulate(const double* __restrict__ T7, V3 pos, const double* __restrict__ f3, Normal3& n) {
  const Pose T = pose_load(T7);
  double R[3][3];
  q_to_matrix(T.q, R);
  const V3 p = pose_act(T, pos);
  const double z_inv = DD(1.0, p.z);
  const double z_inv_sq = DM(z_inv, z_inv);
  const double P[2][3] = {{-z_inv, -0.0, -(DM(-p.x, z_inv_sq))}, {-0.0, -z_inv, -(DM(-p.y, z_inv_sq))}};
  double J[2][3];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) J[r][c] = DA(DA(DM(P[r][0], R[0][c]), DM(P[r][1], R[1][c])), DM(P[r][2], R[2][c]));
  const double e0 = DS(DD(f3[0], f3[2]), DD(p.x, p.z)), e1 = DS(DD(f3[1], f3[2]), DD(p.y, p.z));
  n.chi2 = DA(n.chi2, DA(DM(e0, e0), DM(e1, e1)));
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 3; ++c) n.A[r][c] = DA(n.A[r][c], DA(DM(J[0][r], J[0][c]), DM(J[1][r], J[1][c])));
    n.b[r] = DS(n.b[r], DA(DM(J[0][r], e0), DM(J[1][r], e1)));
  }
}

// x = A.ldlt().solve(b), Eigen/src/Cholesky/LDLT.h (lower, diagonal pivoting, pseudo-inverse of D)
__device__ void ldlt3_solve(const double Ain[3][3], const double b[3], double x[3]) {
  double m[3][3];
  int tr[3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) m[i][j] = Ain[i][j];
  bool done = false;
  for (int k = 0; k < 3 && !done; ++k) {
    int piv = k;
    double big = fabs(m[k][k]);
    for (int i = k + 1; i < 3; ++i)
      if (fabs(m[i][i]) > big) big = fabs(m[i][i]), piv = i;
    tr[k] = piv;
    if (piv != k) {
      for (int j = 0; j < k; ++j) {
        const double t = m[k][j];
        m[k][j] = m[piv][j], m[piv][j] = t;
      }
      for (int i = piv + 1; i < 3; ++i) {
        const double t = m[i][k];
        m[i][k] = m[i][piv], m[i][piv] = t;
      }
      {
        const double t = m[k][k];
        m[k][k] = m[piv][piv], m[piv][piv] = t;
      }
      for (int i = k + 1; i < piv; ++i) {
        const double t = m[i][k];
        m[i][k] = m[piv][i], m[piv][i] = t;
      }
    }
    if (k > 0) {
      double temp[3];
      for (int j = 0; j < k; ++j) temp[j] = DM(m[j][j], m[k][j]);
      double s = 0.0;
      for (int j = 0; j < k; ++j) s = DA(s, DM(m[k][j], temp[j]));
      m[k][k] = DS(m[k][k], s);
      for (int i = k + 1; i < 3; ++i) {
        double s2 = 0.0;
        for (int j = 0; j < k; ++j) s2 = DA(s2, DM(m[i][j], temp[j]));
        m[i][k] = DS(m[i][k], s2);
      }
    }
    const double akk = m[k][k];
    const bool valid = fabs(akk) > 0.0;
    if (k == 0 && !valid) {
      for (int j = 0; j < 3; ++j) tr[j] = j;
      done = true;
    } else if (valid) {
      for (int i = k + 1; i < 3; ++i) m[i][k] = DD(m[i][k], akk);
    }
  }
  for (int i = 0; i < 3; ++i) x[i] = b[i];
  for (int k = 0; k < 3; ++k) {
    const double t = x[k];
    x[k] = x[tr[k]], x[tr[k]] = t;
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < i; ++j) x[i] = DS(x[i], DM(m[i][j], x[j]));
  const double tol = 1.0 / 1.7976931348623157e308;
  for (int i = 0; i < 3; ++i) x[i] = (fabs(m[i][i]) > tol) ? DD(x[i], m[i][i]) : 0.0;
  for (int i = 2; i >= 0; --i)
    for (int j = i + 1; j < 3; ++j) x[i] = DS(x[i], DM(m[j][i], x[j]));
  for (int k = 2; k >= 0; --k) {
    const double t = x[k];
    x[k] = x[tr[k]], x[tr[k]] = t;
  }
}
__device__ __forceinline__ double norm_max3(const double x[3]) { return fmax(fmax(fabs(x[0]), fabs(x[1])), fabs(x[2])); }

constexpr double kEps = 0.0000000001;  // plsvo::EPS (include/plsvo/global.h:92)

__global__ void __launch_bounds__(128) structopt_kernel(const StructOptArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < a.n_points) {
    V3 pos{a.pt_pos[3 * (size_t)i], a.pt_pos[3 * (size_t)i + 1], a.pt_pos[3 * (size_t)i + 2]};
    V3 old_point = pos;
    double chi2 = 0.0;
    int iters = 0;
    const int o0 = a.pt_obs_begin[i], o1 = a.pt_obs_begin[i + 1];
    for (int it = 0; it < a.n_iter_pts; ++it) {
      Normal3 n;
      normal_clear(n);
      for (int o = o0; o < o1; ++o) accumulate(a.T_f_w + 7 * (size_t)a.pt_obs_frame[o], pos, a.pt_obs_f + 3 * (size_t)o, n);
      double dp[3];
      ldlt3_solve(n.A, n.b, dp);
      ++iters;
      if ((it > 0 && n.chi2 > chi2) || isnan(dp[0])) {
        pos = old_point;  // roll-back
        break;
      }
      old_point = pos;
      pos = V3{DA(pos.x, dp[0]), DA(pos.y, dp[1]), DA(pos.z, dp[2])};
      chi2 = n.chi2;
      if (norm_max3(dp) <= kEps) break;
    }
    a.out_pt_pos[3 * (size_t)i] = pos.x, a.out_pt_pos[3 * (size_t)i + 1] = pos.y, a.out_pt_pos[3 * (size_t)i + 2] = pos.z;
    if (a.out_pt_iters) a.out_pt_iters[i] = iters;
    return;
  }
  const int s = i - a.n_points;
  if (s >= a.n_segs) return;
  const size_t S = (size_t)s;
  V3 sp{a.seg_spos[3 * S], a.seg_spos[3 * S + 1], a.seg_spos[3 * S + 2]};
  V3 ep{a.seg_epos[3 * S], a.seg_epos[3 * S + 1], a.seg_epos[3 * S + 2]};
  V3 old_s = sp, old_e = ep;
  double chi2s = 0.0, chi2e = 0.0;
  int iters = 0;
  const int o0 = a.seg_obs_begin[s], o1 = a.seg_obs_begin[s + 1];
  for (int it = 0; it < a.n_iter_segs; ++it) {
    Normal3 ns, ne;
    normal_clear(ns);
    normal_clear(ne);
    for (int o = o0; o < o1; ++o) {
      const double* T7 = a.T_f_w + 7 * (size_t)a.seg_obs_frame[o];
      accumulate(T7, sp, a.seg_obs_sf + 3 * (size_t)o, ns);
      accumulate(T7, ep, a.seg_obs_ef + 3 * (size_t)o, ne);
    }
    double dps[3], dpe[3];
    ldlt3_solve(ns.A, ns.b, dps);
    ldlt3_solve(ne.A, ne.b, dpe);
    ++iters;
    if ((it > 0 && ns.chi2 > chi2s) || isnan(dps[0]) || (it > 0 && ne.chi2 > chi2e) || isnan(dpe[0])) {
      sp = old_s, ep = old_e;
      break;
    }
    old_s = sp, old_e = ep;
    sp = V3{DA(sp.x, dps[0]), DA(sp.y, dps[1]), DA(sp.z, dps[2])};
    ep = V3{DA(ep.x, dpe[0]), DA(ep.y, dpe[1]), DA(ep.z, dpe[2])};
    chi2s = ns.chi2, chi2e = ne.chi2;
    if (norm_max3(dps) <= kEps || norm_max3(dpe) <= kEps) break;
  }
  a.out_seg_spos[3 * S] = sp.x, a.out_seg_spos[3 * S + 1] = sp.y, a.out_seg_spos[3 * S + 2] = sp.z;
  a.out_seg_epos[3 * S] = ep.x, a.out_seg_epos[3 * S + 1] = ep.y, a.out_seg_epos[3 * S + 2] = ep.z;
  if (a.out_seg_iters) a.out_seg_iters[s] = iters;
}

}  // namespace

cudaError_t structopt_kernel_launch(const StructOptArgs& a, cudaStream_t s) {
  const int n = a.n_points + a.n_segs;
  if (n <= 0) return cudaSuccess;
  structopt_kernel<<<(n + 127) / 128, 128, 0, s>>>(a);
  return cudaGetLastError();
}

}  // namespace plsvo
