// exact_math.cuh — double-precision geometry with explicit round-to-nearest intrinsics.
//
// The "next-row" kernels (Matcher::findMatchDirect, Point::optimize / LineSeg::optimize) reproduce the
// reference's scalar double arithmetic bit for bit.  nvcc contracts a*b+c into an FMA by default, which
// changes the last bit; every operation here is therefore an explicit __d*_rn intrinsic, in the
// evaluation order of the CPU code (old non-templated Sophus SE3 on Eigen quaternions, as restated in
// oracle/plsvo_oracle.cpp and checked there against the reference's own translation units).
#pragma once
#include <cuda_runtime.h>

namespace plsvo {
namespace {

struct V3 {
  double x, y, z;
};
struct Q4 {
  double x, y, z, w;
};
struct Pose {
  Q4 q;
  V3 t;
};
#define DM(a, b) __dmul_rn((a), (b))
#define DA(a, b) __dadd_rn((a), (b))
#define DS(a, b) __dsub_rn((a), (b))
#define DD(a, b) __ddiv_rn((a), (b))
__device__ __forceinline__ V3 v_add(V3 a, V3 b) { return {DA(a.x, b.x), DA(a.y, b.y), DA(a.z, b.z)}; }
__device__ __forceinline__ V3 v_sub(V3 a, V3 b) { return {DS(a.x, b.x), DS(a.y, b.y), DS(a.z, b.z)}; }
__device__ __forceinline__ V3 v_scale(V3 a, double s) { return {DM(a.x, s), DM(a.y, s), DM(a.z, s)}; }
__device__ __forceinline__ V3 v_cross(V3 a, V3 b) {
  return {DS(DM(a.y, b.z), DM(a.z, b.y)), DS(DM(a.z, b.x), DM(a.x, b.z)), DS(DM(a.x, b.y), DM(a.y, b.x))};
}
__device__ __forceinline__ double v_norm(V3 a) { return __dsqrt_rn(DA(DA(DM(a.x, a.x), DM(a.y, a.y)), DM(a.z, a.z))); }
__device__ __forceinline__ Q4 q_normalized(Q4 q) {  // Eigen: coeffs() /= coeffs().norm()
  const double n = __dsqrt_rn(DA(DA(DA(DM(q.x, q.x), DM(q.y, q.y)), DM(q.z, q.z)), DM(q.w, q.w)));
  return {DD(q.x, n), DD(q.y, n), DD(q.z, n), DD(q.w, n)};
}
__device__ __forceinline__ Q4 q_mul(Q4 a, Q4 b) {
  return {DS(DA(DA(DM(a.w, b.x), DM(a.x, b.w)), DM(a.y, b.z)), DM(a.z, b.y)),
          DS(DA(DA(DM(a.w, b.y), DM(a.y, b.w)), DM(a.z, b.x)), DM(a.x, b.z)),
          DS(DA(DA(DM(a.w, b.z), DM(a.z, b.w)), DM(a.x, b.y)), DM(a.y, b.x)),
          DS(DS(DS(DM(a.w, b.w), DM(a.x, b.x)), DM(a.y, b.y)), DM(a.z, b.z))};
}
__device__ __forceinline__ V3 q_rot(Q4 q, V3 v) {  // Eigen QuaternionBase::_transformVector
  const V3 qv{q.x, q.y, q.z};
  V3 uv = v_cross(qv, v);
  uv = v_add(uv, uv);
  return v_add(v_add(v, v_scale(uv, q.w)), v_cross(qv, uv));
}
__device__ __forceinline__ Pose pose_load(const double* p) {  // SO3(const Quaterniond&) normalises
  return {q_normalized(Q4{p[0], p[1], p[2], p[3]}), V3{p[4], p[5], p[6]}};
}
__device__ __forceinline__ Pose pose_inverse(Pose a) {  // Sophus SE3::inverse
  Pose r;
  r.q = q_normalized(Q4{-a.q.x, -a.q.y, -a.q.z, a.q.w});
  r.t = q_rot(r.q, v_scale(a.t, -1.0));
  return r;
}
__device__ __forceinline__ Pose pose_mul(Pose a, Pose b) {  // SE3::operator*=
  Pose r;
  r.t = v_add(a.t, q_rot(a.q, b.t));
  r.q = q_normalized(q_mul(a.q, b.q));
  return r;
}
__device__ __forceinline__ V3 pose_act(Pose T, V3 p) { return v_add(q_rot(T.q, p), T.t); }

__device__ __forceinline__ void q_to_matrix(Q4 q, double R[3][3]) {  // Eigen QuaternionBase::toRotationMatrix
  const double tx = DM(2.0, q.x), ty = DM(2.0, q.y), tz = DM(2.0, q.z);
  const double twx = DM(tx, q.w), twy = DM(ty, q.w), twz = DM(tz, q.w);
  const double txx = DM(tx, q.x), txy = DM(ty, q.x), txz = DM(tz, q.x);
  const double tyy = DM(ty, q.y), tyz = DM(tz, q.y), tzz = DM(tz, q.z);
  R[0][0] = DS(1.0, DA(tyy, tzz)), R[0][1] = DS(txy, twz), R[0][2] = DA(txz, twy);
  R[1][0] = DA(txy, twz), R[1][1] = DS(1.0, DA(txx, tzz)), R[1][2] = DS(tyz, twx);
  R[2][0] = DS(txz, twy), R[2][1] = DA(tyz, twx), R[2][2] = DS(1.0, DA(txx, tyy));
}

}  // namespace
}  // namespace plsvo
