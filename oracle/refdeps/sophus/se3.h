// Stand-in for the old non-templated Sophus sophus/se3.{h,cpp} — TEST INFRASTRUCTURE.
#ifndef PLSVO_REFDEPS_SOPHUS_SE3
#define PLSVO_REFDEPS_SOPHUS_SE3
#include "so3.h"

namespace Sophus {

typedef Matrix<double, 6, 1> Vector6d;
typedef Matrix<double, 6, 6> Matrix6d;

class SE3 {
 public:
  SE3() { translation_.setZero(); }
  SE3(const SO3& so3, const Vector3d& t) : so3_(so3), translation_(t) {}
  SE3(const Matrix3d& R, const Vector3d& t) : so3_(R), translation_(t) {}
  SE3(const Quaterniond& q, const Vector3d& t) : so3_(q), translation_(t) {}

  SE3 inverse() const {
    SE3 ret;
    ret.so3_ = so3_.inverse();
    ret.translation_ = ret.so3_ * (translation_ * -1.);
    return ret;
  }
  SE3& operator*=(const SE3& o) {
    translation_ += so3_ * (o.translation_);
    so3_ *= o.so3_;
    return *this;
  }
  SE3 operator*(const SE3& o) const {
    SE3 r(*this);
    r *= o;
    return r;
  }
  Vector3d operator*(const Vector3d& xyz) const { return so3_ * xyz + translation_; }

  static SE3 exp(const Vector6d& update) {
    const Vector3d upsilon = update.head<3>();
    const Vector3d omega = update.tail<3>();
    double theta;
    const SO3 so3 = SO3::expAndTheta(omega, &theta);
    const Matrix3d Omega = SO3::hat(omega);
    const Matrix3d Omega_sq = Omega * Omega;
    Matrix3d V;
    if (theta < SMALL_EPS) {
      V = so3.matrix();
    } else {
      const double theta_sq = theta * theta;
      V = (Matrix3d::Identity() + (1 - std::cos(theta)) / (theta_sq)*Omega +
           (theta - std::sin(theta)) / (theta_sq * theta) * Omega_sq);
    }
    return SE3(so3, V * upsilon);
  }

  // SE3::log (sophus/se3.cpp): [upsilon, omega] with omega = SO3::log and upsilon = V^-1 t
  static Vector6d log(const SE3& se3) {
    Vector6d upsilon_omega;
    const Quaterniond& q = se3.so3_.unit_quaternion();
    const double n = q.vec().norm(), w = q.w(), squared_w = w * w;
    double two_atan_nbyw_by_n;
    if (n < SMALL_EPS) {
      two_atan_nbyw_by_n = 2. / w - 2. * (n * n) / (w * squared_w);
    } else if (std::abs(w) < SMALL_EPS) {
      two_atan_nbyw_by_n = (w > 0 ? M_PI : -M_PI) / n;
    } else {
      two_atan_nbyw_by_n = 2 * std::atan(n / w) / n;
    }
    const double theta = two_atan_nbyw_by_n * n;
    const Vector3d omega = two_atan_nbyw_by_n * q.vec();
    Vector3d upsilon;
    if (theta < SMALL_EPS) {
      const Matrix3d Omega = SO3::hat(omega);
      const Matrix3d V_inv = Matrix3d::Identity() - 0.5 * Omega + (1. / 12.) * (Omega * Omega);
      upsilon = V_inv * se3.translation_;
    } else {
      const Matrix3d Omega = SO3::hat(omega);
      const Matrix3d V_inv = (Matrix3d::Identity() - 0.5 * Omega +
                              (1 - theta / (2 * std::tan(theta / 2))) / (theta * theta) * (Omega * Omega));
      upsilon = V_inv * se3.translation_;
    }
    for (int k = 0; k < 3; ++k) upsilon_omega[k] = upsilon[k], upsilon_omega[3 + k] = omega[k];
    return upsilon_omega;
  }
  Vector6d log() const { return log(*this); }
  Matrix3d rotation_matrix() const { return so3_.matrix(); }
  const Vector3d& translation() const { return translation_; }
  Vector3d& translation() { return translation_; }
  const SO3& so3() const { return so3_; }
  SO3& so3() { return so3_; }
  const Quaterniond& unit_quaternion() const { return so3_.unit_quaternion(); }

 private:
  SO3 so3_;
  Vector3d translation_;
};

}  // namespace Sophus
#endif
