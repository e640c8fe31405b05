// Stand-in for the old non-templated Sophus (strasdat/Sophus, sophus/so3.{h,cpp}) — TEST
// INFRASTRUCTURE, see oracle/refdeps/Eigen/Core.  Rotation is a unit quaternion that is
// re-normalised after every product; exp uses the series the original uses below SMALL_EPS.
#ifndef PLSVO_REFDEPS_SOPHUS_SO3
#define PLSVO_REFDEPS_SOPHUS_SO3
#include <Eigen/Core>
#include <cmath>
#include <list>
#include <string>
#include <vector>

namespace Sophus {
using namespace Eigen;
using namespace std;

const double SMALL_EPS = 1e-10;

class SO3 {
 public:
  SO3() { unit_quaternion_.setIdentity(); }
  SO3(const Quaterniond& q) : unit_quaternion_(q) { unit_quaternion_.normalize(); }
  SO3(const Matrix3d& R) {
    // Eigen quaternion-from-rotation-matrix (Geometry/Quaternion.h quaternionbase_assign_impl)
    double t = R(0, 0) + R(1, 1) + R(2, 2);
    double q[4];  // x y z w
    if (t > 0) {
      t = std::sqrt(t + 1.0);
      q[3] = 0.5 * t;
      t = 0.5 / t;
      q[0] = (R(2, 1) - R(1, 2)) * t, q[1] = (R(0, 2) - R(2, 0)) * t, q[2] = (R(1, 0) - R(0, 1)) * t;
    } else {
      int i = 0;
      if (R(1, 1) > R(0, 0)) i = 1;
      if (R(2, 2) > R(i, i)) i = 2;
      const int j = (i + 1) % 3, k = (j + 1) % 3;
      t = std::sqrt(R(i, i) - R(j, j) - R(k, k) + 1.0);
      q[i] = 0.5 * t;
      t = 0.5 / t;
      q[3] = (R(k, j) - R(j, k)) * t, q[j] = (R(j, i) + R(i, j)) * t, q[k] = (R(k, i) + R(i, k)) * t;
    }
    unit_quaternion_ = Quaterniond(q[3], q[0], q[1], q[2]);
    unit_quaternion_.normalize();
  }
  SO3 inverse() const { return SO3(unit_quaternion_.conjugate()); }
  Matrix3d matrix() const { return unit_quaternion_.toRotationMatrix(); }
  SO3& operator*=(const SO3& o) {
    unit_quaternion_ *= o.unit_quaternion_;
    unit_quaternion_.normalize();
    return *this;
  }
  SO3 operator*(const SO3& o) const {
    SO3 r(*this);
    r *= o;
    return r;
  }
  Vector3d operator*(const Vector3d& xyz) const { return unit_quaternion_ * xyz; }
  const Quaterniond& unit_quaternion() const { return unit_quaternion_; }

  static Matrix3d hat(const Vector3d& v) {
    Matrix3d O;
    O(0, 0) = 0, O(0, 1) = -v(2), O(0, 2) = v(1);
    O(1, 0) = v(2), O(1, 1) = 0, O(1, 2) = -v(0);
    O(2, 0) = -v(1), O(2, 1) = v(0), O(2, 2) = 0;
    return O;
  }
  static SO3 expAndTheta(const Vector3d& omega, double* theta) {
    *theta = omega.norm();
    const double half_theta = 0.5 * (*theta);
    double imag_factor;
    const double real_factor = std::cos(half_theta);
    if ((*theta) < SMALL_EPS) {
      const double theta_sq = (*theta) * (*theta);
      const double theta_po4 = theta_sq * theta_sq;
      imag_factor = 0.5 - 0.0208333 * theta_sq + 0.000260417 * theta_po4;
    } else {
      const double sin_half_theta = std::sin(half_theta);
      imag_factor = sin_half_theta / (*theta);
    }
    return SO3(Quaterniond(real_factor, imag_factor * omega.x(), imag_factor * omega.y(), imag_factor * omega.z()));
  }
  static SO3 exp(const Vector3d& omega) {
    double theta;
    return expAndTheta(omega, &theta);
  }

 protected:
  Quaterniond unit_quaternion_;
};

}  // namespace Sophus
#endif
