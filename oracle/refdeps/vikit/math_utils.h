// Stand-in for uzh-rpg/rpg_vikit vikit_common/math_utils.h — TEST INFRASTRUCTURE (only what the
// hot path calls: project2d, unproject2d, norm_max, getMedian).
#ifndef PLSVO_REFDEPS_VIKIT_MATH_UTILS
#define PLSVO_REFDEPS_VIKIT_MATH_UTILS
#include <Eigen/Core>
#include <sophus/se3.h>
#include <algorithm>
#include <cassert>
#include <cmath>
#include <vector>

namespace vk {
using namespace Eigen;
using namespace std;
using namespace Sophus;

inline Vector2d project2d(const Vector3d& v) { return v.head<2>() / v[2]; }
inline Vector3d unproject2d(const Vector2d& v) { return Vector3d(v[0], v[1], 1.0); }

template <int N>
inline double norm_max(const Eigen::Matrix<double, N, 1>& v) {
  double max = -1;
  for (int i = 0; i < v.size(); i++) {
    double abs = fabs(v[i]);
    if (abs > max) max = abs;
  }
  return max;
}

template <class T>
T getMedian(vector<T>& data_vec) {
  assert(!data_vec.empty());
  typename vector<T>::iterator it = data_vec.begin() + floor(data_vec.size() / 2);
  nth_element(data_vec.begin(), it, data_vec.end());
  return *it;
}

}  // namespace vk
#endif
