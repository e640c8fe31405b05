// Stand-in for rpg_vikit vikit_common/patch_score.h — TEST INFRASTRUCTURE.  ZMSSD (zero-mean sum of
// squared differences over a (2*HALF)^2 patch, integer arithmetic) is only used by the epipolar search
// (src/matcher.cpp:277-420), which is outside the widened path; provided so that matcher.cpp compiles.
#ifndef PLSVO_REFDEPS_VIKIT_PATCH_SCORE
#define PLSVO_REFDEPS_VIKIT_PATCH_SCORE
#include <stdint.h>

namespace vk {
namespace patch_score {

template <int HALF_PATCH_SIZE>
class ZMSSD {
 public:
  static const int patch_size_ = 2 * HALF_PATCH_SIZE;
  static const int patch_area_ = patch_size_ * patch_size_;
  static const int threshold_ = 2000 * patch_area_;
  uint8_t* ref_patch_;
  int sumA_, sumAA_;

  ZMSSD(uint8_t* ref_patch) : ref_patch_(ref_patch) {
    uint32_t sumA_uint = 0, sumAA_uint = 0;
    for (int r = 0; r < patch_area_; r++) {
      uint8_t n = ref_patch_[r];
      sumA_uint += n;
      sumAA_uint += n * n;
    }
    sumA_ = sumA_uint;
    sumAA_ = sumAA_uint;
  }
  static int threshold() { return threshold_; }
  int computeScore(uint8_t* cur_patch) const {
    uint32_t sumB_uint = 0, sumBB_uint = 0, sumAB_uint = 0;
    for (int r = 0; r < patch_area_; r++) {
      const uint8_t cur_pixel = cur_patch[r];
      sumB_uint += cur_pixel;
      sumBB_uint += cur_pixel * cur_pixel;
      sumAB_uint += cur_pixel * ref_patch_[r];
    }
    const int sumB = sumB_uint, sumBB = sumBB_uint, sumAB = sumAB_uint;
    return sumAA_ - 2 * sumAB + sumBB - (sumA_ * sumA_ - 2 * sumA_ * sumB + sumB * sumB) / patch_area_;
  }
  int computeScore(uint8_t* cur_patch, int stride) const {
    int sumB = 0, sumBB = 0, sumAB = 0;
    for (int y = 0, r = 0; y < patch_size_; ++y) {
      uint8_t* cur_patch_ptr = cur_patch + y * stride;
      for (int x = 0; x < patch_size_; ++x, ++r) {
        const int cur_px = cur_patch_ptr[x];
        sumB += cur_px;
        sumBB += cur_px * cur_px;
        sumAB += cur_px * ref_patch_[r];
      }
    }
    return sumAA_ - 2 * sumAB + sumBB - (sumA_ * sumA_ - 2 * sumA_ * sumB + sumB * sumB) / patch_area_;
  }
};

}  // namespace patch_score
}  // namespace vk
#endif
