// Stand-in for rpg_vikit vikit_common/robust_cost.{h,cpp} — TEST INFRASTRUCTURE.
// MAD scale = 1.48f * median; Tukey biweight with b = 4.6851f, all in float as upstream.
#ifndef PLSVO_REFDEPS_VIKIT_ROBUST_COST
#define PLSVO_REFDEPS_VIKIT_ROBUST_COST
#include <vikit/math_utils.h>
#include <memory>
#include <vector>

namespace vk {
namespace robust_cost {

class ScaleEstimator {
 public:
  virtual ~ScaleEstimator() {}
  virtual float compute(std::vector<float>& errors) const = 0;
};
typedef std::shared_ptr<ScaleEstimator> ScaleEstimatorPtr;

class UnitScaleEstimator : public ScaleEstimator {
 public:
  float compute(std::vector<float>&) const { return 1.0f; }
};

class MADScaleEstimator : public ScaleEstimator {
 public:
  float compute(std::vector<float>& errors) const {
    // error must be in absolute values!
    return NORMALIZER() * vk::getMedian(errors);
  }
 private:
  static float NORMALIZER() { return 1.48f; }  // 1 / 0.6745
};

class WeightFunction {
 public:
  virtual ~WeightFunction() {}
  virtual float value(const float& x) const = 0;
  virtual void configure(const float&) {}
};
typedef std::shared_ptr<WeightFunction> WeightFunctionPtr;

class UnitWeightFunction : public WeightFunction {
 public:
  float value(const float&) const { return 1.0f; }
};

class TukeyWeightFunction : public WeightFunction {
 public:
  TukeyWeightFunction(const float b = 4.6851f) { configure(b); }
  float value(const float& x) const {
    const float x_square = x * x;
    if (x_square <= b_square) {
      const float tmp = 1.0f - x_square / b_square;
      return tmp * tmp;
    } else {
      return 0.0f;
    }
  }
  void configure(const float& param) { b_square = param * param; }
 private:
  float b_square;
};

}  // namespace robust_cost
}  // namespace vk
#endif
