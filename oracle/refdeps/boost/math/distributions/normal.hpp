// Stand-in for <boost/math/distributions/normal.hpp> — TEST INFRASTRUCTURE.
// pdf(normal_distribution<T>(mean, sd), x) as boost computes it:
//   exponent = x - mean; exponent *= -exponent; exponent /= 2*sd*sd; result = exp(exponent) / (sd * sqrt(2*pi)), all in T.
#ifndef PLSVO_REFDEPS_BOOST_MATH_NORMAL
#define PLSVO_REFDEPS_BOOST_MATH_NORMAL
#include <cmath>
namespace boost {
namespace math {
template <class RealType = double>
class normal_distribution {
  RealType mean_, sd_;

 public:
  normal_distribution(RealType mean = 0, RealType sd = 1) : mean_(mean), sd_(sd) {}
  RealType mean() const { return mean_; }
  RealType standard_deviation() const { return sd_; }
};
template <class RealType>
inline RealType pdf(const normal_distribution<RealType>& dist, const RealType& x) {
  const RealType sd = dist.standard_deviation();
  const RealType mean = dist.mean();
  if (std::isinf(x)) return 0;
  RealType exponent = x - mean;
  exponent *= -exponent;
  exponent /= 2 * sd * sd;
  RealType result = std::exp(exponent);
  result /= sd * std::sqrt(2 * static_cast<RealType>(3.141592653589793238462643383279502884L));
  return result;
}
}  // namespace math
}  // namespace boost
#endif
