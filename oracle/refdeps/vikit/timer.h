// Stand-in for rpg_vikit vikit_common/timer.h — TEST INFRASTRUCTURE (frame_handler_base.h keeps a vk::Timer member).
#ifndef PLSVO_REFDEPS_VIKIT_TIMER
#define PLSVO_REFDEPS_VIKIT_TIMER
#include <chrono>
namespace vk {
class Timer {
  std::chrono::steady_clock::time_point start_;
  double time_, accumulated_;

 public:
  Timer() : time_(0), accumulated_(0) { start(); }
  inline void start() {
    accumulated_ = 0.0;
    start_ = std::chrono::steady_clock::now();
  }
  inline void resume() { start_ = std::chrono::steady_clock::now(); }
  inline double stop() {
    time_ = std::chrono::duration<double>(std::chrono::steady_clock::now() - start_).count() + accumulated_;
    accumulated_ = time_;
    return time_;
  }
  inline double getTime() const { return time_; }
  inline void reset() { time_ = 0.0, accumulated_ = 0.0; }
  static double getCurrentTime() { return std::chrono::duration<double>(std::chrono::system_clock::now().time_since_epoch()).count(); }
  static double getCurrentSecond() { return getCurrentTime(); }
};
}  // namespace vk
#endif
