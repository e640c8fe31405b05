// Stand-in for <boost/thread.hpp> over the C++11 standard library — TEST INFRASTRUCTURE.
// plsvo::DepthFilter owns an optional worker thread (src/depth_filter.cpp:98-113,234-260); the harness never starts
// it, so interruption is a plain flag.
#ifndef PLSVO_REFDEPS_BOOST_THREAD
#define PLSVO_REFDEPS_BOOST_THREAD
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <utility>

namespace boost {
using std::condition_variable;
using std::mutex;
using std::unique_lock;
namespace detail {
inline std::atomic<bool>& interrupt_flag() {
  static std::atomic<bool> f(false);
  return f;
}
}  // namespace detail
class thread {
  std::thread t_;

 public:
  template <class F, class... A>
  explicit thread(F&& f, A&&... a) : t_(std::forward<F>(f), std::forward<A>(a)...) {}
  void interrupt() { detail::interrupt_flag() = true; }
  void join() {
    if (t_.joinable()) t_.join();
  }
};
namespace this_thread {
inline bool interruption_requested() { return detail::interrupt_flag(); }
}  // namespace this_thread
}  // namespace boost
#endif
