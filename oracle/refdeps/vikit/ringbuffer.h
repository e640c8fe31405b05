// Stand-in for rpg_vikit vikit_common/ringbuffer.h — TEST INFRASTRUCTURE (frame_handler_base.h keeps two vk::RingBuffer members).
#ifndef PLSVO_REFDEPS_VIKIT_RINGBUFFER
#define PLSVO_REFDEPS_VIKIT_RINGBUFFER
#include <algorithm>
#include <cassert>
#include <numeric>
#include <vector>
namespace vk {
template <typename T>
class RingBuffer {
  std::vector<T> arr_;
  int begin_, end_, num_elem_, arr_size_;

 public:
  RingBuffer(int size) : arr_(size), begin_(0), end_(-1), num_elem_(0), arr_size_(size) {}
  void push_back(const T& elem) {
    if (num_elem_ < arr_size_) {
      end_++;
      arr_[end_] = elem;
      num_elem_++;
    } else {
      end_ = (end_ + 1) % arr_size_;
      begin_ = (begin_ + 1) % arr_size_;
      arr_[end_] = elem;
    }
  }
  bool empty() const { return arr_.empty(); }
  T get(int i) {
    assert(i < num_elem_);
    return arr_[(begin_ + i) % arr_size_];
  }
  T getSum() const {
    T sum = 0;
    for (int i = 0; i < num_elem_; ++i) sum += arr_[i];
    return sum;
  }
  T getMean() const {
    if (num_elem_ == 0) return 0;
    return getSum() / num_elem_;
  }
  int size() { return num_elem_; }
};
}  // namespace vk
#endif
