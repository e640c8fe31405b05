// Stand-in for <opencv2/opencv.hpp> — TEST INFRASTRUCTURE, see oracle/refdeps/Eigen/Core.
// A dense row-major cv::Mat (owned or wrapping caller memory), ROI views, and the few geometry
// types the reference's feature.h / feature.cpp / sparse_img_align.cpp name.  No image
// processing is provided: the hot path only reads pixels through Mat::ptr / Mat::data.
#ifndef PLSVO_REFDEPS_OPENCV
#define PLSVO_REFDEPS_OPENCV

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

typedef unsigned char uchar;

#define CV_8U 0
#define CV_32F 5
#define CV_8UC1 0
#define CV_32FC1 5

namespace cv {

template <class T>
struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T xx, T yy) : x(xx), y(yy) {}
  template <class U>
  Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {}
};
typedef Point_<int> Point;
typedef Point_<int> Point2i;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;

template <class T>
struct Size_ {
  T width, height;
  Size_() : width(0), height(0) {}
  Size_(T w, T h) : width(w), height(h) {}
};
typedef Size_<int> Size;
typedef Size_<float> Size2f;

template <class T>
struct Rect_ {
  T x, y, width, height;
  Rect_() : x(0), y(0), width(0), height(0) {}
  Rect_(T xx, T yy, T w, T h) : x(xx), y(yy), width(w), height(h) {}
};
typedef Rect_<int> Rect;
typedef Rect_<float> Rect2f;

struct Scalar {
  double val[4];
  Scalar(double v0 = 0, double v1 = 0, double v2 = 0, double v3 = 0) { val[0] = v0, val[1] = v1, val[2] = v2, val[3] = v3; }
};

struct RotatedRect {
  Point2f center;
  Size2f size;
  float angle;
  RotatedRect() : angle(0) {}
  RotatedRect(const Point2f& c, const Size2f& s, float a) : center(c), size(s), angle(a) {}
  Rect boundingRect() const {  // axis-aligned bounding box of the four corners
    const double a = angle * M_PI / 180.0;
    const float b = (float)std::cos(a) * 0.5f, s = (float)std::sin(a) * 0.5f;
    float xs[4], ys[4];
    xs[0] = center.x - s * size.height - b * size.width, ys[0] = center.y + b * size.height - s * size.width;
    xs[1] = center.x + s * size.height - b * size.width, ys[1] = center.y - b * size.height - s * size.width;
    xs[2] = 2 * center.x - xs[0], ys[2] = 2 * center.y - ys[0];
    xs[3] = 2 * center.x - xs[1], ys[3] = 2 * center.y - ys[1];
    float x0 = xs[0], x1 = xs[0], y0 = ys[0], y1 = ys[0];
    for (int k = 1; k < 4; ++k) {
      x0 = std::fmin(x0, xs[k]), x1 = std::fmax(x1, xs[k]);
      y0 = std::fmin(y0, ys[k]), y1 = std::fmax(y1, ys[k]);
    }
    Rect r((int)std::floor(x0), (int)std::floor(y0), 0, 0);
    r.width = (int)std::ceil(x1) - r.x + 1;
    r.height = (int)std::ceil(y1) - r.y + 1;
    return r;
  }
};

template <class T>
class MatIterator_ {
  uchar* base_;
  size_t step_;
  int cols_, i_, j_;

 public:
  MatIterator_() : base_(nullptr), step_(0), cols_(0), i_(0), j_(0) {}
  MatIterator_(uchar* base, size_t step, int cols) : base_(base), step_(step), cols_(cols), i_(0), j_(0) {}
  T& operator*() const { return *reinterpret_cast<T*>(base_ + (size_t)i_ * step_ + (size_t)j_ * sizeof(T)); }
  MatIterator_& operator++() {
    if (++j_ >= cols_) j_ = 0, ++i_;
    return *this;
  }
};

class Mat {
  static size_t elem_size(int type) { return type == CV_32F ? 4 : 1; }
  std::shared_ptr<std::vector<uchar>> owner_;
  int type_;

 public:
  struct Step {
    size_t p[2];
    Step() { p[0] = p[1] = 0; }
    size_t operator[](int i) const { return p[i]; }
    operator size_t() const { return p[0]; }
  };

  uchar* data;
  int rows, cols;
  Step step;

  Mat() : type_(CV_8U), data(nullptr), rows(0), cols(0) {}
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(Size s, int type) { create(s.height, s.width, type); }
  Mat(Size s, int type, const Scalar& v) {
    create(s.height, s.width, type);
    fill(v);
  }
  Mat(int r, int c, int type, const Scalar& v) {
    create(r, c, type);
    fill(v);
  }
  // wrap caller-owned pixels (no copy), as cv::Mat(rows, cols, type, data, step)
  Mat(int r, int c, int type, void* ext, size_t step_bytes) : type_(type), data((uchar*)ext), rows(r), cols(c) {
    step.p[0] = step_bytes ? step_bytes : (size_t)c * elem_size(type);
    step.p[1] = elem_size(type);
  }
  // region of interest: shares pixels and keeps the parent's row step
  Mat(const Mat& m, const Rect& roi) : owner_(m.owner_), type_(m.type_), rows(roi.height), cols(roi.width), step(m.step) {
    data = m.data + (ptrdiff_t)roi.y * (ptrdiff_t)m.step.p[0] + (ptrdiff_t)roi.x * (ptrdiff_t)m.step.p[1];
  }
  Mat operator()(const Rect& roi) const { return Mat(*this, roi); }

  void create(int r, int c, int type) {
    type_ = type, rows = r, cols = c;
    step.p[1] = elem_size(type);
    step.p[0] = (size_t)c * step.p[1];
    owner_ = std::make_shared<std::vector<uchar>>((size_t)r * step.p[0] + 16);
    data = owner_->data();
  }
  void fill(const Scalar& v) {
    for (int i = 0; i < rows; ++i)
      for (int j = 0; j < cols; ++j) {
        if (type_ == CV_32F)
          *reinterpret_cast<float*>(data + i * step.p[0] + j * 4) = (float)v.val[0];
        else
          data[i * step.p[0] + j] = (uchar)v.val[0];
      }
  }
  int type() const { return type_; }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  Size size() const { return Size(cols, rows); }
  uchar* ptr(int i = 0) { return data + (ptrdiff_t)i * (ptrdiff_t)step.p[0]; }
  const uchar* ptr(int i = 0) const { return data + (ptrdiff_t)i * (ptrdiff_t)step.p[0]; }
  template <class T>
  T* ptr(int i = 0) {
    return reinterpret_cast<T*>(ptr(i));
  }
  template <class T>
  T& at(int i, int j) {
    return *reinterpret_cast<T*>(data + (size_t)i * step.p[0] + (size_t)j * sizeof(T));
  }
  template <class T>
  const T& at(int i, int j) const {
    return *reinterpret_cast<const T*>(data + (size_t)i * step.p[0] + (size_t)j * sizeof(T));
  }
  template <class T>
  MatIterator_<T> begin() {
    return MatIterator_<T>(data, step.p[0], cols);
  }
  void copyTo(Mat& dst) const { dst = clone(); }
  Mat clone() const {
    Mat m(rows, cols, type_);
    for (int i = 0; i < rows; ++i) std::memcpy(m.ptr(i), ptr(i), (size_t)cols * step.p[1]);
    return m;
  }
};

// drawing / colour conversion named by the reference's display code (frame_handler_mono.cpp:280-300): declared so that
// the callers compile; nothing on the measured path calls them
enum { COLOR_GRAY2BGR = 8 };
inline void cvtColor(const Mat& src, Mat& dst, int) { dst = src.clone(); }
inline void rectangle(Mat&, Point, Point, const Scalar&, int = 1) {}
inline void rectangle(Mat&, Rect, const Scalar&, int = 1) {}
inline void line(Mat&, Point, Point, const Scalar&, int = 1) {}
inline void circle(Mat&, Point, int, const Scalar&, int = 1) {}

}  // namespace cv
#endif
