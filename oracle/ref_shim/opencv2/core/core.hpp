// Minimal stand-in for <opencv2/core/core.hpp> -- TEST INFRASTRUCTURE (oracle/_ref).
//
// OpenCV's C++ headers and libraries are not available in this image, so the reference's own sources
// (/root/reference/include/DistanceTransform.hpp, include/Math.hpp, src/HOGFeatures.cpp, src/DynamicProgram.cpp,
// include/Candidate.hpp) are compiled UNMODIFIED against this header: it declares just the slice of the cv:: API
// those files touch, with the documented OpenCV semantics (reference-counted Mat headers, saturate_cast = round
// half to even, Mat +/+=/> elementwise in the Mat's own depth, transpose that tolerates src == dst, Rect |, &).
// Everything numeric that matters (the DT, the max reductions, the HOG arithmetic, the DP) is the reference's code;
// this file only moves bytes.  cv::resize / cv::pyrDown are provided by oracle/ref_driver.cpp from the oracle's
// restatements, which tests/test_oracle_pins.py pins bit for bit against cv2.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#define CV_MAJOR_VERSION 2
#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_CN_SHIFT 3
#define CV_MAT_DEPTH(t) ((t) & 7)
#define CV_MAT_CN(t) ((((t) >> CV_CN_SHIFT) & 511) + 1)
#define CV_MAKETYPE(depth, cn) (CV_MAT_DEPTH(depth) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_StsUnsupportedFormat (-210)
#define CV_Error(code, msg) throw cv::Exception(code, msg)
#define CV_Assert(expr) do { if (!(expr)) throw cv::Exception(-215, #expr); } while (0)

namespace cv {

typedef unsigned char uchar;
typedef unsigned short ushort;

struct Exception : std::runtime_error {
  int code;
  Exception(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

inline int cvRound(double v) { return (int)std::lrint(v); }          // round half to even (default rounding mode), as OpenCV's SSE2 path
inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
template <typename T> inline T saturate_cast(double v) { return (T)v; }
template <> inline int saturate_cast<int>(double v) { return cvRound(v); }
template <> inline uchar saturate_cast<uchar>(double v) { int i = cvRound(v); return (uchar)(i < 0 ? 0 : i > 255 ? 255 : i); }
template <typename T> inline T saturate_cast(float v) { return saturate_cast<T>((double)v); }
template <typename T> inline T saturate_cast(int v) { return (T)v; }

template <typename T> struct DataType;
template <> struct DataType<uchar> { enum { depth = CV_8U, type = CV_8U }; };
template <> struct DataType<ushort> { enum { depth = CV_16U, type = CV_16U }; };
template <> struct DataType<int> { enum { depth = CV_32S, type = CV_32S }; };
template <> struct DataType<float> { enum { depth = CV_32F, type = CV_32F }; };
template <> struct DataType<double> { enum { depth = CV_64F, type = CV_64F }; };
inline size_t depth_bytes(int depth) { static const size_t s[7] = {1, 1, 2, 2, 4, 4, 8}; return s[depth]; }

template <typename T> struct Size_;
template <typename T> struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
  template <typename U> Point_(const Point_<U>& p) : x(saturate_cast<T>(p.x)), y(saturate_cast<T>(p.y)) {}
};
template <typename T> inline Point_<T> operator+(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(saturate_cast<T>(a.x + b.x), saturate_cast<T>(a.y + b.y)); }
template <typename T> inline Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(saturate_cast<T>(a.x - b.x), saturate_cast<T>(a.y - b.y)); }
template <typename T> inline Point_<T> operator*(const Point_<T>& a, int b) { return Point_<T>(saturate_cast<T>(a.x * b), saturate_cast<T>(a.y * b)); }
template <typename T> inline Point_<T> operator*(const Point_<T>& a, float b) { return Point_<T>(saturate_cast<T>(a.x * b), saturate_cast<T>(a.y * b)); }
template <typename T> inline Point_<T> operator*(const Point_<T>& a, double b) { return Point_<T>(saturate_cast<T>(a.x * b), saturate_cast<T>(a.y * b)); }
typedef Point_<int> Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;

template <typename T> struct Point3_ {
  T x, y, z;
  Point3_() : x(0), y(0), z(0) {}
  Point3_(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
  Point3_& operator*=(double s) { x = saturate_cast<T>(x * s); y = saturate_cast<T>(y * s); z = saturate_cast<T>(z * s); return *this; }
};
template <typename T> inline Point3_<T> operator+(const Point3_<T>& a, const Point3_<T>& b) { return Point3_<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename T> inline Point3_<T> operator-(const Point3_<T>& a, const Point3_<T>& b) { return Point3_<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
typedef Point3_<int> Point3i;
typedef Point3_<double> Point3d;

template <typename T> struct Size_ {
  T width, height;
  Size_() : width(0), height(0) {}
  Size_(T w, T h) : width(w), height(h) {}
  template <typename U> Size_(const Size_<U>& s) : width(saturate_cast<T>(s.width)), height(saturate_cast<T>(s.height)) {}
  T area() const { return width * height; }
  bool operator==(const Size_& o) const { return width == o.width && height == o.height; }
  bool operator!=(const Size_& o) const { return !(*this == o); }
};
template <typename T> inline Size_<T> operator*(const Size_<T>& a, T b) { return Size_<T>(a.width * b, a.height * b); }
typedef Size_<int> Size;

template <typename T> struct Rect_ {
  T x, y, width, height;
  Rect_() : x(0), y(0), width(0), height(0) {}
  Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
  Rect_(const Point_<T>& a, const Point_<T>& b) : x(std::min(a.x, b.x)), y(std::min(a.y, b.y)), width(std::max(a.x, b.x) - x), height(std::max(a.y, b.y) - y) {}
  Point_<T> tl() const { return Point_<T>(x, y); }
  Point_<T> br() const { return Point_<T>(x + width, y + height); }
  T area() const { return width * height; }
  bool contains(const Point_<T>& p) const { return x <= p.x && p.x < x + width && y <= p.y && p.y < y + height; }
};
template <typename T> inline Rect_<T>& operator&=(Rect_<T>& a, const Rect_<T>& b) {
  T x1 = std::max(a.x, b.x), y1 = std::max(a.y, b.y);
  a.width = std::min(a.x + a.width, b.x + b.width) - x1;
  a.height = std::min(a.y + a.height, b.y + b.height) - y1;
  a.x = x1; a.y = y1;
  if (a.width <= 0 || a.height <= 0) a = Rect_<T>();
  return a;
}
template <typename T> inline Rect_<T>& operator|=(Rect_<T>& a, const Rect_<T>& b) {
  if (a.width <= 0 || a.height <= 0) { a = b; return a; }            // OpenCV >= 3.3 ("empty" operands); 2.4 has no such guard but no shipped box is empty
  if (b.width <= 0 || b.height <= 0) return a;
  T x1 = std::min(a.x, b.x), y1 = std::min(a.y, b.y);
  a.width = std::max(a.x + a.width, b.x + b.width) - x1;
  a.height = std::max(a.y + a.height, b.y + b.height) - y1;
  a.x = x1; a.y = y1;
  return a;
}
template <typename T> inline Rect_<T> operator&(const Rect_<T>& a, const Rect_<T>& b) { Rect_<T> c = a; return c &= b; }
template <typename T> inline Rect_<T> operator|(const Rect_<T>& a, const Rect_<T>& b) { Rect_<T> c = a; return c |= b; }
template <typename T> inline Rect_<T> operator+(const Rect_<T>& a, const Size_<T>& s) { return Rect_<T>(a.x, a.y, a.width + s.width, a.height + s.height); }
template <typename T> inline Rect_<T> operator+(const Rect_<T>& a, const Point_<T>& p) { return Rect_<T>(a.x + p.x, a.y + p.y, a.width, a.height); }
typedef Rect_<int> Rect;

template <typename T> struct Scalar_ {
  T val[4];
  Scalar_() { val[0] = val[1] = val[2] = val[3] = 0; }
  Scalar_(T v0, T v1 = 0, T v2 = 0, T v3 = 0) { val[0] = v0; val[1] = v1; val[2] = v2; val[3] = v3; }
  T& operator[](int i) { return val[i]; }
  const T& operator[](int i) const { return val[i]; }
  T operator()(int i) const { return val[i]; }
};
typedef Scalar_<double> Scalar;

template <typename T, int m, int n> struct Matx {
  T val[m * n];
  Matx() { for (int i = 0; i < m * n; ++i) val[i] = 0; }
  Matx(T a, T b, T c) { val[0] = a; val[1] = b; val[2] = c; }
};

template <typename T> class Ptr : public std::shared_ptr<T> {
 public:
  Ptr() {}
  Ptr(T* p) : std::shared_ptr<T>(p) {}
  operator T*() const { return this->get(); }
};
class FilterEngine;     // only named in types.hpp (vectorFilterEngine)

template <typename T> class Mat_;
template <typename T> class MatIterator_;

struct Range {
  int start, end;
  Range() : start(0), end(0) {}
  Range(int s, int e) : start(s), end(e) {}
  int size() const { return end - start; }
};
// the slice of cv::MatExpr the reference uses: Mat_<T>::zeros / ones (optionally scaled).  Assigned to an existing matrix of the
// same size and type it fills that matrix IN PLACE (OpenCV's MatOp_Initializer::assign), which is how `roi = Mat_::zeros(...)`
// clears a region of interest in src/nms.cpp:114.
struct MatInit {
  int rows, cols, type;
  double value;
};
inline MatInit operator*(double s, MatInit m) { m.value *= s; return m; }
inline MatInit operator*(MatInit m, double s) { m.value *= s; return m; }

// Reference-counted 2-D array header (continuous rows or a region of interest inside a parent buffer).
class Mat {
 public:
  int flags = 0, rows = 0, cols = 0;
  uchar* data = nullptr;
  size_t step = 0;                           // bytes per row
  std::shared_ptr<uchar> buf;

  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(Size s, int type) { create(s.height, s.width, type); }
  Mat(int r, int c, int type, void* ext, size_t step_ = 0) : flags(type), rows(r), cols(c), data((uchar*)ext) {
    step = step_ ? step_ : (size_t)c * elemSize();
  }
  Mat(const MatInit& e) { create(e.rows, e.cols, e.type); setTo(Scalar_<double>(e.value)); }                 // NOLINT: implicit, as MatExpr -> Mat
  Mat& operator=(const MatInit& e) {
    if (!(data && rows == e.rows && cols == e.cols && flags == e.type)) create(e.rows, e.cols, e.type);
    return setTo(Scalar_<double>(e.value));
  }
  Mat operator()(const Range& r, const Range& c) const;
  Mat mul(const Mat& o) const;
  void create(int r, int c, int type) {
    if (data && r == rows && c == cols && type == flags && isContinuous()) return;
    flags = type; rows = r; cols = c;
    step = (size_t)c * elemSize();
    const size_t bytes = std::max<size_t>(step * (size_t)r, 1);
    buf = std::shared_ptr<uchar>(new uchar[bytes + 64], std::default_delete<uchar[]>());
    data = buf.get();
  }
  void create(Size s, int type) { create(s.height, s.width, type); }
  static Mat zeros(int r, int c, int type) { Mat m(r, c, type); for (int y = 0; y < r; ++y) std::memset(m.data + y * m.step, 0, (size_t)c * m.elemSize()); return m; }
  static Mat zeros(Size s, int type) { return zeros(s.height, s.width, type); }
  int type() const { return flags; }
  int depth() const { return CV_MAT_DEPTH(flags); }
  int channels() const { return CV_MAT_CN(flags); }
  size_t elemSize() const { return depth_bytes(depth()) * channels(); }
  size_t elemSize1() const { return depth_bytes(depth()); }
  size_t step1() const { return step / elemSize1(); }
  Size size() const { return Size(cols, rows); }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  bool isContinuous() const { return step == (size_t)cols * elemSize() || rows <= 1; }
  size_t total() const { return (size_t)rows * cols; }
  template <typename T> T* ptr(int y = 0) { return reinterpret_cast<T*>(data + (size_t)y * step); }
  template <typename T> const T* ptr(int y = 0) const { return reinterpret_cast<const T*>(data + (size_t)y * step); }
  uchar* ptr(int y = 0) { return data + (size_t)y * step; }
  const uchar* ptr(int y = 0) const { return data + (size_t)y * step; }
  template <typename T> T& at(int y, int x) { return ptr<T>(y)[x]; }
  template <typename T> const T& at(int y, int x) const { return ptr<T>(y)[x]; }
  template <typename T> T& at(Point p) { return ptr<T>(p.y)[p.x]; }
  template <typename T> const T& at(Point p) const { return ptr<T>(p.y)[p.x]; }
  void copyTo(Mat& dst) const {
    if (dst.data == data && dst.rows == rows && dst.cols == cols && dst.flags == flags) return;
    Mat out;
    if (dst.data && dst.rows == rows && dst.cols == cols && dst.flags == flags) out = dst; else out.create(rows, cols, flags);
    for (int y = 0; y < rows; ++y) std::memcpy(out.data + (size_t)y * out.step, data + (size_t)y * step, (size_t)cols * elemSize());
    dst = out;
  }
  Mat clone() const { Mat m; copyTo(m); return m; }
  void convertTo(Mat& dst, int rtype) const;
  Mat reshape(int cn) const {                 // same rows, channel count changed (continuous rows only)
    assert((cols * channels()) % cn == 0);
    Mat m = *this;
    m.cols = cols * channels() / cn;
    m.flags = CV_MAKETYPE(depth(), cn);
    return m;
  }
  Mat operator()(const Rect& r) const {
    Mat m = *this;
    m.rows = r.height; m.cols = r.width;
    m.data = data + (size_t)r.y * step + (size_t)r.x * elemSize();
    return m;
  }
  Mat& operator=(const Scalar_<double>& s) { return setTo(s); }
  Mat& setTo(const Scalar_<double>& s);
  Mat& setTo(const Scalar_<double>& s, const Mat& mask);
  template <typename T> MatIterator_<T> begin();
  template <typename T> MatIterator_<T> end();
};

template <typename T> class MatIterator_ {
 public:
  Mat* m = nullptr; size_t i = 0;
  typedef std::random_access_iterator_tag iterator_category;
  typedef T value_type; typedef ptrdiff_t difference_type; typedef T* pointer; typedef T& reference;
  MatIterator_() {}
  MatIterator_(Mat* m_, size_t i_) : m(m_), i(i_) {}
  T& operator*() const { return m->ptr<T>((int)(i / m->cols))[i % m->cols]; }
  T& operator[](ptrdiff_t d) const { return *(*this + d); }
  MatIterator_& operator++() { ++i; return *this; }
  MatIterator_ operator++(int) { MatIterator_ t = *this; ++i; return t; }
  MatIterator_& operator--() { --i; return *this; }
  MatIterator_& operator+=(ptrdiff_t d) { i += d; return *this; }
  MatIterator_& operator-=(ptrdiff_t d) { i -= d; return *this; }
  MatIterator_ operator+(ptrdiff_t d) const { return MatIterator_(m, i + d); }
  MatIterator_ operator-(ptrdiff_t d) const { return MatIterator_(m, i - d); }
  ptrdiff_t operator-(const MatIterator_& o) const { return (ptrdiff_t)i - (ptrdiff_t)o.i; }
  bool operator==(const MatIterator_& o) const { return i == o.i; }
  bool operator!=(const MatIterator_& o) const { return i != o.i; }
  bool operator<(const MatIterator_& o) const { return i < o.i; }
  bool operator<=(const MatIterator_& o) const { return i <= o.i; }
  bool operator>(const MatIterator_& o) const { return i > o.i; }
  bool operator>=(const MatIterator_& o) const { return i >= o.i; }
};
template <typename T> MatIterator_<T> Mat::begin() { return MatIterator_<T>(this, 0); }
template <typename T> MatIterator_<T> Mat::end() { return MatIterator_<T>(this, total()); }

namespace shim {
template <typename F> inline void by_depth(int depth, F f) {
  switch (depth) {
    case CV_8U: f((uchar)0); break;
    case CV_16U: f((ushort)0); break;
    case CV_32S: f((int)0); break;
    case CV_32F: f((float)0); break;
    case CV_64F: f((double)0); break;
    default: throw Exception(-210, "shim: unsupported depth");
  }
}
}  // namespace shim

inline Mat Mat::operator()(const Range& r, const Range& c) const {
  Mat m = *this;
  m.rows = r.size(); m.cols = c.size();
  m.data = data + (size_t)r.start * step + (size_t)c.start * elemSize();
  return m;
}
inline Mat Mat::mul(const Mat& o) const {                // elementwise product, saturated to the depth (8U masks: 255 * 255 -> 255)
  assert(size() == o.size() && type() == o.type());
  Mat out(rows, cols, flags);
  shim::by_depth(depth(), [&](auto t) {
    typedef decltype(t) T;
    for (int y = 0; y < rows; ++y) { const T* a = ptr<T>(y); const T* b = o.ptr<T>(y); T* d = out.ptr<T>(y); for (int x = 0; x < cols * channels(); ++x) d[x] = saturate_cast<T>((double)a[x] * (double)b[x]); }
  });
  return out;
}
inline void Mat::convertTo(Mat& dst, int rtype) const {
  const int ddepth = rtype < 0 ? depth() : CV_MAT_DEPTH(rtype);
  Mat out(rows, cols, CV_MAKETYPE(ddepth, channels()));
  const int n = cols * channels();
  shim::by_depth(depth(), [&](auto s) {
    typedef decltype(s) S;
    shim::by_depth(ddepth, [&](auto d) {
      typedef decltype(d) D;
      for (int y = 0; y < rows; ++y) {
        const S* sp = ptr<S>(y); D* dp = out.ptr<D>(y);
        for (int x = 0; x < n; ++x) dp[x] = saturate_cast<D>(sp[x]);
      }
    });
  });
  dst = out;
}
inline Mat& Mat::setTo(const Scalar& s) {
  const int n = cols * channels();
  shim::by_depth(depth(), [&](auto t) {
    typedef decltype(t) T;
    for (int y = 0; y < rows; ++y) { T* p = ptr<T>(y); for (int x = 0; x < n; ++x) p[x] = saturate_cast<T>(s.val[x % channels()]); }
  });
  return *this;
}

inline Mat& Mat::setTo(const Scalar& s, const Mat& mask) {
  shim::by_depth(depth(), [&](auto t) {
    typedef decltype(t) T;
    for (int y = 0; y < rows; ++y) { T* p = ptr<T>(y); const uchar* m = mask.ptr<uchar>(y); for (int x = 0; x < cols; ++x) if (m[x]) p[x] = saturate_cast<T>(s.val[0]); }
  });
  return *this;
}

template <typename T> class Mat_ : public Mat {
 public:
  Mat_() { flags = DataType<T>::type; }
  Mat_(int r, int c) : Mat(r, c, DataType<T>::type) {}
  explicit Mat_(Size s) : Mat(s, DataType<T>::type) {}
  Mat_(const Mat& m) { assign(m); }
  Mat_(const Mat_& m) : Mat(m) {}
  Mat_(const MatInit& e) : Mat(e) {}                         // NOLINT
  Mat_& operator=(const MatInit& e) { Mat::operator=(e); return *this; }
  static MatInit zeros(Size s) { return MatInit{s.height, s.width, DataType<T>::type, 0.0}; }
  static MatInit ones(Size s) { return MatInit{s.height, s.width, DataType<T>::type, 1.0}; }
  Mat_ operator()(const Range& r, const Range& c) const { return Mat_(Mat::operator()(r, c)); }
  Mat_ operator()(const Rect& r) const { return Mat_(Mat::operator()(r)); }
  Mat_& operator=(const Mat& m) { assign(m); return *this; }
  Mat_& operator=(const Mat_& m) { Mat::operator=(m); return *this; }
  void create(Size s) { Mat::create(s, DataType<T>::type); }
  void create(int r, int c) { Mat::create(r, c, DataType<T>::type); }
  T* operator[](int y) { return ptr<T>(y); }
  const T* operator[](int y) const { return ptr<T>(y); }
  T& operator()(int y, int x) { return ptr<T>(y)[x]; }
  const T& operator()(int y, int x) const { return ptr<T>(y)[x]; }
  T& operator()(int i) { return rows == 1 ? ptr<T>(0)[i] : ptr<T>(i / cols)[i % cols]; }
  const T& operator()(int i) const { return rows == 1 ? ptr<T>(0)[i] : ptr<T>(i / cols)[i % cols]; }
  void push_back(const T& v) {                // column vector growth (Candidate::boundingBox3D; not on the detect() path)
    Mat_<T> grown(rows + 1, 1);
    for (int y = 0; y < rows; ++y) grown(y, 0) = (*this)(y, 0);
    grown(rows, 0) = v;
    *this = grown;
  }
  MatIterator_<T> begin() { return Mat::begin<T>(); }
  MatIterator_<T> end() { return Mat::end<T>(); }
 private:
  void assign(const Mat& m) {                 // same type: shares the data (reference semantics); else converts
    if (m.empty() || m.type() == DataType<T>::type) { Mat::operator=(m); if (m.empty()) flags = DataType<T>::type; }
    else m.convertTo(*this, DataType<T>::type);
  }
};

std::ostream& operator<<(std::ostream& os, const Mat& m);

// ---- elementwise operations in the Mat's own depth (what cv::add / cv::compare do for same-depth operands; a scalar is
//      converted to the working depth first: arithm_op / compare in OpenCV's core/src/arithm.cpp) ----
inline Mat operator+(const Mat& a, double s) {
  Mat out(a.rows, a.cols, a.type());
  const int n = a.cols * a.channels();
  shim::by_depth(a.depth(), [&](auto t) {
    typedef decltype(t) T;
    const T sv = saturate_cast<T>(s);
    for (int y = 0; y < a.rows; ++y) { const T* p = a.ptr<T>(y); T* o = out.ptr<T>(y); for (int x = 0; x < n; ++x) o[x] = (T)(p[x] + sv); }
  });
  return out;
}
inline Mat& operator+=(Mat& a, const Mat& b) {
  assert(a.size() == b.size() && a.type() == b.type());
  const int n = a.cols * a.channels();
  shim::by_depth(a.depth(), [&](auto t) {
    typedef decltype(t) T;
    for (int y = 0; y < a.rows; ++y) { T* p = a.ptr<T>(y); const T* q = b.ptr<T>(y); for (int x = 0; x < n; ++x) p[x] = (T)(p[x] + q[x]); }
  });
  return a;
}
inline Mat operator>(const Mat& a, double s) {
  Mat out(a.rows, a.cols, CV_8U);
  shim::by_depth(a.depth(), [&](auto t) {
    typedef decltype(t) T;
    const T sv = (T)s;
    for (int y = 0; y < a.rows; ++y) { const T* p = a.ptr<T>(y); uchar* o = out.ptr<uchar>(y); for (int x = 0; x < a.cols; ++x) o[x] = p[x] > sv ? 255 : 0; }
  });
  return out;
}
inline Mat operator==(const Mat& a, double s) {
  Mat out(a.rows, a.cols, CV_8U);
  shim::by_depth(a.depth(), [&](auto t) {
    typedef decltype(t) T;
    const T sv = (T)s;
    for (int y = 0; y < a.rows; ++y) { const T* p = a.ptr<T>(y); uchar* o = out.ptr<uchar>(y); for (int x = 0; x < a.cols; ++x) o[x] = p[x] == sv ? 255 : 0; }
  });
  return out;
}
inline void transpose(const Mat& src, Mat& dst) {
  Mat out(src.cols, src.rows, src.type());
  const size_t es = src.elemSize();
  for (int y = 0; y < src.rows; ++y)
    for (int x = 0; x < src.cols; ++x) std::memcpy(out.data + (size_t)x * out.step + (size_t)y * es, src.data + (size_t)y * src.step + (size_t)x * es, es);
  dst = out;
}
template <typename T> inline void transpose(const Mat_<T>& src, Mat_<T>& dst) { Mat out; transpose(static_cast<const Mat&>(src), out); dst = out; }
inline Mat noArray() { return Mat(); }
// cv::minMaxLoc: first minimum / maximum in row-major order (strict comparisons); with a mask only its non-zero elements take part,
// and if there is none the values are 0 and the locations (-1, -1) (core/src/stat.cpp, minMaxIdx)
inline void minMaxLoc(const Mat& m, double* minv, double* maxv, Point* minloc = nullptr, Point* maxloc = nullptr, const Mat& mask = Mat()) {
  double lo = 0, hi = 0;
  Point lp(-1, -1), hp(-1, -1);
  bool any = false;
  shim::by_depth(m.depth(), [&](auto t) {
    typedef decltype(t) T;
    for (int y = 0; y < m.rows; ++y) {
      const T* p = m.ptr<T>(y);
      const uchar* k = mask.empty() ? nullptr : mask.ptr<uchar>(y);
      for (int x = 0; x < m.cols * m.channels(); ++x) {
        if (k && !k[x]) continue;
        const double v = (double)p[x];
        if (!any) { lo = hi = v; lp = hp = Point(x, y); any = true; continue; }
        if (v < lo) { lo = v; lp = Point(x, y); }
        if (v > hi) { hi = v; hp = Point(x, y); }
      }
    }
  });
  if (minv) *minv = lo;
  if (maxv) *maxv = hi;
  if (minloc) *minloc = lp;
  if (maxloc) *maxloc = hp;
}
template <typename T> inline double norm(const Point_<T>& p) { return std::sqrt((double)p.x * p.x + (double)p.y * p.y); }
inline Scalar sum(const Mat& m) {
  Scalar s;
  shim::by_depth(m.depth(), [&](auto t) {
    typedef decltype(t) T;
    for (int y = 0; y < m.rows; ++y) { const T* p = m.ptr<T>(y); for (int x = 0; x < m.cols * m.channels(); ++x) s.val[x % m.channels()] += (double)p[x]; }
  });
  return s;
}
inline void split(const Mat& m, std::vector<Mat>& out) {
  const int cn = m.channels();
  out.resize(cn);
  const size_t es = m.elemSize1();
  for (int c = 0; c < cn; ++c) {
    out[c].create(m.rows, m.cols, m.depth());
    for (int y = 0; y < m.rows; ++y)
      for (int x = 0; x < m.cols; ++x) std::memcpy(out[c].data + (size_t)y * out[c].step + (size_t)x * es, m.data + (size_t)y * m.step + ((size_t)x * cn + c) * es, es);
  }
}
void meanStdDev(const Mat& m, Scalar& mean, Scalar& stddev);

}  // namespace cv
