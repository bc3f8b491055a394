// pbd_oracle.cpp -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
//
// A plain C++17 + OpenMP restatement of the reference detector's hot path,
// PartsBasedDetector<T>::detect() (reference src/PartsBasedDetector.cpp:69-95).
// It exists only so that tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs can check and time the CUDA path against
// the reference's arithmetic.  Nothing under partsbaseddetector_b200/ may link,
// import or call it.
//
// The reference itself cannot be compiled in this image (it needs OpenCV C++
// headers/libs and Boost, neither is installed and there is no network), so
// this file follows the cited reference lines statement by statement instead.
// Every function names the reference file:line it restates.  OpenCV routines
// the reference calls (cv::resize, cv::pyrDown, FilterEngine/Filter2D,
// cv::transpose, Mat + scalar, cvRound) are restated from their published
// algorithm and pinned against the cv2 4.13 wheel in tests/test_oracle_pins.py.
//
// Pinning status ("parity pinned by cv2 + brute force, unpinned by reference
// tests"): the reference ships no golden vectors or unit tests for this path
// (test/CMakeLists.txt only instantiates the ORK pipeline), so the pins are:
// (all in tests/test_oracle_pins.py)
//   * image pyramid  == cv2.resize / cv2.pyrDown, bit-exact
//   * responses      == sum_c cv2.filter2D to 1e-5 (1e-11 in double)
//   * DT             == brute-force max, value and 1-D argmax
//   * DP             == exhaustive enumeration on a toy model
//   * HOG            == independent numpy formulation (1e-9 in double)
//   * committed golden vectors tests/golden/oracle_golden.npz
//
// Build: see oracle/Makefile (-O2 -fopenmp -ffp-contract=off, no -march: the
// reference sets no arch flags, so x86-64 baseline SSE2 => no FMA contraction).
//
// Deliberate deviations from literal HEAD (both are undefined behaviour there):
//   T2  defid is read as the full sequence (reference src/FileStorageModel.cpp:148-152
//       replaces multi-valued defid by [0] and then indexes out of bounds).
//   T4  bias(mm)[m] is read as biasw[biasid[p][mm] + m] (reference include/Parts.hpp:172-175
//       builds a temporary of length nmix(child) and indexes it by the parent mixture).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ---------------------------------------------------------------------------
// Model in flat form (built by the tests from cv2.FileStorage, the authority
// for the XML format; reference src/FileStorageModel.cpp:96-159).
// ---------------------------------------------------------------------------
struct OPart {
  int parentid;
  std::vector<int> filterid, biasid, defid;
};
struct OModel {
  int interval = 0, sbin = 0, norient = 0, flen = 0;
  float thresh = 0.f;
  std::vector<int> frows, fcols;              // filter kh, kw (Mat is kh x kw*flen)
  std::vector<std::vector<double>> filters;   // HWC, as stored in the XML (f64)
  std::vector<float> biasw;
  std::vector<int> anchors;                   // x,y pairs
  std::vector<float> defs;                    // 4 per def
  std::vector<std::vector<OPart>> comps;
};

template <typename T>
struct Map {
  int rows = 0, cols = 0;
  std::vector<T> d;
  Map() {}
  Map(int r, int c, T v = T(0)) : rows(r), cols(c), d((size_t)r * c, v) {}
  bool empty() const { return d.empty(); }
  T* row(int r) { return d.data() + (size_t)r * cols; }
  const T* row(int r) const { return d.data() + (size_t)r * cols; }
};

struct Image8 {
  int rows = 0, cols = 0, ch = 3;
  std::vector<uint8_t> d;
};

inline int cv_round_f(float v) { return (int)lrintf(v); }    // cvRound: SSE cvtss2si, half-to-even
inline int cv_round_d(double v) { return (int)lrint(v); }
inline int cv_floor_f(float v) { int i = (int)v; return i - (i > v); }
inline int16_t sat_short(float v) { int i = cv_round_f(v); return (int16_t)std::min(32767, std::max(-32768, i)); }
inline uint8_t sat_u8(int v) { return (uint8_t)std::min(255, std::max(0, v)); }

// ---------------------------------------------------------------------------
// cv::resize, INTER_LINEAR, CV_8U (OpenCV imgproc resize.cpp: HResizeLinear /
// VResizeLinear fixed-point path, INTER_RESIZE_COEF_BITS = 11).  Called by the
// reference at src/HOGFeatures.cpp:116.  Pinned bit-exact against cv2 4.13.
// ---------------------------------------------------------------------------
void resize_linear_u8(const Image8& src, Image8& dst, int dw, int dh) {
  const int cn = src.ch, sw = src.cols, sh = src.rows;
  dst.rows = dh; dst.cols = dw; dst.ch = cn;
  dst.d.assign((size_t)dw * dh * cn, 0);
  if (dw == sw && dh == sh) { dst.d = src.d; return; }      // resize.cpp: same size => copyTo
  const double inv_scale_x = (double)dw / sw, inv_scale_y = (double)dh / sh;
  const double scale_x = 1. / inv_scale_x, scale_y = 1. / inv_scale_y;
  std::vector<int> xofs(dw), yofs(dh);
  std::vector<int16_t> ialpha((size_t)dw * 2), ibeta((size_t)dh * 2);
  for (int dx = 0; dx < dw; ++dx) {
    float fx = (float)((dx + 0.5) * scale_x - 0.5);
    int sx = cv_floor_f(fx);
    fx -= sx;
    if (sx < 0) { fx = 0; sx = 0; }
    if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
    xofs[dx] = sx;
    ialpha[dx * 2] = sat_short((1.f - fx) * 2048.f);
    ialpha[dx * 2 + 1] = sat_short(fx * 2048.f);
  }
  for (int dy = 0; dy < dh; ++dy) {
    float fy = (float)((dy + 0.5) * scale_y - 0.5);
    int sy = cv_floor_f(fy);
    fy -= sy;
    yofs[dy] = sy;
    ibeta[dy * 2] = sat_short((1.f - fy) * 2048.f);
    ibeta[dy * 2 + 1] = sat_short(fy * 2048.f);
  }
  std::vector<int> r0((size_t)dw * cn), r1((size_t)dw * cn);
  auto hrow = [&](int sy, std::vector<int>& out) {
    sy = std::min(std::max(sy, 0), sh - 1);                 // clip(sy0 - ksize2 + 1 + k, 0, ssize.height)
    const uint8_t* S = src.d.data() + (size_t)sy * sw * cn;
    for (int dx = 0; dx < dw; ++dx) {
      const int sx = xofs[dx], sx1 = std::min(sx + 1, sw - 1);
      const int a0 = ialpha[dx * 2], a1 = ialpha[dx * 2 + 1];
      for (int c = 0; c < cn; ++c) out[dx * cn + c] = S[sx * cn + c] * a0 + S[sx1 * cn + c] * a1;
    }
  };
  for (int dy = 0; dy < dh; ++dy) {
    hrow(yofs[dy], r0);
    hrow(yofs[dy] + 1, r1);
    const int b0 = ibeta[dy * 2], b1 = ibeta[dy * 2 + 1];
    uint8_t* D = dst.d.data() + (size_t)dy * dw * cn;
    for (int i = 0; i < dw * cn; ++i)
      D[i] = sat_u8((((b0 * (r0[i] >> 4)) >> 16) + ((b1 * (r1[i] >> 4)) >> 16) + 2) >> 2);
  }
}

// ---------------------------------------------------------------------------
// cv::pyrDown, CV_8U, BORDER_REFLECT_101 (OpenCV imgproc pyramids.cpp).  Called
// by the reference at src/HOGFeatures.cpp:122.  Pinned bit-exact against cv2.
// ---------------------------------------------------------------------------
inline int reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) { if (p < 0) p = -p; else p = 2 * n - 2 - p; }
  return p;
}
void pyrdown_u8(const Image8& src, Image8& dst) {
  const int cn = src.ch, sw = src.cols, sh = src.rows;
  const int dw = (sw + 1) / 2, dh = (sh + 1) / 2;
  dst.rows = dh; dst.cols = dw; dst.ch = cn;
  dst.d.assign((size_t)dw * dh * cn, 0);
  static const int k[5] = {1, 4, 6, 4, 1};
  std::vector<int> hrows((size_t)5 * dw * cn);
  for (int y = 0; y < dh; ++y) {
    for (int i = 0; i < 5; ++i) {
      const int sy = reflect101(2 * y + i - 2, sh);
      const uint8_t* S = src.d.data() + (size_t)sy * sw * cn;
      int* H = hrows.data() + (size_t)i * dw * cn;
      for (int x = 0; x < dw; ++x)
        for (int c = 0; c < cn; ++c) {
          int s = 0;
          for (int j = 0; j < 5; ++j) s += k[j] * S[reflect101(2 * x + j - 2, sw) * cn + c];
          H[x * cn + c] = s;
        }
    }
    uint8_t* D = dst.d.data() + (size_t)y * dw * cn;
    for (int i = 0; i < dw * cn; ++i) {
      int s = 0;
      for (int r = 0; r < 5; ++r) s += k[r] * hrows[(size_t)r * dw * cn + i];
      D[i] = (uint8_t)((s + 128) >> 8);
    }
  }
}

// ---------------------------------------------------------------------------
// HOGFeatures<T>::pyramid geometry (reference src/HOGFeatures.cpp:95-127,
// include/HOGFeatures.hpp:74-81).  With <cmath> and `using namespace std`
// (src/HOGFeatures.cpp:46-47) log/pow/floor on float arguments resolve to the
// float overloads, pow(float,int) promotes to double.
// ---------------------------------------------------------------------------
struct LevelGeom { int img_w, img_h; float scale; };
int pyramid_geometry(int h, int w, int sbin, int interval, std::vector<LevelGeom>& lv) {
  const float sfactor = std::pow(2.0f, 1.0f / (float)interval);             // HOGFeatures.hpp:78
  const float fw = (float)w, fh = (float)h;                                  // Size_<float> imsize, :98
  const float ns = 1 + std::floor(std::log(std::min(fh, fw) / (5.0f * (float)sbin)) / std::log(sfactor));  // :99
  const int nscales = ns > 0 ? (int)ns : 0;
  lv.assign(nscales, LevelGeom{0, 0, 0.f});
  for (int i = 0; i < interval && i < nscales; ++i) {       // guard: reference writes past the vector if interval > nscales
    const float s = (float)(1.0f / std::pow((double)sfactor, i));            // :116
    int cw = cv_round_f(fw * s), ch = cv_round_f(fh * s);                    // Size_<float> -> Size: saturate_cast = cvRound
    lv[i] = LevelGeom{cw, ch, (float)(std::pow((double)sfactor, i) * sbin)}; // :118
    for (int j = i + interval; j < nscales; j += interval) {                 // :120-126
      cw = (cw + 1) / 2; ch = (ch + 1) / 2;
      lv[j] = LevelGeom{cw, ch, 2 * lv[j - interval].scale};
    }
  }
  return nscales;
}

// ---------------------------------------------------------------------------
// HOGFeatures<T>::features<uint8_t> (reference src/HOGFeatures.cpp:168-341).
// ---------------------------------------------------------------------------
template <typename T> inline T sq(const T& x) { return x * x; }

template <typename T>
void hog_features(const Image8& imm, int binsize, int norient, int flen, Map<T>& featm, int& out_w, int& out_h) {
  const bool color = imm.ch == 3;
  const int cols = imm.cols, rows = imm.rows;
  const int bw = (int)std::round((float)cols / (float)binsize);              // :174
  const int bh = (int)std::round((float)rows / (float)binsize);
  const int ow = std::max(bw - 2, 0), oh = std::max(bh - 2, 0);              // :175
  const int vis_w = bw * binsize, vis_h = bh * binsize;                      // :176
  out_w = ow; out_h = oh;
  std::vector<T> hist((size_t)bw * norient * bh, T(0));
  std::vector<T> norm((size_t)bw * bh, T(0));
  featm = Map<T>(oh, ow * flen, T(0));
  const size_t imstride = (size_t)cols * imm.ch;
  const size_t histstride = (size_t)bw * norient, normstride = bw, featstride = (size_t)ow * flen;
  const double eps = 0.0001;                                                 // :189
  const T uu[9] = {(T)1.000, (T)0.9397, (T)0.7660, (T)0.5000, (T)0.1736, (T)-0.1736, (T)-0.5000, (T)-0.7660, (T)-0.9397};
  const T vv[9] = {(T)0.000, (T)0.3420, (T)0.6428, (T)0.8660, (T)0.9848, (T)0.9848, (T)0.8660, (T)0.6428, (T)0.3420};
  const uint8_t* im = imm.d.data();
  const int half = norient / 2;

  for (int y = 1; y < vis_h - 1; ++y) {                                      // :202
    for (int x = 1; x < vis_w - 1; ++x) {
      T dx, dy, v;
      const int sx = std::min(x, cols - 2), sy = std::min(y, rows - 2);
      if (!color) {                                                          // :207-212
        const uint8_t* s = im + sx + (size_t)sy * imstride;
        dy = (T)(*(s + imstride) - *(s - imstride));
        dx = (T)(*(s + 1) - *(s - 1));
        v = dx * dx + dy * dy;
      } else {                                                               // :217-240
        const uint8_t* s = im + 3 * sx + (size_t)sy * imstride;
        T dyb = (T)(*(s + imstride) - *(s - imstride));
        T dxb = (T)(*(s + 3) - *(s - 3));
        T vb = dxb * dxb + dyb * dyb;
        s += 1;
        T dyg = (T)(*(s + imstride) - *(s - imstride));
        T dxg = (T)(*(s + 3) - *(s - 3));
        T vg = dxg * dxg + dyg * dyg;
        s += 1;
        dy = (T)(*(s + imstride) - *(s - imstride));
        dx = (T)(*(s + 3) - *(s - 3));
        v = dx * dx + dy * dy;
        if (vg > v) { v = vg; dx = dxg; dy = dyg; }
        if (vb > v) { v = vb; dx = dxb; dy = dyb; }
      }
      T best_dot = 0;                                                        // :243-249
      int best_o = 0;
      for (int o = 0; o < half; ++o) {
        T dot = uu[o] * dx + vv[o] * dy;
        if (dot > best_dot) { best_dot = dot; best_o = o; }
        else if (-dot > best_dot) { best_dot = -dot; best_o = o + half; }
      }
      T yp = (T)(((T)y + 0.5) / (T)binsize - 0.5);                           // :252-259 (double intermediates)
      T xp = (T)(((T)x + 0.5) / (T)binsize - 0.5);
      int iyp = (int)std::floor(yp);
      int ixp = (int)std::floor(xp);
      T vy0 = yp - iyp, vx0 = xp - ixp;
      T vy1 = (T)(1.0 - vy0), vx1 = (T)(1.0 - vx0);
      v = std::sqrt(v);                                                      // :260
      if (iyp >= 0 && ixp >= 0)           hist[iyp * histstride + ixp * norient + best_o] += vy1 * vx1 * v;           // :262
      if (iyp >= 0 && ixp + 1 < bw)       hist[iyp * histstride + (ixp + 1) * norient + best_o] += vx0 * vy1 * v;     // :263
      if (iyp + 1 < bh && ixp >= 0)       hist[(iyp + 1) * histstride + ixp * norient + best_o] += vy0 * vx1 * v;     // :264
      if (iyp + 1 < bh && ixp + 1 < bw)   hist[(iyp + 1) * histstride + (ixp + 1) * norient + best_o] += vy0 * vx0 * v; // :265
    }
  }
  for (int y = 0; y < bh; ++y) {                                             // :270-283
    const T* src = hist.data() + y * histstride;
    T* dst = norm.data() + y * normstride;
    for (int x = 0; x < bw; ++x) {
      T acc = 0;
      for (int o = 0; o < half; ++o) { acc += sq<T>(*src + *(src + half)); src++; }
      *dst++ = acc;
      src += half;
    }
  }
  for (int y = 0; y < oh; ++y) {                                             // :286-340
    for (int x = 0; x < ow; ++x) {
      T* dst = featm.d.data() + y * featstride + (size_t)x * flen;
      const T* p;
      T n1, n2, n3, n4;
      p = norm.data() + (y + 1) * normstride + (x + 1);
      n1 = (T)(1.0f / std::sqrt(*p + *(p + 1) + *(p + normstride) + *(p + normstride + 1) + eps));
      p = norm.data() + y * normstride + (x + 1);
      n2 = (T)(1.0f / std::sqrt(*p + *(p + 1) + *(p + normstride) + *(p + normstride + 1) + eps));
      p = norm.data() + (y + 1) * normstride + x;
      n3 = (T)(1.0f / std::sqrt(*p + *(p + 1) + *(p + normstride) + *(p + normstride + 1) + eps));
      p = norm.data() + y * normstride + x;
      n4 = (T)(1.0f / std::sqrt(*p + *(p + 1) + *(p + normstride) + *(p + normstride + 1) + eps));
      T t1 = 0, t2 = 0, t3 = 0, t4 = 0;
      const T* src = hist.data() + (y + 1) * histstride + (size_t)(x + 1) * norient;
      for (int o = 0; o < norient; ++o) {                                    // :305-317
        T val = *src;
        T h1 = std::min(val * n1, (T)0.2), h2 = std::min(val * n2, (T)0.2);
        T h3 = std::min(val * n3, (T)0.2), h4 = std::min(val * n4, (T)0.2);
        *(dst++) = (T)(0.5 * (h1 + h2 + h3 + h4));
        src++;
        t1 += h1; t2 += h2; t3 += h3; t4 += h4;
      }
      src = hist.data() + (y + 1) * histstride + (size_t)(x + 1) * norient;
      for (int o = 0; o < half; ++o) {                                       // :321-329
        T sum = *src + *(src + half);
        T h1 = std::min(sum * n1, (T)0.2), h2 = std::min(sum * n2, (T)0.2);
        T h3 = std::min(sum * n3, (T)0.2), h4 = std::min(sum * n4, (T)0.2);
        *(dst++) = (T)(0.5 * (h1 + h2 + h3 + h4));
        src++;
      }
      *(dst++) = (T)(0.2357 * t1);                                           // :332-335
      *(dst++) = (T)(0.2357 * t2);
      *(dst++) = (T)(0.2357 * t3);
      *(dst++) = (T)(0.2357 * t4);
      *dst = 0;                                                              // :338
    }
  }
}

// ---------------------------------------------------------------------------
// SpatialConvolutionEngine::convolve (reference src/SpatialConvolutionEngine.cpp:70-94)
// over cv::Filter2D (src/filter.cpp:3879-3924, SSE path :2186-2225): per channel
// plane c a "same"-size correlation with anchor (kw/2, kh/2)
// (include/filterengine.hpp:310-318), BORDER_CONSTANT 0 for c < flen-1 and 1 for
// the last channel (src/SpatialConvolutionEngine.cpp:147-156), accumulator starts
// at delta = 0 and adds f[k]*src[k] over the non-zero taps in row-major order
// (preprocess2DKernel, src/filter.cpp:3808-3857) with separate multiply and add;
// then pdf = 0; pdf += R_c for c = 0..flen-1.
// ---------------------------------------------------------------------------
template <typename T>
void convolve(const Map<T>& feat, int oh, int ow, int flen, const std::vector<T>& filt, int kh, int kw, Map<T>& pdf) {
  pdf = Map<T>(oh, ow, T(0));
  const int ax = kw / 2, ay = kh / 2;
  const int pw = ow + kw - 1, ph = oh + kh - 1;
  std::vector<T> plane((size_t)pw * ph);
  std::vector<T> acc(ow);
  for (int c = 0; c < flen; ++c) {
    const T border = (c == flen - 1) ? T(1) : T(0);
    std::fill(plane.begin(), plane.end(), border);
    for (int y = 0; y < oh; ++y) {
      const T* f = feat.row(y);
      T* p = plane.data() + (size_t)(y + ay) * pw + ax;
      for (int x = 0; x < ow; ++x) p[x] = f[(size_t)x * flen + c];
    }
    for (int y = 0; y < oh; ++y) {
      std::fill(acc.begin(), acc.end(), T(0));                // s0 = delta = 0
      for (int ky = 0; ky < kh; ++ky) {
        for (int kx = 0; kx < kw; ++kx) {
          const T w = filt[((size_t)ky * kw + kx) * flen + c];
          if (w == 0) continue;                               // preprocess2DKernel keeps non-zero taps only
          const T* s = plane.data() + (size_t)(y + ky) * pw + kx;
          for (int x = 0; x < ow; ++x) acc[x] = acc[x] + w * s[x];
        }
      }
      T* d = pdf.row(y);
      for (int x = 0; x < ow; ++x) d[x] = d[x] + acc[x];      // pdf += pdfc, :92
    }
  }
}

// ---------------------------------------------------------------------------
// DistanceTransform<T> (reference include/DistanceTransform.hpp:89-105 Quadratic,
// :152-182 computeRow, :203-245 compute).
// ---------------------------------------------------------------------------
struct Quadratic {
  double a, b;
  static int square(int x) { return x * x; }
  double isect(int x0, int x1, double y0, double y1) const {                 // :98-100
    return ((y1 - y0) - b * (x1 - x0) + a * (square(x1) - square(x0))) / (2 * a * (x1 - x0));
  }
  double val(int x, double y) const { return a * square(x) + b * x + y; }    // :102-104
};

template <typename T>
void dt_row(const T* src, T* dst, int* ptr, int N, const Quadratic& f, int os) {  // :152-182
  std::vector<int> v(N);
  std::vector<T> z(N + 1);
  int k = 0;
  v[0] = 0;
  z[0] = -std::numeric_limits<T>::infinity();
  z[1] = +std::numeric_limits<T>::infinity();
  for (int q = 1; q < N; ++q) {
    T s = (T)f.isect(v[k], q, src[v[k]], src[q]);
    while (s <= z[k] && k > 0) {
      k--;
      s = (T)f.isect(v[k], q, src[v[k]], src[q]);
    }
    k++;
    v[k] = q;
    z[k] = s;
    z[k + 1] = +std::numeric_limits<T>::infinity();
  }
  k = 0;
  for (int q = 0; q < N; ++q) {
    while (z[k + 1] < os) k++;
    dst[q] = (T)f.val(os - v[k], src[v[k]]);
    ptr[q] = v[k];
    os++;
  }
}

// backptr_mode 0 = reference composition (T1), 1 = exact 2-D argmax composition
template <typename T>
void dt_2d(const Map<T>& in, const Quadratic& fx, const Quadratic& fy, int osx, int osy,
           Map<T>& out, Map<int>& Ix, Map<int>& Iy, int backptr_mode) {       // :203-245
  const int M = in.rows, N = in.cols;
  Map<T> tmp(M, N);
  Ix = Map<int>(M, N);
  for (int m = 0; m < M; ++m) dt_row(in.row(m), tmp.row(m), Ix.row(m), N, fx, osx);
  Map<T> tmpT(N, M), outT(N, M);
  Map<int> IyT(N, M);
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) tmpT.d[(size_t)n * M + m] = tmp.d[(size_t)m * N + n];
  for (int n = 0; n < N; ++n) dt_row(tmpT.row(n), outT.row(n), IyT.row(n), M, fy, osy);
  out = Map<T>(M, N);
  Iy = Map<int>(M, N);
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
    out.d[(size_t)m * N + n] = outT.d[(size_t)n * M + m];
    Iy.d[(size_t)m * N + n] = IyT.d[(size_t)n * M + m];
  }
  if (backptr_mode == 0) {
    std::vector<int> row(N);                                                 // :233-244
    for (int m = 0; m < M; ++m) {
      int* Iy_ptr = Iy.row(m);
      const int* Ix_ptr = Ix.row(m);
      for (int n = 0; n < N; ++n) row[n] = Iy_ptr[Ix_ptr[n]];
      for (int n = 0; n < N; ++n) Iy_ptr[n] = row[n];
    }
  } else {
    Map<int> Ix2(M, N);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) Ix2.d[(size_t)m * N + n] = Ix.d[(size_t)Iy.d[(size_t)m * N + n] * N + n];
    Ix = Ix2;
  }
}

// Math::reduceMax (reference include/Math.hpp:149-185)
template <typename T>
void reduce_max(const std::vector<Map<T>>& in, Map<T>& maxv, Map<int>& maxi) {
  const size_t K = in.size();
  maxv = Map<T>(in[0].rows, in[0].cols);
  maxi = Map<int>(in[0].rows, in[0].cols, 0);
  if (K == 1) { maxv = in[0]; return; }
  const size_t n = in[0].d.size();
  for (size_t i = 0; i < n; ++i) {
    T v = -std::numeric_limits<T>::infinity();
    int idx = 0;
    for (size_t k = 0; k < K; ++k) if (in[k].d[i] > v) { idx = (int)k; v = in[k].d[i]; }
    maxi.d[i] = idx;
    maxv.d[i] = v;
  }
}
// Math::reducePickIndex (reference include/Math.hpp:109-135)
void reduce_pick(const std::vector<Map<int>>& in, const Map<int>& idx, Map<int>& out) {
  if (in.size() == 1) { out = in[0]; return; }
  out = Map<int>(in[0].rows, in[0].cols);
  for (size_t i = 0; i < out.d.size(); ++i) out.d[i] = in[idx.d[i]].d[i];
}

struct OCandidate {
  int level, component;
  float score;
  std::vector<int> x, y, m;            // part locations in cells + mixture ids (not stored by the reference Candidate)
  std::vector<int> rect;               // x,y,w,h per part (cv::Rect), reference include/Candidate.hpp:72
};

template <typename T>
struct Detector {
  OModel model;
  std::vector<std::vector<T>> filtersT;       // convertTo(T), reference src/PartsBasedDetector.cpp:115-117
  int backptr_mode = 0;
  double thresh = 0;
  int max_levels = 0;                         // 0 = all
  // last run
  std::vector<LevelGeom> geom;
  std::vector<Image8> pyr;
  std::vector<Map<T>> feats;
  std::vector<int> ow, oh;
  std::vector<std::vector<Map<T>>> resp;      // [level][filter]
  // [level][comp][part][parent mixture]
  std::vector<std::vector<std::vector<std::vector<Map<int>>>>> Ix, Iy, Ik;
  std::vector<std::vector<Map<T>>> rootv;
  std::vector<std::vector<Map<int>>> rooti;
  std::vector<OCandidate> cands;
  double t_pyr = 0, t_hog = 0, t_pdf = 0, t_min = 0, t_argmin = 0;
};

double now_s() {
#ifdef _OPENMP
  return omp_get_wtime();
#else
  return 0;
#endif
}

// HOGFeatures<T>::pyramid (reference src/HOGFeatures.cpp:95-151)
template <typename T>
void stage_pyramid(Detector<T>& D, const Image8& im) {
  const OModel& mo = D.model;
  double t0 = now_s();
  int ns = pyramid_geometry(im.rows, im.cols, mo.sbin, mo.interval, D.geom);
  D.pyr.assign(ns, Image8());
  const int interval = mo.interval;
#pragma omp parallel for schedule(dynamic, 1)
  for (int i = 0; i < interval; ++i) {                                       // :111-127
    if (i >= ns) continue;
    resize_linear_u8(im, D.pyr[i], D.geom[i].img_w, D.geom[i].img_h);
    for (int j = i + interval; j < ns; j += interval) pyrdown_u8(D.pyr[j - interval], D.pyr[j]);
  }
  if (D.max_levels > 0 && ns > D.max_levels) { ns = D.max_levels; D.geom.resize(ns); D.pyr.resize(ns); }
  double t1 = now_s();
  D.feats.assign(ns, Map<T>());
  D.ow.assign(ns, 0); D.oh.assign(ns, 0);
#pragma omp parallel for schedule(dynamic, 1)
  for (int n = 0; n < ns; ++n)                                               // :130-150
    hog_features<T>(D.pyr[n], mo.sbin, mo.norient, mo.flen, D.feats[n], D.ow[n], D.oh[n]);
  double t2 = now_s();
  D.t_pyr = t1 - t0; D.t_hog = t2 - t1;
}

// SpatialConvolutionEngine::pdf (reference src/SpatialConvolutionEngine.cpp:106-124)
template <typename T>
void stage_pdf(Detector<T>& D) {
  double t0 = now_s();
  const int M = (int)D.feats.size(), N = (int)D.filtersT.size();
  D.resp.assign(M, std::vector<Map<T>>(N));
#pragma omp parallel for schedule(dynamic, 1)
  for (int n = 0; n < N; ++n)
    for (int m = 0; m < M; ++m)
      convolve<T>(D.feats[m], D.oh[m], D.ow[m], D.model.flen, D.filtersT[n], D.model.frows[n], D.model.fcols[n], D.resp[m][n]);
  D.t_pdf = now_s() - t0;
}

// DynamicProgram<T>::min (reference src/DynamicProgram.cpp:67-173)
template <typename T>
void stage_min(Detector<T>& D) {
  double t0 = now_s();
  const OModel& mo = D.model;
  const int nscales = (int)D.resp.size(), ncomp = (int)mo.comps.size();
  D.Ix.assign(nscales, {}); D.Iy.assign(nscales, {}); D.Ik.assign(nscales, {});
  D.rootv.assign(nscales, std::vector<Map<T>>(ncomp));
  D.rooti.assign(nscales, std::vector<Map<int>>(ncomp));
  for (int n = 0; n < nscales; ++n) { D.Ix[n].resize(ncomp); D.Iy[n].resize(ncomp); D.Ik[n].resize(ncomp); }
#pragma omp parallel for schedule(dynamic, 1)
  for (int nc = 0; nc < nscales * ncomp; ++nc) {                             // :83
    const int n = nc / ncomp, c = nc % ncomp;
    const std::vector<OPart>& parts = mo.comps[c];
    const int nparts = (int)parts.size();
    D.Ix[n][c].assign(nparts, {}); D.Iy[n][c].assign(nparts, {}); D.Ik[n][c].assign(nparts, {});
    std::vector<Map<T>> ncscores(D.resp[n].size());                          // :93
    for (int p = nparts - 1; p > 0; --p) {                                   // :95
      const OPart& cp = parts[p];
      const OPart& par = parts[cp.parentid];
      const int nmix = (int)cp.filterid.size(), pnmix = (int)par.filterid.size();
      D.Ix[n][c][p].resize(pnmix); D.Iy[n][c][p].resize(pnmix); D.Ik[n][c][p].resize(pnmix);
      std::vector<Map<T>> scoresp(nmix);
      std::vector<Map<int>> Ixp(nmix), Iyp(nmix);
      for (int m = 0; m < nmix; ++m) {                                       // :110-132
        const int fid = cp.filterid[m];
        const Map<T>& score_in = ncscores[fid].empty() ? D.resp[n][fid] : ncscores[fid];
        const int did = cp.defid[m];
        const float* w = &mo.defs[(size_t)did * 4];
        Quadratic fx{-(double)w[0], -(double)w[1]}, fy{-(double)w[2], -(double)w[3]};   // :126-127 (-float -> double)
        dt_2d<T>(score_in, fx, fy, mo.anchors[did * 2], mo.anchors[did * 2 + 1], scoresp[m], Ixp[m], Iyp[m], D.backptr_mode);
      }
      for (int m = 0; m < pnmix; ++m) {                                      // :134-160
        std::vector<Map<T>> weighted(nmix);
        for (int mm = 0; mm < nmix; ++mm) {
          const T b = (T)mo.biasw[cp.biasid[mm] + m];                        // :139 (T4: flat indexing)
          weighted[mm] = scoresp[mm];
          for (auto& e : weighted[mm].d) e = e + b;
        }
        Map<T> maxv; Map<int> maxi;
        reduce_max<T>(weighted, maxv, maxi);                                 // :143
        reduce_pick(Ixp, maxi, D.Ix[n][c][p][m]);                            // :147-148
        reduce_pick(Iyp, maxi, D.Iy[n][c][p][m]);
        D.Ik[n][c][p][m] = maxi;
        const int pfid = par.filterid[m];
        if (ncscores[pfid].empty()) ncscores[pfid] = D.resp[n][pfid];        // :155
        for (size_t i = 0; i < maxv.d.size(); ++i) ncscores[pfid].d[i] = ncscores[pfid].d[i] + maxv.d[i];   // :156
      }
    }
    const OPart& root = parts[0];                                            // :163-171
    const T bias = (T)mo.biasw[root.biasid[0]];
    std::vector<Map<T>> weighted(root.filterid.size());
    for (size_t m = 0; m < root.filterid.size(); ++m) {
      const int fid = root.filterid[m];
      weighted[m] = ncscores[fid].empty() ? D.resp[n][fid] : ncscores[fid];
      // NB: for a single-part component the reference adds the bias to an empty Mat;
      // the evident intent (the root's own response) is used here.
      for (auto& e : weighted[m].d) e = e + bias;
    }
    reduce_max<T>(weighted, D.rootv[n][c], D.rooti[n][c]);
  }
  D.t_min = now_s() - t0;
}

// DynamicProgram<T>::argmin (reference src/DynamicProgram.cpp:190-255).  Candidates
// are emitted in the deterministic single-threaded order (level, component, row-major).
template <typename T>
void stage_argmin(Detector<T>& D) {
  double t0 = now_s();
  const OModel& mo = D.model;
  D.cands.clear();
  const int nscales = (int)D.rootv.size();
  for (int n = 0; n < nscales; ++n) {
    const T scale = (T)D.geom[n].scale;                                      // :198
    for (int c = 0; c < (int)mo.comps.size(); ++c) {
      const std::vector<OPart>& parts = mo.comps[c];
      const int nparts = (int)parts.size();
      const Map<T>& rv = D.rootv[n][c];
      const T th = (T)D.thresh;                                              // compare(Mat, scalar): scalar converted to the Mat depth
      for (int yy = 0; yy < rv.rows; ++yy) for (int xx = 0; xx < rv.cols; ++xx) {   // Math::find, include/Math.hpp:84-93
        if (!(rv.d[(size_t)yy * rv.cols + xx] > th)) continue;               // :208
        OCandidate cand;
        cand.level = n; cand.component = c;
        cand.x.assign(nparts, 0); cand.y.assign(nparts, 0); cand.m.assign(nparts, 0);
        for (int p = 0; p < nparts; ++p) {                                   // :219-245
          if (p == 0) {
            cand.x[0] = xx; cand.y[0] = yy; cand.m[0] = D.rooti[n][c].d[(size_t)yy * rv.cols + xx];
          } else {
            const int idx = parts[p].parentid;
            const int x = cand.x[idx], y = cand.y[idx], m = cand.m[idx];
            const size_t o = (size_t)y * rv.cols + x;
            cand.x[p] = D.Ix[n][c][p][m].d[o];
            cand.y[p] = D.Iy[n][c][p][m].d[o];
            cand.m[p] = D.Ik[n][c][p][m].d[o];
          }
          const int ks = mo.frows[parts[p].filterid[cand.m[p]]];             // xsize == ysize == rows (T5), Parts.hpp:185-187
          int x1, y1, sx, sy;
          if (sizeof(T) == 4) {                                              // Point * T: saturate_cast<int> = cvRound
            x1 = cv_round_f((float)(cand.x[p] - 1) * (float)scale); y1 = cv_round_f((float)(cand.y[p] - 1) * (float)scale);
            sx = cv_round_f((float)ks * (float)scale); sy = sx;
          } else {
            x1 = cv_round_d((double)(cand.x[p] - 1) * (double)scale); y1 = cv_round_d((double)(cand.y[p] - 1) * (double)scale);
            sx = cv_round_d((double)ks * (double)scale); sy = sx;
          }
          const int x2 = x1 + sx - 1, y2 = y1 + sy - 1;                      // :240
          const int rx = std::min(x1, x2), ry = std::min(y1, y2);            // cv::Rect(pt1, pt2)
          cand.rect.push_back(rx); cand.rect.push_back(ry);
          cand.rect.push_back(std::max(x1, x2) - rx); cand.rect.push_back(std::max(y1, y2) - ry);
        }
        cand.score = (float)rv.d[(size_t)yy * rv.cols + xx];                 // :242, addPart(Rect, float)
        D.cands.push_back(cand);
      }
    }
  }
  D.t_argmin = now_s() - t0;
}

template <typename T>
Detector<T>* make_detector(const OModel& mo) {
  Detector<T>* D = new Detector<T>();
  D->model = mo;
  D->thresh = (double)mo.thresh;
  D->filtersT.resize(mo.filters.size());
  for (size_t i = 0; i < mo.filters.size(); ++i) {
    D->filtersT[i].resize(mo.filters[i].size());
    for (size_t j = 0; j < mo.filters[i].size(); ++j) D->filtersT[i][j] = (T)mo.filters[i][j];
  }
  return D;
}

struct Handle {
  int prec;                 // 32 or 64
  Detector<float>* f = nullptr;
  Detector<double>* d = nullptr;
};

#define DISPATCH(h, expr_f, expr_d) do { if ((h)->prec == 32) { auto& D = *(h)->f; expr_f; } else { auto& D = *(h)->d; expr_d; } } while (0)

template <typename T, typename U>
void copy_map(const Map<T>& m, U* dst) { for (size_t i = 0; i < m.d.size(); ++i) dst[i] = (U)m.d[i]; }

}  // namespace

extern "C" {

// ---- primitives (used to pin the restatement against cv2 / brute force) ----
void orc_resize_u8(const uint8_t* src, int sh, int sw, int cn, uint8_t* dst, int dh, int dw) {
  Image8 s; s.rows = sh; s.cols = sw; s.ch = cn; s.d.assign(src, src + (size_t)sh * sw * cn);
  Image8 d; resize_linear_u8(s, d, dw, dh);
  std::memcpy(dst, d.d.data(), d.d.size());
}
void orc_pyrdown_u8(const uint8_t* src, int sh, int sw, int cn, uint8_t* dst) {
  Image8 s; s.rows = sh; s.cols = sw; s.ch = cn; s.d.assign(src, src + (size_t)sh * sw * cn);
  Image8 d; pyrdown_u8(s, d);
  std::memcpy(dst, d.d.data(), d.d.size());
}
// returns nscales; fills up to cap entries of (img_w, img_h) and scale
int orc_pyramid_geometry(int h, int w, int sbin, int interval, int cap, int* wh, float* scales) {
  std::vector<LevelGeom> lv;
  int ns = pyramid_geometry(h, w, sbin, interval, lv);
  for (int i = 0; i < ns && i < cap; ++i) { wh[2 * i] = lv[i].img_w; wh[2 * i + 1] = lv[i].img_h; scales[i] = lv[i].scale; }
  return ns;
}
void orc_hog_dims(int h, int w, int sbin, int* oh, int* ow) {
  const int bw = (int)std::round((float)w / (float)sbin), bh = (int)std::round((float)h / (float)sbin);
  *ow = std::max(bw - 2, 0); *oh = std::max(bh - 2, 0);
}
void orc_hog_f32(const uint8_t* img, int h, int w, int cn, int sbin, int norient, int flen, float* out) {
  Image8 s; s.rows = h; s.cols = w; s.ch = cn; s.d.assign(img, img + (size_t)h * w * cn);
  Map<float> f; int ow, oh; hog_features<float>(s, sbin, norient, flen, f, ow, oh);
  std::memcpy(out, f.d.data(), f.d.size() * sizeof(float));
}
void orc_hog_f64(const uint8_t* img, int h, int w, int cn, int sbin, int norient, int flen, double* out) {
  Image8 s; s.rows = h; s.cols = w; s.ch = cn; s.d.assign(img, img + (size_t)h * w * cn);
  Map<double> f; int ow, oh; hog_features<double>(s, sbin, norient, flen, f, ow, oh);
  std::memcpy(out, f.d.data(), f.d.size() * sizeof(double));
}
void orc_convolve_f32(const float* feat, int oh, int ow, int flen, const float* filt, int kh, int kw, float* out) {
  Map<float> F(oh, ow * flen); std::memcpy(F.d.data(), feat, F.d.size() * sizeof(float));
  std::vector<float> W(filt, filt + (size_t)kh * kw * flen);
  Map<float> R; convolve<float>(F, oh, ow, flen, W, kh, kw, R);
  std::memcpy(out, R.d.data(), R.d.size() * sizeof(float));
}
void orc_convolve_f64(const double* feat, int oh, int ow, int flen, const double* filt, int kh, int kw, double* out) {
  Map<double> F(oh, ow * flen); std::memcpy(F.d.data(), feat, F.d.size() * sizeof(double));
  std::vector<double> W(filt, filt + (size_t)kh * kw * flen);
  Map<double> R; convolve<double>(F, oh, ow, flen, W, kh, kw, R);
  std::memcpy(out, R.d.data(), R.d.size() * sizeof(double));
}
void orc_dt1d_f32(const float* src, int N, double a, double b, int os, float* dst, int* ptr) {
  Quadratic f{a, b}; dt_row<float>(src, dst, ptr, N, f, os);
}
void orc_dt1d_f64(const double* src, int N, double a, double b, int os, double* dst, int* ptr) {
  Quadratic f{a, b}; dt_row<double>(src, dst, ptr, N, f, os);
}
// w4 = model deformation weights (w0..w3) as floats; a = -w0, b = -w1 etc. as in DynamicProgram.cpp:126-127
void orc_dt2d_f32(const float* src, int M, int N, const float* w4, int osx, int osy, int backptr_mode, float* out, int* Ix, int* Iy) {
  Map<float> in(M, N); std::memcpy(in.d.data(), src, in.d.size() * sizeof(float));
  Quadratic fx{-(double)w4[0], -(double)w4[1]}, fy{-(double)w4[2], -(double)w4[3]};
  Map<float> o; Map<int> ix, iy; dt_2d<float>(in, fx, fy, osx, osy, o, ix, iy, backptr_mode);
  std::memcpy(out, o.d.data(), o.d.size() * sizeof(float));
  std::memcpy(Ix, ix.d.data(), ix.d.size() * sizeof(int));
  std::memcpy(Iy, iy.d.data(), iy.d.size() * sizeof(int));
}
void orc_dt2d_f64(const double* src, int M, int N, const float* w4, int osx, int osy, int backptr_mode, double* out, int* Ix, int* Iy) {
  Map<double> in(M, N); std::memcpy(in.d.data(), src, in.d.size() * sizeof(double));
  Quadratic fx{-(double)w4[0], -(double)w4[1]}, fy{-(double)w4[2], -(double)w4[3]};
  Map<double> o; Map<int> ix, iy; dt_2d<double>(in, fx, fy, osx, osy, o, ix, iy, backptr_mode);
  std::memcpy(out, o.d.data(), o.d.size() * sizeof(double));
  std::memcpy(Ix, ix.d.data(), ix.d.size() * sizeof(int));
  std::memcpy(Iy, iy.d.data(), iy.d.size() * sizeof(int));
}

// Candidate::nonMaximaSuppression (reference include/Candidate.hpp:277-304) over boundingBox() (:104-110) with
// cv::Rect's | and & operators; rects = n x nparts x (x, y, width, height); keep receives the kept indices.
int orc_nms(const int* rects, int n, int nparts, int im_h, int im_w, float overlap, int* keep) {
  std::vector<uint8_t> scratch((size_t)im_h * im_w, 0);          // cv::Mat::zeros(im.size(), CV_8U)
  int nk = 0;
  for (int i = 0; i < n; ++i) {
    const int* r = rects + (size_t)i * nparts * 4;
    int hx = r[0], hy = r[1], hw = r[2], hh = r[3];              // hull = parts_[0]
    for (int p = 0; p < nparts; ++p) {                           // hull = hull | parts_[n]
      const int* q = r + 4 * p;
      if (hw <= 0 || hh <= 0) { hx = q[0]; hy = q[1]; hw = q[2]; hh = q[3]; }
      else if (q[2] > 0 && q[3] > 0) {
        const int x1 = std::min(hx, q[0]), y1 = std::min(hy, q[1]);
        hw = std::max(hx + hw, q[0] + q[2]) - x1; hh = std::max(hy + hh, q[1] + q[3]) - y1; hx = x1; hy = y1;
      }
    }
    int bx = std::max(hx, 0), by = std::max(hy, 0);              // box = hull & Rect(0,0,w,h)
    int bw = std::min(hx + hw, im_w) - bx, bh = std::min(hy + hh, im_h) - by;
    if (bw <= 0 || bh <= 0) { bx = by = bw = bh = 0; }
    double sum = 0;                                               // cv::sum(scratch(box))[0]
    for (int y = by; y < by + bh; ++y) for (int x = bx; x < bx + bw; ++x) sum += scratch[(size_t)y * im_w + x];
    if (sum / (bw * bh) > overlap) continue;                     // boxsum[0] / box.area() > overlap
    for (int y = by; y < by + bh; ++y) for (int x = bx; x < bx + bw; ++x) scratch[(size_t)y * im_w + x] = 1;
    keep[nk++] = i;
  }
  return nk;
}

// nonMaximaSuppression(src, sz, dst, mask) of reference src/nms.cpp:84-129 (Neubeck / Van Gool block-wise strict local maxima):
// the map is cut into (sz+1) x (sz+1) blocks; the block's first maximum (row-major, strict >) is a local maximum iff it is strictly
// greater than every element of the (2 sz + 1)^2 window centred on it that lies outside the block.  mask (optional, non-zero =
// eligible) restricts both searches; cv::minMaxLoc over an empty selection yields the value 0 at (-1, -1), which the reference
// then offsets and may write out of bounds -- guarded here (such a block has no eligible element and produces no maximum unless
// the reference's bogus comparison 0 > vnmax fires inside the map, which is replicated when the offset location is in range).
void orc_rootmap_nms(const float* src, int M, int N, int sz, const uint8_t* mask, uint8_t* dst) {
  for (size_t i = 0; i < (size_t)M * N; ++i) dst[i] = 0;
  auto sel_max = [&](int y0, int y1, int x0, int x1, int by0, int by1, int bx0, int bx1, bool exclude, double& vmax, int& ym, int& xm) {
    bool any = false;
    vmax = 0; ym = -1; xm = -1;
    for (int y = y0; y < y1; ++y)
      for (int x = x0; x < x1; ++x) {
        if (mask && !mask[(size_t)y * N + x]) continue;
        if (exclude && y >= by0 && y < by1 && x >= bx0 && x < bx1) continue;
        const double v = (double)src[(size_t)y * N + x];
        if (!any || v > vmax) { vmax = v; ym = y; xm = x; any = true; }
      }
    return any;
  };
  for (int m = 0; m < M; m += sz + 1)
    for (int n = 0; n < N; n += sz + 1) {
      const int i1 = std::min(m + sz + 1, M), j1 = std::min(n + sz + 1, N);
      double vc, vn;
      int yc, xc, yn, xn;
      const bool any = sel_max(m, i1, n, j1, 0, 0, 0, 0, false, vc, yc, xc);
      if (!any) { yc = m - 1; xc = n - 1; }                        // ijmax = (-1,-1) + block origin (:101)
      const int in0 = std::max(yc - sz, 0), in1 = std::min(yc + sz + 1, M), jn0 = std::max(xc - sz, 0), jn1 = std::min(xc + sz + 1, N);
      // blockmask: zero over the block's rows/columns relative to the window (:111-114)
      const int iis0 = m - in0, iis1 = std::min(m - in0 + sz + 1, in1 - in0), jis0 = n - jn0, jis1 = std::min(n - jn0 + sz + 1, jn1 - jn0);
      sel_max(in0, in1, jn0, jn1, in0 + iis0, in0 + iis1, jn0 + jis0, jn0 + jis1, true, vn, yn, xn);
      if (vc > vn && yc >= 0 && xc >= 0) dst[(size_t)yc * N + xc] = 255;
    }
}

// SearchSpacePruning<T>::filterCandidatesByDepth (reference src/SearchSpacePruning.cpp:73-95) with Math::median
// (include/Math.hpp:62-72: std::nth_element at size/2): a candidate survives iff for no part p = nparts-1 .. 1 both the child's and
// the parent's box have a positive median depth that differ by more than |anchor(p, mixture 0)| * zfactor.  rects = n x nparts x
// (x, y, w, h); parent / anchor0 = per part; boxes are clipped to the depth image here (the reference takes depth(box) unclipped
// and asserts inside OpenCV when a box crosses the border); an empty box has median 0.  Candidates of a one-part component are
// all dropped (the reference's loop never reaches its push_back).
int orc_filter_by_depth(const int* rects, int n, int nparts, const int* parent, const int* anchor0_xy, const float* depth, int im_h, int im_w,
                        float zfactor, int* keep) {
  auto median = [&](const int* r) -> float {
    int x0 = std::max(r[0], 0), y0 = std::max(r[1], 0), x1 = std::min(r[0] + r[2], im_w), y1 = std::min(r[1] + r[3], im_h);
    if (x1 <= x0 || y1 <= y0) return 0.f;
    std::vector<float> v;
    v.reserve((size_t)(x1 - x0) * (y1 - y0));
    for (int y = y0; y < y1; ++y) for (int x = x0; x < x1; ++x) v.push_back(depth[(size_t)y * im_w + x]);
    std::nth_element(v.begin(), v.begin() + v.size() / 2, v.end());
    return v[v.size() / 2];
  };
  int nk = 0;
  for (int i = 0; i < n; ++i) {
    const int* R = rects + (size_t)i * nparts * 4;
    bool kept = false;
    for (int p = nparts - 1; p >= 1; --p) {
      const float cm = median(R + 4 * p), pm = median(R + 4 * parent[p]);
      if (cm > 0 && pm > 0) {
        const double nrm = std::sqrt((double)anchor0_xy[2 * p] * anchor0_xy[2 * p] + (double)anchor0_xy[2 * p + 1] * anchor0_xy[2 * p + 1]);
        if (std::abs(cm - pm) > nrm * zfactor) break;
      }
      if (p == 1) kept = true;
    }
    keep[i] = kept ? 1 : 0;
    nk += kept;
  }
  return nk;
}

// ---- detector handle ----
// hdr = {interval, sbin, norient, flen, nfilters, nbias, ndefs, ncomp}; fdims = (kh,kw) per filter;
// indexers = for c: nparts, then for p: parentid, nf, nb, nd, filterid[nf], biasid[nb], defid[nd]
void* orc_create(const int* hdr, float thresh, const int* fdims, const double* filters, const float* biasw,
                 const int* anchors, const float* defs, const int* indexers, int precision) {
  OModel mo;
  mo.interval = hdr[0]; mo.sbin = hdr[1]; mo.norient = hdr[2]; mo.flen = hdr[3];
  const int nf = hdr[4], nb = hdr[5], nd = hdr[6], nc = hdr[7];
  mo.thresh = thresh;
  size_t off = 0;
  for (int i = 0; i < nf; ++i) {
    mo.frows.push_back(fdims[2 * i]); mo.fcols.push_back(fdims[2 * i + 1]);
    const size_t n = (size_t)fdims[2 * i] * fdims[2 * i + 1] * mo.flen;
    mo.filters.emplace_back(filters + off, filters + off + n);
    off += n;
  }
  mo.biasw.assign(biasw, biasw + nb);
  mo.anchors.assign(anchors, anchors + 2 * nd);
  mo.defs.assign(defs, defs + 4 * nd);
  const int* ip = indexers;
  mo.comps.resize(nc);
  for (int c = 0; c < nc; ++c) {
    const int np = *ip++;
    mo.comps[c].resize(np);
    for (int p = 0; p < np; ++p) {
      OPart& P = mo.comps[c][p];
      P.parentid = *ip++;
      const int a = *ip++, b = *ip++, d = *ip++;
      P.filterid.assign(ip, ip + a); ip += a;
      P.biasid.assign(ip, ip + b); ip += b;
      P.defid.assign(ip, ip + d); ip += d;
      if (P.defid.empty()) P.defid.push_back(0);        // root: <defid></defid>
    }
  }
  Handle* h = new Handle();
  h->prec = precision;
  if (precision == 32) h->f = make_detector<float>(mo); else h->d = make_detector<double>(mo);
  return h;
}
void orc_destroy(void* hv) { Handle* h = (Handle*)hv; delete h->f; delete h->d; delete h; }
void orc_set_thresh(void* hv, double t) { Handle* h = (Handle*)hv; DISPATCH(h, D.thresh = t, D.thresh = t); }
void orc_set_backptr_mode(void* hv, int m) { Handle* h = (Handle*)hv; DISPATCH(h, D.backptr_mode = m, D.backptr_mode = m); }
void orc_set_max_levels(void* hv, int m) { Handle* h = (Handle*)hv; DISPATCH(h, D.max_levels = m, D.max_levels = m); }
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// stages: 1 = pyramid+HOG, 2 = pdf, 3 = min, 4 = argmin (runs stages [from, to])
int orc_run(void* hv, const uint8_t* img, int h, int w, int cn, int from, int to) {
  Handle* H = (Handle*)hv;
  Image8 im;
  if (from <= 1) { im.rows = h; im.cols = w; im.ch = cn; im.d.assign(img, img + (size_t)h * w * cn); }
  DISPATCH(H,
    { if (from <= 1 && to >= 1) stage_pyramid<float>(D, im); if (from <= 2 && to >= 2) stage_pdf<float>(D);
      if (from <= 3 && to >= 3) stage_min<float>(D); if (from <= 4 && to >= 4) stage_argmin<float>(D); },
    { if (from <= 1 && to >= 1) stage_pyramid<double>(D, im); if (from <= 2 && to >= 2) stage_pdf<double>(D);
      if (from <= 3 && to >= 3) stage_min<double>(D); if (from <= 4 && to >= 4) stage_argmin<double>(D); });
  return 0;
}
// inject stage inputs (for stage-isolated parity tests)
void orc_set_levels(void* hv, int nlevels, const int* ohow, const float* scales) {
  Handle* H = (Handle*)hv;
  DISPATCH(H,
    { D.geom.assign(nlevels, LevelGeom{0,0,0}); D.oh.resize(nlevels); D.ow.resize(nlevels); D.feats.assign(nlevels, Map<float>());
      D.resp.assign(nlevels, std::vector<Map<float>>(D.filtersT.size()));
      for (int i = 0; i < nlevels; ++i) { D.oh[i] = ohow[2*i]; D.ow[i] = ohow[2*i+1]; D.geom[i].scale = scales[i]; } },
    { D.geom.assign(nlevels, LevelGeom{0,0,0}); D.oh.resize(nlevels); D.ow.resize(nlevels); D.feats.assign(nlevels, Map<double>());
      D.resp.assign(nlevels, std::vector<Map<double>>(D.filtersT.size()));
      for (int i = 0; i < nlevels; ++i) { D.oh[i] = ohow[2*i]; D.ow[i] = ohow[2*i+1]; D.geom[i].scale = scales[i]; } });
}
void orc_set_features(void* hv, int level, const double* src) {
  Handle* H = (Handle*)hv;
  DISPATCH(H,
    { D.feats[level] = Map<float>(D.oh[level], D.ow[level] * D.model.flen); for (size_t i = 0; i < D.feats[level].d.size(); ++i) D.feats[level].d[i] = (float)src[i]; },
    { D.feats[level] = Map<double>(D.oh[level], D.ow[level] * D.model.flen); for (size_t i = 0; i < D.feats[level].d.size(); ++i) D.feats[level].d[i] = src[i]; });
}
void orc_set_response(void* hv, int level, int filter, const double* src) {
  Handle* H = (Handle*)hv;
  DISPATCH(H,
    { D.resp[level][filter] = Map<float>(D.oh[level], D.ow[level]); for (size_t i = 0; i < D.resp[level][filter].d.size(); ++i) D.resp[level][filter].d[i] = (float)src[i]; },
    { D.resp[level][filter] = Map<double>(D.oh[level], D.ow[level]); for (size_t i = 0; i < D.resp[level][filter].d.size(); ++i) D.resp[level][filter].d[i] = src[i]; });
}
int orc_nlevels(void* hv) { Handle* H = (Handle*)hv; int n = 0; DISPATCH(H, n = (int)D.geom.size(), n = (int)D.geom.size()); return n; }
void orc_level_info(void* hv, int l, int* img_h, int* img_w, int* oh, int* ow, float* scale) {
  Handle* H = (Handle*)hv;
  DISPATCH(H,
    { *img_h = D.geom[l].img_h; *img_w = D.geom[l].img_w; *oh = D.oh[l]; *ow = D.ow[l]; *scale = D.geom[l].scale; },
    { *img_h = D.geom[l].img_h; *img_w = D.geom[l].img_w; *oh = D.oh[l]; *ow = D.ow[l]; *scale = D.geom[l].scale; });
}
void orc_get_image(void* hv, int l, uint8_t* dst) {
  Handle* H = (Handle*)hv;
  DISPATCH(H, std::memcpy(dst, D.pyr[l].d.data(), D.pyr[l].d.size()), std::memcpy(dst, D.pyr[l].d.data(), D.pyr[l].d.size()));
}
// all numeric getters return double so one binding serves both precisions (float -> double is exact)
void orc_get_features(void* hv, int l, double* dst) { Handle* H = (Handle*)hv; DISPATCH(H, copy_map(D.feats[l], dst), copy_map(D.feats[l], dst)); }
void orc_get_response(void* hv, int l, int f, double* dst) { Handle* H = (Handle*)hv; DISPATCH(H, copy_map(D.resp[l][f], dst), copy_map(D.resp[l][f], dst)); }
void orc_get_rootv(void* hv, int l, int c, double* dst) { Handle* H = (Handle*)hv; DISPATCH(H, copy_map(D.rootv[l][c], dst), copy_map(D.rootv[l][c], dst)); }
void orc_get_rooti(void* hv, int l, int c, int* dst) { Handle* H = (Handle*)hv; DISPATCH(H, copy_map(D.rooti[l][c], dst), copy_map(D.rooti[l][c], dst)); }
void orc_get_backptr(void* hv, int l, int c, int p, int m, int* ix, int* iy, int* ik) {
  Handle* H = (Handle*)hv;
  DISPATCH(H,
    { copy_map(D.Ix[l][c][p][m], ix); copy_map(D.Iy[l][c][p][m], iy); copy_map(D.Ik[l][c][p][m], ik); },
    { copy_map(D.Ix[l][c][p][m], ix); copy_map(D.Iy[l][c][p][m], iy); copy_map(D.Ik[l][c][p][m], ik); });
}
int orc_ncandidates(void* hv) { Handle* H = (Handle*)hv; int n = 0; DISPATCH(H, n = (int)D.cands.size(), n = (int)D.cands.size()); return n; }
// per candidate: level, component, score, nparts, then xs/ys/ms/rects copied into caller arrays sized by nparts
int orc_candidate_nparts(void* hv, int i) { Handle* H = (Handle*)hv; int n = 0; DISPATCH(H, n = (int)D.cands[i].x.size(), n = (int)D.cands[i].x.size()); return n; }
void orc_get_candidate(void* hv, int i, int* level, int* comp, float* score, int* xs, int* ys, int* ms, int* rects) {
  Handle* H = (Handle*)hv;
  const OCandidate* c = nullptr;
  DISPATCH(H, c = &D.cands[i], c = &D.cands[i]);
  *level = c->level; *comp = c->component; *score = c->score;
  const size_t np = c->x.size();
  std::memcpy(xs, c->x.data(), np * sizeof(int)); std::memcpy(ys, c->y.data(), np * sizeof(int));
  std::memcpy(ms, c->m.data(), np * sizeof(int)); std::memcpy(rects, c->rect.data(), np * 4 * sizeof(int));
}
void orc_get_timings(void* hv, double* t5) {
  Handle* H = (Handle*)hv;
  DISPATCH(H,
    { t5[0] = D.t_pyr; t5[1] = D.t_hog; t5[2] = D.t_pdf; t5[3] = D.t_min; t5[4] = D.t_argmin; },
    { t5[0] = D.t_pyr; t5[1] = D.t_hog; t5[2] = D.t_pdf; t5[3] = D.t_min; t5[4] = D.t_argmin; });
}

}  // extern "C"
