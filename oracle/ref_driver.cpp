// oracle/ref_driver.cpp -- C entry points over the REFERENCE'S OWN SOURCES, compiled unmodified from /root/reference against the
// minimal cv:: stand-in in oracle/ref_shim (OpenCV's C++ library is not available in this image).  TEST INFRASTRUCTURE: built by
// `make -C oracle ref` into oracle/_ref/libpbd_ref.so (git-ignored; it travels to the GPU box like every built .so) and used by
// tests/test_oracle_ref.py to pin the restated oracle (oracle/pbd_oracle.cpp) bit for bit to the reference's code:
//   include/DistanceTransform.hpp   DistanceTransform<T>::compute          (1-D envelopes, transposes, pointer composition)
//   include/Math.hpp                Math::reduceMax / reducePickIndex / find
//   src/HOGFeatures.cpp             HOGFeatures<T>::pyramid -> features<uint8_t>   (cv::resize / cv::pyrDown: the oracle's
//                                   restatements of the OpenCV functions, themselves pinned bit-exact to cv2)
//   src/DynamicProgram.cpp          DynamicProgram<T>::min / argmin over include/Parts.hpp (T4: the reference indexes
//                                   bias(mm)[m] out of bounds when the child has fewer mixtures than the parent; callers only pass
//                                   models for which the reference is defined)
//   include/Candidate.hpp           Candidate::sort / nonMaximaSuppression
//   src/nms.cpp                     nonMaximaSuppression(src, sz, dst, mask): block-wise strict local maxima of a score map
//   src/SearchSpacePruning.cpp      SearchSpacePruning<T>::filterCandidatesByDepth (+ Math::median)
// Not compiled: src/SpatialConvolutionEngine.cpp + src/filter.cpp (a 4 kLoC copy of OpenCV's FilterEngine that needs OpenCV's
// internal headers; the response restatement stays pinned to cv2.filter2D at 1e-5 / 1e-11, tests/test_oracle_pins.py).
#include <cstdint>
#include <cstring>
#include <vector>

#include <opencv2/core/core.hpp>
#include <opencv2/imgproc/imgproc.hpp>

#include "DistanceTransform.hpp"
#include "DynamicProgram.hpp"
#include "HOGFeatures.hpp"
#include "Math.hpp"
#include "SearchSpacePruning.hpp"
#include "nms.hpp"

extern "C" {
// the oracle's restatements of cv::resize (INTER_LINEAR, 8U) and cv::pyrDown (oracle/pbd_oracle.cpp)
void orc_resize_u8(const uint8_t* src, int sh, int sw, int cn, uint8_t* dst, int dh, int dw);
void orc_pyrdown_u8(const uint8_t* src, int sh, int sw, int cn, uint8_t* dst);
}

namespace cv {
void resize(const Mat& src, Mat& dst, Size dsize, double, double, int) {
  CV_Assert(src.depth() == CV_8U && src.isContinuous());
  Mat out(dsize, src.type());
  orc_resize_u8(src.data, src.rows, src.cols, src.channels(), out.data, dsize.height, dsize.width);
  dst = out;
}
void pyrDown(const Mat& src, Mat& dst, const Size&, int) {
  CV_Assert(src.depth() == CV_8U && src.isContinuous());
  Mat out((src.rows + 1) / 2, (src.cols + 1) / 2, src.type());
  orc_pyrdown_u8(src.data, src.rows, src.cols, src.channels(), out.data);
  dst = out;
}
Mat getGaussianKernel(int, double, int) { throw Exception(-213, "shim: getGaussianKernel is not implemented"); }
void filter2D(const Mat&, Mat&, int, const Mat&, Point, double, int) { throw Exception(-213, "shim: filter2D is not implemented"); }
void meanStdDev(const Mat&, Scalar&, Scalar&) { throw Exception(-213, "shim: meanStdDev is not implemented"); }
std::ostream& operator<<(std::ostream& os, const Mat& m) { return os << "[Mat " << m.rows << "x" << m.cols << "]"; }
}  // namespace cv

namespace {
template <typename T>
void dt2d(const T* src, int M, int N, const float* w4, int osx, int osy, T* out, int* Ix, int* Iy) {
  cv::Mat_<T> in(M, N), o;
  std::memcpy(in.data, src, sizeof(T) * (size_t)M * N);
  cv::Mat_<int> ix, iy;
  Quadratic fx(-w4[0], -w4[1]), fy(-w4[2], -w4[3]);             // src/DynamicProgram.cpp:125-127
  DistanceTransform<T> dt;
  dt.compute(in, fx, fy, cv::Point(osx, osy), o, ix, iy);
  for (int m = 0; m < M; ++m) {
    std::memcpy(out + (size_t)m * N, o[m], sizeof(T) * N);
    std::memcpy(Ix + (size_t)m * N, ix[m], sizeof(int) * N);
    std::memcpy(Iy + (size_t)m * N, iy[m], sizeof(int) * N);
  }
}

// a Model whose arrays are filled from flat buffers (the reference's FileStorageModel needs cv::FileStorage)
struct FlatModel : Model {
  bool serialize(const std::string&) const override { return false; }
  bool deserialize(const std::string&) override { return false; }
  using Model::filtersw_; using Model::filtersi_; using Model::defw_; using Model::defi_; using Model::biasw_; using Model::biasi_;
  using Model::anchors_; using Model::biasid_; using Model::filterid_; using Model::defid_; using Model::parentid_;
  using Model::thresh_; using Model::binsize_; using Model::flen_; using Model::norient_; using Model::nscales_;
};

// DynamicProgram<T>::min + argmin on caller-supplied response maps (reference src/DynamicProgram.cpp:67-255)
struct RefDP {
  int precision = 32;
  FlatModel model;
  Parts parts;
  std::vector<int> level_h, level_w;
  vectorf scales;
  vector2DMat scores;                       // [level][filter]
  vector2DMat rootv, rooti;
  vector4DMat Ix, Iy, Ik;
  vectorCandidate cands;
};
}  // namespace

extern "C" {
// hdr / fdims / filters / indexers: the flat model layout of partsbaseddetector_b200/flatmodel.py (same as orc_create)
void* ref_dp_create(const int* hdr, const int* fdims, const double* filters, const float* biasw, const int* anchors, const float* defs,
                    const int* indexers, int precision) {
  RefDP* D = new RefDP();
  D->precision = precision;
  FlatModel& m = D->model;
  const int flen = hdr[3], nf = hdr[4], nb = hdr[5], nd = hdr[6], nc = hdr[7];
  size_t off = 0;
  for (int i = 0; i < nf; ++i) {
    const int kh = fdims[2 * i], kw = fdims[2 * i + 1];
    cv::Mat f(kh, kw * flen, CV_64F);
    std::memcpy(f.data, filters + off, sizeof(double) * (size_t)kh * kw * flen);
    off += (size_t)kh * kw * flen;
    cv::Mat ft;
    f.convertTo(ft, precision == 64 ? CV_64F : CV_32F);          // src/PartsBasedDetector.cpp:115-117
    m.filtersw_.push_back(ft);
    m.filtersi_.push_back(i);
  }
  m.biasw_.assign(biasw, biasw + nb);
  for (int i = 0; i < nb; ++i) m.biasi_.push_back(i);
  for (int i = 0; i < nd; ++i) {
    m.anchors_.push_back(cv::Point(anchors[2 * i], anchors[2 * i + 1]));
    m.defw_.push_back(vectorf(defs + 4 * i, defs + 4 * i + 4));
    m.defi_.push_back(i);
  }
  const int* ip = indexers;
  m.parentid_.resize(nc); m.filterid_.resize(nc); m.biasid_.resize(nc); m.defid_.resize(nc);
  for (int c = 0; c < nc; ++c) {
    const int np = *ip++;
    for (int p = 0; p < np; ++p) {
      m.parentid_[c].push_back(*ip++);
      const int a = *ip++, b = *ip++, d = *ip++;
      m.filterid_[c].push_back(vectori(ip, ip + a)); ip += a;
      m.biasid_[c].push_back(vectori(ip, ip + b)); ip += b;
      vectori did(ip, ip + d); ip += d;
      if (did.empty()) did.push_back(0);                          // root: <defid></defid>, src/FileStorageModel.cpp:148-152
      m.defid_[c].push_back(did);
    }
  }
  // PartsBasedDetector<T>::distributeModel, src/PartsBasedDetector.cpp:119-121
  D->parts = Parts(m.filters(), m.filtersi(), m.def(), m.defi(), m.bias(), m.biasi(), m.anchors(), m.biasid(), m.filterid(), m.defid(), m.parentid());
  return D;
}
void ref_dp_destroy(void* h) { delete (RefDP*)h; }
void ref_dp_set_levels(void* h, int nlevels, const int* ohow, const float* scales) {
  RefDP* D = (RefDP*)h;
  D->level_h.clear(); D->level_w.clear();
  for (int l = 0; l < nlevels; ++l) { D->level_h.push_back(ohow[2 * l]); D->level_w.push_back(ohow[2 * l + 1]); }
  D->scales.assign(scales, scales + nlevels);
  D->scores.assign(nlevels, vectorMat(D->model.filters().size()));
}
void ref_dp_set_response(void* h, int level, int filter, const double* src) {
  RefDP* D = (RefDP*)h;
  cv::Mat m64(D->level_h[level], D->level_w[level], CV_64F);
  std::memcpy(m64.data, src, sizeof(double) * m64.total());
  cv::Mat m;
  m64.convertTo(m, D->precision == 64 ? CV_64F : CV_32F);
  D->scores[level][filter] = m;
}
// runs min() and argmin(); returns the number of candidates
int ref_dp_run(void* h, double thresh) {
  RefDP* D = (RefDP*)h;
  D->Ix.clear(); D->Iy.clear(); D->Ik.clear(); D->rootv.clear(); D->rooti.clear(); D->cands.clear();
  vector2DMat scores = D->scores;                                // min() takes the responses by non-const reference
  if (D->precision == 64) {
    DynamicProgram<double> dp(thresh);
    dp.min(D->parts, scores, D->Ix, D->Iy, D->Ik, D->rootv, D->rooti);
    dp.argmin(D->parts, D->rootv, D->rooti, D->scales, D->Ix, D->Iy, D->Ik, D->cands);
  } else {
    DynamicProgram<float> dp(thresh);
    dp.min(D->parts, scores, D->Ix, D->Iy, D->Ik, D->rootv, D->rooti);
    dp.argmin(D->parts, D->rootv, D->rooti, D->scales, D->Ix, D->Iy, D->Ik, D->cands);
  }
  return (int)D->cands.size();
}
void ref_dp_get_root(void* h, int level, int comp, double* rootv, int* rooti) {
  RefDP* D = (RefDP*)h;
  const cv::Mat& v = D->rootv[level][comp];
  const cv::Mat& i = D->rooti[level][comp];
  for (int y = 0; y < v.rows; ++y)
    for (int x = 0; x < v.cols; ++x) {
      rootv[(size_t)y * v.cols + x] = D->precision == 64 ? v.at<double>(y, x) : (double)v.at<float>(y, x);
      rooti[(size_t)y * v.cols + x] = i.at<int>(y, x);
    }
}
void ref_dp_get_backptr(void* h, int level, int comp, int part, int pm, int* ix, int* iy, int* ik) {
  RefDP* D = (RefDP*)h;
  const cv::Mat &X = D->Ix[level][comp][part][pm], &Y = D->Iy[level][comp][part][pm], &K = D->Ik[level][comp][part][pm];
  for (int y = 0; y < X.rows; ++y)
    for (int x = 0; x < X.cols; ++x) {
      const size_t o = (size_t)y * X.cols + x;
      ix[o] = X.at<int>(y, x); iy[o] = Y.at<int>(y, x); ik[o] = K.at<int>(y, x);
    }
}
// nonMaximaSuppression of src/nms.cpp:84-129 on one float map; mask may be null (the unmasked call)
void ref_rootmap_nms(const float* src, int h, int w, int sz, const uint8_t* mask, uint8_t* dst) {
  cv::Mat m(h, w, CV_32F), k, out;
  std::memcpy(m.data, src, sizeof(float) * (size_t)h * w);
  if (mask) { k = cv::Mat(h, w, CV_8U); std::memcpy(k.data, mask, (size_t)h * w); }
  nonMaximaSuppression(m, sz, out, k);
  for (int y = 0; y < h; ++y) std::memcpy(dst + (size_t)y * w, out.ptr(y), w);
}
// SearchSpacePruning<float>::filterCandidatesByDepth on the candidates of the last run (boxes must lie inside the depth image: the
// reference takes depth(box) without clipping); keep[i] = 1 for the survivors, returns their number.  Math::median prints every box
// to std::cerr (include/Math.hpp:70): the caller redirects fd 2.
int ref_dp_filter_by_depth(void* h, const float* depth, int im_h, int im_w, float zfactor, int* keep) {
  RefDP* D = (RefDP*)h;
  cv::Mat d(im_h, im_w, CV_32F);
  std::memcpy(d.data, depth, sizeof(float) * (size_t)im_h * im_w);
  vectorCandidate c = D->cands;
  for (size_t i = 0; i < c.size(); ++i) c[i].setScore((float)i);          // tag: the survivors are identified by their score slot
  SearchSpacePruning<float> ssp;
  ssp.filterCandidatesByDepth(D->parts, c, d, zfactor);
  for (size_t i = 0; i < D->cands.size(); ++i) keep[i] = 0;
  for (size_t i = 0; i < c.size(); ++i) keep[(int)c[i].score()] = 1;
  return (int)c.size();
}
int ref_dp_candidate_nparts(void* h, int i) { return (int)((RefDP*)h)->cands[i].parts().size(); }
// rects = nparts x (x, y, width, height); conf = nparts confidences; the reference's Candidate keeps neither level nor part indices
void ref_dp_get_candidate(void* h, int i, int* comp, int* rects, float* conf) {
  Candidate c = ((RefDP*)h)->cands[i];
  *comp = c.component();
  for (size_t p = 0; p < c.parts().size(); ++p) {
    const cv::Rect& r = c.parts()[p];
    rects[4 * p] = r.x; rects[4 * p + 1] = r.y; rects[4 * p + 2] = r.width; rects[4 * p + 3] = r.height;
    conf[p] = c.confidence()[p];
  }
}
// Candidate::sort + Candidate::nonMaximaSuppression (include/Candidate.hpp:97-99, 277-304) on the candidates of the last run;
// returns the survivors' indices into the (stable-sorted by descending score) list through keep_rects = their rectangles
int ref_dp_sort_nms(void* h, int im_h, int im_w, float overlap, int max_keep, int* keep_rects /* [max_keep][nparts][4] */, float* keep_scores) {
  RefDP* D = (RefDP*)h;
  vectorCandidate c = D->cands;
  std::stable_sort(c.begin(), c.end(), Candidate::descending);   // Candidate::sort uses std::sort: the test feeds distinct scores
  cv::Mat im(im_h, im_w, CV_8U);
  Candidate::nonMaximaSuppression(im, c, overlap);
  int n = 0;
  for (size_t i = 0; i < c.size() && n < max_keep; ++i, ++n) {
    const size_t np = c[i].parts().size();
    for (size_t p = 0; p < np; ++p) {
      const cv::Rect& r = c[i].parts()[p];
      int* o = keep_rects + ((size_t)n * np + p) * 4;
      o[0] = r.x; o[1] = r.y; o[2] = r.width; o[3] = r.height;
    }
    keep_scores[n] = c[i].score();
  }
  return (int)c.size();
}
}  // extern "C"

namespace {
}  // namespace

extern "C" {
void ref_dt2d_f32(const float* src, int M, int N, const float* w4, int osx, int osy, float* out, int* Ix, int* Iy) { dt2d<float>(src, M, N, w4, osx, osy, out, Ix, Iy); }
void ref_dt2d_f64(const double* src, int M, int N, const float* w4, int osx, int osy, double* out, int* Ix, int* Iy) { dt2d<double>(src, M, N, w4, osx, osy, out, Ix, Iy); }

// Math::reduceMax over K float maps of h x w, then Math::reducePickIndex of K int maps by the arg-max
void ref_reduce_max_pick_f32(const float* in, const int* pick_in, int K, int h, int w, float* maxv, int* maxi, int* picked) {
  vectorMat v(K), pv(K);
  for (int k = 0; k < K; ++k) {
    v[k] = cv::Mat(h, w, CV_32F); std::memcpy(v[k].data, in + (size_t)k * h * w, sizeof(float) * h * w);
    pv[k] = cv::Mat(h, w, CV_32S); std::memcpy(pv[k].data, pick_in + (size_t)k * h * w, sizeof(int) * h * w);
  }
  cv::Mat mv, mi, pk;
  Math::reduceMax<float>(v, mv, mi);
  Math::reducePickIndex<int>(pv, mi, pk);
  std::memcpy(maxv, mv.data, sizeof(float) * h * w);
  std::memcpy(maxi, mi.data, sizeof(int) * h * w);
  std::memcpy(picked, pk.data, sizeof(int) * h * w);
}

// HOGFeatures<T>::pyramid on an 8-bit image: returns the number of levels; level l's features (oh x ow*flen, T) are copied by
// ref_hog_level.  One static result slot per precision (single-threaded test use).
static std::vector<cv::Mat> g_feat[2];
static vectorf g_scales[2];
int ref_hog_pyramid(const uint8_t* img, int h, int w, int cn, int sbin, int interval, int flen, int norient, int precision) {
  cv::Mat im(h, w, CV_MAKETYPE(CV_8U, cn));
  std::memcpy(im.data, img, (size_t)h * w * cn);
  const int s = precision == 64;
  g_feat[s].clear();
  if (s) { HOGFeatures<double> f(sbin, interval, flen, norient); f.pyramid(im, g_feat[s]); g_scales[s] = f.scales(); }
  else { HOGFeatures<float> f(sbin, interval, flen, norient); f.pyramid(im, g_feat[s]); g_scales[s] = f.scales(); }
  return (int)g_feat[s].size();
}
void ref_hog_level_dims(int level, int precision, int* rows, int* cols, float* scale) {
  const int s = precision == 64;
  *rows = g_feat[s][level].rows; *cols = g_feat[s][level].cols; *scale = g_scales[s][level];
}
void ref_hog_level(int level, int precision, void* dst) {
  const int s = precision == 64;
  const cv::Mat& m = g_feat[s][level];
  for (int y = 0; y < m.rows; ++y) std::memcpy((char*)dst + (size_t)y * m.cols * m.elemSize(), m.ptr(y), (size_t)m.cols * m.elemSize());
}
}  // extern "C"
