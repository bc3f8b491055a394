// pbd_b200_plugins.hpp -- the reference's two stage plug-ins, backed by the CUDA path:
//
//   pbd_b200::CudaHOGFeatures<T>      : IFeatures            (reference include/IFeatures.hpp:49-73;  replaces HOGFeatures<T>,
//                                                              src/HOGFeatures.cpp, in PartsBasedDetector<T>::distributeModel :109)
//   pbd_b200::CudaConvolutionEngine   : IConvolutionEngine   (reference include/IConvolutionEngine.hpp:44-68; replaces
//                                                              SpatialConvolutionEngine, src/PartsBasedDetector.cpp:111-118)
//
// so that the reference's own PartsBasedDetector (boost::scoped_ptr<IFeatures> features_, scoped_ptr<IConvolutionEngine>
// convolution_engine_, include/PartsBasedDetector.hpp:158-160) can host ONE accelerated stage and keep the rest of its pipeline:
// same method names, argument meaning and container types (cv::Mat feature maps of rows = oh, cols = ow * flen, channel fastest;
// responses[level][filter] of the feature map's size).  Inside the reference tree the reference's own interface headers are used
// (put its include/ directory on the include path); elsewhere the two interfaces are declared here with identical signatures.
//
// Needs a cv::Mat: OpenCV's <opencv2/core/core.hpp>, or the minimal stand-in under oracle/ref_shim that the tests compile against.
// The device computes in single precision.  T = double (reference ros/Node.hpp:121, cells/detect.cpp:180) is accepted at this
// boundary -- Mats are converted on the way in and out -- but the numbers are the fp32 path's: scores agree with the reference's
// double pipeline to ~1e-6 relative (north star: 1e-4), bit-identical integer outputs are only claimed for T = float.
#ifndef PBD_B200_PLUGINS_HPP_
#define PBD_B200_PLUGINS_HPP_
#include <opencv2/core/core.hpp>

#include "pbd_b200.hpp"

#if defined(__has_include) && __has_include("IFeatures.hpp") && __has_include("IConvolutionEngine.hpp")
#include "IConvolutionEngine.hpp"
#include "IFeatures.hpp"
#else
// ---- the reference's interfaces, signature for signature (include/types.hpp:52-70, IFeatures.hpp:49-73, IConvolutionEngine.hpp:44-68)
typedef std::vector<float> vectorf;
typedef std::vector<cv::Mat> vectorMat;
typedef std::vector<vectorMat> vector2DMat;
class IFeatures {
 public:
  virtual ~IFeatures() {}
  virtual size_t binsize(void) const = 0;
  virtual size_t nscales(void) const = 0;
  virtual vectorf scales(void) const = 0;
  virtual void pyramid(const cv::Mat& im, vectorMat& pyrafeatures) = 0;
};
class IConvolutionEngine {
 public:
  virtual ~IConvolutionEngine() {}
  virtual void pdf(const vectorMat& features, vector2DMat& responses) = 0;
  virtual void setFilters(const vectorMat& filters) = 0;
};
#endif

namespace pbd_b200 {
namespace detail {
// a one-part model around a set of filters: what the stage kernels need of a model (header fields + filter bank)
inline pbd_model* stage_model(int sbin, int interval, int norient, int flen, const std::vector<int32_t>& fdims, const std::vector<double>& filters) {
  const int nf = (int)(fdims.size() / 2);
  const int32_t hdr[8] = {interval, sbin, norient, flen, nf, 1, 1, 1};
  const float biasw[1] = {0.f}, defs[4] = {0.01f, 0.f, 0.01f, 0.f};
  const int32_t anchors[2] = {0, 0};
  const int32_t indexers[] = {1, /*root:*/ -1, 1, 1, 0, /*filterid*/ 0, /*biasid*/ 0};
  pbd_model* m = nullptr;
  check(pbd_model_create("stage", hdr, 0.f, fdims.data(), filters.data(), biasw, anchors, defs, indexers, &m));
  return m;
}
template <typename T> struct CvType;
template <> struct CvType<float> { enum { type = CV_32F }; };
template <> struct CvType<double> { enum { type = CV_64F }; };
}  // namespace detail

// HOGFeatures<T> (src/HOGFeatures.cpp): image pyramid + HOG cells of every level on the GPU
template <typename T>
class CudaHOGFeatures : public IFeatures {
  size_t binsize_, interval_, flen_, norient_, nscales_;
  vectorf scales_;
  pbd_detector* d_ = nullptr;

 public:
  // same arguments as HOGFeatures(binsize, nscales, flen, norient), include/HOGFeatures.hpp:77-84 (nscales = the model's interval)
  CudaHOGFeatures(size_t binsize, size_t nscales, size_t flen, size_t norient, int device = 0, void* cuda_stream = nullptr)
      : binsize_(binsize), interval_(nscales), flen_(flen), norient_(norient), nscales_(nscales) {
    const std::vector<int32_t> fdims{1, 1};
    const std::vector<double> filt(flen, 0.0);
    pbd_model* m = detail::stage_model((int)binsize, (int)nscales, (int)norient, (int)flen, fdims, filt);
    const int rc = pbd_create(m, device, cuda_stream, &d_);
    pbd_model_free(m);
    check(rc);
  }
  ~CudaHOGFeatures() override { pbd_destroy(d_); }
  CudaHOGFeatures(const CudaHOGFeatures&) = delete;
  CudaHOGFeatures& operator=(const CudaHOGFeatures&) = delete;

  size_t binsize(void) const override { return binsize_; }
  size_t nscales(void) const override { return nscales_; }          // the level count of the last pyramid() (src/HOGFeatures.cpp:99)
  vectorf scales(void) const override { return scales_; }
  void pyramid(const cv::Mat& im, vectorMat& pyrafeatures) override {
    if (im.empty()) throw Error(PBD_E_ARG, "empty image");
    if (im.depth() != CV_8U) throw Error(PBD_E_UNSUPPORTED, "Unsupported image type");   // the device path takes 8-bit frames (CV_Error at :141-145 for the rest)
    check(pbd_stage_pyramid(d_, im.data, 1, im.rows, im.cols, im.channels(), im.step, 0));
    const int n = pbd_num_levels(d_);
    nscales_ = (size_t)n;
    scales_.assign(n, 0.f);
    pyrafeatures.clear();
    pyrafeatures.resize(n);
    std::vector<float> buf;
    for (int l = 0; l < n; ++l) {
      int32_t ih, iw, oh, ow;
      check(pbd_level_info(d_, l, &ih, &iw, &oh, &ow, &scales_[l]));
      cv::Mat f(oh, ow * (int)flen_, CV_32F);                          // rows = oh, cols = ow * flen, channel fastest (:180)
      if (oh > 0 && ow > 0) check(pbd_get_features(d_, 0, l, f.ptr<float>(0)));
      if ((int)detail::CvType<T>::type == CV_32F) pyrafeatures[l] = f;
      else f.convertTo(pyrafeatures[l], detail::CvType<T>::type);
    }
  }
};

// SpatialConvolutionEngine (src/SpatialConvolutionEngine.cpp): dense part-filter responses of every level on the GPU
class CudaConvolutionEngine : public IConvolutionEngine {
  int type_;
  size_t flen_;
  int device_;
  void* stream_;
  pbd_detector* d_ = nullptr;
  int nfilters_ = 0;

 public:
  // same arguments as SpatialConvolutionEngine(type, flen), include/SpatialConvolutionEngine.hpp:52
  CudaConvolutionEngine(int type, size_t flen, int device = 0, void* cuda_stream = nullptr) : type_(type), flen_(flen), device_(device), stream_(cuda_stream) {
    if (type != CV_32F && type != CV_64F) throw Error(PBD_E_UNSUPPORTED, "convolution engine type must be CV_32F or CV_64F");
  }
  ~CudaConvolutionEngine() override { pbd_destroy(d_); }
  CudaConvolutionEngine(const CudaConvolutionEngine&) = delete;
  CudaConvolutionEngine& operator=(const CudaConvolutionEngine&) = delete;
  void setOption(const char* key, double v) { if (!d_) throw Error(PBD_E_STATE, "setFilters() has not been called"); check(pbd_set_option(d_, key, v)); }

  // filters: kh x (kw * flen) Mats, channel fastest (what Model::filters() holds); moved to the device once (:133-159)
  void setFilters(const vectorMat& filters) override {
    if (filters.empty()) throw Error(PBD_E_ARG, "no filters");
    std::vector<int32_t> fdims;
    std::vector<double> flat;
    for (const cv::Mat& f : filters) {
      if (f.empty() || f.channels() != 1 || f.cols % (int)flen_) throw Error(PBD_E_ARG, "filter must be a single-channel kh x (kw*flen) Mat");
      cv::Mat f64;
      f.convertTo(f64, CV_64F);
      fdims.push_back(f.rows); fdims.push_back(f.cols / (int)flen_);
      for (int y = 0; y < f64.rows; ++y) flat.insert(flat.end(), f64.ptr<double>(y), f64.ptr<double>(y) + f64.cols);
    }
    const int norient = ((int)flen_ - 5) * 2 / 3;
    pbd_model* m = detail::stage_model(4, 1, norient, (int)flen_, fdims, flat);
    pbd_destroy(d_);
    d_ = nullptr;
    const int rc = pbd_create(m, device_, stream_, &d_);
    pbd_model_free(m);
    check(rc);
    nfilters_ = (int)filters.size();
  }

  // responses[level][filter]: the correlation of every filter with every level's feature map, same size as the map (:106-124)
  void pdf(const vectorMat& features, vector2DMat& responses) override {
    if (!d_) throw Error(PBD_E_STATE, "setFilters() must be called before pdf()");      // "must necessarily be called before pdf()"
    const int M = (int)features.size();
    responses.clear();
    responses.resize(M, vectorMat(nfilters_));
    if (!M) return;
    std::vector<int32_t> ohow;
    std::vector<float> scales(M, 1.f);
    for (const cv::Mat& f : features) {
      if (f.depth() != CV_32F && f.depth() != CV_64F) throw Error(PBD_E_UNSUPPORTED, "feature maps must be CV_32F or CV_64F");   // `assert(feature.depth() == type_)`
      if (f.rows <= 0 || f.cols <= 0 || f.cols % (int)flen_) throw Error(PBD_E_ARG, "feature map must be oh x (ow*flen)");
      ohow.push_back(f.rows); ohow.push_back(f.cols / (int)flen_);
    }
    check(pbd_set_levels(d_, 1, M, ohow.data(), scales.data()));
    for (int l = 0; l < M; ++l) {
      cv::Mat f32;
      if (features[l].depth() == CV_32F && features[l].isContinuous()) f32 = features[l];
      else if (features[l].depth() == CV_32F) features[l].copyTo(f32);
      else features[l].convertTo(f32, CV_32F);
      check(pbd_set_features(d_, 0, l, f32.ptr<float>(0)));
    }
    check(pbd_stage_pdf(d_));
    for (int l = 0; l < M; ++l)
      for (int f = 0; f < nfilters_; ++f) {
        cv::Mat r(ohow[2 * l], ohow[2 * l + 1], CV_32F);
        check(pbd_get_response(d_, 0, l, f, r.ptr<float>(0)));
        if (type_ == CV_32F) responses[l][f] = r;
        else r.convertTo(responses[l][f], type_);
      }
  }
};

}  // namespace pbd_b200
#endif  // PBD_B200_PLUGINS_HPP_
