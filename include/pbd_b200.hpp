// pbd_b200.hpp -- header-only C++ mirror of the reference's interface for the detect() path, over the C-ABI
// (pbd_b200.h).  Class and method names follow wg-perception/PartsBasedDetector so that caller code such as
// src/demo.cpp:55-118 keeps its shape:
//
//   FileStorageModel model;  model.deserialize("Person_26parts.xml");          // src/demo.cpp:64-82
//   PartsBasedDetector<float> pbd;  pbd.distributeModel(model);                  // :85-86
//   vectorCandidate candidates;  pbd.detect(im, depth, candidates);              // :103
//   Candidate::sort(candidates);                                                 // :111
//
// Images are passed as `pbd_b200::Mat` (rows, cols, channels, step, 8-bit data pointer); when OpenCV headers are
// available a cv::Mat converts implicitly.  All numerics run in libpbd_b200.so on the GPU in single precision (the reference's
// PartsBasedDetector<float>, src/demo.cpp:85).  PartsBasedDetector<double> (ros/Node.hpp:121, cells/detect.cpp:180) is accepted
// and runs the same fp32 path: scores agree with the reference's double pipeline to ~1e-6 relative (north star: 1e-4), but
// bit-identical integer outputs are only claimed against PartsBasedDetector<float>.
#ifndef PBD_B200_HPP_
#define PBD_B200_HPP_
#include <algorithm>
#include <cstdint>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

#include "pbd_b200.h"

#if defined(__has_include)
#if __has_include(<opencv2/core/core.hpp>)
#include <opencv2/core/core.hpp>
#define PBD_B200_HAVE_OPENCV 1
#endif
#endif

namespace pbd_b200 {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
inline void check(int rc) {
  if (rc != PBD_OK) throw Error(rc, pbd_last_error());     // the reference throws cv::Exception / asserts
}

// minimal 8-bit image view (cv::Mat CV_8UC1 / CV_8UC3, BGR interleaved)
struct Mat {
  int rows = 0, cols = 0, channels = 3;
  size_t step = 0;                 // bytes per row
  const uint8_t* data = nullptr;
  Mat() {}
  Mat(int r, int c, int ch, const uint8_t* d, size_t s = 0) : rows(r), cols(c), channels(ch), step(s ? s : (size_t)c * ch), data(d) {}
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
#ifdef PBD_B200_HAVE_OPENCV
  Mat(const cv::Mat& m) : rows(m.rows), cols(m.cols), channels(m.channels()), step(m.step), data(m.data) {   // NOLINT: implicit on purpose
    if (!m.empty() && m.depth() != CV_8U) throw Error(PBD_E_UNSUPPORTED, "Unsupported image type");         // src/HOGFeatures.cpp:141-145
  }
#endif
};

struct Rect { int x = 0, y = 0, width = 0, height = 0; };

typedef std::vector<float> vectorf;

// reference include/Model.hpp:49-122 (getters) + include/FileStorageModel.hpp
class Model {
 protected:
  pbd_model* h_ = nullptr;
  void need() const { if (!h_) throw Error(PBD_E_STATE, "model is empty"); }
  void header(int32_t hdr[8], float* th) const { need(); check(pbd_model_header(h_, hdr, th)); }

 public:
  Model() {}
  virtual ~Model() { pbd_model_free(h_); }
  Model(const Model&) = delete;
  Model& operator=(const Model&) = delete;
  const pbd_model* handle() const { need(); return h_; }
  std::string name() const { need(); return pbd_model_name(h_); }
  int nscales() const { int32_t h[8]; header(h, nullptr); return h[0]; }   // the reference keeps `interval` here (Model.hpp:112)
  int binsize() const { int32_t h[8]; header(h, nullptr); return h[1]; }
  int norient() const { int32_t h[8]; header(h, nullptr); return h[2]; }
  int flen() const { int32_t h[8]; header(h, nullptr); return h[3]; }
  int ncomponents() const { int32_t h[8]; header(h, nullptr); return h[7]; }
  float thresh() const { int32_t h[8]; float t; header(h, &t); return t; }
  virtual bool serialize(const std::string& filename) const = 0;
  virtual bool deserialize(const std::string& filename) = 0;
};

class FileStorageModel : public Model {     // reference src/FileStorageModel.cpp:42-159
 public:
  bool serialize(const std::string& filename) const override { need(); check(pbd_model_save_storage(h_, filename.c_str())); return true; }
  bool deserialize(const std::string& filename) override {
    pbd_model* m = nullptr;
    const int rc = pbd_model_load_storage(filename.c_str(), &m);
    if (rc == PBD_E_IO) return false;           // cannot open: `if (!ok) return false;` (:100-101)
    check(rc);
    pbd_model_free(h_);
    h_ = m;
    return true;
  }
};

class MatlabIOModel : public Model {        // reference src/MatlabIOModel.cpp:71-195 (native MAT-v5 reader, no cvmatio)
 public:
  bool serialize(const std::string&) const override { return false; }   // "TODO: implement" in the reference as well (:191-195)
  bool deserialize(const std::string& filename) override {
    pbd_model* m = nullptr;
    const int rc = pbd_model_load_mat(filename.c_str(), &m);
    if (rc == PBD_E_IO) return false;           // cannot open: `if (!ok) return false;` (:76-77)
    check(rc);
    pbd_model_free(h_);
    h_ = m;
    return true;
  }
};

// reference include/Candidate.hpp:56-99 (+ frame, level and per-part cell locations / mixtures)
class Candidate {
  std::vector<Rect> parts_;
  vectorf confidence_;
  int component_ = 0;

 public:
  int frame = 0, level = 0;
  std::vector<int> x, y, mixture;
  const std::vector<Rect>& parts() const { return parts_; }
  const vectorf& confidence() const { return confidence_; }
  void addPart(Rect r, float confidence) { parts_.push_back(r); confidence_.push_back(confidence); }
  float score() const { return confidence_.size() > 0 ? confidence_[0] : -std::numeric_limits<float>::infinity(); }
  void setComponent(int c) { component_ = c; }
  int component() const { return component_; }
  static bool descending(const Candidate& c1, const Candidate& c2) { return c1.score() > c2.score(); }
  static void sort(std::vector<Candidate>& candidates) { std::stable_sort(candidates.begin(), candidates.end(), descending); }
  // reference include/Candidate.hpp:277-304: greedy box painting in the given order (sort first)
  static void nonMaximaSuppression(const Mat& im, std::vector<Candidate>& candidates, const float overlap = 0.0f) {
    const int W = im.cols, H = im.rows;
    std::vector<uint8_t> scratch((size_t)W * H, 0);
    size_t keep = 0;
    for (size_t n = 0; n < candidates.size(); ++n) {
      const Rect hull = candidates[n].boundingBox();
      int bx = std::max(hull.x, 0), by = std::max(hull.y, 0);
      int bw = std::min(hull.x + hull.width, W) - bx, bh = std::min(hull.y + hull.height, H) - by;
      if (bw <= 0 || bh <= 0) { bx = by = bw = bh = 0; }
      double sum = 0;
      for (int y = by; y < by + bh; ++y) for (int x = bx; x < bx + bw; ++x) sum += scratch[(size_t)y * W + x];
      if (sum / (bw * bh) > overlap) continue;
      for (int y = by; y < by + bh; ++y) std::fill(scratch.begin() + (size_t)y * W + bx, scratch.begin() + (size_t)y * W + bx + bw, (uint8_t)1);
      if (keep != n) candidates[keep] = candidates[n];
      keep++;
    }
    candidates.resize(keep);
  }
  Rect boundingBox() const {
    Rect hull = parts_.at(0);
    for (const Rect& r : parts_) {
      const int x0 = std::min(hull.x, r.x), y0 = std::min(hull.y, r.y);
      const int x1 = std::max(hull.x + hull.width, r.x + r.width), y1 = std::max(hull.y + hull.height, r.y + r.height);
      hull.x = x0; hull.y = y0; hull.width = x1 - x0; hull.height = y1 - y0;
    }
    return hull;
  }
};
typedef std::vector<Candidate> vectorCandidate;

// reference include/PartsBasedDetector.hpp:152-175
template <typename T>
class PartsBasedDetector {
  static_assert(sizeof(T) == sizeof(float) || sizeof(T) == sizeof(double), "T is float or double (the device computes in single precision either way)");
  pbd_detector* d_ = nullptr;
  std::string name_;

  static void append(pbd_candidates* c, vectorCandidate& out) {
    const int n = pbd_candidates_count(c);
    std::vector<int32_t> xs, ys, ms, rc;
    for (int i = 0; i < n; ++i) {
      const int np = pbd_candidates_nparts(c, i);
      xs.resize(np); ys.resize(np); ms.resize(np); rc.resize(4 * (size_t)np);
      int32_t frame, level, comp;
      float score;
      check(pbd_candidates_get(c, i, &frame, &level, &comp, &score, xs.data(), ys.data(), ms.data(), rc.data()));
      Candidate cand;
      cand.setComponent(comp);
      cand.frame = frame; cand.level = level;
      cand.x.assign(xs.begin(), xs.end()); cand.y.assign(ys.begin(), ys.end()); cand.mixture.assign(ms.begin(), ms.end());
      for (int p = 0; p < np; ++p) {
        Rect r; r.x = rc[4 * p]; r.y = rc[4 * p + 1]; r.width = rc[4 * p + 2]; r.height = rc[4 * p + 3];
        cand.addPart(r, p == 0 ? score : 0.0f);            // src/DynamicProgram.cpp:241-244
      }
      out.push_back(cand);                                  // appended, never cleared (:250)
    }
    pbd_candidates_free(c);
  }

 public:
  explicit PartsBasedDetector(int device = 0, void* cuda_stream = nullptr) : device_(device), stream_(cuda_stream) {}
  ~PartsBasedDetector() { pbd_destroy(d_); }
  PartsBasedDetector(const PartsBasedDetector&) = delete;
  PartsBasedDetector& operator=(const PartsBasedDetector&) = delete;

  void distributeModel(Model& model) {                      // src/PartsBasedDetector.cpp:102-127
    pbd_destroy(d_);
    d_ = nullptr;
    check(pbd_create(model.handle(), device_, stream_, &d_));
    name_ = model.name();
  }
  const std::string& name() const { return name_; }
  void setOption(const char* key, double v) { need(); check(pbd_set_option(d_, key, v)); }

  void detect(const Mat& im, vectorCandidate& candidates) { detect(im, Mat(), candidates); }   // :54-56
  void detect(const Mat& im, const Mat& depth, vectorCandidate& candidates) {                  // :69-95 (depth is ignored there too)
    (void)depth;
    need();
    if (im.empty()) throw Error(PBD_E_ARG, "empty image");
    pbd_candidates* c = nullptr;
    check(pbd_detect_batch_u8(d_, im.data, 1, im.rows, im.cols, im.channels, im.step, 0, &c));
    append(c, candidates);
  }
  // batch of n equally sized frames, frame_stride bytes apart (0 = tightly packed)
  void detectBatch(const Mat& first, int n, size_t frame_stride, vectorCandidate& candidates) {
    need();
    pbd_candidates* c = nullptr;
    check(pbd_detect_batch_u8(d_, first.data, n, first.rows, first.cols, first.channels, first.step, frame_stride, &c));
    append(c, candidates);
  }
  // pipelined variant (pbd_submit_batch_u8 / pbd_collect_ticket): submit batch i+1 before collecting batch i; `first` must be
  // tightly packed, should be pinned and must stay untouched until its ticket has been collected
  int submitBatch(const Mat& first, int n) {
    need();
    int ticket = -1;
    check(pbd_submit_batch_u8(d_, first.data, n, first.rows, first.cols, first.channels, &ticket));
    return ticket;
  }
  void collectTicket(int ticket, vectorCandidate& candidates) {
    need();
    pbd_candidates* c = nullptr;
    check(pbd_collect_ticket(d_, ticket, &c));
    append(c, candidates);
  }
  pbd_detector* handle() { need(); return d_; }

 private:
  void need() const { if (!d_) throw Error(PBD_E_STATE, "distributeModel() has not been called"); }
  int device_;
  void* stream_;
};

// depth image view (cv::Mat CV_32FC1), metres; 0 = no reading
struct DepthMat {
  int rows = 0, cols = 0;
  size_t step = 0;                 // bytes per row
  const float* data = nullptr;
  DepthMat() {}
  DepthMat(int r, int c, const float* d, size_t s = 0) : rows(r), cols(c), step(s ? s : (size_t)c * sizeof(float)), data(d) {}
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
#ifdef PBD_B200_HAVE_OPENCV
  DepthMat(const cv::Mat& m) : rows(m.rows), cols(m.cols), step(m.step), data(reinterpret_cast<const float*>(m.data)) {   // NOLINT
    if (!m.empty() && m.type() != CV_32F) throw Error(PBD_E_UNSUPPORTED, "depth image must be CV_32FC1");
  }
#endif
};

// reference include/SearchSpacePruning.hpp / src/SearchSpacePruning.cpp:73-95 (the call detect() keeps commented out, zfactor 0.03)
template <typename T>
class SearchSpacePruning {
 public:
  // candidates of ONE frame; the model replaces the reference's `Parts&` argument (it holds the same tree and anchors)
  void filterCandidatesByDepth(const Model& model, vectorCandidate& candidates, const DepthMat& depth, const float zfactor) {
    if (depth.empty()) throw Error(PBD_E_ARG, "empty depth image");
    const int n = (int)candidates.size();
    size_t mp = 1;
    for (const Candidate& c : candidates) mp = std::max(mp, c.parts().size());
    std::vector<int32_t> meta((size_t)n * 4), parts((size_t)n * mp * 7, 0);
    std::vector<float> scores(n);
    for (int i = 0; i < n; ++i) {
      const Candidate& c = candidates[i];
      meta[4 * i] = c.frame; meta[4 * i + 1] = c.level; meta[4 * i + 2] = c.component(); meta[4 * i + 3] = (int32_t)c.parts().size();
      scores[i] = c.score();
      for (size_t p = 0; p < c.parts().size(); ++p) {
        int32_t* o = &parts[((size_t)i * mp + p) * 7];
        o[0] = i;                                                    // tag: which input candidate this is
        o[3] = c.parts()[p].x; o[4] = c.parts()[p].y; o[5] = c.parts()[p].width; o[6] = c.parts()[p].height;
      }
    }
    pbd_candidates* h = nullptr;
    check(pbd_candidates_create(n, (int)mp, meta.data(), scores.data(), parts.data(), &h));
    const int rc = pbd_candidates_filter_by_depth(h, model.handle(), depth.data, depth.rows, depth.cols, depth.step, zfactor);
    if (rc != PBD_OK) { pbd_candidates_free(h); check(rc); }
    const int k = pbd_candidates_count(h);
    std::vector<int32_t> m2((size_t)k * 4), p2((size_t)k * mp * 7);
    std::vector<float> s2(k);
    if (k) check(pbd_candidates_export(h, m2.data(), s2.data(), p2.data(), (int)mp));
    pbd_candidates_free(h);
    vectorCandidate out;
    for (int j = 0; j < k; ++j) out.push_back(candidates[p2[(size_t)j * mp * 7]]);
    candidates.swap(out);
  }
};

// cv::imread for the containers the library decodes itself (PNG, binary PNM): packed BGR pixels (src/demo.cpp:90)
inline std::vector<uint8_t> imread(const std::string& path, int& rows, int& cols) {
  int32_t h = 0, w = 0;
  check(pbd_imread_bgr8(path.c_str(), nullptr, 0, &h, &w));
  std::vector<uint8_t> px((size_t)h * w * 3);
  check(pbd_imread_bgr8(path.c_str(), px.data(), px.size(), &h, &w));
  rows = h; cols = w;
  return px;
}

// the fields of a sensor_msgs/Image message (ros/Node.cpp:144-183 receives them as ImageConstPtr): converts like
// cv_bridge::toCvCopy(msg, enc::BGR8) / toCvCopy(msg, enc::TYPE_32FC1) without ROS or OpenCV
struct RosImage {
  uint32_t height = 0, width = 0;
  std::string encoding;
  uint8_t is_bigendian = 0;
  uint32_t step = 0;
  const uint8_t* data = nullptr;
  std::vector<uint8_t> toBgr8() const {
    std::vector<uint8_t> out((size_t)height * width * 3);
    check(pbd_ros_image_to_bgr8(encoding.c_str(), (int)height, (int)width, step, is_bigendian, data, out.data()));
    return out;
  }
  std::vector<float> toDepth32F() const {
    std::vector<float> out((size_t)height * width);
    check(pbd_ros_depth_to_f32(encoding.c_str(), (int)height, (int)width, step, is_bigendian, data, out.data()));
    return out;
  }
};

}  // namespace pbd_b200
#endif  // PBD_B200_HPP_
