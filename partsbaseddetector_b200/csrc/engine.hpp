// engine.hpp -- host orchestration of the detector's stage pipeline on one GPU / one stream.
// Plays the role of PartsBasedDetector<float> after distributeModel() (reference
// src/PartsBasedDetector.cpp:102-127): owns the feature engine, the convolution engine, the part
// tables and the dynamic program -- here as device buffers, tables and kernel launches.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "kernels.cuh"
#include "model.hpp"

namespace pbd {

struct CudaError : std::runtime_error { using std::runtime_error::runtime_error; };
struct StateError : std::runtime_error { using std::runtime_error::runtime_error; };
struct ArgError : std::runtime_error { using std::runtime_error::runtime_error; };
struct UnsupportedError : std::runtime_error { using std::runtime_error::runtime_error; };

struct CandidateRec {
  int frame, level, component;
  float score;
  std::vector<int> x, y, m;       // per part: cell location and mixture id
  std::vector<int> rect;          // per part: cv::Rect x, y, width, height
};

class Engine {
 public:
  Engine(const Model& m, int device, cudaStream_t stream);
  ~Engine();
  Engine(const Engine&) = delete;
  Engine& operator=(const Engine&) = delete;

  // options
  double thresh;
  int exact = 1, backptr = 0, max_levels = 0, max_candidates = 65536, timing = 0;

  // ---- batch set-up ----
  void set_frames_geometry(int n, int h, int w, int c);                 // pyramid geometry of HOGFeatures::pyramid
  void set_levels_manual(int n, int nlevels, const int32_t* ohow, const float* scales);
  void upload_frames(const uint8_t* frames, size_t row_stride, size_t frame_stride);   // host -> device (async, pinned staging)
  void use_device_frames(const uint8_t* d_frames);

  // ---- stages (enqueue only) ----
  void run_pyramid();      // image pyramid + HOG  (IFeatures::pyramid)
  void run_pdf();          // IConvolutionEngine::pdf
  void run_dp_min();       // DynamicProgram::min
  void run_argmin();       // DynamicProgram::argmin (device part: hits + backtrack)
  void collect(std::vector<CandidateRec>& out);   // syncs, downloads and orders the candidates

  // ---- accessors (sync) ----
  const Geometry& geom() const { return g_; }
  const Model& model() const { return model_; }
  void get_pyramid_image(int frame, int level, uint8_t* dst);
  void get_features(int frame, int level, float* dst);
  void get_response(int frame, int level, int filter, float* dst);
  void get_rootv(int frame, int level, int comp, float* dst);
  void get_rooti(int frame, int level, int comp, int32_t* dst);
  void get_backptr(int frame, int level, int comp, int part, int pm, int32_t* ix, int32_t* iy, int32_t* ik);
  void set_features(int frame, int level, const float* src);
  void set_response(int frame, int level, int filter, const float* src);

  long long launches() const { return launches_; }
  size_t device_bytes() const { return dev_bytes_; }
  void stage_times(float ms[6]);
  cudaStream_t stream() const { return stream_; }
  int device() const { return device_; }

 private:
  void check_cuda(cudaError_t e, const char* what) const;
  template <typename T> void ensure(T*& p, size_t& cap, size_t n);
  void alloc_batch();
  void build_tables();
  void build_batch_tables();
  void need(int stage, const char* who) const;

  Model model_;
  int device_;
  cudaStream_t stream_;
  long long launches_ = 0;
  size_t dev_bytes_ = 0;

  // model-derived device data
  FilterBank fb_{};
  float* d_wpacked_ = nullptr;
  float* d_wgeneric_ = nullptr;
  int *d_foff_ = nullptr, *d_fkh_ = nullptr, *d_fkw_ = nullptr;
  std::vector<PartJob> jobs_;                 // ordered by wave
  std::vector<int> wave_first_, wave_count_, wave_maxmix_;
  std::vector<RootJob> roots_;
  PartJob* d_jobs_ = nullptr;
  RootJob* d_roots_ = nullptr;
  int nwork_ = 0, ncm_ = 0, npm_ = 0, tmp_maps_ = 0, max_parts_ = 0;
  std::vector<int> h_parent_, h_nparts_, h_cm_slot_, h_pm_slot_;
  int *d_parent_ = nullptr, *d_nparts_ = nullptr, *d_cm_slot_ = nullptr, *d_pm_slot_ = nullptr;

  // batch geometry + buffers
  Geometry g_{};
  Geometry* d_g_ = nullptr;
  bool have_images_ = false;
  bool feat_from_hog_ = false;
  int stage_ = 0;                              // 0 none, 1 geometry, 2 features, 3 responses, 4 dp, 5 argmin
  DeviceBuffers b_{};
  uint8_t* d_frames_own_ = nullptr; size_t cap_frames_ = 0;
  uint8_t* h_pinned_ = nullptr; size_t cap_pinned_ = 0;
  size_t cap_pyr_ = 0, cap_hist_ = 0, cap_norm_ = 0, cap_feat_ = 0, cap_resp_ = 0, cap_work_ = 0, cap_tmp_ = 0, cap_val_ = 0,
         cap_ixdt_ = 0, cap_iyraw_ = 0, cap_ik_ = 0, cap_rootv_ = 0, cap_rooti_ = 0;
  // tables depending on the batch geometry
  int *d_xofs_ = nullptr, *d_yofs_ = nullptr; short *d_xalpha_ = nullptr, *d_ybeta_ = nullptr;
  size_t cap_xofs_ = 0, cap_yofs_ = 0, cap_xalpha_ = 0, cap_ybeta_ = 0;
  int *d_tile_level_ = nullptr, *d_tile_first_ = nullptr; size_t cap_tile_level_ = 0, cap_tile_first_ = 0; int ntiles_ = 0;
  int *d_rg_level_ = nullptr, *d_rg_row0_ = nullptr; size_t cap_rg_level_ = 0, cap_rg_row0_ = 0; int nrg_ = 0;
  int *d_cg_level_ = nullptr, *d_cg_col0_ = nullptr; size_t cap_cg_level_ = 0, cap_cg_col0_ = 0; int ncg_ = 0;
  int max_ow_ = 0, max_oh_ = 0;
  PassGeom pg_rows_{}, pg_cols_{};
  PassGeom* d_pg_ = nullptr;                   // [rows, cols]
  PassMap *d_maps_rows_ = nullptr, *d_maps_cols_ = nullptr; size_t cap_maps_rows_ = 0, cap_maps_cols_ = 0;
  std::vector<int> wave_map_first_, wave_map_count_;
  // candidates
  Hit* d_hits_ = nullptr; size_t cap_hits_ = 0;
  int* d_nhits_ = nullptr;
  int* d_xym_ = nullptr; size_t cap_xym_ = 0;
  int* d_scratch_i_ = nullptr; size_t cap_scratch_i_ = 0;
  // timing
  cudaEvent_t ev_[7] = {};
  bool ev_valid_[7] = {};
};

// geometry helpers shared with the ABI (pyramid level table of HOGFeatures::pyramid)
int compute_pyramid_levels(int h, int w, int sbin, int interval, int max_levels, Geometry& g);

}  // namespace pbd
