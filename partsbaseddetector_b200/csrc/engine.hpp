// engine.hpp -- host orchestration of the detector's stage pipeline on one GPU / one stream.
// Plays the role of PartsBasedDetector<float> after distributeModel() (reference
// src/PartsBasedDetector.cpp:102-127): owns the feature engine, the convolution engine, the part
// tables and the dynamic program -- here as device buffers, tables and kernel launches.
#pragma once
#include <algorithm>
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

// Flat (SoA) candidate set: one allocation per field instead of per candidate.
struct CandidateSet {
  int n = 0, stride = 1;           // stride = part slots per candidate
  std::vector<int> meta;           // n x 4: frame, level, component, nparts
  std::vector<float> score;        // n: root score (reference confidence_[0])
  std::vector<int> parts;          // n x stride x 7: x, y, mixture, rect.x, rect.y, rect.width, rect.height
  void resize(int n_, int stride_) {
    n = n_; stride = stride_ > 0 ? stride_ : 1;
    meta.assign((size_t)n * 4, 0); score.assign(n, 0.f); parts.assign((size_t)n * stride * 7, 0);
  }
  int* part(int i, int p) { return parts.data() + ((size_t)i * stride + p) * 7; }
  const int* part(int i, int p) const { return parts.data() + ((size_t)i * stride + p) * 7; }
  // keep the candidates listed in `order`, in that order
  void select(const std::vector<int>& order) {
    CandidateSet o;
    o.resize((int)order.size(), stride);
    for (size_t k = 0; k < order.size(); ++k) {
      const int i = order[k];
      for (int t = 0; t < 4; ++t) o.meta[k * 4 + t] = meta[(size_t)i * 4 + t];
      o.score[k] = score[i];
      std::copy(parts.begin() + (size_t)i * stride * 7, parts.begin() + (size_t)(i + 1) * stride * 7, o.parts.begin() + k * (size_t)stride * 7);
    }
    *this = std::move(o);
  }
};

class Engine {
 public:
  Engine(const Model& m, int device, cudaStream_t stream);
  ~Engine();
  Engine(const Engine&) = delete;
  Engine& operator=(const Engine&) = delete;

  // options
  double thresh;
  // resp_mode: 0 exact (separately rounded multiply/add in the reference's order, bit-identical scores), 1 fused multiply-add,
  // 2 tensor cores (tf32x3 split products, fp32 accumulate; models the tensor kernel does not cover run the bit-exact kernel of mode 0),
  // 3 tensor cores with fp16 split operands (the same 11 + 11 significand bits, power-of-two pre-scaling; half the MMAs of mode 2)
  int resp_mode = 0, tc_taps_per_partial = 0, backptr = 0, max_levels = 0, max_candidates = 65536, timing = 0;
  // the DP stage runs the batch as this many groups of frames on concurrent streams (1 = single stream)
  int dp_streams = 2;
  // 1: enqueue_device() replays a CUDA graph of the whole path (captured on the second call with the same frames pointer, geometry and
  // options; timing must be 0).  Removes the ~100 launch overheads per batch: what a single-frame call is bound by.
  int use_graph = 0;
  // >= 0: detect() returns, per frame, the candidates sorted by score and greedily suppressed on the device (Candidate::sort +
  // Candidate::nonMaximaSuppression with this overlap); < 0 (default): the raw candidate list, as the reference's detect()
  double nms_overlap = -1.0;
  // > 0: root-map non-maxima suppression with this window before the backtrack (reference src/nms.cpp, the call commented out at
  // src/PartsBasedDetector.cpp:86); 0 (default): every root cell above the threshold is a candidate, as the reference's detect()
  int root_nms = 0;
  // which response kernel the last pdf stage ran: 0 generic exact (filters of different sizes), 1 tiled exact, 3 tensor tf32x3, 4 tensor
  // fp16x3, 5 generic FFMA, 6 tiled FFMA (read-only, option "response_kernel")
  int last_response_kernel = -1;
  // dt_pass variant: 0 eager emission with double break points (default), 1 eager emission with certified fp32 break points (12 %
  // slower on B200), 2 lagged-scan emission (15 % slower on real score maps, 6.8x faster on white noise); results are identical
  int dt_segment = -1;                         // dt_variant 3: lines cut into segments for launches that cannot fill the GPU: -1 automatic, 0 off, n = steps per segment
  int dt_scan = 3;                             // dt_variant: 3 = windowed certified evaluation with replay (dt_pass_win), 0 / 1 / 2 = stack kernels
  long long dt_replayed_lines();               // lines the windowed transform handed to the stack algorithm since the last call (synchronises)
  static constexpr int kMinFramesPerDpGroup = 4;

  // ---- batch set-up ----
  void set_frames_geometry(int n, int h, int w, int c);                 // pyramid geometry of HOGFeatures::pyramid
  void set_levels_manual(int n, int nlevels, const int32_t* ohow, const float* scales);
  void upload_frames(const uint8_t* frames, size_t row_stride, size_t frame_stride);   // host -> device (async)
  void use_device_frames(const uint8_t* d_frames);
  void upload_and_pyramid(const uint8_t* frames, size_t row_stride, size_t frame_stride);   // chunked H2D overlapped with pyramid + HOG

  // ---- stages (enqueue only) ----
  void run_pyramid();      // image pyramid + HOG  (IFeatures::pyramid)
  void run_pdf();          // IConvolutionEngine::pdf
  void run_dp_min();       // DynamicProgram::min
  void run_argmin();       // DynamicProgram::argmin (device part: hits + backtrack)
  void collect(CandidateSet& out);
  // pipelined API: submit() enqueues H2D (own copy stream, double-buffered frames) + all stages and returns a ticket (0/1);
  // collect_ticket() waits for that batch only and downloads its candidates on a separate stream.
  // all stages on device-resident frames, enqueue only (graph replay when use_graph is set); results via collect()
  void enqueue_device(const uint8_t* d_frames, int n, int h, int w, int c);
  int submit(const uint8_t* frames, int n, int h, int w, int c);
  void collect_ticket(int ticket, CandidateSet& out);   // syncs, downloads and orders the candidates

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

  void invalidate_geometry() { stage_ = 0; }     // an option that shapes the batch tables changed: rebuild them with the next batch
  long long launches() const { return launches_; }
  size_t device_bytes() const { return dev_bytes_; }
  void stage_times(float ms[6]);
  // per-kernel device time of the last batch run with timing == 2 (summed over the waves of the DP):
  // 0 feat_split, 1 part_response, 2 dt_pass rows, 3 dt_pass cols, 4 mix_max, 5 root_select
  static constexpr int kKernelTimes = 6;
  void kernel_times(float ms[kKernelTimes]);
  cudaStream_t stream() const { return stream_; }
  int device() const { return device_; }

 private:
  struct ResultSlot;
  void check_cuda(cudaError_t e, const char* what) const;
  void release();
  template <typename T> void ensure(T*& p, size_t& cap, size_t n);
  void alloc_batch();
  void build_tables();
  void build_batch_tables();
  void need(int stage, const char* who) const;
  void ensure_slot(ResultSlot& S);
  void ensure_tc(bool f16);
  void download_slot(ResultSlot& S, cudaStream_t st, CandidateSet& out);
  void chunked_upload_pyramid(const uint8_t* frames, uint8_t* d_dst, cudaEvent_t wait_before_copy, cudaEvent_t record_after, bool pipeline_busy = false);

  Model model_;
  int device_;
  cudaStream_t stream_;
  long long launches_ = 0;
  size_t dev_bytes_ = 0;

  // model-derived device data
  FilterBank fb_{};
  float* d_wpacked_ = nullptr;
  float* d_wgeneric_ = nullptr;
  // tensor-core response path: packed tf32 hi/lo weight slabs, padded strip copies of the HOG cells, work list
  float* d_wtc_ = nullptr;
  float *d_fhi_ = nullptr, *d_flo_ = nullptr; size_t cap_fhi_ = 0, cap_flo_ = 0;
  uint16_t* d_f16_ = nullptr; size_t cap_f16_ = 0;             // response mode 3: fp16 [hi | lo'] strips
  uint16_t* d_wtc16_ = nullptr; std::vector<float> wtc16_scales_;   // ... and weight slabs (+ the per-filter output scales, kernel parameters)
  long long tc16_serial_ = -1;
  TcLevel* d_tc_levels_ = nullptr; size_t cap_tc_levels_ = 0;
  TcTile* d_tc_tiles_ = nullptr; size_t cap_tc_tiles_ = 0;
  int tc_ntiles_ = 0; long long tc_frame_rows_ = 0;
  long long geom_serial_ = 0, tc_serial_ = -1;
  int num_sms_ = 0;
  int *d_foff_ = nullptr, *d_fkh_ = nullptr, *d_fkw_ = nullptr;
  std::vector<PartJob> jobs_;                 // ordered by wave
  std::vector<int> wave_first_, wave_count_, wave_maxmix_;
  std::vector<RootJob> roots_;
  PartJob* d_jobs_ = nullptr;
  RootJob* d_roots_ = nullptr;
  int nwork_ = 0, ncm_ = 0, npm_ = 0, tmp_maps_ = 0, max_parts_ = 0;
  std::vector<int> h_parent_, h_nparts_, h_cm_slot_, h_pm_slot_;
  int *d_parent_ = nullptr, *d_nparts_ = nullptr, *d_cm_slot_ = nullptr, *d_pm_slot_ = nullptr;
  std::vector<int> h_ksize_; int* d_ksize_ = nullptr;         // filter rows of (component, part, mixture): candidate rectangles
  NmsBuffers nms_{};
  size_t cap_nms_boxes_ = 0, cap_nms_keys_ = 0, cap_nms_skeys_ = 0, cap_nms_sidx_ = 0, cap_nms_kept_ = 0, cap_nms_fc_ = 0, cap_nms_fill_ = 0,
         cap_nms_kc_ = 0, cap_nms_oo_ = 0, cap_nms_so_ = 0, cap_nms_scratch_ = 0;
  void ensure_nms(ResultSlot& S);

  // batch geometry + buffers
  Geometry g_{};
  Geometry* d_g_ = nullptr;
  bool have_images_ = false;
  bool feat_from_hog_ = false;
  int stage_ = 0;                              // 0 none, 1 geometry, 2 features, 3 responses, 4 dp, 5 argmin
  DeviceBuffers b_{};
  uint8_t* d_frames_own_ = nullptr; size_t cap_frames_ = 0;
  size_t cap_pyr_ = 0, cap_hist_ = 0, cap_norm_ = 0, cap_feat_ = 0, cap_resp_ = 0, cap_work_ = 0, cap_tmp_ = 0, cap_val_ = 0,
         cap_ixdt_ = 0, cap_iyraw_ = 0, cap_ik_ = 0, cap_rootv_ = 0, cap_rooti_ = 0;
  // tables depending on the batch geometry
  int *d_xofs_ = nullptr, *d_yofs_ = nullptr; short *d_xalpha_ = nullptr, *d_ybeta_ = nullptr;
  size_t cap_xofs_ = 0, cap_yofs_ = 0, cap_xalpha_ = 0, cap_ybeta_ = 0;
  int *d_tile_level_ = nullptr, *d_tile_first_ = nullptr; size_t cap_tile_level_ = 0, cap_tile_first_ = 0; int ntiles_ = 0, ntiles0_ = 0;   // tiles [0, ntiles0_) : 8 x 4-quad tiles, the rest 6 x 5 (response.cu)
  int max_ow_ = 0, max_oh_ = 0;
  PassGeom pg_rows_{}, pg_cols_{};
  PassGeom* d_pg_ = nullptr;                   // [rows, cols]
  PassMap *d_maps_rows_ = nullptr, *d_maps_cols_ = nullptr; size_t cap_maps_rows_ = 0, cap_maps_cols_ = 0;
  double* d_etab_ = nullptr; size_t cap_etab_ = 0;   // per-map parabola tables of both passes
  std::vector<int> wave_map_first_, wave_map_count_;
  // dt_variant 3 (dt_pass_win): per-map window parameters, parallel to d_maps_rows_ / d_maps_cols_; counters of replayed lines
  dtw::WinParams *d_wp_rows_ = nullptr, *d_wp_cols_ = nullptr; size_t cap_wp_rows_ = 0, cap_wp_cols_ = 0;
  int* d_dtw_ctr_ = nullptr;                   // [64] one per frame group
  static constexpr int kSegGroups = 8;         // dp_streams <= 8
  size_t seg_lines_per_group() const { return (size_t)num_sms_ * 20 * 32; }
  int* d_seg_ctr_ = nullptr;                   // [kSegGroups][seg_lines_per_group()] per-line counters of the segmented windowed walk (dt.cu)
  // candidates
  // Result slots: the synchronous API uses slot 0; the pipelined submit/collect API alternates between the two so that the
  // candidates of batch i can be downloaded while batch i+1 is being computed.
  struct ResultSlot {
    Hit* d_hits = nullptr; size_t cap_hits = 0;
    int* d_nhits = nullptr;
    int* d_xym = nullptr; size_t cap_xym = 0;
    Hit* d_hits_out = nullptr; size_t cap_hits_out = 0;      // device NMS: compacted survivors of the batch
    int* d_xym_out = nullptr; size_t cap_xym_out = 0;
    int* d_total = nullptr;
    bool nms = false;                         // the batch in this slot went through the device NMS
    cudaEvent_t done = nullptr;               // recorded after the backtrack of the batch that filled the slot
    std::vector<float> scales;                // per-level scales of that batch (candidate rects)
    int max_candidates = 0;
    bool pending = false;
  };
  ResultSlot slots_[2];
  int cur_slot_ = 0;
  cudaStream_t d2h_stream_ = nullptr;
  std::vector<cudaStream_t> dp_aux_;           // extra streams of the DP stage's frame groups
  std::vector<cudaEvent_t> dp_join_;
  cudaEvent_t dp_fork_ = nullptr;
  uint8_t* d_frames_alt_ = nullptr; size_t cap_frames_alt_ = 0;   // second frame buffer of the pipelined API
  cudaEvent_t frames_free_ev_[2] = {};         // recorded when the pyramid stage has consumed frame buffer 0 / 1
  int frames_buf_ = 0;
  int* d_scratch_i_ = nullptr; size_t cap_scratch_i_ = 0;
  unsigned char* d_rootkeep_ = nullptr; size_t cap_rootkeep_ = 0;
  unsigned char* d_orient_lut_ = nullptr;      // snapped orientation of every (dx, dy) gradient (hog.cu)
  std::vector<Hit> h_hits_;                    // host staging reused across batches
  std::vector<int> h_xym_;
  cudaStream_t copy_stream_ = nullptr;
  cudaEvent_t copy_ev_[4] = {}, main_ev_ = nullptr;
  // timing
  cudaEvent_t ev_[7] = {};
  bool ev_valid_[7] = {};
  // timing == 2: one event after every kernel of the pdf / dp_min stages; kev_tag_[i] = kernel id of the interval ending at event i
  std::vector<cudaEvent_t> kev_;
  std::vector<int> kev_tag_;
  size_t kev_n_ = 0;
  void kmark(int tag);
  // CUDA-graph replay of enqueue_device(): the key is everything the captured launch sequence depends on
  struct GraphKey {
    const uint8_t* frames = nullptr;
    long long geom_serial = -1;
    int n = 0, resp_mode = -1, backptr = -1, max_candidates = 0, dp_streams = 0, root_nms = 0, dt_scan = 0, dt_segment = 0;
    double thresh = 0, nms_overlap = 0;
    bool operator==(const GraphKey& o) const {
      return frames == o.frames && geom_serial == o.geom_serial && n == o.n && resp_mode == o.resp_mode && backptr == o.backptr &&
             max_candidates == o.max_candidates && dp_streams == o.dp_streams && root_nms == o.root_nms && dt_scan == o.dt_scan && dt_segment == o.dt_segment && thresh == o.thresh && nms_overlap == o.nms_overlap;
    }
  };
  GraphKey graph_key_{}, warm_key_{};
  cudaGraphExec_t graph_exec_ = nullptr;
  cudaStream_t capture_stream_ = nullptr;
  long long graph_launches_ = 0;
  int graph_slot_ = 0;
  void run_stages();
};

// geometry helpers shared with the ABI (pyramid level table of HOGFeatures::pyramid)
int compute_pyramid_levels(int h, int w, int sbin, int interval, int max_levels, Geometry& g);

}  // namespace pbd
