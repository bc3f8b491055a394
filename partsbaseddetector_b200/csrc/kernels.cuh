// kernels.cuh -- device-side data layout shared by the stage kernels (sm_100a).
//
// HBM layout for one batch of n equally sized frames ("map-major over all levels"):
//   pyr    u8   [n][img_bytes]            pyramid images, level l at LevelDesc::img_off
//   hist   f32  [n][blocks_total][18]     orientation histograms per HOG block (+norm [n][blocks_total])
//   feat   f32  [n][cells_total][flen]    HOG features, HWC (reference featm layout, src/HOGFeatures.cpp:180)
//   resp   f32  [n][nfilters][cells_total]  part-filter responses, planar per filter
//   work   f32  [n][nwork][cells_total]   working scores of non-leaf parts (response + child messages)
//   tmp    f32  [n][njobs*MAXMIX][cells_total]  row-pass output of the part(s) currently processed (val: column-pass output)
//   ixdt   u16  [n][ncm][cells_total]     row-pass argmax per (component, part, child mixture), stored TRANSPOSED [x][y] per level
//   iyraw  u16  [n][ncm][cells_total]     column-pass argmax (not yet composed, see dt.cu)
//   ik     u8   [n][npm][cells_total]     best child mixture per (component, part, parent mixture)
//   rootv  f32  [n][ncomp][cells_total], rooti u8 [n][ncomp][cells_total]
// where a cell index inside a map is LevelDesc::cell_off + y*ow + x.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include <vector>

namespace pbd {

constexpr int kMaxLevels = 96;
constexpr int kMaxMix = 8;       // max mixtures per part (shipped models: <= 6)
constexpr int kMaxParts = 80;    // max parts per component (shipped models: <= 68)

struct LevelDesc {
  int img_w, img_h;      // pyramid image size (pixels)
  int bw, bh;            // HOG blocks  (reference `blocks`, src/HOGFeatures.cpp:174)
  int ow, oh;            // HOG cells   (reference `outsize`, :175)
  float scale;           // reference scales_[l], :118,124
  int src_level;         // -1: resized from the input frame, else pyrDown of that level
  int identity;          // 1: the level IS the input frame (same size: cv::resize copies): read in place from `frames`, no pyramid copy
  long long img_off;     // byte offset of the level image inside one frame's pyramid buffer
  int block_off;         // offset (in blocks) inside one frame's hist/norm arrays
  int cell_off;          // offset (in cells) inside one map
  int xofs_off, yofs_off;  // offsets into the resize tables (levels with src_level == -1)
};

struct Geometry {
  int n_frames, n_levels;
  int in_h, in_w, in_c;
  long long img_bytes;     // per frame
  int blocks_total, cells_total;
  LevelDesc lv[kMaxLevels];
};

// One DP job = one non-root part of one component (reference src/DynamicProgram.cpp:95-161).
struct PartJob {
  int nmix, pnmix;
  int in_slot[kMaxMix];      // child score map for mixture mm: response filter id, or work slot if in_is_work
  int in_is_work[kMaxMix];
  float w[kMaxMix][4];       // deformation weights (model defs), reference defw(mm)
  int ax[kMaxMix], ay[kMaxMix];  // anchor
  int cm_slot[kMaxMix];      // ixdt / iyraw slot of (c, p, mm)
  float bias[kMaxMix][kMaxMix];  // [mm][pm] = biasw[biasid[p][mm] + pm]
  int out_work_slot[kMaxMix];    // parent's working-score slot for parent mixture pm
  int out_resp_fid[kMaxMix];     // parent's filter id (first-touch initialisation, reference :155)
  int pm_slot[kMaxMix];      // ik slot of (c, p, pm)
  int first_touch;           // 1: work = resp + msg, 0: work += msg
  int tmp_base;              // first tmp map of this job (job_in_wave * kMaxMix)
};

struct RootJob {               // reference src/DynamicProgram.cpp:163-171
  int nmix;
  int in_slot[kMaxMix];
  int in_is_work[kMaxMix];
  float bias;
};

struct DeviceBuffers {
  const uint8_t* frames;   // [n][h][w][c] input
  uint8_t* pyr;
  float* hist;
  float* norm;
  float* feat;
  float* resp;
  float* work;
  float* tmp;      // row-pass output
  float* val;      // column-pass output (fully transformed child maps of the current wave)
  uint16_t* ixdt;
  uint16_t* iyraw;
  uint8_t* ik;
  float* rootv;
  uint8_t* rooti;
};

// candidate records produced by root_select / backtrack
struct Hit {
  int frame, level, comp, y, x;
  float score;
};

// ---- launchers (each enqueues on `s`, returns the number of kernels launched) ----
int launch_pyramid(const Geometry& g, const Geometry* d_g, const DeviceBuffers& b, const int* d_xofs, const short* d_xalpha,
                   const int* d_yofs, const short* d_ybeta, int frame0, int nframes, cudaStream_t s);
// d_orient_lut: hog_orient_lut_bytes() bytes filled once by launch_hog_orient_lut (the snapped orientation of every possible gradient)
size_t hog_orient_lut_bytes();
int launch_hog_orient_lut(unsigned char* d_lut, cudaStream_t s);
int launch_hog(const Geometry& g, const Geometry* d_g, const DeviceBuffers& b, const unsigned char* d_orient_lut, int sbin, int frame0, int nframes,
               cudaStream_t s);

// device copy of the model's filters (converted to float, reference src/PartsBasedDetector.cpp:115-117)
struct FilterBank {
  int nfilters, flen;
  int uniform, kh, kw;         // uniform: all filters kh x kw -> packed fast path
  int ngroups;                 // ceil(nfilters / 8)
  const float* w;              // packed  [group][c][ky][kx][8], zero padded
  const float* wg;             // generic [f] -> [c][ky][kx] at foff[f]
  const int* foff;
  const int* fkh;
  const int* fkw;
  int khm, kwm;                // largest filter (generic path halo)
};
int response_plan_tiles(const Geometry& g, const FilterBank& fb, std::vector<int>& tile_level, std::vector<int>& level_first);   // returns the number of shape-0 tiles
bool response_has_fast_path(const FilterBank& fb);
int launch_response_tiles(const Geometry& g, const Geometry* d_g, const DeviceBuffers& b, const FilterBank& fb, const int* d_tile_level,
                          const int* d_tile_first, int ntiles0, int ntiles, int exact, int trunc_zero, cudaStream_t s);

// ---- tensor-core response path (response_tc.cu): padded strip layout of the HOG cells + work list ----
struct TcLevel { int R, Wp, ow, oh, cell_off; };   // R: strip row of real cell (0,0); Wp: strip rows per image row (ow + ax)
struct TcTile { int level, q0; };                  // work item: 128 consecutive strip rows starting at q0
}  // namespace pbd
#include <vector>
namespace pbd {
bool response_tc_supported(const FilterBank& fb);
int response_tc_np(int nfilters);
void response_tc_pack_weights(const std::vector<std::vector<float>>& filters, int taps, std::vector<float>& out);
long long response_tc_plan(const Geometry& g, int kh, int kw, std::vector<TcLevel>& levels, std::vector<TcTile>& tiles, long long* slack_rows);
int launch_feat_split(const Geometry& g, const Geometry* d_g, const TcLevel* d_levels, const float* feat, float* fhi, float* flo,
                      long long frame_rows, cudaStream_t s);
int launch_tc_border_init(float* fhi, float* flo, long long rows, cudaStream_t s);
// fp16 flavour of the tensor path (response mode 3)
void response_tc_pack_weights_f16(const std::vector<std::vector<float>>& filters, int taps, std::vector<uint16_t>& out, std::vector<float>& out_scales);
int launch_feat_split_f16(const Geometry& g, const Geometry* d_g, const TcLevel* d_levels, const float* feat, uint16_t* strips, long long frame_rows,
                          cudaStream_t s);
int launch_tc_border_init_f16(uint16_t* strips, long long rows, cudaStream_t s);
int launch_response_tc(const Geometry& g, const DeviceBuffers& b, const FilterBank& fb, const float* fhi, const float* flo, const float* wpk,
                       const TcLevel* d_levels, const TcTile* d_tiles, int n_tiles, long long frame_rows, int num_sms, int taps_per_partial, cudaStream_t s,
                       const float* f16_scales = nullptr);

// Geometry of one separable-transform pass: per level the number of lines, their length and the map offset.
struct PassGeom {
  int n_levels;
  int nlines[kMaxLevels];
  int N[kMaxLevels];
  int cell_off[kMaxLevels];
};
// One map of a pass (frame-independent): element offsets from the buffers' frame base.
struct PassMap {
  unsigned long long in_off, out_off, ptr_off;
  int in_buf;                  // 0: buffer A, 1: buffer B (rows pass: A = responses, B = working scores)
  float w_sq, w_lin;           // deformation weights of this direction
  int os;                      // anchor of this direction
  // device table of the position-independent part of the parabola, etab[x + tab_bias] = a x^2 + b x for every
  // x = pos - v a line of up to maxn samples can produce (+ kDtTabPad look-ahead entries), followed by kDtRcp reciprocals
  // 1/(2a*dd); tab_len = dt_table_len(maxn) (filled by launch_dt_tables)
  double* etab;
  int tab_bias, tab_len;
};
// doubles a pass over lines of at most maxn samples needs per map, and the bias that goes with anchor `os`
constexpr int kDtTabPad = 2, kDtRcp = 32;     // look-ahead entries past the largest offset; tabulated reciprocals (dt_envelope.cuh)
inline int dt_table_len(int maxn) { return 2 * maxn - 1 + kDtTabPad + kDtRcp; }
inline int dt_table_bias(int maxn, int os) { return maxn - 1 - os; }
int launch_dt_tables(const PassMap* d_maps, int nmaps, cudaStream_t s);
// Geometry of one pass of the parallel-in-q transform (dt_lines.cuh): lines of a level are processed in batches of b[l] lines per warp
// (power of two <= 32, chosen so that a batch's state fits the warp's shared-memory region).  Rows pass: line = image row (element
// (line, q) at cell_off + line*N + q); columns pass: line = image column (element at cell_off + q*nlines + line).
struct LineGeom {
  int n_levels;
  int alias;                   // 1: every line has at most 256 samples, the ownership slots alias the break points (dt_lines.cuh)
  int region_bytes;            // shared memory per warp
  int nlines[kMaxLevels];
  int N[kMaxLevels];
  int cell_off[kMaxLevels];
  int b[kMaxLevels];
  int nblk[kMaxLevels];        // ceil(nlines / b)
};
// fills b / nblk / kreg / region_bytes from nlines / N for a per-warp shared-memory budget
void plan_line_geom(LineGeom& lg, int budget_bytes, bool stash);   // stash: the kernel parks u16 arg-maxes per sample (column kernels)
long long line_tasks(const LineGeom& lg, int units);     // warps (rows pass, units = maps) or CTAs (fused columns pass, units = jobs)
namespace dtw { struct WinParams; }
constexpr int kDtWindowW = 5;       // half window of dt_variant 3: candidates within 5 samples of the position (anchors up to +-5 are eligible)
int launch_dt_wave(const Geometry& g, const Geometry* d_g, const DeviceBuffers& b, const PassGeom& pg_rows, const PassGeom* d_pg_rows,
                   const PassGeom& pg_cols, const PassGeom* d_pg_cols, const PassMap* d_maps_rows, const PassMap* d_maps_cols, int nmaps,
                   int max_ow, int max_oh, const PartJob* d_jobs, int njobs, int nfilters, int nwork, int ncm, int npm, int tmp_maps,
                   cudaStream_t s, void (*mark)(void*, int) = nullptr, void* mark_ctx = nullptr,   // mark(ctx, kernel id) after each kernel
                   int scan = 0,                        // dt_pass variant: 0 double break points, 1 certified fp32 break points, 2 lagged-scan emission,
                                                        // 3 windowed certified evaluation (dt_pass_win; needs the per-map window parameters)
                   const dtw::WinParams* d_wp_rows = nullptr, const dtw::WinParams* d_wp_cols = nullptr, int* d_replayed = nullptr,
                   int warp_slots = 0,                  // resident warps of dt_pass_win on the device (SMs x CTAs per SM x warps per CTA)
                   int seg_mode = -1,                   // dt_variant 3, segmented walk: -1 automatic, 0 off, else steps per segment (dt.cu)
                   int* d_seg_ctr = nullptr);           // warp_slots * 32 zeroed ints private to this stream: per-line verdict counters of the segmented walk
int launch_root(const Geometry& g, const Geometry* d_g, const DeviceBuffers& b, const RootJob* d_roots, int ncomp, int nfilters, int nwork,
                cudaStream_t s);

// root_nms_sz > 0: only strict local maxima of the root map (window sz, reference src/nms.cpp) become hits; d_keep = [n][ncomp][cells] scratch
int launch_hits(const Geometry& g, const Geometry* d_g, const DeviceBuffers& b, int ncomp, float thresh, Hit* d_hits, int* d_nhits,
                int max_hits, cudaStream_t s, int root_nms_sz = 0, unsigned char* d_keep = nullptr);

struct BacktrackTables {       // per component, flattened with strides kMaxParts / kMaxMix
  const int* parent;           // [ncomp][kMaxParts]
  const int* nparts;           // [ncomp]
  const int* cm_slot;          // [ncomp][kMaxParts][kMaxMix]  ixdt/iyraw slot of (c, p, child mixture)
  const int* pm_slot;          // [ncomp][kMaxParts][kMaxMix]  ik slot of (c, p, parent mixture)
};
int launch_backtrack(const Geometry& g, const Geometry* d_g, const DeviceBuffers& b, const BacktrackTables& t, int ncomp, int ncm, int npm,
                     const Hit* d_hits, const int* d_nhits, int max_hits, int backptr_mode, int out_parts, int* d_out_xym, cudaStream_t s);
// materialise the reference's Ix/Iy/Ik Mats for one (frame, level, comp, part, parent mixture)
int launch_expand_backptr(const Geometry& g, const DeviceBuffers& b, int frame, int level, int ncm, int npm, const int* d_cm_slots,
                          int pm_slot, int backptr_mode, int* d_ix, int* d_iy, int* d_ik, cudaStream_t s);

// standalone 2-D DT over n_maps maps of h x w (config 5 microbenchmark / pbd_dt2d_f32); d_pg2 = {rows, cols} geometry,
// d_maps2 = n_maps row-pass maps followed by n_maps column-pass maps
int launch_dt2d_standalone(const float* d_in, int n_maps, int h, int w, const PassGeom* d_pg2, const PassMap* d_maps2, float* d_tmp,
                           float* d_out, uint16_t* d_ix, uint16_t* d_iy, uint16_t* d_ixraw, uint16_t* d_iyraw, int backptr_mode,
                           cudaStream_t s, int scan = 0,   // 0 eager streaming, 1 lagged scan, 3 (with d_wp2) windowed certified
                           const dtw::WinParams* d_wp2 = nullptr, int* d_replayed = nullptr,
                           int seg_steps = 0, int* d_seg_ctr = nullptr);  // scan 3: lines cut into segments (one zeroed int per line in d_seg_ctr)
// the same transform through the parallel-in-q kernels (all maps [y][x]); d_lg2 = {rows, cols}
int launch_dt2d_lines(const float* d_in, int n_maps, int h, int w, const LineGeom& lg_rows, const LineGeom& lg_cols, const LineGeom* d_lg2,
                      const PassMap* d_maps2, float* d_tmp, float* d_out, uint16_t* d_ix, uint16_t* d_iy, uint16_t* d_ixraw, uint16_t* d_iyraw,
                      int backptr_mode, cudaStream_t s);
constexpr int kLinesMaxN = 1024;   // longest line the parallel-in-q kernels are used for (longer: the streaming dt_pass)
// device-side Candidate::sort + nonMaximaSuppression (nms.cu): work buffers (shared by all batches of a detector) and the
// compacted result of one batch (hits_out / xym_out / total: per result slot)
struct NmsBuffers {
  int4* boxes;                    // [max_hits] clipped bounding box per raw candidate
  unsigned long long* keys;       // [max_hits]
  unsigned long long* skeys;      // [2*max_hits] per-frame power-of-two sort segments
  int *sidx, *kept_idx;           // [2*max_hits]
  int *frame_count, *fill, *kept_count, *out_off;   // [n_frames]
  int* seg_off;                   // [n_frames + 1]
  unsigned int* scratch;          // [n_frames][in_h][ceil(in_w / 32)] painted pixels, one bit each
  int* total;                     // [1] kept candidates of the batch
  Hit* hits_out;                  // [max_hits]
  int* xym_out;                   // [max_hits][3][out_parts]
};
size_t nms_scratch_words(const Geometry& g);
int launch_device_nms(const Geometry& g, const Geometry* d_g, const NmsBuffers& nb, const Hit* d_hits, const int* d_nhits, int max_hits,
                      const int* d_xym, int out_parts, const int* d_nparts, const int* d_ksize, float overlap, cudaStream_t s);
constexpr int kMaxDim = 1024;  // largest level width/height (cells) the DP kernels are instantiated for

}  // namespace pbd
