/* pbd_b200.h -- C-ABI of the B200-native parts-based detector hot path.
 *
 * This is the drop-in boundary for ONE path of wg-perception/PartsBasedDetector:
 * PartsBasedDetector<T>::detect() (reference src/PartsBasedDetector.cpp:69-95), i.e.
 *   HOGFeatures<T>::pyramid()            reference src/HOGFeatures.cpp:95-151, 168-341
 *   SpatialConvolutionEngine::pdf()      reference src/SpatialConvolutionEngine.cpp:106-124
 *   DynamicProgram<T>::min() / argmin()  reference src/DynamicProgram.cpp:67-173, 190-255
 *   DistanceTransform<T>::compute()      reference include/DistanceTransform.hpp:203-245
 * plus the model container/loader either side of it (Model / FileStorageModel,
 * reference include/Model.hpp:49-122, src/FileStorageModel.cpp:42-159).
 *
 * Conventions: plain pointers and sizes only; every function returns PBD_OK (0) or a
 * negative pbd_status and never throws; pbd_last_error() gives the message for the
 * calling thread.  One pbd_detector per host thread (the reference detector is not
 * re-entrant either: pyramid() mutates members, src/HOGFeatures.cpp:99,106-107).
 * The CUDA path is the only implementation: there is no CPU fallback, creating a
 * detector without a usable CUDA device fails with PBD_E_CUDA.
 */
#ifndef PBD_B200_H_
#define PBD_B200_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum pbd_status {
  PBD_OK = 0,
  PBD_E_ARG = -1,      /* bad argument (null pointer, index out of range, bad shape)           */
  PBD_E_IO = -2,       /* file cannot be opened (FileStorageModel::deserialize returns false)  */
  PBD_E_FORMAT = -3,   /* file opened but is not a model in the expected format                */
  PBD_E_CUDA = -4,     /* CUDA runtime error / no device                                        */
  PBD_E_STATE = -5,    /* stage called out of order (e.g. pdf before pyramid)                   */
  PBD_E_UNSUPPORTED = -6 /* e.g. image depth other than 8U (reference CV_Error, HOGFeatures.cpp:141-145) */
} pbd_status;

typedef struct pbd_model pbd_model;         /* reference: class Model (include/Model.hpp:49-122)          */
typedef struct pbd_detector pbd_detector;   /* reference: PartsBasedDetector<float> after distributeModel */
typedef struct pbd_candidates pbd_candidates; /* reference: std::vector<Candidate> (include/Candidate.hpp) */

const char* pbd_last_error(void);
const char* pbd_version(void);

/* ------------------------------------------------------------------ model ---
 * replaces FileStorageModel::deserialize / serialize (src/FileStorageModel.cpp:96-159 / 42-94).
 * pbd_model_load_xml reads the `opencv_storage` XML written by cv::FileStorage.  One deliberate
 * deviation from literal HEAD: a multi-valued <defid> is read in full (HEAD replaces it by [0] and
 * then indexes out of bounds, src/FileStorageModel.cpp:148-152); empty/root <defid> -> [0].
 * The .pbdm binary format is this library's own compact container of the same fields. */
int pbd_model_load_xml(const char* path, pbd_model** out);
int pbd_model_save_xml(const pbd_model* m, const char* path);
/* FileStorageModel as cv::FileStorage behaves: XML or YAML (%YAML:1.0 as OpenCV writes it), the parser chosen from the content, the
 * writer from the extension (.yml / .yaml -> YAML, otherwise XML) */
int pbd_model_load_storage(const char* path, pbd_model** out);
int pbd_model_save_storage(const pbd_model* m, const char* path);
/* replaces MatlabIOModel::deserialize (src/MatlabIOModel.cpp:71-188): reads the training code's `model` struct from a MATLAB
 * Level-5 MAT-file (native reader: compressed variables, small elements, any numeric storage type; no cvmatio).  1-based ids
 * become 0-based, filters w(m, n, c) are flattened to rows of n*C + c as the reference does; an empty (root) defid -> [0] as
 * in pbd_model_load_xml.  PBD_E_IO if the file cannot be opened (the reference returns false), PBD_E_FORMAT otherwise. */
int pbd_model_load_mat(const char* path, pbd_model** out);
int pbd_model_load_bin(const char* path, pbd_model** out);
int pbd_model_save_bin(const pbd_model* m, const char* path);
/* hdr = {interval, sbin, norient, flen, nfilters, nbias, ndefs, ncomponents};
 * fdims = (rows kh, kw) per filter; filters = concatenated kh*kw*flen doubles (HWC, channel fastest);
 * indexers = for each component: nparts, then per part: parentid, nf, nb, nd, filterid[nf], biasid[nb], defid[nd] */
int pbd_model_create(const char* name, const int32_t* hdr, float thresh, const int32_t* fdims,
                     const double* filters, const float* biasw, const int32_t* anchors,
                     const float* defs, const int32_t* indexers, pbd_model** out);
void pbd_model_free(pbd_model* m);

/* getters mirroring Model's (include/Model.hpp:98-118) */
const char* pbd_model_name(const pbd_model* m);
int pbd_model_header(const pbd_model* m, int32_t hdr[8], float* thresh);
int pbd_model_filter(const pbd_model* m, int i, int32_t* rows, int32_t* kw, const double** data);
int pbd_model_bias(const pbd_model* m, const float** data, int32_t* n);
int pbd_model_anchors(const pbd_model* m, const int32_t** xy, int32_t* n);
int pbd_model_defs(const pbd_model* m, const float** w4, int32_t* n);
int pbd_model_nparts(const pbd_model* m, int component);
/* copies up to cap ints of the requested list; *n receives its full length.  which: 0 filterid, 1 biasid, 2 defid */
int pbd_model_part(const pbd_model* m, int component, int part, int32_t* parentid, int which,
                   int32_t* dst, int32_t cap, int32_t* n);

/* --------------------------------------------------------------- detector ---
 * replaces PartsBasedDetector<float>::distributeModel (src/PartsBasedDetector.cpp:102-127): uploads the
 * filters (converted double->float as :115-117), the part tables and creates the stage pipelines.
 * `stream` is a cudaStream_t (0 = the legacy default stream); all work of this detector is issued on it. */
int pbd_create(const pbd_model* m, int device, void* stream, pbd_detector** out);
void pbd_destroy(pbd_detector* d);

/* options (defaults reproduce the reference):
 *   "thresh"      detection threshold (default model thresh; DynamicProgram::thresh_)
 *   "response_mode"  how the part-filter responses (IConvolutionEngine::pdf) are computed:
 *                 0 (default): separately rounded multiply/add in the reference's summation order => bit-identical scores;
 *                 1: fused multiply-add on the FP32 pipes (scores differ in the last ulps);
 *                 2: 5th-generation tensor cores (tcgen05, tf32x3 split products with fp32 accumulation; scores within ~4e-7
 *                    relative of the reference; over 256 frames 42 517 of 42 518 candidates identical to the reference's, one on the
 *                    other side of the threshold; 8x faster than mode 0).
 *                 3: the same three split products as tcgen05 kind::f16 MMAs: the fp32 operands are pre-scaled by powers of two (features
 *                    2^12, each filter so that its largest weight lies in [2^13, 2^14)) and split into fp16 hi + fp16 residual -- the same
 *                    11 + 11 significand bits as the tf32 split -- which halves the MMAs and the operand bytes; the same accuracy and
 *                    parity tests as mode 2, 1.4x faster.  Features must stay below 16 in magnitude (HOG features are <= 1).
 *                    Models with non-uniform filter sizes run the bit-exact kernel of mode 0 instead (modes 2 and 3).
 *   "exact"       alias kept for compatibility: 1 = response_mode 0, 0 = response_mode 1
 *   "tc_taps_per_partial"  response_mode 2: 0 (default) sums the hi*hi products of one filter row on the tensor core before the
 *                 fp32 round-to-nearest summation (score bias ~4 ulp), 1 sums tap by tap (bias < 1 ulp, 25 % slower)
 *   "backptr"     0 (default): reference back-pointer composition (include/DistanceTransform.hpp:232-244),
 *                 1: true 2-D argmax composition
 *   "max_levels"  0 (default) = all pyramid levels, n = only the first n (finest) levels
 *   "max_candidates" capacity of the candidate buffer per batch (default 65536)
 *   "timing"      1: record per-stage CUDA events (pbd_stage_times_ms); disables the chunked H2D/pyramid overlap;
 *                 2: additionally one event after every kernel of the pdf / dp_min stages (pbd_kernel_times_ms)
 *   "nms_overlap" < 0 (default): detect returns the raw candidate list in the reference's order, as PartsBasedDetector::detect;
 *                 0 <= v < 1: what the reference's callers do next (ros/Node.cpp:192-196) happens on the device -- per frame the
 *                 candidates are sorted by descending score (ties: raw-list order) and greedily suppressed by
 *                 Candidate::nonMaximaSuppression (include/Candidate.hpp:277-304) with overlap v; only the survivors are
 *                 downloaded, frame by frame in that order.  Needs real frames (not pbd_set_levels) and <= 32 components.
 *   "dp_streams"  1..8 (default 2): the DP stage processes the batch as this many groups of frames on concurrent CUDA streams
 *                 (forked from and joined back into the detector's stream) so that kernel tails of one group overlap the
 *                 other's work; results do not depend on it.  Batches under 8 frames and timing == 2 use one stream.
 * Environment defaults read at pbd_create: PBD_EXACT=0|1, PBD_RESPONSE_MODE=exact|ffma|tensor|tensor16, PBD_BACKPTR=reference|exact,
 * PBD_MAX_LEVELS=n, PBD_DP_STREAMS=n.  Further keys: "graph" (1: CUDA-graph replay of pbd_enqueue_batch_u8_device), "root_nms" (window
 * sz > 0: root-map non-maxima suppression of src/nms.cpp before the backtrack; 0 = off, the reference's detect()), "dt_variant" (3, the
 * default: windowed certified evaluation with in-kernel replay of the lines it cannot certify; 0 / 1 / 2: the stack-algorithm kernels
 * -- double break points, certified fp32 break points, lagged scan; identical results for all four), "dt_segment" (dt_variant 3: -1, the
 * default: launches that cannot fill the GPU -- single frames, small batches -- cut every line into segments walked by separate lanes, a
 * line being accepted only if all of its segments are; 0: never; n, a multiple of 16: n steps per segment; identical results),
 * get-only "dt_replayed_lines"
 * (lines variant 3 handed to the stack algorithm since the last query; synchronises).  Response modes 2 / 3 need a bank of equally sized square filters; for any
 * other model they run the bit-exact FP32 kernel (mode 0).  pbd_get_option("response_kernel") tells which kernel the last pdf stage
 * ran: 0 generic exact, 1 tiled exact, 3 tensor tf32x3, 4 tensor fp16x3, 5 generic FFMA, 6 tiled FFMA. */
int pbd_set_option(pbd_detector* d, const char* key, double value);
int pbd_get_option(const pbd_detector* d, const char* key, double* value);

/* whole path = PartsBasedDetector<T>::detect (src/PartsBasedDetector.cpp:69-95) on a batch of n equally
 * sized 8-bit frames (c = 1 or 3 channels, BGR interleaved as cv::Mat 8UC3).  Host buffers; the call
 * copies the frames to the device, runs all stages and returns the candidates of every frame.
 * Like the reference, candidates are in (level, component, row-major hit) order per frame. */
int pbd_detect_batch_u8(pbd_detector* d, const uint8_t* frames, int n, int h, int w, int c,
                        size_t row_stride, size_t frame_stride, pbd_candidates** out);
/* same, frames already resident in device memory (tightly packed n*h*w*c); result stays queued on the
 * stream until pbd_candidates_* is called (which synchronises). */
int pbd_detect_batch_u8_device(pbd_detector* d, const uint8_t* d_frames, int n, int h, int w, int c,
                               pbd_candidates** out);
/* enqueue only (no host sync, no candidate download): used for device-side timing */
int pbd_enqueue_batch_u8_device(pbd_detector* d, const uint8_t* d_frames, int n, int h, int w, int c);
int pbd_collect_candidates(pbd_detector* d, pbd_candidates** out);

/* pipelined variant for streams of batches: pbd_submit_batch_u8 enqueues the H2D copy (own copy stream, double-buffered
 * frame storage) and all stages of a batch and returns a ticket without waiting; pbd_collect_ticket waits for that batch
 * only and downloads its candidates on a separate stream, so transfer, compute and candidate download of consecutive batches
 * overlap.  At most two batches may be in flight.  `frames` must be tightly packed, should be pinned host memory and must stay
 * untouched until its ticket has been collected.  Results equal pbd_detect_batch_u8's. */
int pbd_submit_batch_u8(pbd_detector* d, const uint8_t* frames, int n, int h, int w, int c, int* ticket);
int pbd_collect_ticket(pbd_detector* d, int ticket, pbd_candidates** out);

/* candidates: reference Candidate (include/Candidate.hpp:56-99) = per part cv::Rect + confidence,
 * component; additionally the frame, pyramid level and per-part (x, y, mixture) in cell coordinates. */
int pbd_candidates_count(const pbd_candidates* c);
int pbd_candidates_nparts(const pbd_candidates* c, int i);
int pbd_candidates_get(const pbd_candidates* c, int i, int32_t* frame, int32_t* level, int32_t* component,
                       float* score, int32_t* xs, int32_t* ys, int32_t* ms, int32_t* rects_xywh);
/* bulk export of all candidates in one call: meta[i] = {frame, level, component, nparts}, scores[i] = root score,
 * parts[i][p] = {x, y, mixture, rect.x, rect.y, rect.width, rect.height} for p < nparts (row stride max_nparts*7). */
int pbd_candidates_export(const pbd_candidates* c, int32_t* meta4, float* scores, int32_t* parts7, int max_nparts);
void pbd_candidates_free(pbd_candidates* c);
/* Candidate::sort (include/Candidate.hpp:97-99): descending root score, stable */
int pbd_candidates_sort(pbd_candidates* c);
/* Candidate::nonMaximaSuppression (include/Candidate.hpp:277-304): greedy box painting over the candidates in their
 * current order (callers sort first, ros/Node.cpp:192-196), applied independently per frame of the batch.  A candidate is
 * dropped when more than `overlap` of its (image-clipped) bounding box has already been painted by kept candidates. */
int pbd_candidates_nms(pbd_candidates* c, int im_h, int im_w, float overlap);
/* SearchSpacePruning<T>::filterCandidatesByDepth (src/SearchSpacePruning.cpp:73-95; the call the reference keeps commented out at
 * src/PartsBasedDetector.cpp:91-93, zfactor 0.03): drops every candidate with a part whose median depth differs from its parent's by
 * more than |anchor| * zfactor (both medians > 0).  depth = im_h x im_w float image (row_stride_bytes 0 = packed) of the frame the
 * candidates come from; boxes are clipped to it (the reference asserts on boxes that cross the border). */
int pbd_candidates_filter_by_depth(pbd_candidates* c, const pbd_model* m, const float* depth, int im_h, int im_w, size_t row_stride_bytes,
                                   float zfactor);
/* build a candidate set from arrays (layout of pbd_candidates_export); for callers that post-process their own lists */
int pbd_candidates_create(int n, int max_nparts, const int32_t* meta4, const float* scores, const int32_t* parts7,
                          pbd_candidates** out);

/* ------------------------------------------------------------ stage level ---
 * The reference's plugin interfaces, one call per stage, operating on the detector's device buffers:
 *   pbd_stage_pyramid  = IFeatures::pyramid        (include/IFeatures.hpp:49-73)
 *   pbd_stage_pdf      = IConvolutionEngine::pdf   (include/IConvolutionEngine.hpp:44-68)
 *   pbd_stage_dp_min   = DynamicProgram<T>::min    (include/DynamicProgram.hpp:74)
 *   pbd_stage_dp_argmin= DynamicProgram<T>::argmin (include/DynamicProgram.hpp:75)
 * Stages must run in this order after pbd_stage_pyramid (or after the corresponding pbd_set_* injection). */
int pbd_stage_pyramid(pbd_detector* d, const uint8_t* frames, int n, int h, int w, int c,
                      size_t row_stride, size_t frame_stride);
int pbd_stage_pdf(pbd_detector* d);
int pbd_stage_dp_min(pbd_detector* d);
int pbd_stage_dp_argmin(pbd_detector* d, pbd_candidates** out);

/* level table HOGFeatures<T>::pyramid would build for an h x w image (src/HOGFeatures.cpp:95-127 and :174-176), without a
 * device: returns the number of levels and fills up to `cap` entries of (img_h, img_w, oh, ow) and the scales. */
int pbd_pyramid_geometry(int h, int w, int sbin, int interval, int max_levels, int cap, int32_t* dims4, float* scales);
/* geometry of the current batch (IFeatures::nscales / scales, include/IFeatures.hpp:60-66) */
int pbd_num_frames(const pbd_detector* d);
int pbd_num_levels(const pbd_detector* d);
int pbd_level_info(const pbd_detector* d, int level, int32_t* img_h, int32_t* img_w, int32_t* oh,
                   int32_t* ow, float* scale);

/* stage outputs copied to host (synchronises the stream) */
int pbd_get_pyramid_image(pbd_detector* d, int frame, int level, uint8_t* dst);          /* img_h*img_w*c   */
int pbd_get_features(pbd_detector* d, int frame, int level, float* dst);                 /* oh*ow*flen, HWC  */
int pbd_get_response(pbd_detector* d, int frame, int level, int filter, float* dst);     /* oh*ow            */
int pbd_get_rootv(pbd_detector* d, int frame, int level, int component, float* dst);     /* oh*ow            */
int pbd_get_rooti(pbd_detector* d, int frame, int level, int component, int32_t* dst);   /* oh*ow            */
/* Ix/Iy/Ik[level][component][part][parent mixture] as the reference's CV_32S Mats */
int pbd_get_backptr(pbd_detector* d, int frame, int level, int component, int part, int parent_mixture,
                    int32_t* ix, int32_t* iy, int32_t* ik);

/* stage inputs injected from host (stage-isolated parity tests): define a batch of n frames with the given
 * level table, then upload features and/or responses. */
int pbd_set_levels(pbd_detector* d, int n_frames, int n_levels, const int32_t* ohow, const float* scales);
int pbd_set_features(pbd_detector* d, int frame, int level, const float* src);
int pbd_set_response(pbd_detector* d, int frame, int level, int filter, const float* src);

/* ------------------------------------------------- standalone DT (config 5) ---
 * Generalised distance transform of n_maps score maps of h x w (device pointers), one (w0..w3, ax, ay)
 * per map: out[m] = DT(in[m]), ix/iy = back-pointers (uint16) with the reference composition.
 * Reference: DistanceTransform<float>::compute, include/DistanceTransform.hpp:203-245.  Samples may be any float: a line with a
 * NaN or an infinity is redone with the reference's two loops, literally, so the result is whatever the reference's comparisons make
 * of such a sample (plan impl 2, the parallel-in-q alternative, is the exception: finite samples only). */
int pbd_dt2d_f32_device(void* stream, const float* d_in, int n_maps, int h, int w, const float* h_defw4,
                        const int32_t* h_anchor_xy, float* d_out, uint16_t* d_ix, uint16_t* d_iy, int backptr_mode);
/* host-buffer convenience wrapper of the above (allocates, copies, runs, copies back) */
int pbd_dt2d_f32(const float* in, int n_maps, int h, int w, const float* defw4, const int32_t* anchor_xy,
                 float* out, int32_t* ix, int32_t* iy, int backptr_mode);
/* The same transform with every table and scratch buffer owned by a plan, so that pbd_dt2d_plan_run only enqueues kernels on
 * `stream` (no allocation, no synchronisation): what the DT microbenchmark times.  impl: 0 = default (1), 1 = streaming envelope
 * (one lane per line, any length <= 4096; the kernels the detector runs), 2 = parallel-in-q (a warp per batch of lines in shared
 * memory, lines <= 1024; bit-identical, kept as a measured alternative), 3 = streaming envelope with lagged-scan emission (one store
 * per position: the variant for rough inputs such as the white-noise maps of the microbenchmark), 4 = windowed certified evaluation
 * with replay of the lines it cannot certify (the detector's default transform, dt_variant 3: fast on score-like maps whose arg-max
 * stays within 4 samples of the position, slower than 1 on maps where it does not); pbd_dt2d_plan_replayed: lines impl 4 handed to
 * the stack algorithm since the last call (synchronises; 0 for the other impls).  All impls give identical results. */
typedef struct pbd_dt2d_plan pbd_dt2d_plan;
int pbd_dt2d_plan_create(int n_maps, int h, int w, const float* defw4, const int32_t* anchor_xy, int impl, pbd_dt2d_plan** out);
int pbd_dt2d_plan_impl(const pbd_dt2d_plan* p);
/* impl 4 only: cut every line into segments of `steps` walk steps (a multiple of 16, 0 = off: the default) -- the form the detector
 * chooses by itself for launches that cannot fill the GPU (option "dt_segment"); identical results. */
int pbd_dt2d_plan_set_segment(pbd_dt2d_plan* p, int steps);
long long pbd_dt2d_plan_replayed(pbd_dt2d_plan* p);
int pbd_dt2d_plan_run(pbd_dt2d_plan* p, void* stream, const float* d_in, float* d_out, uint16_t* d_ix, uint16_t* d_iy, int backptr_mode);
void pbd_dt2d_plan_destroy(pbd_dt2d_plan* p);

/* --------------------------------------------------------------- ingest ---
 * What the reference's callers do before detect(): cv::imread (src/demo.cpp:88-99) and cv_bridge::toCvCopy (ros/Node.cpp:165-176).
 * Containers: PNG (zlib) and binary PNM; JPEG is not decoded here (PBD_E_UNSUPPORTED).  All host code. */
/* header only: size, channels as stored (1, 2, 3, 4) and bits per sample (8 / 16) */
int pbd_image_info(const uint8_t* bytes, size_t n, int32_t* h, int32_t* w, int32_t* channels, int32_t* bits);
/* cv::imread(..., IMREAD_COLOR): packed 8-bit BGR (grey replicated, alpha dropped, 16-bit samples >> 8); dst_capacity in bytes */
int pbd_image_decode_bgr8(const uint8_t* bytes, size_t n, uint8_t* dst, size_t dst_capacity, int32_t* h, int32_t* w);
/* cv::imread(..., IMREAD_ANYDEPTH) of a depth image, times `scale` (the demo: 1/1000, mm -> m); dst_capacity in floats */
int pbd_image_decode_depth_f32(const uint8_t* bytes, size_t n, float scale, float* dst, size_t dst_capacity, int32_t* h, int32_t* w);
/* file variant; dst_capacity = 0 only queries the size */
int pbd_imread_bgr8(const char* path, uint8_t* dst, size_t dst_capacity, int32_t* h, int32_t* w);
/* sensor_msgs/Image payload (encoding bgr8 / rgb8 / bgra8 / rgba8 / mono8 / mono16 / bgr16 / rgb16 / 8UC1 / 8UC3 / 8UC4 / 16UC1) -> packed
 * BGR8, as cv_bridge::toCvCopy(msg, enc::BGR8); step = bytes per row (0 = packed) */
int pbd_ros_image_to_bgr8(const char* encoding, int h, int w, size_t step, int is_bigendian, const uint8_t* data, uint8_t* dst_bgr8);
/* sensor_msgs/Image depth payload (32FC1 / 16UC1) -> packed float, as cv_bridge::toCvCopy(msg, enc::TYPE_32FC1) */
int pbd_ros_depth_to_f32(const char* encoding, int h, int w, size_t step, int is_bigendian, const uint8_t* data, float* dst);
/* pinned (page-locked, portable) host memory for the frame ring of pbd_submit_batch_u8 */
int pbd_host_alloc_pinned(size_t bytes, void** out);
void pbd_host_free_pinned(void* p);

/* ----------------------------------------------------------- measurement --- */
/* number of kernels launched by this detector since creation (bench.py's gpu_launches) */
long long pbd_launch_count(const pbd_detector* d);
/* per-stage device time of the last pbd_detect_batch / pbd_enqueue call, measured with CUDA events on the
 * detector's stream: ms[0..5] = h2d, pyramid, hog, pdf, dp_min, argmin.  Enabled by option "timing"=1. */
int pbd_stage_times_ms(pbd_detector* d, float ms[6]);
/* "timing" = 2 additionally records an event after every kernel of the pdf and dp_min stages; device time of the last batch per
 * kernel (summed over the waves of the DP): 0 feat_split, 1 part_response, 2 dt_pass rows, 3 dt_pass columns, 4 mix_max, 5 root_select */
int pbd_kernel_times_ms(pbd_detector* d, float ms[6]);
/* device bytes currently held by the detector */
size_t pbd_device_bytes(const pbd_detector* d);

#ifdef __cplusplus
}
#endif
#endif /* PBD_B200_H_ */
