// abi.cpp -- extern "C" boundary (include/pbd_b200.h).  Exceptions never cross it.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <string>

#include "../../include/pbd_b200.h"
#include "engine.hpp"
#include "dt_window.cuh"
#include "ingest.hpp"
#include "model.hpp"

using namespace pbd;

struct pbd_model { Model m; };
struct pbd_detector { std::unique_ptr<Engine> e; };
struct pbd_candidates { CandidateSet s; };

namespace {
thread_local std::string g_err;

template <typename F>
int guarded(F f) {
  try { f(); g_err.clear(); return PBD_OK; }
  catch (const IoError& e) { g_err = e.what(); return PBD_E_IO; }
  catch (const FormatError& e) { g_err = e.what(); return PBD_E_FORMAT; }
  catch (const CudaError& e) { g_err = e.what(); return PBD_E_CUDA; }
  catch (const StateError& e) { g_err = e.what(); return PBD_E_STATE; }
  catch (const UnsupportedError& e) { g_err = e.what(); return PBD_E_UNSUPPORTED; }
  catch (const ArgError& e) { g_err = e.what(); return PBD_E_ARG; }
  catch (const std::bad_alloc&) { g_err = "out of host memory"; return PBD_E_ARG; }
  catch (const std::exception& e) { g_err = e.what(); return PBD_E_ARG; }
  catch (...) { g_err = "unknown error"; return PBD_E_ARG; }
}
#define REQUIRE(cond, msg) do { if (!(cond)) throw ArgError(msg); } while (0)

// Every entry point that touches a detector makes the detector's device current for the duration of the call and restores the
// caller's device afterwards: a process may hold detectors on several devices, and a host thread's current device is whatever its
// last caller left behind.
struct DeviceGuard {
  int prev = -1, dev = -1;
  explicit DeviceGuard(int device) : dev(device) {
    if (dev < 0) return;
    if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) throw CudaError("cudaSetDevice failed");
  }
  ~DeviceGuard() { if (dev >= 0 && prev >= 0 && prev != dev) cudaSetDevice(prev); }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};
inline int dev_of(const pbd_detector* d) { return d && d->e ? d->e->device() : -1; }
}  // namespace

extern "C" {

const char* pbd_last_error(void) { return g_err.c_str(); }
const char* pbd_version(void) { return "pbd_b200 0.1 (sm_100a)"; }

int pbd_model_load_xml(const char* path, pbd_model** out) {
  return guarded([&] { REQUIRE(path && out, "null argument"); auto m = std::make_unique<pbd_model>(); load_xml(path, m->m); *out = m.release(); });
}
int pbd_model_save_xml(const pbd_model* m, const char* path) { return guarded([&] { REQUIRE(m && path, "null argument"); save_xml(m->m, path); }); }
int pbd_model_load_storage(const char* path, pbd_model** out) {
  return guarded([&] { REQUIRE(path && out, "null argument"); auto m = std::make_unique<pbd_model>(); load_storage(path, m->m); *out = m.release(); });
}
int pbd_model_save_storage(const pbd_model* m, const char* path) { return guarded([&] { REQUIRE(m && path, "null argument"); save_storage(m->m, path); }); }
int pbd_model_load_mat(const char* path, pbd_model** out) {
  return guarded([&] { REQUIRE(path && out, "null argument"); auto m = std::make_unique<pbd_model>(); load_mat(path, m->m); *out = m.release(); });
}
int pbd_model_load_bin(const char* path, pbd_model** out) {
  return guarded([&] { REQUIRE(path && out, "null argument"); auto m = std::make_unique<pbd_model>(); load_bin(path, m->m); *out = m.release(); });
}
int pbd_model_save_bin(const pbd_model* m, const char* path) { return guarded([&] { REQUIRE(m && path, "null argument"); save_bin(m->m, path); }); }

int pbd_model_create(const char* name, const int32_t* hdr, float thresh, const int32_t* fdims, const double* filters,
                     const float* biasw, const int32_t* anchors, const float* defs, const int32_t* indexers, pbd_model** out) {
  return guarded([&] {
    REQUIRE(hdr && fdims && filters && biasw && anchors && defs && indexers && out, "null argument");
    auto pm = std::make_unique<pbd_model>();
    Model& m = pm->m;
    m.name = name ? name : "";
    m.interval = hdr[0]; m.sbin = hdr[1]; m.norient = hdr[2]; m.flen = hdr[3];
    const int nf = hdr[4], nb = hdr[5], nd = hdr[6], nc = hdr[7];
    REQUIRE(nf > 0 && nb > 0 && nd >= 0 && nc > 0 && m.flen > 0, "bad header");
    m.thresh = thresh;
    size_t off = 0;
    for (int i = 0; i < nf; ++i) {
      REQUIRE(fdims[2 * i] > 0 && fdims[2 * i + 1] > 0, "bad filter dims");
      m.frows.push_back(fdims[2 * i]); m.fkw.push_back(fdims[2 * i + 1]);
      const size_t n = (size_t)fdims[2 * i] * fdims[2 * i + 1] * m.flen;
      m.filters.emplace_back(filters + off, filters + off + n);
      off += n;
    }
    m.biasw.assign(biasw, biasw + nb);
    m.anchors.assign(anchors, anchors + 2 * (size_t)nd);
    m.defs.assign(defs, defs + 4 * (size_t)nd);
    const int32_t* ip = indexers;
    m.comps.resize(nc);
    for (int c = 0; c < nc; ++c) {
      const int np = *ip++;
      REQUIRE(np > 0, "component without parts");
      m.comps[c].resize(np);
      for (int p = 0; p < np; ++p) {
        Part& P = m.comps[c][p];
        P.parentid = *ip++;
        const int a = *ip++, b = *ip++, d = *ip++;
        REQUIRE(a >= 0 && b >= 0 && d >= 0, "bad indexer lengths");
        P.filterid.assign(ip, ip + a); ip += a;
        P.biasid.assign(ip, ip + b); ip += b;
        P.defid.assign(ip, ip + d); ip += d;
        if (P.defid.empty()) P.defid.push_back(0);
      }
    }
    m.validate();
    *out = pm.release();
  });
}
void pbd_model_free(pbd_model* m) { delete m; }

const char* pbd_model_name(const pbd_model* m) { return m ? m->m.name.c_str() : ""; }
int pbd_model_header(const pbd_model* m, int32_t hdr[8], float* thresh) {
  return guarded([&] {
    REQUIRE(m && hdr, "null argument");
    const Model& M = m->m;
    hdr[0] = M.interval; hdr[1] = M.sbin; hdr[2] = M.norient; hdr[3] = M.flen; hdr[4] = M.nfilters(); hdr[5] = (int)M.biasw.size();
    hdr[6] = M.ndefs(); hdr[7] = M.ncomponents();
    if (thresh) *thresh = M.thresh;
  });
}
int pbd_model_filter(const pbd_model* m, int i, int32_t* rows, int32_t* kw, const double** data) {
  return guarded([&] {
    REQUIRE(m && i >= 0 && i < m->m.nfilters(), "filter index out of range");
    if (rows) *rows = m->m.frows[i];
    if (kw) *kw = m->m.fkw[i];
    if (data) *data = m->m.filters[i].data();
  });
}
int pbd_model_bias(const pbd_model* m, const float** data, int32_t* n) {
  return guarded([&] { REQUIRE(m && data && n, "null argument"); *data = m->m.biasw.data(); *n = (int)m->m.biasw.size(); });
}
int pbd_model_anchors(const pbd_model* m, const int32_t** xy, int32_t* n) {
  return guarded([&] { REQUIRE(m && xy && n, "null argument"); *xy = m->m.anchors.data(); *n = (int)m->m.anchors.size() / 2; });
}
int pbd_model_defs(const pbd_model* m, const float** w4, int32_t* n) {
  return guarded([&] { REQUIRE(m && w4 && n, "null argument"); *w4 = m->m.defs.data(); *n = m->m.ndefs(); });
}
int pbd_model_nparts(const pbd_model* m, int component) {
  if (!m || component < 0 || component >= m->m.ncomponents()) { g_err = "component out of range"; return PBD_E_ARG; }
  return (int)m->m.comps[component].size();
}
int pbd_model_part(const pbd_model* m, int component, int part, int32_t* parentid, int which, int32_t* dst, int32_t cap, int32_t* n) {
  return guarded([&] {
    REQUIRE(m && component >= 0 && component < m->m.ncomponents(), "component out of range");
    const auto& parts = m->m.comps[component];
    REQUIRE(part >= 0 && part < (int)parts.size(), "part out of range");
    REQUIRE(which >= 0 && which <= 2, "bad list selector");
    const Part& P = parts[part];
    if (parentid) *parentid = P.parentid;
    const std::vector<int>& v = which == 0 ? P.filterid : which == 1 ? P.biasid : P.defid;
    if (n) *n = (int)v.size();
    if (dst) for (int i = 0; i < cap && i < (int)v.size(); ++i) dst[i] = v[i];
  });
}

int pbd_create(const pbd_model* m, int device, void* stream, pbd_detector** out) {
  return guarded([&] {
    REQUIRE(m && out, "null argument");
    auto d = std::make_unique<pbd_detector>();
    d->e = std::make_unique<Engine>(m->m, device, (cudaStream_t)stream);
    // process-wide defaults from the environment (SURVEY.md section 5 "config / flags"); pbd_set_option overrides them
    if (const char* v = getenv("PBD_EXACT")) d->e->resp_mode = atoi(v) != 0 ? 0 : 1;
    if (const char* v = getenv("PBD_RESPONSE_MODE")) {
      if (!strcmp(v, "exact") || !strcmp(v, "0")) d->e->resp_mode = 0;
      else if (!strcmp(v, "ffma") || !strcmp(v, "1")) d->e->resp_mode = 1;
      else if (!strcmp(v, "tensor") || !strcmp(v, "2")) d->e->resp_mode = 2;
      else if (!strcmp(v, "tensor16") || !strcmp(v, "3")) d->e->resp_mode = 3;
    }
    if (const char* v = getenv("PBD_BACKPTR")) d->e->backptr = (strcmp(v, "exact") == 0 || strcmp(v, "1") == 0) ? 1 : 0;
    if (const char* v = getenv("PBD_MAX_LEVELS")) d->e->max_levels = std::max(0, atoi(v));
    if (const char* v = getenv("PBD_DP_STREAMS")) d->e->dp_streams = std::min(8, std::max(1, atoi(v)));
    *out = d.release();
  });
}
void pbd_destroy(pbd_detector* d) {
  if (!d) return;
  try { DeviceGuard dg_(dev_of(d)); delete d; } catch (...) { delete d; }
}

int pbd_set_option(pbd_detector* d, const char* key, double value) {
  return guarded([&] { DeviceGuard dg_(dev_of(d));
    REQUIRE(d && key, "null argument");
    Engine& e = *d->e;
    const std::string k(key);
    if (k == "thresh") e.thresh = value;
    else if (k == "exact") e.resp_mode = value != 0 ? 0 : 1;
    else if (k == "response_mode") { REQUIRE(value == 0 || value == 1 || value == 2 || value == 3, "response_mode must be 0 (exact), 1 (ffma), 2 (tensor tf32) or 3 (tensor fp16)"); e.resp_mode = (int)value; }
    else if (k == "tc_taps_per_partial") { REQUIRE(value >= 0 && value <= 1024, "tc_taps_per_partial out of range"); e.tc_taps_per_partial = (int)value; }
    else if (k == "backptr") { REQUIRE(value == 0 || value == 1, "backptr must be 0 or 1"); e.backptr = (int)value; }
    else if (k == "max_levels") { REQUIRE(value >= 0, "max_levels must be >= 0"); e.max_levels = (int)value; }
    else if (k == "max_candidates") { REQUIRE(value >= 1 && value <= (1 << 24), "max_candidates out of range"); e.max_candidates = (int)value; }
    else if (k == "timing") e.timing = value >= 2 ? 2 : (value != 0);
    else if (k == "nms_overlap") { REQUIRE(value < 1.0, "nms_overlap must be < 1 (negative = off)"); e.nms_overlap = value; }
    else if (k == "dp_streams") { REQUIRE(value >= 1 && value <= 8, "dp_streams must be 1..8"); e.dp_streams = (int)value; }
    else if (k == "graph") e.use_graph = value != 0;
    else if (k == "dt_variant") { REQUIRE(value == 0 || value == 1 || value == 2 || value == 3, "dt_variant must be 0 (double break points), 1 (certified fp32 break points), 2 (lagged scan) or 3 (windowed certified transform)"); e.dt_scan = (int)value; }
    else if (k == "dt_segment") { REQUIRE(value == -1 || value == 0 || (value >= 16 && value <= 4096 && (int)value % 16 == 0), "dt_segment must be -1 (automatic), 0 (off) or a multiple of 16 (steps per line segment)"); e.dt_segment = (int)value; }
    else if (k == "root_nms") { REQUIRE(value >= 0 && value <= 64, "root_nms window must be 0 (off) .. 64"); e.root_nms = (int)value; }
    else throw ArgError("unknown option '" + k + "'");
  });
}
int pbd_get_option(const pbd_detector* d, const char* key, double* value) {
  return guarded([&] { DeviceGuard dg_(dev_of(d));
    REQUIRE(d && key && value, "null argument");
    const Engine& e = *d->e;
    const std::string k(key);
    if (k == "thresh") *value = e.thresh;
    else if (k == "exact") *value = e.resp_mode == 0;
    else if (k == "response_mode") *value = e.resp_mode;
    else if (k == "tc_taps_per_partial") *value = e.tc_taps_per_partial;
    else if (k == "backptr") *value = e.backptr;
    else if (k == "max_levels") *value = e.max_levels;
    else if (k == "max_candidates") *value = e.max_candidates;
    else if (k == "timing") *value = e.timing;
    else if (k == "dp_streams") *value = e.dp_streams;
    else if (k == "graph") *value = e.use_graph;
    else if (k == "response_kernel") *value = e.last_response_kernel;
    else if (k == "dt_variant") *value = e.dt_scan;
    else if (k == "dt_segment") *value = e.dt_segment;
    else if (k == "dt_replayed_lines") *value = (double)const_cast<Engine&>(e).dt_replayed_lines();
    else if (k == "root_nms") *value = e.root_nms;
    else if (k == "nms_overlap") *value = e.nms_overlap;
    else throw ArgError("unknown option '" + k + "'");
  });
}

static void run_all(Engine& e) { e.run_pyramid(); e.run_pdf(); e.run_dp_min(); e.run_argmin(); }

int pbd_detect_batch_u8(pbd_detector* d, const uint8_t* frames, int n, int h, int w, int c, size_t row_stride, size_t frame_stride,
                        pbd_candidates** out) {
  return guarded([&] { DeviceGuard dg_(dev_of(d));
    REQUIRE(d && frames && out, "null argument");
    Engine& e = *d->e;
    e.set_frames_geometry(n, h, w, c);
    e.upload_and_pyramid(frames, row_stride, frame_stride);
    e.run_pdf(); e.run_dp_min(); e.run_argmin();
    auto cs = std::make_unique<pbd_candidates>();
    e.collect(cs->s);
    *out = cs.release();
  });
}
int pbd_detect_batch_u8_device(pbd_detector* d, const uint8_t* d_frames, int n, int h, int w, int c, pbd_candidates** out) {
  return guarded([&] { DeviceGuard dg_(dev_of(d));
    REQUIRE(d && d_frames && out, "null argument");
    Engine& e = *d->e;
    e.set_frames_geometry(n, h, w, c);
    e.use_device_frames(d_frames);
    run_all(e);
    auto cs = std::make_unique<pbd_candidates>();
    e.collect(cs->s);
    *out = cs.release();
  });
}
int pbd_enqueue_batch_u8_device(pbd_detector* d, const uint8_t* d_frames, int n, int h, int w, int c) {
  return guarded([&] { DeviceGuard dg_(dev_of(d));
    REQUIRE(d && d_frames, "null argument");
    d->e->enqueue_device(d_frames, n, h, w, c);
  });
}
int pbd_collect_candidates(pbd_detector* d, pbd_candidates** out) {
  return guarded([&] { DeviceGuard dg_(dev_of(d)); REQUIRE(d && out, "null argument"); auto cs = std::make_unique<pbd_candidates>(); d->e->collect(cs->s); *out = cs.release(); });
}

int pbd_submit_batch_u8(pbd_detector* d, const uint8_t* frames, int n, int h, int w, int c, int* ticket) {
  return guarded([&] { DeviceGuard dg_(dev_of(d)); REQUIRE(d && frames && ticket, "null argument"); *ticket = d->e->submit(frames, n, h, w, c); });
}
int pbd_collect_ticket(pbd_detector* d, int ticket, pbd_candidates** out) {
  return guarded([&] { DeviceGuard dg_(dev_of(d));
    REQUIRE(d && out, "null argument");
    auto cs = std::make_unique<pbd_candidates>();
    d->e->collect_ticket(ticket, cs->s);
    *out = cs.release();
  });
}

int pbd_candidates_count(const pbd_candidates* c) { return c ? c->s.n : 0; }
int pbd_candidates_nparts(const pbd_candidates* c, int i) {
  if (!c || i < 0 || i >= c->s.n) { g_err = "candidate index out of range"; return PBD_E_ARG; }
  return c->s.meta[(size_t)i * 4 + 3];
}
int pbd_candidates_get(const pbd_candidates* c, int i, int32_t* frame, int32_t* level, int32_t* component, float* score, int32_t* xs,
                       int32_t* ys, int32_t* ms, int32_t* rects_xywh) {
  return guarded([&] {
    REQUIRE(c && i >= 0 && i < c->s.n, "candidate index out of range");
    const CandidateSet& S = c->s;
    if (frame) *frame = S.meta[(size_t)i * 4];
    if (level) *level = S.meta[(size_t)i * 4 + 1];
    if (component) *component = S.meta[(size_t)i * 4 + 2];
    if (score) *score = S.score[i];
    const int np = S.meta[(size_t)i * 4 + 3];
    for (int p = 0; p < np; ++p) {
      const int* o = S.part(i, p);
      if (xs) xs[p] = o[0];
      if (ys) ys[p] = o[1];
      if (ms) ms[p] = o[2];
      if (rects_xywh) { rects_xywh[4 * p] = o[3]; rects_xywh[4 * p + 1] = o[4]; rects_xywh[4 * p + 2] = o[5]; rects_xywh[4 * p + 3] = o[6]; }
    }
  });
}
int pbd_candidates_export(const pbd_candidates* c, int32_t* meta4, float* scores, int32_t* parts7, int max_nparts) {
  return guarded([&] {
    REQUIRE(c && meta4 && scores && parts7 && max_nparts > 0, "null argument");
    const CandidateSet& S = c->s;
    if (S.n == 0) return;
    memcpy(meta4, S.meta.data(), (size_t)S.n * 4 * sizeof(int));
    memcpy(scores, S.score.data(), (size_t)S.n * sizeof(float));
    if (max_nparts == S.stride) { memcpy(parts7, S.parts.data(), S.parts.size() * sizeof(int)); return; }
    for (int i = 0; i < S.n; ++i) {
      const int np = S.meta[(size_t)i * 4 + 3];
      REQUIRE(np <= max_nparts, "max_nparts too small");
      memcpy(parts7 + (size_t)i * max_nparts * 7, S.part(i, 0), (size_t)np * 7 * sizeof(int));
    }
  });
}
void pbd_candidates_free(pbd_candidates* c) { delete c; }
int pbd_candidates_sort(pbd_candidates* c) {
  return guarded([&] {
    REQUIRE(c, "null argument");
    std::vector<int> order(c->s.n);
    for (int i = 0; i < c->s.n; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return c->s.score[a] > c->s.score[b]; });
    c->s.select(order);
  });
}

namespace {
struct IRect { int x, y, w, h; };
inline bool rect_empty(const IRect& r) { return r.w <= 0 || r.h <= 0; }
inline IRect rect_or(IRect a, const IRect& b) {            // cv::Rect operator|
  if (rect_empty(a)) return b;
  if (rect_empty(b)) return a;
  const int x1 = std::min(a.x, b.x), y1 = std::min(a.y, b.y);
  a.w = std::max(a.x + a.w, b.x + b.w) - x1; a.h = std::max(a.y + a.h, b.y + b.h) - y1; a.x = x1; a.y = y1;
  return a;
}
inline IRect rect_and(IRect a, const IRect& b) {           // cv::Rect operator&
  const int x1 = std::max(a.x, b.x), y1 = std::max(a.y, b.y);
  a.w = std::min(a.x + a.w, b.x + b.w) - x1; a.h = std::min(a.y + a.h, b.y + b.h) - y1; a.x = x1; a.y = y1;
  if (a.w <= 0 || a.h <= 0) a = IRect{0, 0, 0, 0};
  return a;
}
}  // namespace

int pbd_candidates_nms(pbd_candidates* c, int im_h, int im_w, float overlap) {
  return guarded([&] {
    REQUIRE(c && im_h > 0 && im_w > 0, "bad argument");
    const CandidateSet& S = c->s;
    // Candidate::nonMaximaSuppression works on one image's candidates in list order.  A batch's list may interleave frames (a global
    // score sort does), so the candidates are grouped by frame (stable: list order inside a frame is kept), the greedy painting runs
    // per frame on its own scratch image, and the survivors are returned in the original list order.
    std::vector<int> by_frame(S.n);
    for (int i = 0; i < S.n; ++i) by_frame[i] = i;
    std::stable_sort(by_frame.begin(), by_frame.end(), [&](int a, int b) { return S.meta[(size_t)a * 4] < S.meta[(size_t)b * 4]; });
    std::vector<char> keep_flag(S.n, 0);
    std::vector<uint8_t> scratch((size_t)im_h * im_w);
    const IRect bounds{0, 0, im_w, im_h};
    int cur_frame = 0;
    bool first = true;
    for (int k = 0; k < S.n; ++k) {
      const int i = by_frame[k];
      const int frame = S.meta[(size_t)i * 4], np = S.meta[(size_t)i * 4 + 3];
      if (first || frame != cur_frame) { std::fill(scratch.begin(), scratch.end(), 0); cur_frame = frame; first = false; }
      const int* p0 = S.part(i, 0);
      IRect hull{p0[3], p0[4], p0[5], p0[6]};                                        // Candidate::boundingBox, :104-110
      for (int p = 0; p < np; ++p) { const int* o = S.part(i, p); hull = rect_or(hull, IRect{o[3], o[4], o[5], o[6]}); }
      const IRect box = rect_and(hull, bounds);
      long long sum = 0;
      for (int y = box.y; y < box.y + box.h; ++y) { const uint8_t* r = &scratch[(size_t)y * im_w + box.x]; for (int x = 0; x < box.w; ++x) sum += r[x]; }
      const double ratio = (double)sum / (box.w * box.h);                             // boxsum[0] / box.area() (NaN for an empty box => kept)
      if (ratio > (double)overlap) continue;
      for (int y = box.y; y < box.y + box.h; ++y) memset(&scratch[(size_t)y * im_w + box.x], 1, (size_t)box.w);
      keep_flag[i] = 1;
    }
    std::vector<int> kept;
    kept.reserve(S.n);
    for (int i = 0; i < S.n; ++i) if (keep_flag[i]) kept.push_back(i);
    c->s.select(kept);
  });
}

// SearchSpacePruning<T>::filterCandidatesByDepth (reference src/SearchSpacePruning.cpp:73-95, the call commented out at
// src/PartsBasedDetector.cpp:91-93) with Math::median (include/Math.hpp:62-72).  Host code: candidates and depth image are host data.
int pbd_candidates_filter_by_depth(pbd_candidates* c, const pbd_model* m, const float* depth, int im_h, int im_w, size_t row_stride_bytes, float zfactor) {
  return guarded([&] {
    REQUIRE(c && m && depth && im_h > 0 && im_w > 0, "bad argument");
    if (row_stride_bytes == 0) row_stride_bytes = (size_t)im_w * sizeof(float);
    REQUIRE(row_stride_bytes >= (size_t)im_w * sizeof(float) && row_stride_bytes % sizeof(float) == 0, "bad depth row stride");
    const size_t pitch = row_stride_bytes / sizeof(float);
    const CandidateSet& S = c->s;
    const Model& M = m->m;
    std::vector<float> buf;
    auto median = [&](const int* o) -> float {                    // o = part record: rect at o[3..6]; the box is clipped to the image
      const int x0 = std::max(o[3], 0), y0 = std::max(o[4], 0), x1 = std::min(o[3] + o[5], im_w), y1 = std::min(o[4] + o[6], im_h);
      if (x1 <= x0 || y1 <= y0) return 0.f;
      buf.clear();
      for (int y = y0; y < y1; ++y) buf.insert(buf.end(), depth + (size_t)y * pitch + x0, depth + (size_t)y * pitch + x1);
      std::nth_element(buf.begin(), buf.begin() + buf.size() / 2, buf.end());
      return buf[buf.size() / 2];
    };
    std::vector<int> kept;
    for (int i = 0; i < S.n; ++i) {
      const int comp = S.meta[(size_t)i * 4 + 2], np = S.meta[(size_t)i * 4 + 3];
      REQUIRE(comp >= 0 && comp < M.ncomponents() && np == (int)M.comps[comp].size(), "candidate does not belong to this model");
      const auto& parts = M.comps[comp];
      bool keep = false;
      for (int p = np - 1; p >= 1; --p) {
        const float cm = median(S.part(i, p)), pm = median(S.part(i, parts[p].parentid));
        if (cm > 0 && pm > 0) {
          const int did = parts[p].defid[0];                       // part.anchor(0)
          const double ax = M.anchors[did * 2], ay = M.anchors[did * 2 + 1];
          if (std::abs(cm - pm) > std::sqrt(ax * ax + ay * ay) * zfactor) break;
        }
        if (p == 1) keep = true;
      }
      if (keep) kept.push_back(i);
    }
    c->s.select(kept);
  });
}

int pbd_candidates_create(int n, int max_nparts, const int32_t* meta4, const float* scores, const int32_t* parts7, pbd_candidates** out) {
  return guarded([&] {
    REQUIRE(n >= 0 && max_nparts > 0 && out && (n == 0 || (meta4 && scores && parts7)), "bad argument");
    auto cs = std::make_unique<pbd_candidates>();
    cs->s.resize(n, max_nparts);
    for (int i = 0; i < n; ++i) REQUIRE(meta4[4 * i + 3] > 0 && meta4[4 * i + 3] <= max_nparts, "bad part count");
    if (n) {
      memcpy(cs->s.meta.data(), meta4, (size_t)n * 4 * sizeof(int));
      memcpy(cs->s.score.data(), scores, (size_t)n * sizeof(float));
      memcpy(cs->s.parts.data(), parts7, (size_t)n * max_nparts * 7 * sizeof(int));
    }
    *out = cs.release();
  });
}

int pbd_stage_pyramid(pbd_detector* d, const uint8_t* frames, int n, int h, int w, int c, size_t row_stride, size_t frame_stride) {
  return guarded([&] { DeviceGuard dg_(dev_of(d));
    REQUIRE(d && frames, "null argument");
    Engine& e = *d->e;
    e.set_frames_geometry(n, h, w, c);
    e.upload_frames(frames, row_stride, frame_stride);
    e.run_pyramid();
  });
}
int pbd_stage_pdf(pbd_detector* d) { return guarded([&] { DeviceGuard dg_(dev_of(d)); REQUIRE(d, "null argument"); d->e->run_pdf(); }); }
int pbd_stage_dp_min(pbd_detector* d) { return guarded([&] { DeviceGuard dg_(dev_of(d)); REQUIRE(d, "null argument"); d->e->run_dp_min(); }); }
int pbd_stage_dp_argmin(pbd_detector* d, pbd_candidates** out) {
  return guarded([&] { DeviceGuard dg_(dev_of(d));
    REQUIRE(d && out, "null argument");
    d->e->run_argmin();
    auto cs = std::make_unique<pbd_candidates>();
    d->e->collect(cs->s);
    *out = cs.release();
  });
}

int pbd_pyramid_geometry(int h, int w, int sbin, int interval, int max_levels, int cap, int32_t* dims4, float* scales) {
  if (h <= 0 || w <= 0 || sbin <= 0 || interval <= 0 || cap < 0 || (cap > 0 && (!dims4 || !scales))) { g_err = "bad argument"; return PBD_E_ARG; }
  Geometry g{};
  const int n = compute_pyramid_levels(h, w, sbin, interval, max_levels, g);
  for (int l = 0; l < n && l < cap; ++l) {
    dims4[4 * l] = g.lv[l].img_h; dims4[4 * l + 1] = g.lv[l].img_w; dims4[4 * l + 2] = g.lv[l].oh; dims4[4 * l + 3] = g.lv[l].ow;
    scales[l] = g.lv[l].scale;
  }
  return n;
}

int pbd_num_frames(const pbd_detector* d) { return d ? d->e->geom().n_frames : 0; }
int pbd_num_levels(const pbd_detector* d) { return d ? d->e->geom().n_levels : 0; }
int pbd_level_info(const pbd_detector* d, int level, int32_t* img_h, int32_t* img_w, int32_t* oh, int32_t* ow, float* scale) {
  return guarded([&] { DeviceGuard dg_(dev_of(d));
    REQUIRE(d && level >= 0 && level < d->e->geom().n_levels, "level out of range");
    const LevelDesc& L = d->e->geom().lv[level];
    if (img_h) *img_h = L.img_h;
    if (img_w) *img_w = L.img_w;
    if (oh) *oh = L.oh;
    if (ow) *ow = L.ow;
    if (scale) *scale = L.scale;
  });
}

int pbd_get_pyramid_image(pbd_detector* d, int frame, int level, uint8_t* dst) { return guarded([&] { DeviceGuard dg_(dev_of(d)); REQUIRE(d && dst, "null argument"); d->e->get_pyramid_image(frame, level, dst); }); }
int pbd_get_features(pbd_detector* d, int frame, int level, float* dst) { return guarded([&] { DeviceGuard dg_(dev_of(d)); REQUIRE(d && dst, "null argument"); d->e->get_features(frame, level, dst); }); }
int pbd_get_response(pbd_detector* d, int frame, int level, int filter, float* dst) { return guarded([&] { DeviceGuard dg_(dev_of(d)); REQUIRE(d && dst, "null argument"); d->e->get_response(frame, level, filter, dst); }); }
int pbd_get_rootv(pbd_detector* d, int frame, int level, int component, float* dst) { return guarded([&] { DeviceGuard dg_(dev_of(d)); REQUIRE(d && dst, "null argument"); d->e->get_rootv(frame, level, component, dst); }); }
int pbd_get_rooti(pbd_detector* d, int frame, int level, int component, int32_t* dst) { return guarded([&] { DeviceGuard dg_(dev_of(d)); REQUIRE(d && dst, "null argument"); d->e->get_rooti(frame, level, component, dst); }); }
int pbd_get_backptr(pbd_detector* d, int frame, int level, int component, int part, int parent_mixture, int32_t* ix, int32_t* iy, int32_t* ik) {
  return guarded([&] { DeviceGuard dg_(dev_of(d)); REQUIRE(d && ix && iy && ik, "null argument"); d->e->get_backptr(frame, level, component, part, parent_mixture, ix, iy, ik); });
}
int pbd_set_levels(pbd_detector* d, int n_frames, int n_levels, const int32_t* ohow, const float* scales) {
  return guarded([&] { DeviceGuard dg_(dev_of(d)); REQUIRE(d && ohow && scales, "null argument"); d->e->set_levels_manual(n_frames, n_levels, ohow, scales); });
}
int pbd_set_features(pbd_detector* d, int frame, int level, const float* src) { return guarded([&] { DeviceGuard dg_(dev_of(d)); REQUIRE(d && src, "null argument"); d->e->set_features(frame, level, src); }); }
int pbd_set_response(pbd_detector* d, int frame, int level, int filter, const float* src) { return guarded([&] { DeviceGuard dg_(dev_of(d)); REQUIRE(d && src, "null argument"); d->e->set_response(frame, level, filter, src); }); }

static void cu(cudaError_t e, const char* what) { if (e != cudaSuccess) throw CudaError(std::string(what) + ": " + cudaGetErrorString(e)); }

// ---- standalone 2-D distance transform: a plan owns every table and scratch buffer, so that a run allocates nothing ----
}  // extern "C"
namespace {
struct DevBuf {                                 // RAII device allocation: error paths free what was allocated
  void* p = nullptr;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { if (p) cudaFree(p); }
  void alloc(size_t bytes) { cu(cudaMalloc(&p, bytes ? bytes : 1), "cudaMalloc"); }
  template <typename T> T* as() const { return static_cast<T*>(p); }
};
}  // namespace
extern "C" {
struct pbd_dt2d_plan {
  int n_maps = 0, h = 0, w = 0, impl = 0, device = 0;
  LineGeom lg[2];
  DevBuf geom, maps, etab, tmp, ixr, iyr, wp, ctr;   // wp / ctr: window parameters and the replay counter of impl 4
  DevBuf segctr; int seg_steps = 0;                  // impl 4: per-line verdict counters of the segmented walk (pbd_dt2d_plan_set_segment)
};

int pbd_dt2d_plan_create(int n_maps, int h, int w, const float* defw4, const int32_t* anchor_xy, int impl, pbd_dt2d_plan** out) {
  return guarded([&] {
    REQUIRE(defw4 && anchor_xy && out, "null argument");
    REQUIRE(n_maps > 0 && n_maps <= (1 << 20) && h > 0 && w > 0 && h <= 4096 && w <= 4096, "map shape out of range (<= 4096 x 4096, <= 2^20 maps)");
    REQUIRE(impl >= 0 && impl <= 4, "impl must be 0 (default), 1 (streaming, eager emission), 2 (parallel-in-q), 3 (streaming, lagged-scan emission) or 4 (windowed certified evaluation with replay)");
    for (int i = 0; i < n_maps; ++i) {
      REQUIRE(defw4[4 * i] > 0.f && defw4[4 * i + 2] > 0.f, "quadratic weights must be > 0");
      REQUIRE(std::abs(anchor_xy[2 * i]) <= 4096 && std::abs(anchor_xy[2 * i + 1]) <= 4096, "anchor out of range");
    }
    auto P = std::make_unique<pbd_dt2d_plan>();
    cu(cudaGetDevice(&P->device), "cudaGetDevice");
    P->n_maps = n_maps; P->h = h; P->w = w;
    P->impl = impl ? impl : 1;                 // measured on B200: the streaming kernels are the faster generation
    REQUIRE(P->impl != 2 || std::max(h, w) <= kLinesMaxN, "impl 2 (parallel-in-q) handles lines of at most 1024 samples");
    const size_t cells = (size_t)n_maps * h * w;
    std::vector<PassMap> maps(2 * (size_t)n_maps);
    P->etab.alloc((size_t)n_maps * (dt_table_len(w) + dt_table_len(h)) * sizeof(double));
    for (int i = 0; i < n_maps; ++i) {
      PassMap R{}, Cc{};
      R.in_buf = 0; R.in_off = R.out_off = R.ptr_off = (unsigned long long)i * h * w;
      R.w_sq = defw4[4 * i]; R.w_lin = defw4[4 * i + 1]; R.os = anchor_xy[2 * i];
      Cc = R; Cc.w_sq = defw4[4 * i + 2]; Cc.w_lin = defw4[4 * i + 3]; Cc.os = anchor_xy[2 * i + 1];
      R.tab_len = dt_table_len(w); R.tab_bias = dt_table_bias(w, R.os);
      Cc.tab_len = dt_table_len(h); Cc.tab_bias = dt_table_bias(h, Cc.os);
      R.etab = P->etab.as<double>() + (size_t)i * dt_table_len(w);
      Cc.etab = P->etab.as<double>() + (size_t)n_maps * dt_table_len(w) + (size_t)i * dt_table_len(h);
      maps[i] = R; maps[n_maps + i] = Cc;
    }
    P->maps.alloc(maps.size() * sizeof(PassMap));
    P->tmp.alloc(cells * sizeof(float)); P->ixr.alloc(cells * 2); P->iyr.alloc(cells * 2);
    if (P->impl == 4) {                        // the detector's default transform (dt_pass_win): certificate parameters per map and direction
      std::vector<dtw::WinParams> wp(2 * (size_t)n_maps);
      for (int i = 0; i < n_maps; ++i) {
        wp[i] = dtw::make_params(maps[i].w_sq, maps[i].w_lin, maps[i].os, w, kDtWindowW);
        wp[n_maps + i] = dtw::make_params(maps[n_maps + i].w_sq, maps[n_maps + i].w_lin, maps[n_maps + i].os, h, kDtWindowW);
      }
      P->wp.alloc(wp.size() * sizeof(dtw::WinParams));
      cu(cudaMemcpy(P->wp.p, wp.data(), wp.size() * sizeof(dtw::WinParams), cudaMemcpyHostToDevice), "H2D");
      P->ctr.alloc(sizeof(int));
      cu(cudaMemset(P->ctr.p, 0, sizeof(int)), "memset");
    }
    cu(cudaMemcpy(P->maps.p, maps.data(), maps.size() * sizeof(PassMap), cudaMemcpyHostToDevice), "H2D");
    if (P->impl != 2) {
      PassGeom pgs[2];
      memset(pgs, 0, sizeof(pgs));
      pgs[0].n_levels = pgs[1].n_levels = 1;
      pgs[0].nlines[0] = h; pgs[0].N[0] = w; pgs[1].nlines[0] = w; pgs[1].N[0] = h;
      P->geom.alloc(sizeof(pgs));
      cu(cudaMemcpy(P->geom.p, pgs, sizeof(pgs), cudaMemcpyHostToDevice), "H2D");
    } else {
      memset(P->lg, 0, sizeof(P->lg));
      P->lg[0].n_levels = P->lg[1].n_levels = 1;
      P->lg[0].nlines[0] = h; P->lg[0].N[0] = w; P->lg[1].nlines[0] = w; P->lg[1].N[0] = h;
      int budget = 13312;
      if (const char* v = getenv("PBD_DT_REGION")) budget = std::min(56 * 1024, std::max(256, atoi(v)));
      plan_line_geom(P->lg[0], budget, false); plan_line_geom(P->lg[1], budget, true);
      REQUIRE(4 * std::max(P->lg[0].region_bytes, P->lg[1].region_bytes) <= 227 * 1024, "line too long for the parallel-in-q kernels");
      P->geom.alloc(sizeof(P->lg));
      cu(cudaMemcpy(P->geom.p, P->lg, sizeof(P->lg), cudaMemcpyHostToDevice), "H2D");
    }
    launch_dt_tables(P->maps.as<PassMap>(), 2 * n_maps, nullptr);
    cu(cudaGetLastError(), "dt table launch");
    cu(cudaDeviceSynchronize(), "dt table sync");
    *out = P.release();
  });
}
void pbd_dt2d_plan_destroy(pbd_dt2d_plan* p) {
  if (!p) return;
  try { DeviceGuard dg_(p->device); delete p; } catch (...) { delete p; }   // the plan's buffers live on the device it was created on
}
int pbd_dt2d_plan_impl(const pbd_dt2d_plan* p) { return p ? p->impl : 0; }
// impl 4: cut every line into segments of `steps` walk steps (a multiple of 16; 0 = one lane per line, the default of a plan) -- the
// form the detector uses by itself for launches that cannot fill the GPU; results are identical
int pbd_dt2d_plan_set_segment(pbd_dt2d_plan* p, int steps) {
  return guarded([&] {
    REQUIRE(p, "null argument");
    REQUIRE(p->impl == 4, "segments exist for impl 4 (windowed certified evaluation) only");
    REQUIRE(steps == 0 || (steps >= 16 && steps <= 4096 && steps % 16 == 0), "steps must be 0 or a multiple of 16");
    DeviceGuard dg_(p->device);
    if (steps && !p->segctr.p) {
      const size_t n = (size_t)p->n_maps * std::max(p->h, p->w);
      p->segctr.alloc(n * sizeof(int));
      cu(cudaMemset(p->segctr.p, 0, n * sizeof(int)), "memset");
    }
    p->seg_steps = steps;
  });
}
// impl 4: lines replayed with the stack algorithm since the last call (synchronises the device); other impls: 0
long long pbd_dt2d_plan_replayed(pbd_dt2d_plan* p) {
  if (!p || p->impl != 4) return 0;
  int v = 0;
  try {
    DeviceGuard dg_(p->device);
    if (cudaDeviceSynchronize() != cudaSuccess || cudaMemcpy(&v, p->ctr.p, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    cudaMemset(p->ctr.p, 0, sizeof(int));
  } catch (...) { return -1; }
  return v;
}

// enqueue only (no allocation, no synchronisation): rows pass, columns pass, back-pointer composition
int pbd_dt2d_plan_run(pbd_dt2d_plan* p, void* stream, const float* d_in, float* d_out, uint16_t* d_ix, uint16_t* d_iy, int backptr_mode) {
  return guarded([&] {
    REQUIRE(p && d_in && d_out && d_ix && d_iy, "null argument");
    REQUIRE(backptr_mode == 0 || backptr_mode == 1, "backptr_mode must be 0 or 1");
    DeviceGuard dg_(p->device);
    cudaStream_t s = (cudaStream_t)stream;
    if (p->impl != 2) {
      launch_dt2d_standalone(d_in, p->n_maps, p->h, p->w, p->geom.as<PassGeom>(), p->maps.as<PassMap>(), p->tmp.as<float>(), d_out, d_ix, d_iy,
                             p->ixr.as<uint16_t>(), p->iyr.as<uint16_t>(), backptr_mode, s, p->impl == 4 ? 3 : (p->impl == 3 ? 1 : 0),
                             p->impl == 4 ? p->wp.as<dtw::WinParams>() : nullptr, p->impl == 4 ? p->ctr.as<int>() : nullptr,
                             p->seg_steps, p->seg_steps ? p->segctr.as<int>() : nullptr);
    } else {
      launch_dt2d_lines(d_in, p->n_maps, p->h, p->w, p->lg[0], p->lg[1], p->geom.as<LineGeom>(), p->maps.as<PassMap>(), p->tmp.as<float>(), d_out,
                        d_ix, d_iy, p->ixr.as<uint16_t>(), p->iyr.as<uint16_t>(), backptr_mode, s);
    }
    cu(cudaGetLastError(), "dt2d launch");
  });
}

int pbd_dt2d_f32_device(void* stream, const float* d_in, int n_maps, int h, int w, const float* h_defw4, const int32_t* h_anchor_xy,
                        float* d_out, uint16_t* d_ix, uint16_t* d_iy, int backptr_mode) {
  pbd_dt2d_plan* plan = nullptr;
  int impl = 0;
  if (const char* v = getenv("PBD_DT_IMPL")) impl = !strcmp(v, "stream") ? 1 : (!strcmp(v, "lines") ? 2 : (!strcmp(v, "scan") ? 3 : 0));
  int rc = pbd_dt2d_plan_create(n_maps, h, w, h_defw4, h_anchor_xy, impl, &plan);
  if (rc != PBD_OK) return rc;
  rc = pbd_dt2d_plan_run(plan, stream, d_in, d_out, d_ix, d_iy, backptr_mode);
  if (rc == PBD_OK) rc = guarded([&] { cu(cudaStreamSynchronize((cudaStream_t)stream), "dt2d sync"); });
  else cudaStreamSynchronize((cudaStream_t)stream);
  pbd_dt2d_plan_destroy(plan);
  return rc;
}
int pbd_dt2d_f32(const float* in, int n_maps, int h, int w, const float* defw4, const int32_t* anchor_xy, float* out, int32_t* ix,
                 int32_t* iy, int backptr_mode) {
  return guarded([&] {
    REQUIRE(in && out && ix && iy && defw4 && anchor_xy, "null argument");
    REQUIRE(n_maps > 0 && n_maps <= (1 << 20) && h > 0 && w > 0 && h <= 4096 && w <= 4096, "map shape out of range (<= 4096 x 4096, <= 2^20 maps)");
    const size_t cells = (size_t)n_maps * h * w;
    DevBuf d_in, d_out, d_ix, d_iy;
    d_in.alloc(cells * 4); d_out.alloc(cells * 4); d_ix.alloc(cells * 2); d_iy.alloc(cells * 2);
    cu(cudaMemcpy(d_in.p, in, cells * 4, cudaMemcpyHostToDevice), "H2D");
    const int rc = pbd_dt2d_f32_device(nullptr, d_in.as<float>(), n_maps, h, w, defw4, anchor_xy, d_out.as<float>(), d_ix.as<uint16_t>(), d_iy.as<uint16_t>(),
                                       backptr_mode);
    if (rc != PBD_OK) throw CudaError(g_err);
    std::vector<uint16_t> hx(cells), hy(cells);
    cu(cudaMemcpy(out, d_out.p, cells * 4, cudaMemcpyDeviceToHost), "D2H");
    cu(cudaMemcpy(hx.data(), d_ix.p, cells * 2, cudaMemcpyDeviceToHost), "D2H"); cu(cudaMemcpy(hy.data(), d_iy.p, cells * 2, cudaMemcpyDeviceToHost), "D2H");
    for (size_t i = 0; i < cells; ++i) { ix[i] = hx[i]; iy[i] = hy[i]; }
  });
}

// ---- ingest: image containers, sensor_msgs/Image encodings, pinned frame buffers (ingest.cpp) ----
int pbd_image_info(const uint8_t* bytes, size_t n, int32_t* h, int32_t* w, int32_t* channels, int32_t* bits) {
  return guarded([&] { REQUIRE(bytes, "null argument"); int hh, ww, cc, bb; image_info(bytes, n, &hh, &ww, &cc, &bb); if (h) *h = hh; if (w) *w = ww; if (channels) *channels = cc < 0 ? -cc : cc; if (bits) *bits = bb; });
}
int pbd_image_decode_bgr8(const uint8_t* bytes, size_t n, uint8_t* dst, size_t dst_capacity, int32_t* h, int32_t* w) {
  return guarded([&] { REQUIRE(bytes && dst, "null argument"); int hh, ww; image_decode_bgr8(bytes, n, dst, dst_capacity, &hh, &ww); if (h) *h = hh; if (w) *w = ww; });
}
int pbd_image_decode_depth_f32(const uint8_t* bytes, size_t n, float scale, float* dst, size_t dst_capacity, int32_t* h, int32_t* w) {
  return guarded([&] { REQUIRE(bytes && dst, "null argument"); int hh, ww; image_decode_depth_f32(bytes, n, scale, dst, dst_capacity, &hh, &ww); if (h) *h = hh; if (w) *w = ww; });
}
int pbd_imread_bgr8(const char* path, uint8_t* dst, size_t dst_capacity, int32_t* h, int32_t* w) {
  return guarded([&] {
    REQUIRE(path && (dst || dst_capacity == 0), "null argument");
    const std::vector<uint8_t> buf = slurp_bytes(path);
    int hh, ww;
    if (dst_capacity == 0) { int cc, bb; image_info(buf.data(), buf.size(), &hh, &ww, &cc, &bb); }     // size query
    else image_decode_bgr8(buf.data(), buf.size(), dst, dst_capacity, &hh, &ww);
    if (h) *h = hh;
    if (w) *w = ww;
  });
}
int pbd_ros_image_to_bgr8(const char* encoding, int h, int w, size_t step, int is_bigendian, const uint8_t* data, uint8_t* dst_bgr8) {
  return guarded([&] { REQUIRE(encoding && data && dst_bgr8 && h > 0 && w > 0, "bad argument"); ros_image_to_bgr8(encoding, h, w, step, is_bigendian, data, dst_bgr8); });
}
int pbd_ros_depth_to_f32(const char* encoding, int h, int w, size_t step, int is_bigendian, const uint8_t* data, float* dst) {
  return guarded([&] { REQUIRE(encoding && data && dst && h > 0 && w > 0, "bad argument"); ros_depth_to_f32(encoding, h, w, step, is_bigendian, data, dst); });
}
int pbd_host_alloc_pinned(size_t bytes, void** out) { return guarded([&] { REQUIRE(out, "null argument"); *out = pinned_alloc(bytes); }); }
void pbd_host_free_pinned(void* p) { pinned_free(p); }

long long pbd_launch_count(const pbd_detector* d) { return d ? d->e->launches() : 0; }
int pbd_stage_times_ms(pbd_detector* d, float ms[6]) { return guarded([&] { DeviceGuard dg_(dev_of(d)); REQUIRE(d && ms, "null argument"); d->e->stage_times(ms); }); }
int pbd_kernel_times_ms(pbd_detector* d, float ms[6]) { return guarded([&] { DeviceGuard dg_(dev_of(d)); REQUIRE(d && ms, "null argument"); d->e->kernel_times(ms); }); }
size_t pbd_device_bytes(const pbd_detector* d) { return d ? d->e->device_bytes() : 0; }

}  // extern "C"
