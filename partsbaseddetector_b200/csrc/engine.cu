// engine.cu -- see engine.hpp.
#include "engine.hpp"
#include "dt_window.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace pbd {

namespace {
inline int cv_round_f(float v) { return (int)lrintf(v); }          // cvRound: round half to even
inline int cv_floor_f(float v) { int i = (int)v; return i - (i > v); }
inline short sat_short(float v) { int i = cv_round_f(v); return (short)std::min(32767, std::max(-32768, i)); }
}  // namespace

// Level table of HOGFeatures<T>::pyramid (reference src/HOGFeatures.cpp:95-127, include/HOGFeatures.hpp:74-81)
// and the per-level HOG sizes of features() (:174-176).  log/pow/floor resolve to the float overloads in the
// reference (<cmath> + using namespace std), pow(float,int) promotes to double.
int compute_pyramid_levels(int h, int w, int sbin, int interval, int max_levels, Geometry& g) {
  const float sfactor = powf(2.0f, 1.0f / (float)interval);
  const float fw = (float)w, fh = (float)h;
  const float ns = 1 + floorf(logf(fminf(fh, fw) / (5.0f * (float)sbin)) / logf(sfactor));
  int nscales = ns > 0 ? (int)ns : 0;
  if (nscales > kMaxLevels) nscales = kMaxLevels;
  std::vector<LevelDesc> lv(nscales);
  for (auto& L : lv) { memset(&L, 0, sizeof(L)); }
  for (int i = 0; i < interval && i < nscales; ++i) {
    const float s = (float)(1.0f / pow((double)sfactor, i));
    int cw = cv_round_f(fw * s), ch = cv_round_f(fh * s);
    lv[i].img_w = cw; lv[i].img_h = ch; lv[i].scale = (float)(pow((double)sfactor, i) * sbin); lv[i].src_level = -1;
    for (int j = i + interval; j < nscales; j += interval) {
      cw = (cw + 1) / 2; ch = (ch + 1) / 2;
      lv[j].img_w = cw; lv[j].img_h = ch; lv[j].scale = 2 * lv[j - interval].scale; lv[j].src_level = j - interval;
    }
  }
  if (max_levels > 0 && nscales > max_levels) nscales = max_levels;
  g.n_levels = nscales;
  for (int l = 0; l < nscales; ++l) {
    LevelDesc L = lv[l];
    L.bw = (int)roundf((float)L.img_w / (float)sbin);
    L.bh = (int)roundf((float)L.img_h / (float)sbin);
    L.ow = std::max(L.bw - 2, 0);
    L.oh = std::max(L.bh - 2, 0);
    g.lv[l] = L;
  }
  return nscales;
}

void Engine::check_cuda(cudaError_t e, const char* what) const {
  if (e != cudaSuccess) throw CudaError(std::string(what) + ": " + cudaGetErrorString(e));
}

template <typename T>
void Engine::ensure(T*& p, size_t& cap, size_t n) {
  if (n <= cap && p) return;
  if (p) { check_cuda(cudaFree(p), "cudaFree"); dev_bytes_ -= cap * sizeof(T); p = nullptr; cap = 0; }
  if (n == 0) n = 1;
  check_cuda(cudaMalloc(&p, n * sizeof(T)), "cudaMalloc");
  cap = n;
  dev_bytes_ += n * sizeof(T);
}

Engine::Engine(const Model& m, int device, cudaStream_t stream) : thresh((double)m.thresh), model_(m), device_(device), stream_(stream) {
  model_.validate();
  if (model_.flen != 32 || model_.norient != 18) throw UnsupportedError("only flen=32 / norient=18 HOG models are supported");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0) throw CudaError(std::string("no CUDA device available: ") + cudaGetErrorString(e));
  if (device < 0 || device >= ndev) throw ArgError("device index out of range");
  // the caller's current device is restored on every exit; a constructor that throws frees what it had allocated (the destructor
  // does not run for a partially constructed object)
  int prev = -1;
  if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
  check_cuda(cudaSetDevice(device), "cudaSetDevice");
  try {
    check_cuda(cudaDeviceGetAttribute(&num_sms_, cudaDevAttrMultiProcessorCount, device), "cudaDeviceGetAttribute");
    build_tables();
    check_cuda(cudaMalloc(&d_g_, sizeof(Geometry)), "cudaMalloc geometry");
    check_cuda(cudaMalloc(&d_orient_lut_, hog_orient_lut_bytes()), "cudaMalloc orientation table");
    launches_ += launch_hog_orient_lut(d_orient_lut_, stream_);
    check_cuda(cudaGetLastError(), "orientation table launch");
    dev_bytes_ += hog_orient_lut_bytes();
    for (ResultSlot& S : slots_) {
      check_cuda(cudaMalloc(&S.d_nhits, sizeof(int)), "cudaMalloc nhits");
      check_cuda(cudaEventCreateWithFlags(&S.done, cudaEventDisableTiming), "cudaEventCreate");
    }
    for (auto& e : frames_free_ev_) check_cuda(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate");
    dev_bytes_ += sizeof(Geometry) + 2 * sizeof(int);
    for (int i = 0; i < 7; ++i) { check_cuda(cudaEventCreate(&ev_[i]), "cudaEventCreate"); }
  } catch (...) {
    release();
    if (prev >= 0 && prev != device) cudaSetDevice(prev);
    throw;
  }
  if (prev >= 0 && prev != device) cudaSetDevice(prev);
}

Engine::~Engine() {
  int prev = -1;
  if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
  cudaSetDevice(device_);
  release();
  if (prev >= 0 && prev != device_) cudaSetDevice(prev);
}

// frees every device allocation, stream and event of the engine (the detector's device is current)
void Engine::release() {
  cudaStreamSynchronize(stream_);
  void* ptrs[] = {d_wtc_, d_wtc16_, d_f16_, d_fhi_, d_flo_, d_tc_levels_, d_tc_tiles_, d_wpacked_, d_wgeneric_, d_foff_, d_fkh_, d_fkw_, d_jobs_, d_roots_, d_parent_, d_nparts_, d_cm_slot_, d_pm_slot_,
                  d_g_, d_frames_own_, b_.pyr, b_.hist, b_.norm, b_.feat, b_.resp, b_.work, b_.tmp, b_.val, b_.ixdt, b_.iyraw, b_.ik,
                  b_.rootv, b_.rooti, d_xofs_, d_yofs_, d_xalpha_, d_ybeta_, d_tile_level_, d_tile_first_,
                  d_scratch_i_, d_rootkeep_, d_orient_lut_, d_pg_, d_maps_rows_, d_maps_cols_, d_etab_, d_wp_rows_, d_wp_cols_, d_dtw_ctr_, d_seg_ctr_, d_frames_alt_, slots_[0].d_hits, slots_[0].d_nhits, slots_[0].d_xym,
                  slots_[1].d_hits, slots_[1].d_nhits, slots_[1].d_xym, d_ksize_, nms_.boxes, nms_.keys, nms_.skeys, nms_.sidx, nms_.kept_idx,
                  nms_.frame_count, nms_.fill, nms_.kept_count, nms_.out_off, nms_.seg_off, nms_.scratch, slots_[0].d_hits_out,
                  slots_[0].d_xym_out, slots_[0].d_total, slots_[1].d_hits_out, slots_[1].d_xym_out, slots_[1].d_total};
  for (void* p : ptrs) if (p) cudaFree(p);
  for (ResultSlot& S : slots_) if (S.done) cudaEventDestroy(S.done);
  for (auto& e : frames_free_ev_) if (e) cudaEventDestroy(e);
  if (graph_exec_) { cudaGraphExecDestroy(graph_exec_); graph_exec_ = nullptr; }
  if (capture_stream_) { cudaStreamDestroy(capture_stream_); capture_stream_ = nullptr; }
  if (d2h_stream_) cudaStreamDestroy(d2h_stream_);
  for (cudaStream_t st : dp_aux_) cudaStreamDestroy(st);
  for (cudaEvent_t ev : dp_join_) cudaEventDestroy(ev);
  if (dp_fork_) cudaEventDestroy(dp_fork_);
  if (copy_stream_) { cudaStreamDestroy(copy_stream_); for (auto& e : copy_ev_) cudaEventDestroy(e); cudaEventDestroy(main_ev_); }
  for (int i = 0; i < 7; ++i) if (ev_[i]) cudaEventDestroy(ev_[i]);
  for (cudaEvent_t e : kev_) cudaEventDestroy(e);
}

// timing == 2: records an event that closes the interval of kernel `tag` (-1 opens the first interval)
void Engine::kmark(int tag) {
  if (timing < 2) return;
  if (kev_n_ == kev_.size()) {
    cudaEvent_t e;
    check_cuda(cudaEventCreate(&e), "cudaEventCreate");
    kev_.push_back(e); kev_tag_.push_back(-1);
  }
  kev_tag_[kev_n_] = tag;
  check_cuda(cudaEventRecord(kev_[kev_n_++], stream_), "event");
}

void Engine::kernel_times(float ms[kKernelTimes]) {
  for (int i = 0; i < kKernelTimes; ++i) ms[i] = 0.f;
  check_cuda(cudaStreamSynchronize(stream_), "sync");
  for (size_t i = 1; i < kev_n_; ++i) {
    const int tag = kev_tag_[i];
    if (tag < 0 || tag >= kKernelTimes) continue;
    float t = 0.f;
    check_cuda(cudaEventElapsedTime(&t, kev_[i - 1], kev_[i]), "event elapsed");
    ms[tag] += t;
  }
}

// Model -> device tables: filters (convertTo float, reference src/PartsBasedDetector.cpp:115-117), the DP job
// list (reference Parts/ComponentPart indirection, include/Parts.hpp:103-189) and the backtrack tables.
void Engine::build_tables() {
  const Model& m = model_;
  const int nf = m.nfilters();
  // ---- filters ----
  fb_.nfilters = nf; fb_.flen = m.flen;
  fb_.uniform = 1; fb_.kh = m.frows[0]; fb_.kw = m.fkw[0]; fb_.khm = 0; fb_.kwm = 0;
  for (int i = 0; i < nf; ++i) {
    if (m.frows[i] != fb_.kh || m.fkw[i] != fb_.kw) fb_.uniform = 0;
    fb_.khm = std::max(fb_.khm, m.frows[i]); fb_.kwm = std::max(fb_.kwm, m.fkw[i]);
  }
  fb_.ngroups = (nf + 7) / 8;
  // generic layout [f][c][ky][kx]
  std::vector<int> foff(nf), fkh(nf), fkw(nf);
  size_t tot = 0;
  for (int i = 0; i < nf; ++i) { foff[i] = (int)tot; fkh[i] = m.frows[i]; fkw[i] = m.fkw[i]; tot += (size_t)m.frows[i] * m.fkw[i] * m.flen; }
  std::vector<float> wg(tot);
  for (int i = 0; i < nf; ++i) {
    const int kh = fkh[i], kw = fkw[i];
    for (int ky = 0; ky < kh; ++ky) for (int kx = 0; kx < kw; ++kx) for (int c = 0; c < m.flen; ++c)
      wg[foff[i] + ((size_t)c * kh + ky) * kw + kx] = (float)m.filters[i][((size_t)ky * kw + kx) * m.flen + c];
  }
  check_cuda(cudaMalloc(&d_wgeneric_, std::max<size_t>(tot, 1) * sizeof(float)), "cudaMalloc filters");
  check_cuda(cudaMemcpy(d_wgeneric_, wg.data(), tot * sizeof(float), cudaMemcpyHostToDevice), "upload filters");
  check_cuda(cudaMalloc(&d_foff_, nf * sizeof(int)), "cudaMalloc"); check_cuda(cudaMalloc(&d_fkh_, nf * sizeof(int)), "cudaMalloc");
  check_cuda(cudaMalloc(&d_fkw_, nf * sizeof(int)), "cudaMalloc");
  check_cuda(cudaMemcpy(d_foff_, foff.data(), nf * sizeof(int), cudaMemcpyHostToDevice), "upload");
  check_cuda(cudaMemcpy(d_fkh_, fkh.data(), nf * sizeof(int), cudaMemcpyHostToDevice), "upload");
  check_cuda(cudaMemcpy(d_fkw_, fkw.data(), nf * sizeof(int), cudaMemcpyHostToDevice), "upload");
  dev_bytes_ += tot * sizeof(float) + 3 * nf * sizeof(int);
  fb_.wg = d_wgeneric_; fb_.foff = d_foff_; fb_.fkh = d_fkh_; fb_.fkw = d_fkw_;
  if (fb_.uniform) {   // packed layout [group][c][ky][kx][8]
    const int kh = fb_.kh, kw = fb_.kw, taps = kh * kw;
    std::vector<float> wp((size_t)fb_.ngroups * m.flen * taps * 8, 0.f);
    for (int i = 0; i < nf; ++i) {
      const int gq = i / 8, j = i % 8;
      for (int c = 0; c < m.flen; ++c) for (int t = 0; t < taps; ++t)
        wp[(((size_t)gq * m.flen + c) * taps + t) * 8 + j] = (float)m.filters[i][(size_t)t * m.flen + c];
    }
    check_cuda(cudaMalloc(&d_wpacked_, wp.size() * sizeof(float)), "cudaMalloc packed filters");
    check_cuda(cudaMemcpy(d_wpacked_, wp.data(), wp.size() * sizeof(float), cudaMemcpyHostToDevice), "upload packed filters");
    dev_bytes_ += wp.size() * sizeof(float);
    fb_.w = d_wpacked_;
  }
  if (response_tc_supported(fb_)) {   // tf32 hi/lo slabs of the tensor-core path
    std::vector<std::vector<float>> ff(nf);
    for (int i = 0; i < nf; ++i) ff[i].assign(m.filters[i].begin(), m.filters[i].end());
    std::vector<float> wt;
    response_tc_pack_weights(ff, fb_.kh * fb_.kw, wt);
    check_cuda(cudaMalloc(&d_wtc_, wt.size() * sizeof(float)), "cudaMalloc tensor filters");
    check_cuda(cudaMemcpy(d_wtc_, wt.data(), wt.size() * sizeof(float), cudaMemcpyHostToDevice), "upload tensor filters");
    dev_bytes_ += wt.size() * sizeof(float);
    std::vector<uint16_t> w16;
    response_tc_pack_weights_f16(ff, fb_.kh * fb_.kw, w16, wtc16_scales_);
    check_cuda(cudaMalloc(&d_wtc16_, w16.size() * sizeof(uint16_t)), "cudaMalloc tensor filters (fp16)");
    check_cuda(cudaMemcpy(d_wtc16_, w16.data(), w16.size() * sizeof(uint16_t), cudaMemcpyHostToDevice), "upload tensor filters (fp16)");
    dev_bytes_ += w16.size() * sizeof(uint16_t);
  }
  // ---- DP slots ----
  const int ncomp = m.ncomponents();
  h_parent_.assign((size_t)ncomp * kMaxParts, -1);
  h_nparts_.assign(ncomp, 0);
  h_ksize_.assign((size_t)ncomp * kMaxParts * kMaxMix, 0);
  h_cm_slot_.assign((size_t)ncomp * kMaxParts * kMaxMix, 0);
  h_pm_slot_.assign((size_t)ncomp * kMaxParts * kMaxMix, 0);
  std::vector<std::vector<std::vector<int>>> work_slot(ncomp);   // [c][p][m] or -1 (leaf)
  nwork_ = ncm_ = npm_ = 0; max_parts_ = 0;
  for (int c = 0; c < ncomp; ++c) {
    const auto& parts = m.comps[c];
    const int np = (int)parts.size();
    if (np > kMaxParts) throw UnsupportedError("component has more parts than kMaxParts");
    max_parts_ = std::max(max_parts_, np);
    h_nparts_[c] = np;
    for (int p = 0; p < np; ++p)
      for (size_t mx = 0; mx < parts[p].filterid.size() && mx < (size_t)kMaxMix; ++mx)
        h_ksize_[((size_t)c * kMaxParts + p) * kMaxMix + mx] = m.frows[parts[p].filterid[mx]];   // xsize == ysize == rows (Parts.hpp:185-187)
    std::vector<char> has_child(np, 0);
    for (int p = 1; p < np; ++p) has_child[parts[p].parentid] = 1;
    work_slot[c].resize(np);
    for (int p = 0; p < np; ++p) {
      const int nm = (int)parts[p].filterid.size();
      if (nm > kMaxMix) throw UnsupportedError("part has more mixtures than kMaxMix");
      h_parent_[(size_t)c * kMaxParts + p] = parts[p].parentid;
      work_slot[c][p].assign(nm, -1);
      if (has_child[p]) for (int mi = 0; mi < nm; ++mi) work_slot[c][p][mi] = nwork_++;
      if (p > 0) {
        for (int mi = 0; mi < nm; ++mi) h_cm_slot_[((size_t)c * kMaxParts + p) * kMaxMix + mi] = ncm_++;
        const int pn = (int)parts[parts[p].parentid].filterid.size();
        for (int pm = 0; pm < pn; ++pm) h_pm_slot_[((size_t)c * kMaxParts + p) * kMaxMix + pm] = npm_++;
      }
    }
  }
  // ---- jobs by wave ----
  // A part can run once all its children have run; siblings must deliver their messages to the parent in
  // decreasing part index (the reference's float accumulation order, src/DynamicProgram.cpp:95,155-156), so a
  // part also waits for its next-higher sibling.  wave(p) = max(max_child wave + 1, wave(next higher sibling) + 1).
  jobs_.clear(); wave_first_.clear(); wave_count_.clear(); wave_maxmix_.clear();
  tmp_maps_ = 0;
  std::vector<std::vector<int>> wave_of(ncomp);
  int nwaves = 0;
  for (int c = 0; c < ncomp; ++c) {
    const auto& parts = m.comps[c];
    const int np = (int)parts.size();
    wave_of[c].assign(np, 0);
    for (int p = np - 1; p >= 1; --p) {                 // children have larger indices than their parents
      int t = 0;
      for (int q = p + 1; q < np; ++q) {
        if (parts[q].parentid == p) t = std::max(t, wave_of[c][q] + 1);                    // my children
        if (parts[q].parentid == parts[p].parentid) t = std::max(t, wave_of[c][q] + 1);    // higher siblings first
      }
      wave_of[c][p] = t;
      nwaves = std::max(nwaves, t + 1);
    }
  }
  for (int k = 0; k < nwaves; ++k) {
    wave_first_.push_back((int)jobs_.size());
    int cnt = 0, mx = 1;
    for (int c = 0; c < ncomp; ++c) {
      const auto& parts = m.comps[c];
      const int np = (int)parts.size();
      for (int p = np - 1; p >= 1; --p) {
      if (wave_of[c][p] != k) continue;
      const Part& P = parts[p];
      const Part& Par = parts[P.parentid];
      PartJob J;
      memset(&J, 0, sizeof(J));
      J.nmix = (int)P.filterid.size(); J.pnmix = (int)Par.filterid.size();
      mx = std::max(mx, std::max(J.nmix, 1));
      bool first_touch = true;                       // is p the highest-index child of its parent?
      for (int q = p + 1; q < np; ++q) if (parts[q].parentid == P.parentid) first_touch = false;
      J.first_touch = first_touch ? 1 : 0;
      for (int mm = 0; mm < J.nmix; ++mm) {
        const int ws = work_slot[c][p][mm];
        J.in_is_work[mm] = ws >= 0; J.in_slot[mm] = ws >= 0 ? ws : P.filterid[mm];
        const int did = P.defid[mm];
        for (int t = 0; t < 4; ++t) J.w[mm][t] = m.defs[(size_t)did * 4 + t];
        J.ax[mm] = m.anchors[did * 2]; J.ay[mm] = m.anchors[did * 2 + 1];
        J.cm_slot[mm] = h_cm_slot_[((size_t)c * kMaxParts + p) * kMaxMix + mm];
        for (int pm = 0; pm < J.pnmix; ++pm) J.bias[mm][pm] = m.biasw[P.biasid[mm] + pm];   // T4: flat indexing
      }
      for (int pm = 0; pm < J.pnmix; ++pm) {
        J.out_work_slot[pm] = work_slot[c][P.parentid][pm];
        J.out_resp_fid[pm] = Par.filterid[pm];
        J.pm_slot[pm] = h_pm_slot_[((size_t)c * kMaxParts + p) * kMaxMix + pm];
      }
      J.tmp_base = cnt * kMaxMix;
      jobs_.push_back(J);
      ++cnt;
      }
    }
    wave_count_.push_back(cnt);
    wave_maxmix_.push_back(mx);
    tmp_maps_ = std::max(tmp_maps_, cnt * kMaxMix);
  }
  roots_.clear();
  for (int c = 0; c < ncomp; ++c) {
    const Part& R = m.comps[c][0];
    RootJob J;
    memset(&J, 0, sizeof(J));
    J.nmix = (int)R.filterid.size();
    for (int mi = 0; mi < J.nmix; ++mi) { const int ws = work_slot[c][0][mi]; J.in_is_work[mi] = ws >= 0; J.in_slot[mi] = ws >= 0 ? ws : R.filterid[mi]; }
    J.bias = m.biasw[R.biasid[0]];
    roots_.push_back(J);
  }
  auto up = [&](auto*& d, const auto& h) {
    using T = std::remove_reference_t<decltype(h[0])>;
    const size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(T);
    check_cuda(cudaMalloc(&d, bytes), "cudaMalloc table");
    if (!h.empty()) check_cuda(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice), "upload table");
    dev_bytes_ += bytes;
  };
  up(d_jobs_, jobs_); up(d_roots_, roots_); up(d_parent_, h_parent_); up(d_nparts_, h_nparts_); up(d_ksize_, h_ksize_); up(d_cm_slot_, h_cm_slot_); up(d_pm_slot_, h_pm_slot_);
}

void Engine::need(int stage, const char* who) const {
  if (stage_ < stage) throw StateError(std::string(who) + ": previous stage has not been run for the current batch");
}

void Engine::set_frames_geometry(int n, int h, int w, int c) {
  if (n <= 0 || h <= 0 || w <= 0) throw ArgError("bad frame batch shape");
  if (c != 1 && c != 3) throw UnsupportedError("frames must have 1 or 3 channels (reference src/HOGFeatures.cpp:171)");
  if (have_images_ && g_.n_frames == n && g_.in_h == h && g_.in_w == w && g_.in_c == c && stage_ >= 1) {
    const int nl = g_.n_levels;
    Geometry probe{};
    compute_pyramid_levels(h, w, model_.sbin, model_.interval, max_levels, probe);
    if (probe.n_levels == nl) { stage_ = 1; return; }          // same geometry: keep buffers and tables
  }
  Geometry g{};
  g.n_frames = n; g.in_h = h; g.in_w = w; g.in_c = c;
  compute_pyramid_levels(h, w, model_.sbin, model_.interval, max_levels, g);
  if (g.n_levels <= 0) throw ArgError("image too small for one pyramid level (min(h,w) < 5*sbin)");
  stage_ = 0;                                   // from here on the old batch is gone: a failure below must not leave stale stages usable
  long long img_off = 0; int block_off = 0, cell_off = 0, xo = 0, yo = 0;
  for (int l = 0; l < g.n_levels; ++l) {
    LevelDesc& L = g.lv[l];
    if (L.img_h < 3 || L.img_w < 3) throw ArgError("pyramid level smaller than 3 pixels");
    L.identity = (L.src_level < 0 && L.img_w == w && L.img_h == h) ? 1 : 0;
    L.img_off = img_off;
    if (!L.identity) img_off += (long long)L.img_w * L.img_h * c;
    img_off = (img_off + 15) / 16 * 16;
    L.block_off = block_off; block_off += L.bw * L.bh;
    L.cell_off = cell_off; cell_off += L.ow * L.oh;
    if (L.src_level < 0) { L.xofs_off = xo; L.yofs_off = yo; xo += L.img_w; yo += L.img_h; }
  }
  g.img_bytes = img_off; g.blocks_total = block_off; g.cells_total = cell_off;
  g_ = g;
  have_images_ = true;
  // resize coefficient tables, built exactly as cv::resize() does for INTER_LINEAR / 8U (fixed point, 11 bits)
  std::vector<int> xofs(std::max(xo, 1)), yofs(std::max(yo, 1));
  std::vector<short> xalpha((size_t)std::max(xo, 1) * 2), ybeta((size_t)std::max(yo, 1) * 2);
  for (int l = 0; l < g.n_levels; ++l) {
    const LevelDesc& L = g.lv[l];
    if (L.src_level >= 0) continue;
    const double scale_x = 1. / ((double)L.img_w / w), scale_y = 1. / ((double)L.img_h / h);
    for (int dx = 0; dx < L.img_w; ++dx) {
      float fx = (float)((dx + 0.5) * scale_x - 0.5);
      int sx = cv_floor_f(fx);
      fx -= sx;
      if (sx < 0) { fx = 0; sx = 0; }
      if (sx >= w - 1) { fx = 0; sx = w - 1; }
      xofs[L.xofs_off + dx] = sx;
      xalpha[2 * (size_t)(L.xofs_off + dx)] = sat_short((1.f - fx) * 2048.f);
      xalpha[2 * (size_t)(L.xofs_off + dx) + 1] = sat_short(fx * 2048.f);
    }
    for (int dy = 0; dy < L.img_h; ++dy) {
      float fy = (float)((dy + 0.5) * scale_y - 0.5);
      int sy = cv_floor_f(fy);
      fy -= sy;
      yofs[L.yofs_off + dy] = sy;
      ybeta[2 * (size_t)(L.yofs_off + dy)] = sat_short((1.f - fy) * 2048.f);
      ybeta[2 * (size_t)(L.yofs_off + dy) + 1] = sat_short(fy * 2048.f);
    }
  }
  ensure(d_xofs_, cap_xofs_, xofs.size()); ensure(d_yofs_, cap_yofs_, yofs.size());
  ensure(d_xalpha_, cap_xalpha_, xalpha.size()); ensure(d_ybeta_, cap_ybeta_, ybeta.size());
  check_cuda(cudaMemcpyAsync(d_xofs_, xofs.data(), xofs.size() * sizeof(int), cudaMemcpyHostToDevice, stream_), "upload xofs");
  check_cuda(cudaMemcpyAsync(d_yofs_, yofs.data(), yofs.size() * sizeof(int), cudaMemcpyHostToDevice, stream_), "upload yofs");
  check_cuda(cudaMemcpyAsync(d_xalpha_, xalpha.data(), xalpha.size() * sizeof(short), cudaMemcpyHostToDevice, stream_), "upload xalpha");
  check_cuda(cudaMemcpyAsync(d_ybeta_, ybeta.data(), ybeta.size() * sizeof(short), cudaMemcpyHostToDevice, stream_), "upload ybeta");
  check_cuda(cudaStreamSynchronize(stream_), "sync tables");     // host vectors go out of scope
  build_batch_tables();
  alloc_batch();
  stage_ = 1;
}

void Engine::set_levels_manual(int n, int nlevels, const int32_t* ohow, const float* scales) {
  if (n <= 0 || nlevels <= 0 || nlevels > kMaxLevels) throw ArgError("bad level table");
  Geometry g{};
  g.n_frames = n; g.n_levels = nlevels; g.in_c = 3;
  int block_off = 0, cell_off = 0;
  for (int l = 0; l < nlevels; ++l) {
    LevelDesc& L = g.lv[l];
    L.oh = ohow[2 * l]; L.ow = ohow[2 * l + 1];
    if (L.oh <= 0 || L.ow <= 0) throw ArgError("level with no cells");
    L.bh = L.oh + 2; L.bw = L.ow + 2; L.scale = scales[l]; L.src_level = -1;
    L.block_off = block_off; block_off += L.bw * L.bh;
    L.cell_off = cell_off; cell_off += L.ow * L.oh;
  }
  g.blocks_total = block_off; g.cells_total = cell_off; g.img_bytes = 0;
  stage_ = 0;
  g_ = g;
  have_images_ = false;
  build_batch_tables();
  alloc_batch();
  stage_ = 1;
}

void Engine::build_batch_tables() {
  const Geometry& g = g_;
  ++geom_serial_;
  max_ow_ = max_oh_ = 0;
  for (int l = 0; l < g.n_levels; ++l) { max_ow_ = std::max(max_ow_, g.lv[l].ow); max_oh_ = std::max(max_oh_, g.lv[l].oh); }
  if (max_ow_ > kMaxDim || max_oh_ > kMaxDim) throw UnsupportedError("pyramid level larger than 1024 cells in one dimension");
  std::vector<int> tl, tf;
  ntiles0_ = response_plan_tiles(g, fb_, tl, tf);
  ntiles_ = (int)tl.size();
  auto up = [&](int*& d, size_t& cap, const std::vector<int>& h) {
    ensure(d, cap, h.size());
    if (!h.empty()) check_cuda(cudaMemcpyAsync(d, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice, stream_), "upload batch table");
  };
  up(d_tile_level_, cap_tile_level_, tl); up(d_tile_first_, cap_tile_first_, tf);
  check_cuda(cudaMemcpyAsync(d_g_, &g_, sizeof(Geometry), cudaMemcpyHostToDevice, stream_), "upload geometry");
  // separable-transform passes: geometry (rows: lines = rows of length ow; cols: lines = columns of length oh) and the
  // per-wave map tables (offsets depend on cells_total)
  memset(&pg_rows_, 0, sizeof(pg_rows_)); memset(&pg_cols_, 0, sizeof(pg_cols_));
  pg_rows_.n_levels = pg_cols_.n_levels = g.n_levels;
  for (int l = 0; l < g.n_levels; ++l) {
    pg_rows_.nlines[l] = g.lv[l].oh; pg_rows_.N[l] = g.lv[l].ow; pg_rows_.cell_off[l] = g.lv[l].cell_off;
    pg_cols_.nlines[l] = g.lv[l].ow; pg_cols_.N[l] = g.lv[l].oh; pg_cols_.cell_off[l] = g.lv[l].cell_off;
  }
  const PassGeom pgs[2] = {pg_rows_, pg_cols_};
  if (!d_pg_) { check_cuda(cudaMalloc(&d_pg_, 2 * sizeof(PassGeom)), "cudaMalloc pass geometry"); dev_bytes_ += 2 * sizeof(PassGeom); }
  check_cuda(cudaMemcpyAsync(d_pg_, pgs, sizeof(pgs), cudaMemcpyHostToDevice, stream_), "upload pass geometry");
  std::vector<PassMap> mr, mc;
  std::vector<dtw::WinParams> wr, wc;                           // dt_variant 3: per-map parameters of the windowed certified evaluation
  wave_map_first_.assign(wave_first_.size(), 0); wave_map_count_.assign(wave_first_.size(), 0);
  const unsigned long long ct = (unsigned long long)g.cells_total;
  for (size_t wv = 0; wv < wave_first_.size(); ++wv) {
    wave_map_first_[wv] = (int)mr.size();
    for (int j = 0; j < wave_count_[wv]; ++j) {
      const PartJob& J = jobs_[wave_first_[wv] + j];
      for (int mm = 0; mm < J.nmix; ++mm) {
        PassMap R{}, C{};
        R.in_buf = J.in_is_work[mm]; R.in_off = (unsigned long long)J.in_slot[mm] * ct;
        R.out_off = (unsigned long long)(J.tmp_base + mm) * ct; R.ptr_off = (unsigned long long)J.cm_slot[mm] * ct;
        R.w_sq = J.w[mm][0]; R.w_lin = J.w[mm][1]; R.os = J.ax[mm];
        C.in_buf = 0; C.in_off = R.out_off; C.out_off = R.out_off; C.ptr_off = R.ptr_off;
        C.w_sq = J.w[mm][2]; C.w_lin = J.w[mm][3]; C.os = J.ay[mm];
        R.tab_len = dt_table_len(max_ow_); R.tab_bias = dt_table_bias(max_ow_, R.os);
        C.tab_len = dt_table_len(max_oh_); C.tab_bias = dt_table_bias(max_oh_, C.os);
        mr.push_back(R); mc.push_back(C);
        // a map / direction the window cannot serve (anchor beyond +-W, a >= 0) has ok = 0: its lines go straight to the stack algorithm
        wr.push_back(dtw::make_params(R.w_sq, R.w_lin, R.os, std::max(max_ow_, 1), kDtWindowW));
        wc.push_back(dtw::make_params(C.w_sq, C.w_lin, C.os, std::max(max_oh_, 1), kDtWindowW));
      }
    }
    wave_map_count_[wv] = (int)mr.size() - wave_map_first_[wv];
  }
  ensure(d_wp_rows_, cap_wp_rows_, wr.size()); ensure(d_wp_cols_, cap_wp_cols_, wc.size());
  if (!wr.empty()) {
    check_cuda(cudaMemcpyAsync(d_wp_rows_, wr.data(), wr.size() * sizeof(dtw::WinParams), cudaMemcpyHostToDevice, stream_), "upload window parameters");
    check_cuda(cudaMemcpyAsync(d_wp_cols_, wc.data(), wc.size() * sizeof(dtw::WinParams), cudaMemcpyHostToDevice, stream_), "upload window parameters");
  }
  if (!d_seg_ctr_) {                                              // segmented walk: per-line verdict counters, one region per frame group (self-cleaning)
    const size_t n = (size_t)kSegGroups * seg_lines_per_group();
    check_cuda(cudaMalloc(&d_seg_ctr_, n * sizeof(int)), "cudaMalloc segment counters"); dev_bytes_ += n * sizeof(int);
    check_cuda(cudaMemsetAsync(d_seg_ctr_, 0, n * sizeof(int), stream_), "clear segment counters");
  }
  if (!d_dtw_ctr_) {
    check_cuda(cudaMalloc(&d_dtw_ctr_, 64 * sizeof(int)), "cudaMalloc replay counters"); dev_bytes_ += 64 * sizeof(int);
    check_cuda(cudaMemsetAsync(d_dtw_ctr_, 0, 64 * sizeof(int), stream_), "clear replay counters");
  }
  ensure(d_maps_rows_, cap_maps_rows_, mr.size()); ensure(d_maps_cols_, cap_maps_cols_, mc.size());
  ensure(d_etab_, cap_etab_, mr.size() * (size_t)(dt_table_len(max_ow_) + dt_table_len(max_oh_)));
  for (size_t i = 0; i < mr.size(); ++i) {
    mr[i].etab = d_etab_ + i * (size_t)dt_table_len(max_ow_);
    mc[i].etab = d_etab_ + mr.size() * (size_t)dt_table_len(max_ow_) + i * (size_t)dt_table_len(max_oh_);
  }
  if (!mr.empty()) {
    check_cuda(cudaMemcpyAsync(d_maps_rows_, mr.data(), mr.size() * sizeof(PassMap), cudaMemcpyHostToDevice, stream_), "upload pass maps");
    check_cuda(cudaMemcpyAsync(d_maps_cols_, mc.data(), mc.size() * sizeof(PassMap), cudaMemcpyHostToDevice, stream_), "upload pass maps");
    launch_dt_tables(d_maps_rows_, (int)mr.size(), stream_);
    launch_dt_tables(d_maps_cols_, (int)mc.size(), stream_);
    check_cuda(cudaGetLastError(), "DT table launch");
  }
  check_cuda(cudaStreamSynchronize(stream_), "sync batch tables");
}

void Engine::alloc_batch() {
  const Geometry& g = g_;
  const size_t n = g.n_frames, ct = g.cells_total, bt = g.blocks_total;
  const int ncomp = model_.ncomponents();
  if (have_images_) {
    ensure(b_.pyr, cap_pyr_, n * (size_t)std::max<long long>(g.img_bytes, 16));
    ensure(b_.hist, cap_hist_, n * bt * 18);
    ensure(b_.norm, cap_norm_, n * bt);
  }
  ensure(b_.feat, cap_feat_, n * ct * 32);
  ensure(b_.resp, cap_resp_, n * ct * model_.nfilters());
  ensure(b_.work, cap_work_, n * ct * std::max(nwork_, 1));
  ensure(b_.tmp, cap_tmp_, n * ct * std::max(tmp_maps_, 1));
  ensure(b_.val, cap_val_, n * ct * std::max(tmp_maps_, 1));
  ensure(b_.ixdt, cap_ixdt_, n * ct * std::max(ncm_, 1));
  ensure(b_.iyraw, cap_iyraw_, n * ct * std::max(ncm_, 1));
  ensure(b_.ik, cap_ik_, n * ct * std::max(npm_, 1));
  ensure(b_.rootv, cap_rootv_, n * ct * ncomp);
  ensure(b_.rooti, cap_rooti_, n * ct * ncomp);
}

void Engine::ensure_slot(ResultSlot& S) {
  ensure(S.d_hits, S.cap_hits, (size_t)max_candidates);
  ensure(S.d_xym, S.cap_xym, (size_t)max_candidates * 3 * std::max(max_parts_, 1));
  S.max_candidates = max_candidates;
}

void Engine::ensure_nms(ResultSlot& S) {
  const size_t mh = (size_t)max_candidates, nf = (size_t)g_.n_frames;
  ensure(nms_.boxes, cap_nms_boxes_, mh); ensure(nms_.keys, cap_nms_keys_, mh);
  ensure(nms_.skeys, cap_nms_skeys_, 2 * mh); ensure(nms_.sidx, cap_nms_sidx_, 2 * mh); ensure(nms_.kept_idx, cap_nms_kept_, 2 * mh);
  ensure(nms_.frame_count, cap_nms_fc_, nf); ensure(nms_.fill, cap_nms_fill_, nf); ensure(nms_.kept_count, cap_nms_kc_, nf);
  ensure(nms_.out_off, cap_nms_oo_, nf); ensure(nms_.seg_off, cap_nms_so_, nf + 1);
  ensure(nms_.scratch, cap_nms_scratch_, nms_scratch_words(g_));
  ensure(S.d_hits_out, S.cap_hits_out, mh);
  ensure(S.d_xym_out, S.cap_xym_out, mh * 3 * std::max(max_parts_, 1));
  if (!S.d_total) { check_cuda(cudaMalloc(&S.d_total, sizeof(int)), "cudaMalloc"); dev_bytes_ += sizeof(int); }
}

void Engine::upload_frames(const uint8_t* frames, size_t row_stride, size_t frame_stride) {
  need(1, "upload_frames");
  if (!have_images_) throw StateError("upload_frames: batch was defined by pbd_set_levels");
  const size_t row = (size_t)g_.in_w * g_.in_c, fb = row * g_.in_h;
  if (row_stride == 0) row_stride = row;
  if (frame_stride == 0) frame_stride = row_stride * g_.in_h;
  ensure(d_frames_own_, cap_frames_, fb * g_.n_frames);
  if (timing) { check_cuda(cudaEventRecord(ev_[0], stream_), "event"); ev_valid_[0] = true; }
  if (row_stride == row && frame_stride == fb) {
    check_cuda(cudaMemcpyAsync(d_frames_own_, frames, fb * g_.n_frames, cudaMemcpyHostToDevice, stream_), "H2D frames");
  } else {
    for (int i = 0; i < g_.n_frames; ++i)
      check_cuda(cudaMemcpy2DAsync(d_frames_own_ + fb * i, row, frames + frame_stride * i, row_stride, row, g_.in_h,
                                   cudaMemcpyHostToDevice, stream_), "H2D frames (strided)");
  }
  b_.frames = d_frames_own_;
}

// Host frames: the H2D copy is issued in chunks on a copy stream and the image pyramid + HOG of each chunk start as soon
// as its frames have landed, so the transfer of chunk i+1 overlaps the feature stage of chunk i.
void Engine::chunked_upload_pyramid(const uint8_t* frames, uint8_t* d_dst, cudaEvent_t wait_before_copy, cudaEvent_t record_after, bool pipeline_busy) {
  const size_t fb = (size_t)g_.in_w * g_.in_c * g_.in_h;
  const int n = g_.n_frames;
  b_.frames = d_dst;
  if (!copy_stream_) {
    check_cuda(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking), "cudaStreamCreate");
    for (auto& e : copy_ev_) check_cuda(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate");
    check_cuda(cudaEventCreateWithFlags(&main_ev_, cudaEventDisableTiming), "cudaEventCreate");
  }
  // the copy stream must not overwrite frames still being read by work already queued
  check_cuda(cudaStreamWaitEvent(copy_stream_, wait_before_copy, 0), "wait");
  // With another batch still computing, the whole copy hides behind that batch: one chunk, so that the pyramid / HOG kernels run over
  // the full batch (four quarter-size launches each cost 0.2 ms more per 64 frames); otherwise chunks let the feature stage start early.
  const int nchunks = (n >= 8 && !pipeline_busy) ? 4 : 1;
  for (int c = 0; c < nchunks; ++c) {
    const int f0 = (int)((long long)n * c / nchunks), f1 = (int)((long long)n * (c + 1) / nchunks);
    if (f1 <= f0) continue;
    check_cuda(cudaMemcpyAsync(d_dst + fb * f0, frames + fb * f0, fb * (f1 - f0), cudaMemcpyHostToDevice, copy_stream_), "H2D frames");
    check_cuda(cudaEventRecord(copy_ev_[c], copy_stream_), "event");
    check_cuda(cudaStreamWaitEvent(stream_, copy_ev_[c], 0), "wait");
    launches_ += launch_pyramid(g_, d_g_, b_, d_xofs_, d_xalpha_, d_yofs_, d_ybeta_, f0, f1 - f0, stream_);
    launches_ += launch_hog(g_, d_g_, b_, d_orient_lut_, model_.sbin, f0, f1 - f0, stream_);
  }
  if (record_after) check_cuda(cudaEventRecord(record_after, stream_), "event");   // frames consumed
  check_cuda(cudaGetLastError(), "pyramid/HOG launch");
  feat_from_hog_ = true;
  stage_ = 2;
}

void Engine::upload_and_pyramid(const uint8_t* frames, size_t row_stride, size_t frame_stride) {
  need(1, "upload_and_pyramid");
  if (!have_images_) throw StateError("upload_and_pyramid: batch was defined by pbd_set_levels");
  const size_t row = (size_t)g_.in_w * g_.in_c, fb = row * g_.in_h;
  if (row_stride == 0) row_stride = row;
  if (frame_stride == 0) frame_stride = row_stride * g_.in_h;
  if (timing || g_.n_frames < 8 || row_stride != row || frame_stride != fb) {       // per-stage timing wants separable stages
    upload_frames(frames, row_stride, frame_stride);
    run_pyramid();
    return;
  }
  ensure(d_frames_own_, cap_frames_, fb * g_.n_frames);
  if (!main_ev_) {
    check_cuda(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking), "cudaStreamCreate");
    for (auto& e : copy_ev_) check_cuda(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate");
    check_cuda(cudaEventCreateWithFlags(&main_ev_, cudaEventDisableTiming), "cudaEventCreate");
  }
  check_cuda(cudaEventRecord(main_ev_, stream_), "event");
  chunked_upload_pyramid(frames, d_frames_own_, main_ev_, nullptr);
}

// Pipelined API.  Frames must be tightly packed (and pinned for the copy to be asynchronous) and stay untouched until the
// ticket has been collected.  At most two batches are in flight: the ticket returned is the result slot (0/1).
int Engine::submit(const uint8_t* frames, int n, int h, int w, int c) {
  const int slot = cur_slot_ ^ 1;
  if (slots_[slot].pending) throw StateError("submit: two batches are already in flight; collect a ticket first");
  set_frames_geometry(n, h, w, c);
  cur_slot_ = slot;
  const size_t fb = (size_t)w * c * h;
  frames_buf_ ^= 1;
  uint8_t*& dst = frames_buf_ ? d_frames_alt_ : d_frames_own_;
  size_t& cap = frames_buf_ ? cap_frames_alt_ : cap_frames_;
  ensure(dst, cap, fb * n);
  // this frame buffer was last read by the pyramid stage of the batch before the previous one
  chunked_upload_pyramid(frames, dst, frames_free_ev_[frames_buf_], frames_free_ev_[frames_buf_], slots_[slot ^ 1].pending);
  run_pdf();
  run_dp_min();
  run_argmin();
  slots_[slot].pending = true;
  return slot;
}

void Engine::collect_ticket(int ticket, CandidateSet& out) {
  if (ticket < 0 || ticket > 1 || !slots_[ticket].pending) throw ArgError("collect_ticket: no batch in flight for this ticket");
  ResultSlot& S = slots_[ticket];
  if (!d2h_stream_) check_cuda(cudaStreamCreateWithFlags(&d2h_stream_, cudaStreamNonBlocking), "cudaStreamCreate");
  check_cuda(cudaStreamWaitEvent(d2h_stream_, S.done, 0), "wait");
  download_slot(S, d2h_stream_, out);
  S.pending = false;
}

void Engine::use_device_frames(const uint8_t* d_frames) {
  need(1, "use_device_frames");
  if (timing) { check_cuda(cudaEventRecord(ev_[0], stream_), "event"); ev_valid_[0] = true; }
  b_.frames = d_frames;
}

void Engine::run_pyramid() {
  need(1, "pyramid");
  if (!have_images_ || !b_.frames) throw StateError("pyramid: no frames");
  if (timing) { check_cuda(cudaEventRecord(ev_[1], stream_), "event"); ev_valid_[1] = true; }
  launches_ += launch_pyramid(g_, d_g_, b_, d_xofs_, d_xalpha_, d_yofs_, d_ybeta_, 0, g_.n_frames, stream_);
  if (timing) { check_cuda(cudaEventRecord(ev_[2], stream_), "event"); ev_valid_[2] = true; }
  launches_ += launch_hog(g_, d_g_, b_, d_orient_lut_, model_.sbin, 0, g_.n_frames, stream_);
  check_cuda(cudaGetLastError(), "pyramid/HOG launch");
  feat_from_hog_ = true;
  stage_ = 2;
}

// Strip layout, work list and border cells of the tensor-core path for the current batch geometry.
void Engine::ensure_tc(bool f16) {
  std::vector<TcLevel> lv;
  std::vector<TcTile> su;
  long long slack = 0;
  const long long frame_rows = response_tc_plan(g_, fb_.kh, fb_.kw, lv, su, &slack);
  const size_t rows = (size_t)frame_rows * g_.n_frames + (size_t)slack;
  if (f16) {
    const bool realloc16 = rows * 64 > cap_f16_ || !d_f16_;
    if (tc16_serial_ == geom_serial_ && !realloc16) return;
    ensure(d_f16_, cap_f16_, rows * 64);
    ensure(d_tc_levels_, cap_tc_levels_, lv.size()); ensure(d_tc_tiles_, cap_tc_tiles_, std::max<size_t>(su.size(), 1));
    check_cuda(cudaMemcpyAsync(d_tc_levels_, lv.data(), lv.size() * sizeof(TcLevel), cudaMemcpyHostToDevice, stream_), "upload strip levels");
    if (!su.empty()) check_cuda(cudaMemcpyAsync(d_tc_tiles_, su.data(), su.size() * sizeof(TcTile), cudaMemcpyHostToDevice, stream_), "upload work list");
    launches_ += launch_tc_border_init_f16(d_f16_, (long long)(cap_f16_ / 64), stream_);
    check_cuda(cudaStreamSynchronize(stream_), "sync strip tables");
    tc_ntiles_ = (int)su.size(); tc_frame_rows_ = frame_rows; tc16_serial_ = geom_serial_;
    return;
  }
  const bool realloc = rows * 32 > cap_fhi_ || !d_fhi_;
  if (tc_serial_ == geom_serial_ && !realloc) return;
  ensure(d_fhi_, cap_fhi_, rows * 32); ensure(d_flo_, cap_flo_, rows * 32);
  ensure(d_tc_levels_, cap_tc_levels_, lv.size()); ensure(d_tc_tiles_, cap_tc_tiles_, std::max<size_t>(su.size(), 1));
  check_cuda(cudaMemcpyAsync(d_tc_levels_, lv.data(), lv.size() * sizeof(TcLevel), cudaMemcpyHostToDevice, stream_), "upload strip levels");
  if (!su.empty()) check_cuda(cudaMemcpyAsync(d_tc_tiles_, su.data(), su.size() * sizeof(TcTile), cudaMemcpyHostToDevice, stream_), "upload work list");
  launches_ += launch_tc_border_init(d_fhi_, d_flo_, (long long)(cap_fhi_ / 32), stream_);
  check_cuda(cudaStreamSynchronize(stream_), "sync strip tables");       // host vectors go out of scope
  tc_ntiles_ = (int)su.size(); tc_frame_rows_ = frame_rows; tc_serial_ = geom_serial_;
}

void Engine::run_pdf() {
  need(2, "pdf");
  if ((resp_mode == 2 || resp_mode == 3) && response_tc_supported(fb_)) {
    const bool f16 = resp_mode == 3;
    ensure_tc(f16);
    if (timing) { check_cuda(cudaEventRecord(ev_[3], stream_), "event"); ev_valid_[3] = true; }
    kev_n_ = 0; kmark(-1);
    if (f16) launches_ += launch_feat_split_f16(g_, d_g_, d_tc_levels_, b_.feat, d_f16_, tc_frame_rows_, stream_);
    else launches_ += launch_feat_split(g_, d_g_, d_tc_levels_, b_.feat, d_fhi_, d_flo_, tc_frame_rows_, stream_);
    kmark(0);
    if (f16)
      launches_ += launch_response_tc(g_, b_, fb_, reinterpret_cast<const float*>(d_f16_), nullptr, reinterpret_cast<const float*>(d_wtc16_), d_tc_levels_,
                                      d_tc_tiles_, tc_ntiles_, tc_frame_rows_, num_sms_, tc_taps_per_partial, stream_, wtc16_scales_.data());
    else
      launches_ += launch_response_tc(g_, b_, fb_, d_fhi_, d_flo_, d_wtc_, d_tc_levels_, d_tc_tiles_, tc_ntiles_, tc_frame_rows_, num_sms_, tc_taps_per_partial, stream_);
    kmark(1);
    check_cuda(cudaGetLastError(), "tensor response launch");
    last_response_kernel = f16 ? 4 : 3;
    stage_ = 3;
    return;
  }
  if (timing) { check_cuda(cudaEventRecord(ev_[3], stream_), "event"); ev_valid_[3] = true; }
  kev_n_ = 0; kmark(-1);
  // a tensor mode the filter bank does not qualify for (filters of different sizes) falls back to the bit-exact FP32 kernel, never to
  // a third arithmetic: only response mode 1 asks for fused multiply-adds
  const int exact = resp_mode != 1 ? 1 : 0;
  launches_ += launch_response_tiles(g_, d_g_, b_, fb_, d_tile_level_, d_tile_first_, ntiles0_, ntiles_, exact, feat_from_hog_ ? 1 : 0, stream_);
  last_response_kernel = (response_has_fast_path(fb_) ? 1 : 0) + (exact ? 0 : 5);
  kmark(1);
  check_cuda(cudaGetLastError(), "response launch");
  stage_ = 3;
}

void Engine::run_dp_min() {
  need(3, "dp_min");
  if (timing) { check_cuda(cudaEventRecord(ev_[4], stream_), "event"); ev_valid_[4] = true; }
  const int nf = model_.nfilters();
  if (kev_n_ > 4096) kev_n_ = 0;          // dp_min re-run many times without a pdf stage in between
  kmark(-1);
  auto mark = [](void* self, int tag) { static_cast<Engine*>(self)->kmark(tag); };
  // Frames are independent, so the batch is cut into `dp_streams` groups of frames whose waves run on separate streams: the
  // tail of every launch (the last CTAs of the long level-0 lines) overlaps the other group's kernels instead of leaving SMs
  // idle.  Per-kernel timing (timing == 2) keeps everything on one stream so that its intervals stay meaningful.
  const int nsplit = (timing >= 2 || g_.n_frames < 2 * kMinFramesPerDpGroup) ? 1 : std::max(1, std::min(dp_streams, g_.n_frames / kMinFramesPerDpGroup));
  const size_t ct = (size_t)g_.cells_total;
  if (nsplit > 1) {
    while ((int)dp_aux_.size() < nsplit - 1) {
      cudaStream_t st; cudaEvent_t ev;
      check_cuda(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking), "cudaStreamCreate");
      check_cuda(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming), "cudaEventCreate");
      dp_aux_.push_back(st); dp_join_.push_back(ev);
    }
    if (!dp_fork_) check_cuda(cudaEventCreateWithFlags(&dp_fork_, cudaEventDisableTiming), "cudaEventCreate");
    check_cuda(cudaEventRecord(dp_fork_, stream_), "event");
    for (int i = 0; i + 1 < nsplit; ++i) check_cuda(cudaStreamWaitEvent(dp_aux_[i], dp_fork_, 0), "wait");
  }
  for (size_t wv = 0; wv < wave_first_.size(); ++wv) {
    if (wave_count_[wv] == 0) continue;
    for (int sp = 0; sp < nsplit; ++sp) {
      const int f0 = (int)((long long)g_.n_frames * sp / nsplit), f1 = (int)((long long)g_.n_frames * (sp + 1) / nsplit);
      Geometry gs = g_;
      gs.n_frames = f1 - f0;
      DeviceBuffers bs = b_;
      bs.resp += (size_t)f0 * ct * nf; bs.work += (size_t)f0 * ct * nwork_; bs.tmp += (size_t)f0 * ct * tmp_maps_; bs.val += (size_t)f0 * ct * tmp_maps_;
      bs.ixdt += (size_t)f0 * ct * ncm_; bs.iyraw += (size_t)f0 * ct * ncm_; bs.ik += (size_t)f0 * ct * npm_;
      const int n = launch_dt_wave(gs, d_g_, bs, pg_rows_, d_pg_, pg_cols_, d_pg_ + 1, d_maps_rows_ + wave_map_first_[wv],
                                   d_maps_cols_ + wave_map_first_[wv], wave_map_count_[wv], max_ow_, max_oh_, d_jobs_ + wave_first_[wv],
                                   wave_count_[wv], nf, nwork_, ncm_, npm_, tmp_maps_, sp == 0 ? stream_ : dp_aux_[sp - 1],
                                   timing >= 2 ? +mark : nullptr, this, dt_scan, d_wp_rows_ + wave_map_first_[wv], d_wp_cols_ + wave_map_first_[wv],
                                   d_dtw_ctr_ + (sp & 63), num_sms_ * 20, dt_segment, sp < kSegGroups ? d_seg_ctr_ + (size_t)sp * seg_lines_per_group() : nullptr);
      launches_ += n;
    }
  }
  for (int i = 0; i + 1 < nsplit; ++i) {
    check_cuda(cudaEventRecord(dp_join_[i], dp_aux_[i]), "event");
    check_cuda(cudaStreamWaitEvent(stream_, dp_join_[i], 0), "wait");
  }
  // root scores (reference computes rootv/rooti at the end of min(), src/DynamicProgram.cpp:163-171)
  launches_ += launch_root(g_, d_g_, b_, d_roots_, model_.ncomponents(), nf, nwork_, stream_);
  kmark(5);
  check_cuda(cudaGetLastError(), "DP launch");
  stage_ = 4;
}

long long Engine::dt_replayed_lines() {
  if (!d_dtw_ctr_) return 0;
  int h[64];
  check_cuda(cudaStreamSynchronize(stream_), "sync");
  check_cuda(cudaMemcpy(h, d_dtw_ctr_, sizeof(h), cudaMemcpyDeviceToHost), "D2H replay counters");
  long long tot = 0;
  for (int i = 0; i < 64; ++i) tot += h[i];
  check_cuda(cudaMemset(d_dtw_ctr_, 0, sizeof(h)), "clear replay counters");
  return tot;
}

void Engine::run_argmin() {
  need(4, "argmin");
  if (timing) { check_cuda(cudaEventRecord(ev_[5], stream_), "event"); ev_valid_[5] = true; }
  ResultSlot& S = slots_[cur_slot_];
  ensure_slot(S);
  check_cuda(cudaMemsetAsync(S.d_nhits, 0, sizeof(int), stream_), "memset nhits");
  if (root_nms > 0) ensure(d_rootkeep_, cap_rootkeep_, (size_t)g_.n_frames * model_.ncomponents() * g_.cells_total);
  launches_ += launch_hits(g_, d_g_, b_, model_.ncomponents(), (float)thresh, S.d_hits, S.d_nhits, S.max_candidates, stream_, root_nms, d_rootkeep_);
  BacktrackTables t{d_parent_, d_nparts_, d_cm_slot_, d_pm_slot_};
  launches_ += launch_backtrack(g_, d_g_, b_, t, model_.ncomponents(), ncm_, npm_, S.d_hits, S.d_nhits, S.max_candidates, backptr, max_parts_,
                                S.d_xym, stream_);
  S.nms = nms_overlap >= 0.0;
  if (S.nms) {
    if (!have_images_) throw StateError("nms_overlap needs the frame size: not available for a batch defined by pbd_set_levels");
    if (model_.ncomponents() > 32) throw UnsupportedError("device NMS supports at most 32 components");
    ensure_nms(S);
    NmsBuffers nb = nms_;
    nb.hits_out = S.d_hits_out; nb.xym_out = S.d_xym_out; nb.total = S.d_total;
    launches_ += launch_device_nms(g_, d_g_, nb, S.d_hits, S.d_nhits, S.max_candidates, S.d_xym, max_parts_, d_nparts_, d_ksize_, (float)nms_overlap,
                                   stream_);
  }
  check_cuda(cudaGetLastError(), "argmin launch");
  S.scales.resize(g_.n_levels);
  for (int l = 0; l < g_.n_levels; ++l) S.scales[l] = g_.lv[l].scale;
  check_cuda(cudaEventRecord(S.done, stream_), "event");
  if (timing) { check_cuda(cudaEventRecord(ev_[6], stream_), "event"); ev_valid_[6] = true; }
  stage_ = 5;
}

void Engine::run_stages() {
  run_pyramid();
  run_pdf();
  run_dp_min();
  run_argmin();
}

// Device-resident frames through all stages.  With use_graph the launch sequence (about 100 kernels, the frame groups of the DP on
// their forked streams included) is captured once per (frames pointer, geometry, options) and replayed: the first call runs eagerly
// (it sizes every buffer and table), the second captures, every later call is one cudaGraphLaunch.
void Engine::enqueue_device(const uint8_t* d_frames, int n, int h, int w, int c) {
  set_frames_geometry(n, h, w, c);
  use_device_frames(d_frames);
  if (!use_graph || timing) { run_stages(); return; }
  GraphKey k;
  k.frames = d_frames; k.geom_serial = geom_serial_; k.n = n; k.resp_mode = resp_mode; k.backptr = backptr; k.max_candidates = max_candidates;
  k.dp_streams = dp_streams; k.thresh = thresh; k.nms_overlap = nms_overlap; k.root_nms = root_nms; k.dt_scan = dt_scan; k.dt_segment = dt_segment;
  if (graph_exec_ && k == graph_key_) {
    check_cuda(cudaGraphLaunch(graph_exec_, stream_), "cudaGraphLaunch");
    launches_ += graph_launches_;
    cur_slot_ = graph_slot_;
    stage_ = 5;
    return;
  }
  if (!(k == warm_key_)) { warm_key_ = k; run_stages(); return; }          // first sight of this configuration: eager, allocates
  if (graph_exec_) { cudaGraphExecDestroy(graph_exec_); graph_exec_ = nullptr; }
  const long long l0 = launches_;
  // Capture happens on a stream of our own (the caller's may be the legacy default stream, which cannot be captured): every stage
  // launches on stream_, so it is swapped for the duration of the capture; the graph is then launched into the caller's stream.
  if (!capture_stream_) check_cuda(cudaStreamCreateWithFlags(&capture_stream_, cudaStreamNonBlocking), "cudaStreamCreate");
  cudaStream_t user = stream_;
  cudaGraph_t g = nullptr;
  check_cuda(cudaStreamBeginCapture(capture_stream_, cudaStreamCaptureModeThreadLocal), "cudaStreamBeginCapture");
  stream_ = capture_stream_;
  try {
    run_stages();
  } catch (...) {
    stream_ = user;
    cudaStreamEndCapture(capture_stream_, &g);
    if (g) cudaGraphDestroy(g);
    throw;
  }
  stream_ = user;
  check_cuda(cudaStreamEndCapture(capture_stream_, &g), "cudaStreamEndCapture");
  const cudaError_t ie = cudaGraphInstantiate(&graph_exec_, g, 0);
  cudaGraphDestroy(g);
  check_cuda(ie, "cudaGraphInstantiate");
  graph_key_ = k;
  graph_launches_ = launches_ - l0;
  graph_slot_ = cur_slot_;
  check_cuda(cudaGraphLaunch(graph_exec_, stream_), "cudaGraphLaunch");
}

void Engine::collect(CandidateSet& out) {
  need(5, "collect");
  download_slot(slots_[cur_slot_], stream_, out);
}

// Downloads the hits of a slot on stream `st` (which must already be ordered after the slot's `done` event) and builds the
// candidate set in the reference's order.
void Engine::download_slot(ResultSlot& S, cudaStream_t st, CandidateSet& out) {
  int nh = 0;
  check_cuda(cudaMemcpyAsync(&nh, S.d_nhits, sizeof(int), cudaMemcpyDeviceToHost, st), "D2H nhits");
  check_cuda(cudaStreamSynchronize(st), "sync");
  if (nh > S.max_candidates)
    throw StateError("candidate buffer overflow: " + std::to_string(nh) + " hits > max_candidates=" + std::to_string(S.max_candidates));
  const int ps = max_parts_;
  if (S.nms) {                                       // only the survivors of the device NMS come back, already in their final order
    check_cuda(cudaMemcpyAsync(&nh, S.d_total, sizeof(int), cudaMemcpyDeviceToHost, st), "D2H kept count");
    check_cuda(cudaStreamSynchronize(st), "sync");
  }
  h_hits_.resize(nh);
  h_xym_.resize((size_t)nh * 3 * ps);
  if (nh) {
    check_cuda(cudaMemcpyAsync(h_hits_.data(), S.nms ? S.d_hits_out : S.d_hits, (size_t)nh * sizeof(Hit), cudaMemcpyDeviceToHost, st), "D2H hits");
    check_cuda(cudaMemcpyAsync(h_xym_.data(), S.nms ? S.d_xym_out : S.d_xym, h_xym_.size() * sizeof(int), cudaMemcpyDeviceToHost, st), "D2H parts");
    check_cuda(cudaStreamSynchronize(st), "sync");
  }
  // the reference's deterministic (single-threaded) order: frame, level, component, row-major hit
  std::vector<int> order(nh);
  for (int i = 0; i < nh; ++i) order[i] = i;
  const std::vector<Hit>& hits = h_hits_;
  if (!S.nms) std::sort(order.begin(), order.end(), [&](int a, int b) {
    const Hit &A = hits[a], &B = hits[b];
    if (A.frame != B.frame) return A.frame < B.frame;
    if (A.level != B.level) return A.level < B.level;
    if (A.comp != B.comp) return A.comp < B.comp;
    if (A.y != B.y) return A.y < B.y;
    return A.x < B.x;
  });
  out.resize(nh, ps);
  for (int oi = 0; oi < nh; ++oi) {
    const int i = order[oi];
    const Hit& H = hits[i];
    const auto& parts = model_.comps[H.comp];
    const int np = (int)parts.size();
    out.meta[(size_t)oi * 4] = H.frame; out.meta[(size_t)oi * 4 + 1] = H.level; out.meta[(size_t)oi * 4 + 2] = H.comp; out.meta[(size_t)oi * 4 + 3] = np;
    out.score[oi] = H.score;
    const int* xs = h_xym_.data() + (size_t)i * 3 * ps;
    const float scale = S.scales[H.level];
    for (int p = 0; p < np; ++p) {                  // reference src/DynamicProgram.cpp:238-244
      int* o = out.part(oi, p);
      const int x = xs[p], y = xs[ps + p], mix = xs[2 * ps + p];
      const int ks = model_.frows[parts[p].filterid[mix]];             // xsize == ysize == rows (Parts.hpp:185-187)
      const int x1 = cv_round_f((float)(x - 1) * scale), y1 = cv_round_f((float)(y - 1) * scale);
      const int sz = cv_round_f((float)ks * scale);
      const int x2 = x1 + sz - 1, y2 = y1 + sz - 1;
      const int rx = std::min(x1, x2), ry = std::min(y1, y2);
      o[0] = x; o[1] = y; o[2] = mix; o[3] = rx; o[4] = ry; o[5] = std::max(x1, x2) - rx; o[6] = std::max(y1, y2) - ry;
    }
  }
}

static void check_idx(bool ok, const char* what) { if (!ok) throw ArgError(std::string(what) + ": index out of range"); }

void Engine::get_pyramid_image(int frame, int level, uint8_t* dst) {
  need(2, "get_pyramid_image");
  if (!have_images_) throw StateError("no pyramid images in a manually defined batch");
  check_idx(frame >= 0 && frame < g_.n_frames && level >= 0 && level < g_.n_levels, "get_pyramid_image");
  const LevelDesc& L = g_.lv[level];
  const uint8_t* src = L.identity ? b_.frames + (size_t)frame * g_.in_h * g_.in_w * g_.in_c : b_.pyr + (size_t)frame * g_.img_bytes + L.img_off;
  check_cuda(cudaMemcpyAsync(dst, src, (size_t)L.img_w * L.img_h * g_.in_c, cudaMemcpyDeviceToHost, stream_), "D2H image");
  check_cuda(cudaStreamSynchronize(stream_), "sync");
}
void Engine::get_features(int frame, int level, float* dst) {
  need(2, "get_features");
  check_idx(frame >= 0 && frame < g_.n_frames && level >= 0 && level < g_.n_levels, "get_features");
  const LevelDesc& L = g_.lv[level];
  check_cuda(cudaMemcpyAsync(dst, b_.feat + ((size_t)frame * g_.cells_total + L.cell_off) * 32, (size_t)L.oh * L.ow * 32 * sizeof(float),
                             cudaMemcpyDeviceToHost, stream_), "D2H features");
  check_cuda(cudaStreamSynchronize(stream_), "sync");
}
void Engine::get_response(int frame, int level, int filter, float* dst) {
  need(3, "get_response");
  check_idx(frame >= 0 && frame < g_.n_frames && level >= 0 && level < g_.n_levels && filter >= 0 && filter < model_.nfilters(), "get_response");
  const LevelDesc& L = g_.lv[level];
  check_cuda(cudaMemcpyAsync(dst, b_.resp + ((size_t)frame * model_.nfilters() + filter) * g_.cells_total + L.cell_off,
                             (size_t)L.oh * L.ow * sizeof(float), cudaMemcpyDeviceToHost, stream_), "D2H response");
  check_cuda(cudaStreamSynchronize(stream_), "sync");
}
void Engine::get_rootv(int frame, int level, int comp, float* dst) {
  need(4, "get_rootv");
  check_idx(frame >= 0 && frame < g_.n_frames && level >= 0 && level < g_.n_levels && comp >= 0 && comp < model_.ncomponents(), "get_rootv");
  const LevelDesc& L = g_.lv[level];
  check_cuda(cudaMemcpyAsync(dst, b_.rootv + ((size_t)frame * model_.ncomponents() + comp) * g_.cells_total + L.cell_off,
                             (size_t)L.oh * L.ow * sizeof(float), cudaMemcpyDeviceToHost, stream_), "D2H rootv");
  check_cuda(cudaStreamSynchronize(stream_), "sync");
}
void Engine::get_rooti(int frame, int level, int comp, int32_t* dst) {
  need(4, "get_rooti");
  check_idx(frame >= 0 && frame < g_.n_frames && level >= 0 && level < g_.n_levels && comp >= 0 && comp < model_.ncomponents(), "get_rooti");
  const LevelDesc& L = g_.lv[level];
  std::vector<uint8_t> tmp((size_t)L.oh * L.ow);
  check_cuda(cudaMemcpyAsync(tmp.data(), b_.rooti + ((size_t)frame * model_.ncomponents() + comp) * g_.cells_total + L.cell_off, tmp.size(),
                             cudaMemcpyDeviceToHost, stream_), "D2H rooti");
  check_cuda(cudaStreamSynchronize(stream_), "sync");
  for (size_t i = 0; i < tmp.size(); ++i) dst[i] = tmp[i];
}
void Engine::get_backptr(int frame, int level, int comp, int part, int pm, int32_t* ix, int32_t* iy, int32_t* ik) {
  need(4, "get_backptr");
  check_idx(frame >= 0 && frame < g_.n_frames && level >= 0 && level < g_.n_levels && comp >= 0 && comp < model_.ncomponents(), "get_backptr");
  const auto& parts = model_.comps[comp];
  check_idx(part >= 1 && part < (int)parts.size(), "get_backptr(part)");
  check_idx(pm >= 0 && pm < (int)parts[parts[part].parentid].filterid.size(), "get_backptr(parent mixture)");
  const LevelDesc& L = g_.lv[level];
  const size_t n = (size_t)L.oh * L.ow;
  ensure(d_scratch_i_, cap_scratch_i_, 3 * n);
  launches_ += launch_expand_backptr(g_, b_, frame, level, ncm_, npm_, d_cm_slot_ + ((size_t)comp * kMaxParts + part) * kMaxMix,
                                     h_pm_slot_[((size_t)comp * kMaxParts + part) * kMaxMix + pm], backptr, d_scratch_i_, d_scratch_i_ + n,
                                     d_scratch_i_ + 2 * n, stream_);
  check_cuda(cudaMemcpyAsync(ix, d_scratch_i_, n * sizeof(int), cudaMemcpyDeviceToHost, stream_), "D2H ix");
  check_cuda(cudaMemcpyAsync(iy, d_scratch_i_ + n, n * sizeof(int), cudaMemcpyDeviceToHost, stream_), "D2H iy");
  check_cuda(cudaMemcpyAsync(ik, d_scratch_i_ + 2 * n, n * sizeof(int), cudaMemcpyDeviceToHost, stream_), "D2H ik");
  check_cuda(cudaStreamSynchronize(stream_), "sync");
}
void Engine::set_features(int frame, int level, const float* src) {
  need(1, "set_features");
  check_idx(frame >= 0 && frame < g_.n_frames && level >= 0 && level < g_.n_levels, "set_features");
  const LevelDesc& L = g_.lv[level];
  check_cuda(cudaMemcpyAsync(b_.feat + ((size_t)frame * g_.cells_total + L.cell_off) * 32, src, (size_t)L.oh * L.ow * 32 * sizeof(float),
                             cudaMemcpyHostToDevice, stream_), "H2D features");
  check_cuda(cudaStreamSynchronize(stream_), "sync");
  feat_from_hog_ = false;                         // injected features: channel 31 is not known to be zero
  stage_ = std::max(stage_, 2);
}
void Engine::set_response(int frame, int level, int filter, const float* src) {
  need(1, "set_response");
  check_idx(frame >= 0 && frame < g_.n_frames && level >= 0 && level < g_.n_levels && filter >= 0 && filter < model_.nfilters(), "set_response");
  const LevelDesc& L = g_.lv[level];
  check_cuda(cudaMemcpyAsync(b_.resp + ((size_t)frame * model_.nfilters() + filter) * g_.cells_total + L.cell_off, src,
                             (size_t)L.oh * L.ow * sizeof(float), cudaMemcpyHostToDevice, stream_), "H2D response");
  check_cuda(cudaStreamSynchronize(stream_), "sync");
  stage_ = std::max(stage_, 3);
}

void Engine::stage_times(float ms[6]) {
  for (int i = 0; i < 6; ++i) ms[i] = 0.f;
  check_cuda(cudaStreamSynchronize(stream_), "sync");
  for (int i = 0; i < 6; ++i)
    if (ev_valid_[i] && ev_valid_[i + 1]) check_cuda(cudaEventElapsedTime(&ms[i], ev_[i], ev_[i + 1]), "event elapsed");
}

}  // namespace pbd
