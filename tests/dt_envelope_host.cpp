// Host build of the product's streaming envelope (partsbaseddetector_b200/csrc/dt_envelope.cuh) for the CPU test suite:
// g++ -O2 -ffp-contract=off.  Test infrastructure only -- the product runs this header as device code inside dt_pass.
#include <cstdint>
#include <vector>

// per-step loop statistics of the lock-step (warp) execution model: see envh_stats below
static thread_local int g_stat_pop = 0, g_stat_adv = 0;
static thread_local long long g_stat_exact = 0;
#define PBD_ENV_STAT(name) ++g_stat_##name;
#include "../partsbaseddetector_b200/csrc/dt_envelope.cuh"

using namespace pbd::env;

extern "C" {
// 32 lines of N samples each (src[line][q]) go through one "warp": the lanes run one after the other over the same Ring
// object (each lane only touches its own column).  maxn >= N sizes the tables exactly as the device does
// (dt_table_len / dt_table_bias of kernels.cuh).  dst/ptr are written [line][pos - os]; every position not stored stays
// at the caller's fill value.  *stores (optional) counts the emit calls.
// window = 0: emissions go straight to dst/ptr; window = 1: through the write-back window (OutWindow<8, 5>);
// window = 2 / 3 / 4: the lagged-scan variant envelope_scan with LAG = 4 / 1 / 12; window = 5: envelope_stream_cert (certified fp32
// break points, option dt_variant 1); *stores then also receives, in stores[1], the number of intersections that took the double path
int envh_dt1d(const float* src, int nlines, int N, float w_sq, float w_lin, int os, int maxn, float* dst, uint16_t* ptr, long long* stores,
              int window) {
  if (N < 1 || N > maxn || nlines < 1 || nlines > 32) return -1;
  const int ne = 2 * maxn - 1 + kTabPad, bias = maxn - 1 - os;
  std::vector<double> tab(ne + kRcp);
  const double a = (double)(-w_sq), b = (double)(-w_lin);
  for (int j = 0; j < ne; ++j) tab[j] = table_E(a, b, j - bias);
  for (int j = 0; j < kRcp; ++j) tab[ne + j] = table_rcp(a, j);
  const Quad f = make_quad(w_sq, w_lin, tab.data() + bias, tab.data() + ne);
  Ring R;
  RingE RE;
  std::vector<float> zb(N);
  std::vector<unsigned short> pb(N);
  long long n = 0;
  g_stat_exact = 0;
  for (int lane = 0; lane < nlines; ++lane) {
    const float* s = src + (size_t)lane * N;
    float* d = dst + (size_t)lane * N;
    uint16_t* p = ptr + (size_t)lane * N;
    auto store = [&](int i, float val, unsigned short v) {
      if (i < 0 || i >= N) __builtin_trap();
      d[i] = val; p[i] = v; ++n;
    };
    if (window == 6) {                                  // the literal two-loop form the kernels use for lines with a non-finite sample
      envelope_literal(N, f, os, zb.data(), pb.data(), [&](int q) { return s[q]; }, [&](int i, float val, int v) { store(i, val, (unsigned short)v); });
    } else if (window == 5) {
      auto ld = [&](int q) { return s[q]; };
      envelope_stream_cert(N, f, os, RE, lane, zb.data(), pb.data(), ld, ld, [&](int i, float val, int v) { store(i, val, (unsigned short)v); });
    } else if (window >= 2) {
      auto ld = [&](int q) { return s[q]; };
      auto em = [&](int i, float val, int v) { store(i, val, (unsigned short)v); };
      if (window == 2) envelope_scan<4>(N, f, os, R, lane, zb.data(), pb.data(), ld, ld, em);
      else if (window == 3) envelope_scan<1>(N, f, os, R, lane, zb.data(), pb.data(), ld, ld, em);
      else envelope_scan<12>(N, f, os, R, lane, zb.data(), pb.data(), ld, ld, em);
    } else if (!window) {
      envelope_stream(N, f, os, R, lane, zb.data(), pb.data(), [&](int q) { return s[q]; }, [&](int v) { return s[v]; },
                      [&](int i, float val, int v) { store(i, val, (unsigned short)v); }, [](int) {});
    } else {
      float wval[8 * 32];
      unsigned short wptr[8 * 32];
      OutWindow<8, 5> win;
      win.init(wval + lane, wptr + lane);
      envelope_stream(N, f, os, R, lane, zb.data(), pb.data(), [&](int q) { return s[q]; }, [&](int v) { return s[v]; },
                      [&](int i, float val, int v) { win.put(i, val, v, store); }, [&](int q) { win.step(q, os, store); });
      win.finish(store);
    }
  }
  if (stores) { stores[0] = n; if (window == 5) stores[1] = g_stat_exact; }
  return 0;
}
// Loop iterations per (line, sample step) of one variant (0 = eager envelope_stream, 2 = envelope_scan<4>): pops, emissions and cursor
// advances, each [nlines][N] (step q = everything between loady(q) and loady(q + 1); the tail after the last sample counts for
// step N - 1).  A warp executes max-over-its-32-lines iterations of each loop, which is what tools/dt_warp_model.py evaluates.
int envh_stats(const float* src, int nlines, int N, float w_sq, float w_lin, int os, int variant, int* pops, int* emits, int* advs) {
  const int maxn = N;
  const int ne = 2 * maxn - 1 + kTabPad, bias = maxn - 1 - os;
  std::vector<double> tab(ne + kRcp);
  const double a = (double)(-w_sq), b = (double)(-w_lin);
  for (int j = 0; j < ne; ++j) tab[j] = table_E(a, b, j - bias);
  for (int j = 0; j < kRcp; ++j) tab[ne + j] = table_rcp(a, j);
  const Quad f = make_quad(w_sq, w_lin, tab.data() + bias, tab.data() + ne);
  Ring R;
  std::vector<float> zb(N);
  std::vector<unsigned short> pb(N);
  for (int line = 0; line < nlines; ++line) {
    const float* s = src + (size_t)line * N;
    int* P = pops + (size_t)line * N; int* E = emits + (size_t)line * N; int* A = advs + (size_t)line * N;
    for (int q = 0; q < N; ++q) P[q] = E[q] = A[q] = 0;
    int cur = 0, ne_cur = 0;
    g_stat_pop = g_stat_adv = 0;
    auto flush = [&]() { P[cur] += g_stat_pop; A[cur] += g_stat_adv; E[cur] += ne_cur; g_stat_pop = g_stat_adv = 0; ne_cur = 0; };
    auto ld = [&](int q) { if (q >= 1) { flush(); cur = q; } return s[q]; };   // step q starts with its load
    auto rl = [&](int v) { return s[v]; };
    auto em = [&](int, float, int) { ++ne_cur; };
    if (variant == 0) envelope_stream(N, f, os, R, line & 31, zb.data(), pb.data(), ld, rl, em, [](int) {});
    else envelope_scan<4>(N, f, os, R, line & 31, zb.data(), pb.data(), ld, rl, em);
    flush();
    // attribute the work done before sample 1 (none) and shift: step q's record holds the work triggered by sample q
  }
  return 0;
}
// Prototype of the parallel-in-q schedule of the same algorithm: phase A computes every adjacent intersection
// independently, phase B repairs the pop sites only (implicit stack: pred(q) = q - 1, z(q) = s_q unless q is a site), phase C emits per
// surviving entry.  Written with the product header's primitives; must equal the oracle bit for bit.  stats[0..3] += sites, pop
// iterations, surviving entries, candidate sites (s_q <= s_{q-1}).
int envh_dt1d_parallel(const float* src, int nlines, int N, float w_sq, float w_lin, int os, float* dst, uint16_t* ptr, long long* stats) {
  const int maxn = N;
  const int ne = 2 * maxn - 1 + kTabPad, bias = maxn - 1 - os;
  std::vector<double> tab(ne + kRcp);
  const double a = (double)(-w_sq), b = (double)(-w_lin);
  for (int j = 0; j < ne; ++j) tab[j] = table_E(a, b, j - bias);
  for (int j = 0; j < kRcp; ++j) tab[ne + j] = table_rcp(a, j);
  const Quad f = make_quad(w_sq, w_lin, tab.data() + bias, tab.data() + ne);
  const int pos_last = os + N - 1;
  std::vector<float> s(N), z(N);
  std::vector<int> pred(N), alive;
  for (int line = 0; line < nlines; ++line) {
    const float* y = src + (size_t)line * N;
    float* d = dst + (size_t)line * N;
    uint16_t* p = ptr + (size_t)line * N;
    // phase A: independent per q
    s[0] = -INFINITY;
    for (int q = 1; q < N; ++q) s[q] = isect_adjacent(f, q, (double)y[q - 1], (double)y[q]);
    for (int q = 2; q < N; ++q) if (s[q] <= s[q - 1]) ++stats[3];
    // phase B: sequential only at the pop sites
    z[0] = -INFINITY; pred[0] = -1;
    for (int q = 1; q < N; ++q) {
      int v = q - 1;
      float sq = s[q];
      if (sq <= z[v] && v != 0) {                                 // the top q - 1 is not the bottom entry: :163 `while (s <= z[k])`
        ++stats[0];
        do {
          v = pred[v];
          ++stats[1];
          sq = isect_far(f, v, q, (double)y[v], (double)y[q]);
        } while (sq <= z[v] && v != 0);
      }
      z[q] = sq; pred[q] = v;
    }
    // phase C: per surviving entry (independent)
    alive.clear();
    for (int v = N - 1; v >= 0; v = pred[v]) alive.push_back(v);
    stats[2] += (long long)alive.size();
    for (size_t i = alive.size(); i-- > 0;) {
      const int v = alive[i];
      const float zlo = z[v], zhi = i == 0 ? INFINITY : z[alive[i - 1]];
      const int lo = imax(imin(f2i_floor(zlo), pos_last) + 1, os), hi = imin(f2i_floor(zhi), pos_last);
      for (int pos = lo; pos <= hi; ++pos) { d[pos - os] = (float)dadd(ld_table(f.E, pos - v), (double)y[v]); p[pos - os] = (uint16_t)v; }
    }
  }
  return 0;
}
// the two quotient paths side by side (for the reciprocal / Markstein test)
float envh_quotient_fast(double num, double den) { return quotient_to_float(num, den, drcp(den)); }
float envh_quotient_exact(double num, double den) { return quotient_exact(num, den); }
}
