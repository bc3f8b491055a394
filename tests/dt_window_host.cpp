// Host build of the windowed certified distance transform (partsbaseddetector_b200/csrc/dt_window.cuh): the same tier-1 / tier-2
// functions the device kernels call, driven one line at a time so that the CPU suite can compare every ACCEPTED line with the oracle
// and count the lines handed to the literal stack algorithm.  Test infrastructure only.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../partsbaseddetector_b200/csrc/dt_window.cuh"

using namespace pbd;

static long long g_cause[8];   // refusal causes (diagnostics): 0 non-finite, 1 local replay conditions, 2 edge, 3 more open positions than the list holds, 4 open neighbour

template <int W>
static int line_w(const float* src, int N, const dtw::WinParams& P, float w_sq, float w_lin, int os, float* dst, uint16_t* ptr, long long* tier2) {
  int dirty = 0;
  for (int v = 0; v < N; ++v)
    if (!std::isfinite(src[v])) { dirty = 1; ++g_cause[0]; }                           // the kernels refuse a line with a NaN / +-inf sample while staging it
  // the parabola tables the stack kernels use (tier 3 forms the reference's break point with env::isect_*)
  const double a = (double)(-w_sq), b = (double)(-w_lin);
  double rcp[env::kRcp];
  for (int d = 0; d < env::kRcp; ++d) rcp[d] = env::table_rcp(a, d);
  const env::Quad f = env::make_quad(w_sq, w_lin, nullptr, rcp);
  std::vector<int> open;                                             // positions no tier certified, in increasing order
  std::vector<int> state(N, 0);                                      // 1: certified by tier 1 / 2
  constexpr int kCap = 256;                                          // the device kernel's list capacity
  for (int q = 0; q < N; ++q) {
    const int p = os + q;
    float y[2 * W + 1], c[2 * W + 1];
    for (int j = 0; j <= 2 * W; ++j) {
      const int v = p - W + j;
      y[j] = (v >= 0 && v < N) ? src[v] : -INFINITY;
      c[j] = env::fadd_r(y[j], P.ef[j]);
    }
    constexpr int RIN = W > 2 ? 2 : W;
    float yin[2 * RIN + 1];
    for (int k = 0; k <= 2 * RIN; ++k) yin[k] = y[W - RIN + k];
    const dtw::Pick pk = dtw::pick_walk<W, RIN>(c, yin, P.tau0, P.ylim);
    int j = pk.ok ? pk.j : -1;
    if (pk.ok && std::memcmp(&pk.yv, &y[pk.j], sizeof(float)) != 0) { ++g_cause[5]; return 2; }   // the carried sample must be the winner's, bit for bit
    if (j < 0) {
      if (tier2) ++*tier2;
      j = dtw::pick_exact<W>(y, P.ed, P.margin1, P.cmax, P.ylim);
    }
    if (j < 0) { open.push_back(q); continue; }
    if (!dtw::edge_ok(j, W, q, N)) { dirty = 1; ++g_cause[2]; continue; }
    dst[q] = dtw::value_of(P.ed[j], y[j]);
    ptr[q] = (uint16_t)(p - W + j);
    state[q] = 1;
  }
  if ((int)open.size() > kCap) { dirty = 1; ++g_cause[3]; }
  for (size_t k = 0; k < open.size() && !dirty;) {
    size_t e = k;
    while (e + 1 < open.size() && open[e + 1] == open[e] + 1) ++e;   // a run of consecutive open positions
    const int qa = open[k], qb = open[e];
    if ((qa > 0 && !state[qa - 1]) || (qb < N - 1 && !state[qb + 1])) { dirty = 1; ++g_cause[4]; break; }
    const int uL = qa > 0 ? ptr[qa - 1] : 0, uR = qb < N - 1 ? ptr[qb + 1] : N - 1;
    if (!dtw::local_ok(W, os, N, qa, qb, uL, uR)) { dirty = 1; ++g_cause[1]; break; }
    dtw::local_owners(f, qa + os, qb + os, uL, uR, [&](int u) { return src[u]; }, [&](int p, int v, float yo) {
      dst[p - os] = dtw::value_of(env::table_E(a, b, p - v), yo);
      ptr[p - os] = (uint16_t)v;
    });
    k = e + 1;
  }
  return dirty;
}

extern "C" {
void wnd_causes(long long* out, int reset) { for (int i = 0; i < 8; ++i) { out[i] = g_cause[i]; if (reset) g_cause[i] = 0; } }
// nl lines of N samples; dirty[line] = 1 when the line must go to the stack algorithm (its dst / ptr are then unspecified).
// returns -1 when the map cannot use the window at all (params.ok == 0)
int wnd_dt1d(const float* src, int nl, int N, float w_sq, float w_lin, int os, int W, float* dst, uint16_t* ptr, uint8_t* dirty,
             long long* tier2) {
  const dtw::WinParams P = dtw::make_params(w_sq, w_lin, os, N, W);
  if (!P.ok) return -1;
  for (int l = 0; l < nl; ++l) {
    const float* s = src + (size_t)l * N;
    float* d = dst + (size_t)l * N;
    uint16_t* p = ptr + (size_t)l * N;
    int r;
    switch (W) {
      case 3: r = line_w<3>(s, N, P, w_sq, w_lin, os, d, p, tier2); break;
      case 4: r = line_w<4>(s, N, P, w_sq, w_lin, os, d, p, tier2); break;
      case 5: r = line_w<5>(s, N, P, w_sq, w_lin, os, d, p, tier2); break;
      case 6: r = line_w<6>(s, N, P, w_sq, w_lin, os, d, p, tier2); break;
      case 8: r = line_w<8>(s, N, P, w_sq, w_lin, os, d, p, tier2); break;
      default: return -2;
    }
    dirty[l] = (uint8_t)r;
  }
  return 0;
}
}
