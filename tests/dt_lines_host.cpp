// Host build of the product's parallel-in-q distance transform (partsbaseddetector_b200/csrc/dt_lines.cuh) for the CPU test suite:
// g++ -O2 -std=c++20 -ffp-contract=off -pthread.  Test infrastructure only.  A "warp" is emulated by 32 threads; every collective
// (shuffle, ballot, syncwarp) is a pair of barrier waits, shared memory is a plain buffer, atomicMax a CAS loop -- so the control flow
// of process_lines (batching, aliasing of own onto z, phase ordering) is exactly what the device executes.
#include <atomic>
#include <barrier>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

static thread_local long long g_stat_pop = 0, g_stat_adv = 0, g_stat_exact = 0;
#define PBD_ENV_STAT(name) ++g_stat_##name;
#include "../partsbaseddetector_b200/csrc/dt_lines.cuh"

using namespace pbd;

namespace {
struct WarpShared {
  std::barrier<> bar{32};
  uint64_t slot[32];
};
struct HostWarp {
  int lane_;
  WarpShared* s;
  int lane() const { return lane_; }
  unsigned ballot(bool p) const {
    s->slot[lane_] = p ? 1u : 0u;
    s->bar.arrive_and_wait();
    unsigned m = 0;
    for (int i = 0; i < 32; ++i) m |= (unsigned)s->slot[i] << i;
    s->bar.arrive_and_wait();
    return m;
  }
  template <typename T> T shfl(T v, int src) const {
    uint64_t u = 0; std::memcpy(&u, &v, sizeof(T));
    s->slot[lane_] = u;
    s->bar.arrive_and_wait();
    T r; std::memcpy(&r, &s->slot[src & 31], sizeof(T));
    s->bar.arrive_and_wait();
    return r;
  }
  template <typename T> T shfl_up(T v, int d) const {
    uint64_t u = 0; std::memcpy(&u, &v, sizeof(T));
    s->slot[lane_] = u;
    s->bar.arrive_and_wait();
    T r = v;
    if (lane_ >= d) std::memcpy(&r, &s->slot[lane_ - d], sizeof(T));
    s->bar.arrive_and_wait();
    return r;
  }
  template <typename T> T shfl_xor(T v, int m) const { return shfl(v, lane_ ^ m); }
  void sync() const { s->bar.arrive_and_wait(); }
  void atomic_max(int* p, int v) const {
    std::atomic_ref<int> a(*p);
    int cur = a.load();
    while (cur < v && !a.compare_exchange_weak(cur, v)) {}
  }
};
}  // namespace

extern "C" {
// nlines lines of N samples (src[line][q]) of ONE map go through one emulated warp in batches of `b` lines (1..32), exactly as the
// device kernels do.  kreg: 0 = run-time loops with a separate own array, 8 = the straight-line variants (own aliases z) for N <= 256.
// dst / ptr are [line][pos - os].  pops (optional) receives {pop iterations of phase B, intersections recomputed in double because
// the fp32 certificate failed}.
int dtl_dt1d(const float* src, int nlines, int N, float w_sq, float w_lin, int os, int b, int kreg, float* dst, uint16_t* ptr, long long* pops) {
  if (N < 1 || N > 65535 || nlines < 1 || b < 1 || b > 32 || (kreg != 0 && kreg != 8) || (kreg == 8 && N > 256)) return -1;
  const int maxn = N;
  const int ne = 2 * maxn - 1 + env::kTabPad, bias = maxn - 1 - os;
  std::vector<double> tab(ne + env::kRcp);
  const double a = (double)(-w_sq), bb = (double)(-w_lin);
  for (int j = 0; j < ne; ++j) tab[j] = env::table_E(a, bb, j - bias);
  for (int j = 0; j < env::kRcp; ++j) tab[ne + j] = env::table_rcp(a, j);
  const env::Quad f = env::make_quad(w_sq, w_lin, tab.data() + bias, tab.data() + ne);
  const int LS = dtl::line_stride(N), NW = (N + 31) >> 5;
  std::vector<float> y((size_t)b * LS), z((size_t)b * LS);
  std::vector<int> own_sep((size_t)b * LS);
  std::vector<unsigned> bits((size_t)b * NW);
  WarpShared ws;
  std::atomic<long long> npop{0}, nexact{0};
  auto body = [&](int lane) {
    HostWarp w{lane, &ws};
    g_stat_pop = 0; g_stat_exact = 0;
    for (int l0 = 0; l0 < nlines; l0 += b) {
      const int nb = std::min(b, nlines - l0);
      float ymax = 0.f;
      for (int l = 0; l < nb; ++l)
        for (int q = lane; q < N; q += 32) { const float v = src[(size_t)(l0 + l) * N + q]; y[(size_t)l * LS + q] = v; ymax = std::fmax(ymax, std::fabs(v)); }
      for (int m = 16; m > 0; m >>= 1) ymax = std::fmax(ymax, w.shfl_xor(ymax, m));
      w.sync();
      auto out = [&](int l, int i, float val, int v) {
        if (l < 0 || l >= nb || i < 0 || i >= N) __builtin_trap();
        dst[(size_t)(l0 + l) * N + i] = val; ptr[(size_t)(l0 + l) * N + i] = (uint16_t)v;
      };
      if (kreg == 8) dtl::process_lines_any(w, f, N, os, nb, ymax, y.data(), z.data(), reinterpret_cast<int*>(z.data()), bits.data(), out);
      else dtl::process_lines<0>(w, f, N, os, nb, ymax, y.data(), z.data(), own_sep.data(), bits.data(), out);
      w.sync();
    }
    npop += g_stat_pop; nexact += g_stat_exact;
  };
  std::vector<std::thread> th;
  for (int lane = 0; lane < 32; ++lane) th.emplace_back(body, lane);
  for (auto& t : th) t.join();
  if (pops) { pops[0] = npop.load(); pops[1] = nexact.load(); }
  return 0;
}
}
