// dt_lines.cuh -- the 1-D generalised distance transform (DistanceTransform<float>::computeRow, reference
// include/DistanceTransform.hpp:152-182) scheduled "parallel in q": a warp owns a batch of lines that live in shared memory and
// runs three phases over them.  It is still the reference's stack algorithm with its float-rounded break points (every
// intersection is the reference's double expression, dt_envelope.cuh), only the order of evaluation differs:
//
//   phase A (lane = sample q, all q of a line in parallel)   every sample is pushed, so at step q the top of the stack is always
//       sample q - 1 and the first intersection s_q = f(q - 1, q) needs two neighbouring samples only.  z[q] = s_q, pred[q] = q - 1,
//       and a bit per sample marks the tentative pop sites  s_q <= s_{q-1}.
//   phase B (lane = line, sequential along the line, sites only: ~5 % of the samples of real score maps)   a site pops the top(s):
//       walk pred[], recompute the intersection with the general formula, mark the popped entries dead (z = +inf), store the
//       site's final z / pred, then re-test the next sample against the repaired break point (a repair can turn q + 1 into a
//       site or clear it); all other tentative bits stay valid because their left neighbour's z is untouched.
//   phase C (lane = position)   the reference's scan picks for position pos the entry k with z[k] < pos <= z[k+1] (:171-181), i.e.
//       the LARGEST live entry whose first position  lo = floor(z) + 1  is <= pos (break points increase strictly up the stack, by
//       the loop condition itself).  Every live entry scatters its index to own[lo] with max, a prefix maximum over the positions
//       fills the ranges, and each position evaluates its owner's parabola (same table and double add as dt_envelope.cuh).
//
// The warp-level primitives come from a policy object so that tests/dt_lines_host.cpp can run this very code on the CPU with 32
// threads per "warp" (std::barrier for every collective) against the oracle.
#pragma once
#include "dt_envelope.cuh"

namespace pbd {
namespace dtl {

using env::Quad;

#if defined(__CUDACC__)
struct DevWarp {
  __device__ __forceinline__ int lane() const { return (int)(threadIdx.x & 31u); }
  __device__ __forceinline__ unsigned ballot(bool p) const { return __ballot_sync(0xffffffffu, p); }
  template <typename T> __device__ __forceinline__ T shfl(T v, int src) const { return __shfl_sync(0xffffffffu, v, src); }
  template <typename T> __device__ __forceinline__ T shfl_up(T v, int d) const { return __shfl_up_sync(0xffffffffu, v, d); }   // lanes < d keep v
  __device__ __forceinline__ void sync() const { __syncwarp(); }
  __device__ __forceinline__ void atomic_max(int* p, int v) const { atomicMax(p, v); }
  __device__ __forceinline__ int ctz(unsigned m) const { return __ffs((int)m) - 1; }
};
#endif

// line stride (in elements) of the per-line shared-memory arrays: odd, so that the lanes of phase B (one line each) spread over the banks
#if defined(__CUDACC__)
#define PBD_HD __host__ __device__ __forceinline__
#else
#define PBD_HD inline
#endif
PBD_HD int line_stride(int N) { return N | 1; }
// bytes of one line's state: y + z (float), pred (u16), site bits; `own` aliases z when the line fits the register window
PBD_HD int line_bytes(int N, bool alias) {
  const int LS = line_stride(N);
  return LS * 4 * (alias ? 2 : 3) + ((N + 31) >> 5) * 4 + ((LS + 1) & ~1) * 2;
}

// ---- phase A: one line, lane = q -------------------------------------------------------------------------------------------
template <class W>
PBD_ENV_FN void phase_a(const W& w, const Quad& f, int N, const float* y, float* z, unsigned short* pred, unsigned* bits) {
  const int lane = w.lane();
  float carry = 0.f;                                              // s of the last sample of the previous chunk
  for (int j0 = 0; j0 < N; j0 += 32) {
    const int q = j0 + lane;
    const bool in = q < N;
    const float yq = in ? y[q] : 0.f;
    const float yp = (in && q > 0) ? y[q - 1] : 0.f;
    float s = env::isect_adjacent(f, q, (double)yp, (double)yq); // :161 with the top = sample q - 1
    if (q == 0) s = -INFINITY;                                    // z[0] = -inf, :156
    float sp = w.shfl_up(s, 1);
    if (lane == 0) sp = carry;
    const bool site = in && q >= 2 && s <= sp;                    // :163 `while (s <= z[k] && k > 0)`: sample 1 sits on the bottom entry
    const unsigned word = w.ballot(site);
    if (in) { z[q] = s; pred[q] = (unsigned short)(q - 1); }
    if (lane == 0) bits[j0 >> 5] = word;
    carry = w.shfl(s, 31);
  }
}

// ---- phase B: one line per lane, sequential over the line's sites ----------------------------------------------------------------
template <class W>
PBD_ENV_FN void phase_b(const W& w, const Quad& f, int N, const float* y, float* z, unsigned short* pred, const unsigned* bits) {
  const int nw = (N + 31) >> 5;
  auto next_site = [&](int from) -> int {                         // smallest tentative site >= from, or N
    if (from >= N) return N;
    int wi = from >> 5;
    unsigned m = bits[wi] & (0xffffffffu << (from & 31));
    for (;;) {
      if (m) return (wi << 5) + w.ctz(m);
      if (++wi >= nw) return N;
      m = bits[wi];
    }
  };
  int q = next_site(2);
  while (q < N) {
    const double yq = (double)y[q];
    int v = q - 1;                                                // the top; it is popped (s_q <= z[q-1] and q - 1 != 0 hold here)
    float zq;
    do {
      const int dead = v;
      v = pred[v];
      z[dead] = INFINITY;                                         // popped entries own no position
      PBD_ENV_STAT(pop)
      zq = env::isect_far(f, v, q, (double)y[v], yq);             // :165
    } while (zq <= z[v] && v != 0);                               // :163
    z[q] = zq; pred[q] = (unsigned short)v;                       // :167-169
    const int qn = q + 1;
    if (qn < N && z[qn] <= zq) q = qn;                            // the repaired break point decides whether q + 1 pops q
    else q = next_site(qn + 1);
  }
}

#if defined(__CUDA_ARCH__)
PBD_ENV_FN float bits_to_float(int v) { return __int_as_float(v); }
PBD_ENV_FN int float_to_bits(float v) { return __float_as_int(v); }
#else
PBD_ENV_FN float bits_to_float(int v) { float f; std::memcpy(&f, &v, 4); return f; }
PBD_ENV_FN int float_to_bits(float v) { int i; std::memcpy(&i, &v, 4); return i; }
#endif
// first position index (pos - os) owned by an entry with break point zq: positions pos > zq, clipped to the line; N = none
PBD_ENV_FN int first_index(float zq, int os, int N) {
  return env::imax(env::imin(env::f2i_floor(zq), os + N - 1) + 1, os) - os;
}

// ---- phase C: one line, lane = position ----------------------------------------------------------------------------------------
// KREG > 0: the line has at most 32*KREG samples and `own` IS the z array (z is then read through `own`, so that no access
// depends on type-based alias analysis; the break points are parked in registers before the slots are reused);
// KREG == 0: any length, own is a separate array.  out(i, value, argmax) is called once for every position index i.
template <int KREG, class W, class Out>
PBD_ENV_FN void phase_c(const W& w, const Quad& f, int N, int os, const float* y, const float* z, int* own, Out out) {
  const int lane = w.lane();
  if constexpr (KREG > 0) {
    int idx[KREG > 0 ? KREG : 1];
#pragma unroll
    for (int k = 0; k < KREG; ++k) {
      const int q = k * 32 + lane;
      idx[k] = (q >= 1 && q < N) ? first_index(bits_to_float(own[q]), os, N) : N;  // sample 0 (z = -inf) owns from index 0: the initial value of own[]
      if (q < N) own[q] = 0;                                      // same slot, same lane: z[q] has just been read
    }
    w.sync();
#pragma unroll
    for (int k = 0; k < KREG; ++k) if (idx[k] < N) w.atomic_max(own + idx[k], k * 32 + lane);
  } else {
    for (int q = lane; q < N; q += 32) own[q] = 0;
    w.sync();
    for (int q = lane; q < N; q += 32) {
      if (q >= 1) { const int ix = first_index(z[q], os, N); if (ix < N) w.atomic_max(own + ix, q); }
    }
  }
  w.sync();
  int carry = 0;
  for (int j0 = 0; j0 < N; j0 += 32) {
    const int i = j0 + lane;
    int o = i < N ? own[i] : 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) o = env::imax(o, w.shfl_up(o, d));   // inclusive prefix maximum (lanes < d get their own value back)
    o = env::imax(o, carry);
    carry = w.shfl(o, 31);
    if (i < N) out(i, (float)env::dadd(env::ld_table(f.E, os + i - o), (double)y[o]), o);   // :175-178
  }
}

// ---- a batch of nb <= 32 lines of one map, already staged in y[line][line_stride(N)] ----------------------------------------------
// out(line, i, value, argmax).  own == (int*)z is allowed when KREG > 0.
template <int KREG, class W, class Out>
PBD_ENV_FN void process_lines(const W& w, const Quad& f, int N, int os, int nb, const float* y, float* z, int* own, unsigned short* pred,
                              unsigned* bits, Out out) {
  const int LS = line_stride(N), NW = (N + 31) >> 5, LSP = (LS + 1) & ~1;
  for (int l = 0; l < nb; ++l) phase_a(w, f, N, y + l * LS, z + l * LS, pred + l * LSP, bits + l * NW);
  w.sync();
  if (w.lane() < nb) { const int l = w.lane(); phase_b(w, f, N, y + l * LS, z + l * LS, pred + l * LSP, bits + l * NW); }
  w.sync();
  for (int l = 0; l < nb; ++l)
    phase_c<KREG>(w, f, N, os, y + l * LS, z + l * LS, own + l * LS, [&](int i, float val, int v) { out(l, i, val, v); });
}

}  // namespace dtl
}  // namespace pbd
