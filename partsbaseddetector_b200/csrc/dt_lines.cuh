// dt_lines.cuh -- the 1-D generalised distance transform (DistanceTransform<float>::computeRow, reference
// include/DistanceTransform.hpp:152-182) scheduled "parallel in q": a warp owns a batch of lines that live in shared memory and
// runs three phases over them.  It is still the reference's stack algorithm with its float-rounded break points, only the order of
// evaluation differs:
//
//   phase A (lane = sample q, all q of a line in parallel)   every sample is pushed, so at step q the top of the stack is always
//       sample q - 1 and the first intersection s_q = f(q - 1, q) needs two neighbouring samples only.  z[q] = s_q, and a bit per
//       sample marks the tentative pop sites  s_q <= s_{q-1}.
//   phase B (lane = line, sequential along the line, sites only: ~5 % of the samples of real score maps)   a site pops the top(s):
//       walk down the stack, recompute the intersection with the general formula, mark the popped entries dead (z = +inf), store
//       the site's final z, then re-test the next sample against the repaired break point (a repair can turn q + 1 into a site or
//       clear it); all other tentative bits stay valid because their left neighbour's z is untouched.  No predecessor links are
//       stored: entries are only ever popped from the top, so the entry below a live entry v is the nearest live sample below v.
//   phase C (lane = position)   the reference's scan picks for position pos the entry k with z[k] < pos <= z[k+1] (:171-181), i.e.
//       the LARGEST live entry whose first position  lo = floor(z) + 1  is <= pos (break points increase strictly up the stack, by
//       the loop condition itself).  Every live entry scatters its index to own[lo] with max; because the scattered values increase
//       with the position, a position's owner is the nearest filled slot at or below it (one ballot + find-leading-one per 32
//       positions); each position then evaluates its owner's parabola (same table and double add as dt_envelope.cuh).
//
// Exactness with cheap arithmetic.  Only DECISIONS depend on the break points (s <= z, floor(z)); the values never leave the line.
// Phase A therefore computes s_q in fp32 (one subtraction and one fused multiply-add) together with a rigorous error bound `eps`
// for the whole line, and certifies every decision: floor(s) is taken from the fp32 value only if no integer lies within eps of
// it, a comparison only if the operands differ by more than 2 eps (one exact operand: eps).  Whatever cannot be certified is
// recomputed with the reference's own double expression (env::isect_adjacent / env::isect_far), so every decision equals the
// reference's.  Error bound: with R = 1/(2a), s* = (y1 - y0) R + q + C0, C0 = -1/2 - b R;  s_fast = fma(fl(y1 - y0), fl(R),
// fl(q + fl(C0))) carries at most 4 u (|dR| + |q + C0| + |C0|) of rounding error (u = 2^-24) and the reference's float(double)
// result u |s*| more; eps = 2^-20 (2 |fl(R)| max|y| + N + 2 |C0| + 1) is > 3x that.  Non-finite or huge operands fail every
// certificate and take the exact path.
//
// The warp-level primitives come from a policy object so that tests/dt_lines_host.cpp can run this very code on the CPU with 32
// threads per "warp" (std::barrier for every collective) against the oracle.
#pragma once
#include "dt_envelope.cuh"

namespace pbd {
namespace dtl {

using env::Quad;

#if defined(__CUDACC__)
#define PBD_HD __host__ __device__ __forceinline__
struct DevWarp {
  __device__ __forceinline__ int lane() const { return (int)(threadIdx.x & 31u); }
  __device__ __forceinline__ unsigned ballot(bool p) const { return __ballot_sync(0xffffffffu, p); }
  template <typename T> __device__ __forceinline__ T shfl(T v, int src) const { return __shfl_sync(0xffffffffu, v, src); }
  template <typename T> __device__ __forceinline__ T shfl_up(T v, int d) const { return __shfl_up_sync(0xffffffffu, v, d); }   // lanes < d keep v
  template <typename T> __device__ __forceinline__ T shfl_xor(T v, int m) const { return __shfl_xor_sync(0xffffffffu, v, m); }
  __device__ __forceinline__ void sync() const { __syncwarp(); }
  __device__ __forceinline__ void atomic_max(int* p, int v) const { atomicMax(p, v); }
};
#else
#define PBD_HD inline
#endif

#if defined(__CUDA_ARCH__)
PBD_ENV_FN float bits_to_float(int v) { return __int_as_float(v); }
PBD_ENV_FN int float_to_bits(float v) { return __float_as_int(v); }
PBD_ENV_FN float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
PBD_ENV_FN float fsub(float a, float b) { return __fsub_rn(a, b); }
PBD_ENV_FN float fnear(float v) { return rintf(v); }
PBD_ENV_FN int ctz(unsigned m) { return __ffs((int)m) - 1; }
PBD_ENV_FN int top_bit(unsigned m) { return 31 - __clz((int)m); }      // m != 0
#else
PBD_ENV_FN float bits_to_float(int v) { float f; std::memcpy(&f, &v, 4); return f; }
PBD_ENV_FN int float_to_bits(float v) { int i; std::memcpy(&i, &v, 4); return i; }
PBD_ENV_FN float ffma(float a, float b, float c) { return std::fmaf(a, b, c); }
PBD_ENV_FN float fsub(float a, float b) { return a - b; }
PBD_ENV_FN float fnear(float v) { return std::nearbyintf(v); }
PBD_ENV_FN int ctz(unsigned m) { return __builtin_ctz(m); }
PBD_ENV_FN int top_bit(unsigned m) { return 31 - __builtin_clz(m); }
#endif

// line stride (in elements) of the per-line shared-memory arrays: odd, so that the lanes of phase B (one line each) spread over the banks
PBD_HD int line_stride(int N) { return N | 1; }
// bytes of one line's state: y + z (float; `own` aliases z when the line fits the register window), site bits, optional u16 stash
PBD_HD int line_bytes(int N, bool alias, bool stash) {
  const int LS = line_stride(N);
  return LS * 4 * (alias ? 2 : 3) + ((N + 31) >> 5) * 4 + (stash ? ((LS + 1) & ~1) * 2 : 0);
}

// per-line constants of the certified fp32 intersection
struct Fast {
  float r;        // fl(1 / (2a))
  float c0;       // fl(-1/2 - b / (2a))
  float eps;      // error bound of s_fast against the reference's float break point
};
PBD_ENV_FN Fast make_fast(const Quad& f, int N, float ymax) {
  Fast F;
  const double R = f.r1;                                          // correctly rounded 1 / (2a)
  F.r = (float)R;
  F.c0 = (float)env::dsub(-0.5, env::dmul(f.b, R));
  const float mag = 2.f * fabsf(F.r) * ymax + (float)N + 2.f * fabsf(F.c0) + 1.f;
  F.eps = mag * 9.5367431640625e-07f;                              // 2^-20
  return F;
}
// s_fast for adjacent samples q - 1, q
PBD_ENV_FN float isect_fast(const Fast& F, int q, float y0, float y1) { return ffma(fsub(y1, y0), F.r, (float)q + F.c0); }
// true if floor() of the reference value is certainly floor(s): no integer within eps (false for NaN / inf / huge values)
PBD_ENV_FN bool floor_certain(float s, float eps) { return fabsf(fsub(s, fnear(s))) > eps; }

// ---- phase A: one line, lane = q; NC = number of 32-sample chunks (0: run-time loop) ----------------------------------------------
template <int NC, class W>
PBD_ENV_FN void phase_a(const W& w, const Quad& f, const Fast& F, int N, const float* y, float* z, unsigned* bits) {
  const int lane = w.lane();
  const float eps = F.eps, eps2 = eps + eps;
  float carry_s = 0.f, carry_y = 0.f;                             // s and y of the last sample of the previous chunk
  auto chunk = [&](int j0) {
    const int q = j0 + lane;
    const bool in = q < N;
    const float yq = in ? y[q] : 0.f;
    float yp = w.shfl_up(yq, 1);
    if (lane == 0) yp = carry_y;
    float s = isect_fast(F, q, yp, yq);
    if (q == 0) s = -INFINITY;                                    // z[0] = -inf, :156
    float sp = w.shfl_up(s, 1);
    if (lane == 0) sp = carry_s;
    const float diff = fsub(s, sp);
    // certificates: floor(s) for phase C, and the sign of s_q - s_{q-1} (pop test, :163) unless both are certain
    const bool live = in && q >= 1;
    const bool ok = !live || (floor_certain(s, eps) && (q < 2 || fabsf(diff) > eps2));
    bool site = live && q >= 2 && diff <= 0.f;
    if (w.ballot(!ok)) {                                          // rare: the reference's own double expressions
      if (!ok) {
        PBD_ENV_STAT(exact)
        s = env::isect_adjacent(f, q, (double)yp, (double)yq);
        if (q >= 2) site = s <= env::isect_adjacent(f, q - 1, (double)y[q - 2], (double)yp);
      }
    }
    const unsigned word = w.ballot(site);
    if (in) z[q] = s;
    if (lane == 0) bits[j0 >> 5] = word;
    carry_s = w.shfl(s, 31);
    carry_y = w.shfl(yq, 31);
  };
  if (NC > 0) {
#pragma unroll
    for (int k = 0; k < NC; ++k) chunk(k * 32);
  } else {
    for (int j0 = 0; j0 < N; j0 += 32) chunk(j0);
  }
}

// ---- phase B: one line per lane, sequential over the line's sites ----------------------------------------------------------------
// z[] holds fp32-certified or exact break points (error <= eps); the intersections computed here are exact, so a comparison against
// a stored value is certain when the two differ by more than eps, and otherwise the stored value is recomputed exactly first.
PBD_ENV_FN void phase_b(const Quad& f, float eps, int N, const float* y, float* z, const unsigned* bits) {
  const int nw = (N + 31) >> 5;
  auto next_site = [&](int from) -> int {                         // smallest tentative site >= from, or N
    if (from >= N) return N;
    int wi = from >> 5;
    unsigned m = bits[wi] & (0xffffffffu << (from & 31));
    for (;;) {
      if (m) return (wi << 5) + ctz(m);
      if (++wi >= nw) return N;
      m = bits[wi];
    }
  };
  auto below = [&](int v) -> int {                                // the stack entry below the live entry v: nearest live sample
    int u = v - 1;
    while (u > 0 && z[u] == INFINITY) --u;
    return u;
  };
  auto exact_z = [&](int v) -> float {                            // the reference's break point of the live entry v > 0
    const int u = below(v);
    return u == v - 1 ? env::isect_adjacent(f, v, (double)y[u], (double)y[v]) : env::isect_far(f, u, v, (double)y[u], (double)y[v]);
  };
  int q = next_site(2);
  while (q < N) {
    const double yq = (double)y[q];
    int v = q - 1;                                                // the top; it is popped (s_q <= z[q-1] and q - 1 != 0 hold here)
    float zq;
    bool pop;
    do {
      const int dead = v;
      v = below(v);
      z[dead] = INFINITY;                                         // popped entries own no position
      PBD_ENV_STAT(pop)
      zq = env::isect_far(f, v, q, (double)y[v], yq);             // :165
      pop = false;
      if (v != 0) {                                               // :163 `while (s <= z[k] && k > 0)`
        float zv = z[v];
        if (!(fabsf(fsub(zq, zv)) > eps)) { PBD_ENV_STAT(exact) zv = exact_z(v); z[v] = zv; }
        pop = zq <= zv;
      }
    } while (pop);
    z[q] = zq;                                                    // :167-169 (exact)
    const int qn = q + 1;
    bool next_pops = false;
    if (qn < N) {
      float zn = z[qn];                                           // adjacent intersection of q, q + 1 (phase A)
      if (!(fabsf(fsub(zn, zq)) > eps)) { PBD_ENV_STAT(exact) zn = env::isect_adjacent(f, qn, yq, (double)y[qn]); z[qn] = zn; }
      next_pops = zn <= zq;
    }
    q = next_pops ? qn : next_site(qn + 1);
  }
}

// first position index (pos - os) owned by an entry with break point zq: positions pos > zq, clipped to the line; N = none
PBD_ENV_FN int first_index(float zq, int os, int N) {
  return env::imax(env::imin(env::f2i_floor(zq), os + N - 1) + 1, os) - os;
}

// ---- phase C: one line, lane = position ----------------------------------------------------------------------------------------
// NC > 0: the line has at most 32*NC samples and `own` IS the z array (z is read through `own`, so that no access depends on
// type-based alias analysis; the first indices are parked in registers before the slots are reused);
// NC == 0: any length, own is a separate array.  out(i, value, argmax) is called once for every position index i.
// Slots hold sample + 1 (0 = nobody starts here); slot 0 always holds at least sample 0 (z[0] = -inf).
template <int NC, class W, class Out>
PBD_ENV_FN void phase_c(const W& w, const Quad& f, int N, int os, const float* y, const float* z, int* own, Out out) {
  const int lane = w.lane();
  if constexpr (NC > 0) {
    int idx[NC > 0 ? NC : 1];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      const int q = k * 32 + lane;
      idx[k] = (q >= 1 && q < N) ? first_index(bits_to_float(own[q]), os, N) : N;
      if (q < N) own[q] = q == 0 ? 1 : 0;                         // same slot, same lane: z[q] has just been read
    }
    w.sync();
#pragma unroll
    for (int k = 0; k < NC; ++k) if (idx[k] < N) w.atomic_max(own + idx[k], k * 32 + lane + 1);
  } else {
    for (int q = lane; q < N; q += 32) own[q] = q == 0 ? 1 : 0;
    w.sync();
    for (int q = lane; q < N; q += 32) {
      if (q >= 1) { const int ix = first_index(z[q], os, N); if (ix < N) w.atomic_max(own + ix, q + 1); }
    }
  }
  w.sync();
  int carry = 1;
  const unsigned le = 0xffffffffu >> (31 - lane);                 // lanes <= mine
  auto chunk = [&](int j0) {
    const int i = j0 + lane;
    const int mine = i < N ? own[i] : 0;
    const unsigned heads = w.ballot(mine != 0) & le;              // filled slots at or below this position (values increase with i)
    int o = w.shfl(mine, heads ? top_bit(heads) : 0);
    if (!heads) o = carry;
    carry = w.shfl(o, 31);
    o -= 1;
    if (i < N) out(i, (float)env::dadd(env::ld_table(f.E, os + i - o), (double)y[o]), o);   // :175-178
  };
  if (NC > 0) {
#pragma unroll
    for (int k = 0; k < NC; ++k) chunk(k * 32);
  } else {
    for (int j0 = 0; j0 < N; j0 += 32) chunk(j0);
  }
}

// ---- a batch of nb <= 32 lines of one map, already staged in y[line][line_stride(N)]; ymax >= |y| over the batch ------------------
// out(line, i, value, argmax).  own == (int*)z is required when NC > 0.
template <int NC, class W, class Out>
PBD_ENV_FN void process_lines(const W& w, const Quad& f, int N, int os, int nb, float ymax, const float* y, float* z, int* own, unsigned* bits,
                              Out out) {
  const int LS = line_stride(N), NW = (N + 31) >> 5;
  const Fast F = make_fast(f, N, ymax);
  for (int l = 0; l < nb; ++l) phase_a<NC>(w, f, F, N, y + l * LS, z + l * LS, bits + l * NW);
  w.sync();
  if (w.lane() < nb) { const int l = w.lane(); phase_b(f, F.eps, N, y + l * LS, z + l * LS, bits + l * NW); }
  w.sync();
  for (int l = 0; l < nb; ++l)
    phase_c<NC>(w, f, N, os, y + l * LS, z + l * LS, own + l * LS, [&](int i, float val, int v) { out(l, i, val, v); });
}
// run-time dispatch on the number of chunks (1..8: straight-line code, own aliases z; longer lines: loops, separate own array)
template <class W, class Out>
PBD_ENV_FN void process_lines_any(const W& w, const Quad& f, int N, int os, int nb, float ymax, const float* y, float* z, int* own, unsigned* bits,
                                  Out out) {
  switch ((N + 31) >> 5) {
    case 1: process_lines<1>(w, f, N, os, nb, ymax, y, z, own, bits, out); break;
    case 2: process_lines<2>(w, f, N, os, nb, ymax, y, z, own, bits, out); break;
    case 3: process_lines<3>(w, f, N, os, nb, ymax, y, z, own, bits, out); break;
    case 4: process_lines<4>(w, f, N, os, nb, ymax, y, z, own, bits, out); break;
    case 5: process_lines<5>(w, f, N, os, nb, ymax, y, z, own, bits, out); break;
    case 6: process_lines<6>(w, f, N, os, nb, ymax, y, z, own, bits, out); break;
    case 7: process_lines<7>(w, f, N, os, nb, ymax, y, z, own, bits, out); break;
    case 8: process_lines<8>(w, f, N, os, nb, ymax, y, z, own, bits, out); break;
    default: process_lines<0>(w, f, N, os, nb, ymax, y, z, own, bits, out); break;
  }
}
constexpr int kAliasMaxN = 256;   // lines up to this length run the straight-line variants (own aliases z)

}  // namespace dtl
}  // namespace pbd
