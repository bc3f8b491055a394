// dt_envelope.cuh -- one lane's 1-D generalised distance transform (DistanceTransform<float>::computeRow, reference
// include/DistanceTransform.hpp:152-182) in streaming form.  Included by dt.cu (device code) and, compiled by plain g++
// with -ffp-contract=off, by tests/dt_envelope_host.cpp: the CPU suite drives exactly this control flow (ring overflow,
// deep pops, table look-ups, reciprocal quotients) against the oracle before any GPU sees it.  Every floating-point
// operation is an explicit round-to-nearest IEEE operation on both sides, so the two compilations agree bit for bit.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define PBD_ENV_FN __device__ __forceinline__
#define PBD_ENV_NOINLINE static __device__ __noinline__
#else
#define PBD_ENV_FN inline
#define PBD_ENV_NOINLINE inline
#endif

// statistics hook of the host test build (counts loop iterations per step); expands to nothing in the product
#ifndef PBD_ENV_STAT
#define PBD_ENV_STAT(name)
#endif

namespace pbd {
namespace env {

constexpr int kRing = 8;          // ring window (stack entries kept in shared memory per lane)
constexpr int kRcp = 32;          // reciprocals 1/(2a*dd) tabulated for sample distances dd < kRcp
constexpr int kTabPad = 2;        // the two look-ahead entries past the largest offset a line can produce

#if defined(__CUDA_ARCH__)
PBD_ENV_FN double dadd(double a, double b) { return __dadd_rn(a, b); }
PBD_ENV_FN double dsub(double a, double b) { return __dsub_rn(a, b); }
PBD_ENV_FN double dmul(double a, double b) { return __dmul_rn(a, b); }
PBD_ENV_FN double dfma(double a, double b, double c) { return __fma_rn(a, b, c); }
PBD_ENV_FN double ddiv(double a, double b) { return __ddiv_rn(a, b); }
PBD_ENV_FN double drcp(double a) { return __drcp_rn(a); }
PBD_ENV_FN int f2i_floor(float v) { return __float2int_rd(v); }           // saturating, NaN -> 0
PBD_ENV_FN int dlo(double v) { return __double2loint(v); }
PBD_ENV_FN int dhi(double v) { return __double2hiint(v); }
// base[idx] with a 32-bit index: one IMAD.WIDE forms the address
PBD_ENV_FN double ld_table(const double* base, int idx) {
  double v;
  asm("{ .reg .u64 a; mad.wide.s32 a, %2, 8, %1; ld.global.nc.f64 %0, [a]; }" : "=d"(v) : "l"(base), "r"(idx));
  return v;
}
#else
PBD_ENV_FN double dadd(double a, double b) { return a + b; }
PBD_ENV_FN double dsub(double a, double b) { return a - b; }
PBD_ENV_FN double dmul(double a, double b) { return a * b; }
PBD_ENV_FN double dfma(double a, double b, double c) { return std::fma(a, b, c); }
PBD_ENV_FN double ddiv(double a, double b) { return a / b; }
PBD_ENV_FN double drcp(double a) { return 1.0 / a; }
PBD_ENV_FN int f2i_floor(float v) {
  if (v != v) return 0;
  if (v >= 2147483648.f) return INT32_MAX;
  if (v <= -2147483648.f) return INT32_MIN;
  return (int)std::floor(v);
}
PBD_ENV_FN int dlo(double v) { uint64_t u; std::memcpy(&u, &v, 8); return (int)(uint32_t)u; }
PBD_ENV_FN int dhi(double v) { uint64_t u; std::memcpy(&u, &v, 8); return (int)(uint32_t)(u >> 32); }
PBD_ENV_FN double ld_table(const double* base, int idx) { return base[idx]; }
#endif
PBD_ENV_FN int imin(int a, int b) { return a < b ? a : b; }
PBD_ENV_FN int imax(int a, int b) { return a > b ? a : b; }

struct Quad {
  double a, b, a2;     // a2 = 2*a (exact), reference evaluates 2*a*(x1-x0) left to right
  double r1;           // correctly rounded 1/(2a): reciprocal of the divisor for adjacent samples (x1-x0 = 1)
  const double* E;     // E[x] = a*x^2 + b*x (both products and the sum rounded as the reference does), x = pos - v
  const double* Rcp;   // Rcp[dd] = correctly rounded 1/RN(2a*dd), 1 <= dd < kRcp
};
// Quadratic(-w[0], -w[1]), reference src/DynamicProgram.cpp:126-127
PBD_ENV_FN Quad make_quad(float w_sq, float w_lin, const double* E, const double* Rcp) {
  Quad f;
  f.a = (double)(-w_sq);
  f.b = (double)(-w_lin);
  f.a2 = dmul(2.0, f.a);
  f.r1 = drcp(f.a2);
  f.E = E;
  f.Rcp = Rcp;
  return f;
}
// table entries (dt_build_tables on the device, the host test fills its own copy with the same functions)
PBD_ENV_FN double table_E(double a, double b, int x) { return dadd(dmul(a, (double)(x * x)), dmul(b, (double)x)); }
PBD_ENV_FN double table_rcp(double a, int dd) { return dd > 0 ? drcp(dmul(dmul(2.0, a), (double)dd)) : 0.0; }

// the rare exact quotient: kept out of line so that the common path does not carry the division's instructions
PBD_ENV_NOINLINE float quotient_exact(double num, double den) { return (float)ddiv(num, den); }

// (float)(num/den) given r = RN(1/den): q1 is the quotient after one Markstein correction step, within 1 ulp(double) of the
// correctly rounded num/den, so its float rounding equals the reference's double-division-then-float unless it lies within
// 2 ulp of a float rounding boundary (probability ~2^-27): those cases, and anything outside 2^-100..2^100 (zero, infinities,
// NaN and every intermediate under/overflow included), take the exact division.
PBD_ENV_FN float quotient_to_float(double num, double den, double r) {
  const double q0 = dmul(num, r);
  const double q1 = dfma(dfma(-den, q0, num), r, q0);
  const int lo = dlo(q1) & 0x1FFFFFFF;
  const unsigned ex = ((unsigned)dhi(q1) & 0x7ff00000u) - ((1023u - 100u) << 20);
  const int d = lo - 0x10000000;
  if (ex <= (200u << 20) && (d > 2 || d < -2)) return (float)q1;
  return quotient_exact(num, den);
}

// Quadratic::operator()(x0,x1,y0,y1), include/DistanceTransform.hpp:98-100, result rounded to float as `T s = f(...)`,
// for adjacent samples x1 = x0 + 1 (the first intersection of every step: the previous sample is always the top).
// b*1 = b and a*(x1^2-x0^2) = a*(2*x1-1) are the reference's own values.
PBD_ENV_FN float isect_adjacent(const Quad& f, int x1, double y0, double y1) {
  const double num = dadd(dsub(dsub(y1, y0), f.b), dmul(f.a, (double)(2 * x1 - 1)));
  return quotient_to_float(num, f.a2, f.r1);
}
// the general case (after a pop x1 - x0 >= 2): ((y1 - y0) - b (x1 - x0) + a (x1^2 - x0^2)) / (2 a (x1 - x0)), every operation
// rounded as the reference's double expression (the integer factors are exact in double); the divisor's reciprocal is
// tabulated for x1 - x0 < kRcp
PBD_ENV_FN float isect_far(const Quad& f, int x0, int x1, double y0, double y1) {
  const int ddi = x1 - x0;
  const double dd = (double)ddi;
  const double t = dsub(dsub(y1, y0), dmul(f.b, dd));
  const double num = dadd(t, dmul(f.a, (double)(ddi * (x1 + x0))));
  const double den = dmul(f.a2, dd);
  if (ddi < kRcp) return quotient_to_float(num, den, ld_table(f.Rcp, ddi));
  return quotient_exact(num, den);
}

// per-warp shared-memory ring: [slot][lane]; vp = v | (v of the entry below << 16), 0xFFFF = none
struct Ring {
  float z[kRing][32];
  float y[kRing][32];
  unsigned int vp[kRing][32];
};

// One lane's 1-D transform.  loady(q) = src[q] (called for q = 0..N-1 in order, in lock step across the warp); reload(v) =
// src[v] for deep-pop reloads; emit(i, val, v) records dst[i] = val, ptr[i] = v for the position index i = pos - os0 in [0, N)
// (may be called more than once for an index; the last call wins; a run of calls has consecutive indices and never starts
// beyond the highest index emitted so far + 1); tick(q) is called once per sample step q = 1..N-1, in lock step across the warp.
//
// Eager emission: when sample q is pushed with break point s, the previous top P (still in registers) owns exactly the
// integer positions z_P < pos <= s, and they are evaluated and stored at once.  If q (or P) is popped later, the positions
// are simply stored again by their new owner: every position's FINAL owner E_i emits when its final successor E_{i+1} is
// pushed, and since break points increase up the stack no later emission (all above E_{i+1}) can touch a position
// <= z_{E_{i+1}}, so the last store to every position is the reference's scan result (:171-181).
//
// The first position `lo` the top will emit is known as soon as the top is (floor of its own break point + 1), so the two
// table entries most emissions need (a step emits 1.0 positions on average, rarely more than 2) are loaded one step ahead
// (e0, e1): their latency hides behind the next step's intersection arithmetic.
//
// The ring holds the newest 8 stack entries for pops (2-3 % of the lane steps); zb/pb are the backing store of the whole
// envelope as a linked list threaded through the sample index (zb[q] = break point of the parabola pushed at q,
// pb[q] = the sample below it), written in lock step across lanes (coalesced) and read only by pops deeper than the ring.
template <typename LoadY, typename Reload, typename Emit, typename Tick>
PBD_ENV_FN void envelope_stream(int N, const Quad& f, int os0, Ring& R, int lane, float* zb, unsigned short* pb,
                                LoadY loady, Reload reload, Emit emit, Tick tick) {
  const int pos_last = os0 + N - 1;
  // positions lo..hi of parabola v (value y): Quadratic::operator()(pos - v, y), :102-104, = (a x^2 + b x) + y with the
  // parenthesis taken from the map's table; e0 / e1 are the table entries of lo and lo + 1
  auto emit_run = [&](int lo, int hi, int v, double yd, double e0, double e1) {
    if (hi >= lo) {
      int i = lo - os0;
      emit(i, (float)dadd(e0, yd), v);
      if (hi > lo) {
        emit(++i, (float)dadd(e1, yd), v);
        int x = lo - v + 2;
#pragma unroll 1
        for (int pos = lo + 2; pos <= hi; ++pos, ++x) emit(++i, (float)dadd(ld_table(f.E, x), yd), v);
      }
    }
  };
  int k = 0, base = 0;                                            // stack depth of the top; lowest depth still valid in the ring
  int vt = 0, pt = 0xFFFF;                                        // the top's sample and the sample below it
  float ytf = loady(0), zt = -INFINITY;
  double yt = (double)ytf;
  R.z[0][lane] = zt; R.y[0][lane] = ytf; R.vp[0][lane] = 0xFFFF0000u;
  zb[0] = zt; pb[0] = 0xFFFF;
  // first position the top owns: max(min(floor(zt), pos_last) + 1, os0) (the float -> int conversion saturates, so the
  // -inf / +inf break points of the bottom and the top need no special case)
  int lo = os0;
  double e0 = ld_table(f.E, lo - vt), e1 = ld_table(f.E, lo - vt + 1);
  for (int q = 1; q < N; ++q) {                                   // :160-170
    tick(q);
    const float yqf = loady(q);
    const double yq = (double)yqf;
    float s = isect_adjacent(f, q, yt, yq);                       // the top is sample q - 1
    if (s <= zt && k > 0) {
      do {
        --k;
        PBD_ENV_STAT(pop)
        const int slot = k & (kRing - 1);
        if (k < base) {                                           // popped below the ring: reload from the backing store
          base = k;
          const int vv = pt;                                      // the entry below the one just popped
          R.vp[slot][lane] = (unsigned)vv | ((unsigned)pb[vv] << 16); R.z[slot][lane] = zb[vv]; R.y[slot][lane] = reload(vv);
        }
        const unsigned vp = R.vp[slot][lane];
        vt = vp & 0xFFFF; pt = vp >> 16; ytf = R.y[slot][lane]; yt = (double)ytf; zt = R.z[slot][lane];
        s = isect_far(f, vt, q, yt, yq);
      } while (s <= zt && k > 0);
      lo = imax(imin(f2i_floor(zt), pos_last) + 1, os0);
      e0 = ld_table(f.E, lo - vt); e1 = ld_table(f.E, lo - vt + 1);
    }
    const int hi = imin(f2i_floor(s), pos_last);
    emit_run(lo, hi, vt, yt, e0, e1);                             // the top's positions up to the new break point
    ++k;
    base = imax(base, k - (kRing - 1));                           // the slot of depth k - kRing is overwritten
    const int slot = k & (kRing - 1);
    R.vp[slot][lane] = (unsigned)q | ((unsigned)vt << 16); R.y[slot][lane] = yqf; R.z[slot][lane] = s;
    zb[q] = s; pb[q] = (unsigned short)vt;
    pt = vt; vt = q; ytf = yqf; yt = yq; zt = s;
    lo = imax(hi + 1, os0);                                       // = max(min(floor(s), pos_last) + 1, os0)
    e0 = ld_table(f.E, lo - q); e1 = ld_table(f.E, lo - q + 1);
  }
  emit_run(lo, pos_last, vt, yt, e0, e1);
}

// The reference's two loops, LITERALLY (include/DistanceTransform.hpp:160-181): build the stack, then scan it.  For lines with a
// non-finite sample: what the reference's comparisons make of a NaN (every test false: nothing is popped, the scan never advances
// past a NaN break point) is an accident of its control flow that the streaming formulations above do not share, so the kernels hand
// such lines to this function (z / v: the lane's N-entry scratch arrays; y(i) = src[i]; emit(i, val, v) once per position index).
template <typename LoadY, typename Emit>
PBD_ENV_FN void envelope_literal(int N, const Quad& f, int os0, float* z, unsigned short* v, LoadY y, Emit emit) {
  int k = 0;
  v[0] = 0; z[0] = -INFINITY;
  for (int q = 1; q < N; ++q) {                                   // :160-170
    const double yq = (double)y(q);
    auto isect = [&](int vk) -> float {
      const double yv = (double)y(vk);
      return q - vk == 1 ? isect_adjacent(f, q, yv, yq) : isect_far(f, vk, q, yv, yq);
    };
    float s = isect((int)v[k]);
    while (s <= z[k] && k > 0) { --k; s = isect((int)v[k]); }
    ++k;
    v[k] = (unsigned short)q; z[k] = s;
  }
  const int top = k;                                              // z[top + 1] = +inf
  k = 0;
  int p = os0;
  for (int q = 0; q < N; ++q, ++p) {                              // :172-181
    while (k < top && z[k + 1] < (float)p) ++k;
    const int vk = (int)v[k];
    emit(q, (float)dadd(ld_table(f.E, p - vk), (double)y(vk)), vk);
  }
}

// ---------------------------------------------------------------------------------------------------------------------------
// envelope_stream with CERTIFIED fp32 break points (round 2).  Only decisions depend on the break points -- the pop test s <= z and
// the integer ranges floor(z) -- the values never leave the lane.  So s is first computed in fp32,
//     adjacent:  s = (yq - yt) * fl(R) + fl(q + C0),                R = 1/(2a), C0 = -1/2 - b R
//     general:   s = (yq - yt) * fl(R/dd) + fl((q + v)/2 + Cb),     Cb = -b R, dd = q - v < kRcp
// together with a bound e = 2^-20 (|product| + |offset| + |constant|) on its distance to the reference's float(double) value (the
// roundings of the difference, the reciprocal, the product, the offset and the sum add up to < 5 * 2^-24 of that magnitude, the
// reference's own float rounding included).  A comparison is taken from the fp32 values only if they differ by more than the sum of
// their bounds, floor(s) only if no integer lies within e of s; anything else is recomputed with the reference's double
// expression (isect_adjacent / isect_far, bound 0), the compared break point of the top included.  Every decision therefore
// equals the reference's; every stored entry carries its bound (RingE::e; entries reloaded from the backing store are
// recomputed exactly).  Non-finite values fail every certificate.
struct RingE {
  float z[kRing][32];
  float y[kRing][32];
  unsigned int vp[kRing][32];
  float e[kRing][32];
};
#if defined(__CUDA_ARCH__)
PBD_ENV_FN float fmul_r(float a, float b) { return __fmul_rn(a, b); }
PBD_ENV_FN float fadd_r(float a, float b) { return __fadd_rn(a, b); }
PBD_ENV_FN float fsub_r(float a, float b) { return __fsub_rn(a, b); }
PBD_ENV_FN float frint(float v) { return rintf(v); }
#else
PBD_ENV_FN float fmul_r(float a, float b) { return a * b; }
PBD_ENV_FN float fadd_r(float a, float b) { return a + b; }
PBD_ENV_FN float fsub_r(float a, float b) { return a - b; }
PBD_ENV_FN float frint(float v) { return std::nearbyintf(v); }
#endif
PBD_ENV_FN bool floor_is_certain(float s, float e) { return fabsf(fsub_r(s, frint(s))) > e; }   // false for NaN / inf / huge

template <typename LoadY, typename Reload, typename Emit>
PBD_ENV_FN void envelope_stream_cert(int N, const Quad& f, int os0, RingE& R, int lane, float* zb, unsigned short* pb, LoadY loady,
                                     Reload reload, Emit emit) {
  const int pos_last = os0 + N - 1;
  const float kEps = 9.5367431640625e-07f;                          // 2^-20
  const float rf = (float)f.r1;                                     // fl(1 / (2a))
  const double cbd = dsub(0.0, dmul(f.b, f.r1));                    // -b / (2a)
  const float cb = (float)cbd, c0 = (float)dsub(cbd, 0.5);
  const float acb = fabsf(cb), ac0 = fabsf(c0);
  auto emit_run = [&](int lo, int hi, int v, double yd, double e0, double e1) {
    if (hi >= lo) {
      int i = lo - os0;
      emit(i, (float)dadd(e0, yd), v);
      if (hi > lo) {
        emit(++i, (float)dadd(e1, yd), v);
        int x = lo - v + 2;
#pragma unroll 1
        for (int pos = lo + 2; pos <= hi; ++pos, ++x) emit(++i, (float)dadd(ld_table(f.E, x), yd), v);
      }
    }
  };
  int k = 0, base = 0;                                            // stack depth of the top; lowest depth still valid in the ring
  int vt = 0, pt = 0xFFFF;                                        // the top's sample and the sample below it
  float ytf = loady(0), zt = -INFINITY, et = 0.f;                 // ... its value, break point and the bound of the break point
  double yt = (double)ytf;
  R.z[0][lane] = zt; R.y[0][lane] = ytf; R.vp[0][lane] = 0xFFFF0000u; R.e[0][lane] = 0.f;
  zb[0] = zt; pb[0] = 0xFFFF;
  // the reference's own break point of the current top (depth k >= 1, sample vt above sample pt)
  auto exact_top = [&]() -> float {
    const float ypf = (k - 1 >= base) ? R.y[(k - 1) & (kRing - 1)][lane] : reload(pt);
    return vt - pt == 1 ? isect_adjacent(f, vt, (double)ypf, yt) : isect_far(f, pt, vt, (double)ypf, yt);
  };
  int lo = os0;
  double e0 = ld_table(f.E, lo - vt), e1 = ld_table(f.E, lo - vt + 1);
  for (int q = 1; q < N; ++q) {                                   // :160-170
    const float yqf = loady(q);
    const double yq = (double)yqf;
    // ---- first intersection: the top is sample q - 1 ----
    const float t = fmul_r(fsub_r(yqf, ytf), rf), qc = fadd_r((float)q, c0);
    float s = fadd_r(t, qc);
    float es = fmul_r(fadd_r(fadd_r(fabsf(t), fabsf(qc)), ac0), kEps);
    if (!(floor_is_certain(s, es) && (k == 0 || fabsf(fsub_r(s, zt)) > fadd_r(es, et)))) {
      PBD_ENV_STAT(exact)
      s = isect_adjacent(f, q, yt, yq); es = 0.f;
      if (k > 0 && et > 0.f && !(fabsf(fsub_r(s, zt)) > et)) { zt = exact_top(); et = 0.f; R.z[k & (kRing - 1)][lane] = zt; R.e[k & (kRing - 1)][lane] = 0.f; }
    }
    if (s <= zt && k > 0) {
      do {
        --k;
        PBD_ENV_STAT(pop)
        const int slot = k & (kRing - 1);
        bool reloaded = false;
        if (k < base) {                                           // popped below the ring: reload from the backing store
          base = k;
          const int vv = pt;                                      // the entry below the one just popped
          R.vp[slot][lane] = (unsigned)vv | ((unsigned)pb[vv] << 16); R.z[slot][lane] = zb[vv]; R.y[slot][lane] = reload(vv); R.e[slot][lane] = 0.f;
          reloaded = true;
        }
        const unsigned vp = R.vp[slot][lane];
        vt = vp & 0xFFFF; pt = vp >> 16; ytf = R.y[slot][lane]; yt = (double)ytf; zt = R.z[slot][lane]; et = R.e[slot][lane];
        if (reloaded && k > 0) { zt = exact_top(); R.z[slot][lane] = zt; }   // the backing store keeps no bound: recompute the break point
        const int dd = q - vt;
        bool certain = false;
        if (dd < kRcp) {
          const float tt = fmul_r(fsub_r(yqf, ytf), (float)ld_table(f.Rcp, dd)), m = fadd_r(0.5f * (float)(q + vt), cb);
          s = fadd_r(tt, m);
          es = fmul_r(fadd_r(fadd_r(fabsf(tt), fabsf(m)), acb), kEps);
          certain = floor_is_certain(s, es) && (k == 0 || fabsf(fsub_r(s, zt)) > fadd_r(es, et));
        }
        if (!certain) {
          PBD_ENV_STAT(exact)
          s = isect_far(f, vt, q, yt, yq); es = 0.f;
          if (k > 0 && et > 0.f && !(fabsf(fsub_r(s, zt)) > et)) { zt = exact_top(); et = 0.f; R.z[slot][lane] = zt; R.e[slot][lane] = 0.f; }
        }
      } while (s <= zt && k > 0);
      lo = imax(imin(f2i_floor(zt), pos_last) + 1, os0);
      e0 = ld_table(f.E, lo - vt); e1 = ld_table(f.E, lo - vt + 1);
    }
    const int hi = imin(f2i_floor(s), pos_last);
    emit_run(lo, hi, vt, yt, e0, e1);                             // the top's positions up to the new break point
    ++k;
    base = imax(base, k - (kRing - 1));                           // the slot of depth k - kRing is overwritten
    const int slot = k & (kRing - 1);
    R.vp[slot][lane] = (unsigned)q | ((unsigned)vt << 16); R.y[slot][lane] = yqf; R.z[slot][lane] = s; R.e[slot][lane] = es;
    zb[q] = s; pb[q] = (unsigned short)vt;
    pt = vt; vt = q; ytf = yqf; yt = yq; zt = s; et = es;
    lo = imax(hi + 1, os0);                                       // = max(min(floor(s), pos_last) + 1, os0)
    e0 = ld_table(f.E, lo - q); e1 = ld_table(f.E, lo - q + 1);
  }
  emit_run(lo, pos_last, vt, yt, e0, e1);
}

// ---------------------------------------------------------------------------------------------------------------------------
// Lagged-scan variant (an experiment for the next round; not used by dt_pass yet).  Same stack construction as envelope_stream,
// but instead of emitting the top's whole range when its successor is pushed (and again after every pop), every step emits exactly
// the positions up to q - LAG through a cursor that walks up the stack: in steady state ONE position per lane per step, the same
// position index in all lanes of a warp (coalesced stores, no repeated stores).  The invariant is "every position below `pe` has
// been emitted according to the CURRENT stack"; a push or pop that changes the owner of an already emitted position (a break point
// more than LAG positions behind its sample: rare) rewinds `pe` and the cursor, and the lane catches up in the same step.  Emission
// order is ascending except after a rewind; the last emission of an index is the reference's scan result, as in envelope_stream.
template <int LAG, typename LoadY, typename Reload, typename Emit>
PBD_ENV_FN void envelope_scan(int N, const Quad& f, int os0, Ring& R, int lane, float* zb, unsigned short* pb, LoadY loady, Reload reload,
                              Emit emit) {
  const int pos_last = os0 + N - 1;
  int k = 0, base = 0;                                            // stack depth of the top; lowest depth still valid in the ring
  int vt = 0, pt = 0xFFFF;                                        // the top's sample and the sample below it
  float ytf = loady(0), zt = -INFINITY;
  double yt = (double)ytf;
  R.z[0][lane] = zt; R.y[0][lane] = ytf; R.vp[0][lane] = 0xFFFF0000u;
  zb[0] = zt; pb[0] = 0xFFFF;
  // cursor: `pe` = next position to emit; (kc, vc, yc) = a stack entry with z[kc] < pe; zhi = z[kc + 1] (+inf for the top)
  int pe = os0, kc = 0, vc = 0;
  double yc = yt;
  float zhi = INFINITY;
  auto first_after = [&](float z) { return imax(imin(f2i_floor(z), pos_last) + 1, os0); };   // first position > z, clipped (saturating)
  auto emit_upto = [&](int target) {
#pragma unroll 1
    while (pe <= target) {
#pragma unroll 1
      while (zhi < (float)pe) {                                   // :176 `while (z[k+1] < q) k++`
        ++kc;
        PBD_ENV_STAT(adv)
        if (kc == k) { vc = vt; yc = yt; zhi = INFINITY; }
        else if (kc >= base) {
          const int slot = kc & (kRing - 1);
          vc = R.vp[slot][lane] & 0xFFFF; yc = (double)R.y[slot][lane]; zhi = R.z[(kc + 1) & (kRing - 1)][lane];
        } else {                                                  // the cursor is below the ring (rare): walk down the backing store
          int d = base;
          const unsigned vpw = R.vp[base & (kRing - 1)][lane];
          int above = vpw & 0xFFFF, below = vpw >> 16;             // samples at depth d and d - 1
#pragma unroll 1
          while (d > kc + 1) { above = below; below = pb[above]; --d; }
          vc = below; yc = (double)reload(vc); zhi = zb[above];
        }
      }
      emit(pe - os0, (float)dadd(ld_table(f.E, pe - vc), yc), vc);
      ++pe;
    }
  };
  for (int q = 1; q < N; ++q) {                                   // :160-170
    const float yqf = loady(q);
    const double yq = (double)yqf;
    float s = isect_adjacent(f, q, yt, yq);                       // the top is sample q - 1
    if (s <= zt && k > 0) {
      do {
        --k;
        PBD_ENV_STAT(pop)
        const int slot = k & (kRing - 1);
        if (k < base) {                                           // popped below the ring: reload from the backing store
          base = k;
          const int vv = pt;
          R.vp[slot][lane] = (unsigned)vv | ((unsigned)pb[vv] << 16); R.z[slot][lane] = zb[vv]; R.y[slot][lane] = reload(vv);
        }
        const unsigned vp = R.vp[slot][lane];
        vt = vp & 0xFFFF; pt = vp >> 16; ytf = R.y[slot][lane]; yt = (double)ytf; zt = R.z[slot][lane];
        s = isect_far(f, vt, q, yt, yq);
      } while (s <= zt && k > 0);
      // everything above the new top T is gone: positions beyond z_T have to be (re-)emitted by T or by q
      if (kc > k) { kc = k; vc = vt; yc = yt; }
      pe = imin(pe, first_after(zt));
    }
    ++k;                                                          // push q
    base = imax(base, k - (kRing - 1));
    const int slot = k & (kRing - 1);
    R.vp[slot][lane] = (unsigned)q | ((unsigned)vt << 16); R.y[slot][lane] = yqf; R.z[slot][lane] = s;
    zb[q] = s; pb[q] = (unsigned short)vt;
    if (kc == k - 1) zhi = s;                                     // the cursor's entry is the old top: its range now ends at s
    pe = imin(pe, first_after(s));                                // positions beyond s belong to q from now on
    pt = vt; vt = q; ytf = yqf; yt = yq; zt = s;
    emit_upto(imin(q - LAG, pos_last));
  }
  emit_upto(pos_last);
}

// Write-back window between the eager emission and global memory.  The lanes of a warp run in lock step over the samples, but the
// positions they emit at a given step differ by a few (and some are emitted twice), so direct stores hit 3-5 different 128-byte lines
// per instruction -- measured as 30 % of dt_pass.  Each lane therefore parks its last W emissions in a shared-memory column and every
// step writes back exactly the index (q - D) - os0: the 32 lanes of a warp then store 32 consecutive floats (their lines are
// consecutive), one line per instruction.  Purely lane-local bookkeeping: `hi` = highest index emitted, `done` = highest index written
// back; an emission at or below `done` (a late rewrite after a deep pop, or a lane that lags) goes to global memory directly.
template <int W, int D>
struct OutWindow {
  float* sval;              // this lane's column of the window: slot s at sval[s * 32]
  unsigned short* sptr;
  int hi, done;
  PBD_ENV_FN void init(float* sv, unsigned short* sp) { sval = sv; sptr = sp; hi = -1; done = -1; }
  template <typename Store>
  PBD_ENV_FN void write_back_to(int upto, Store store) {
#pragma unroll 1
    while (done < upto) {
      ++done;
      const int s = done & (W - 1);
      store(done, sval[s * 32], sptr[s * 32]);
    }
  }
  template <typename Store>
  PBD_ENV_FN void put(int i, float val, int v, Store store) {
    if (i <= done) { store(i, val, (unsigned short)v); return; }
    if (i > done + W) write_back_to(i - W, store);            // a burst longer than the window (every index up to hi >= i - 1 is valid)
    const int s = i & (W - 1);
    sval[s * 32] = val; sptr[s * 32] = (unsigned short)v;
    hi = imax(hi, i);
  }
  template <typename Store>
  PBD_ENV_FN void step(int q, int os0, Store store) { write_back_to(imin(q - D - os0, hi), store); }
  template <typename Store>
  PBD_ENV_FN void finish(Store store) { write_back_to(hi, store); }
};

}  // namespace env
}  // namespace pbd
