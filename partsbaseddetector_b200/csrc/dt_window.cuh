// dt_window.cuh -- windowed, CERTIFIED evaluation of the 1-D generalised distance transform (DistanceTransform<float>::computeRow,
// reference include/DistanceTransform.hpp:152-182) at one output position, shared by the device kernels (dt_window.cu) and the host
// harness (tests/dt_window_host.cpp, plain g++ with -ffp-contract=off, compared with the oracle on the CPU).
//
// What the reference computes.  With a = -w0 < 0 its stack algorithm builds the UPPER envelope of the parabolas
// f_v(p) = a (p-v)^2 + b (p-v) + y_v and reports, for position p = os + q, the parabola v[k] with z[k] < p <= z[k+1], where every
// break point z is the double expression of Quadratic::operator()(x0,x1,y0,y1) (:98-100) rounded to float.  Near-ties are decided
// by those rounded values, which is why dt_pass replays the stack literally.  Away from near-ties the result is simply the arg-max
// over v, and "away" can be decided locally:
//
//   Lemma.  Let delta bound the distance between any computed (float) break point near p and the real intersection abscissa.  If
//   f_v*(p) - f_u(p) > 2|a| |u - v*| delta for every sample u != v*, the reference reports v* at p.
//   Proof.  f_w - f_u (w > u) is linear in p with slope 2|a|(w-u), so the real intersection of v* with any u lies more than delta
//   from p, on the far side of p.  v* is pushed when its sample is reached; it is popped by q only if fl(s(v*,q)) <= z = fl(s(u,v*))
//   for the entry u below it at that time, but fl(s(v*,q)) > p > fl(s(u,v*)) -- never.  In the final stack (break points strictly
//   increasing) z[k*] = fl(s(pred,v*)) < p and z[k*+1] = fl(s(v*,succ)) > p (or +inf), so the scan (:171-176) stops at k*.
//
//   The margin need only be checked inside a window: if for EVERY position p of a line the best sample within |v - p| <= W leads
//   every other sample of that window by the margin and is not at the window's edge (|v* - p| <= W-1; the edge facing outwards is
//   allowed at the first / last position), and |os| <= W, then it leads every sample of the line.  (Consecutive owners
//   v(p) <= v(p+1) are in each other's windows, so adjacent envelope members beat each other with margin at their boundary positions; a sample between two consecutive owners is in the windows of both
//   boundary positions; differences being linear in p, the margins add up along the chain of owners.  Samples before the first /
//   after the last owner are inside the first / last position's window because |os| <= W.)  Maps with a larger anchor keep dt_pass.
//
// delta.  The float spacing at |p| <= Pmax is d32 = 2^(floor(log2 Pmax) - 23): a real number more than d32 below (above) the integer p
// cannot round to p or beyond.  The double expression itself is within e64 <= 1.01 * 2^-53 (5 Ymax/|a| + 2.5 |b|/|a| + 4 N) of the real
// abscissa (five roundings in the numerator, divisor and quotient one each); samples are required to satisfy |y| <= ylim, chosen so
// that e64 <= d32 / 4, and delta = 1.25 d32.
//
// Tier 1 (fp32, every position): c_j = fl(y_j + fl(E_j)) is within 2^-24 (|E_j| + |c_j|) of f_j(p); the position is accepted when
// exactly one c_j lies at or above  best - tau,  tau = tau0 + 2^-21 |best|,  tau0 = 1.01 (4 |a| W delta) + 2^-22 max|E_j|  (the worst
// margin 2|a| 2W delta plus both evaluation errors plus the rounding of the threshold itself).  Tier 2 (double, the few positions
// tier 1 leaves open): f_j = E_j + y_j in double, accepted when the best leads candidate j by more than
// 1.01 (2 |a| delta) |j - j*| + 2^-50 (|best| + max|E_j|).  A position neither tier accepts marks its LINE for the literal stack
// algorithm (dt_fix in dt_window.cu); nothing is ever guessed.
//
// The reported value is (float)(E[x] + (double)y_v) with E[x] = a x^2 + b x formed like the reference (:102-104): the same expression
// dt_pass evaluates.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>

#include "dt_envelope.cuh"

namespace pbd {
namespace dtw {

constexpr int kWMax = 8;                       // largest half window the parameter block is sized for

// One map, one direction.  Candidate j in [0, 2W] of position p is sample v = p - W + j, i.e. x = p - v = W - j.
struct WinParams {
  double ed[2 * kWMax + 1];                    // E[W - j] = a x^2 + b x, rounded like the reference (env::table_E)
  double margin1;                              // tier 2: required lead per unit of sample distance
  double cmax;                                 // max |ed[j]|
  float ef[2 * kWMax + 1];                     // (float)ed[j]
  float tau0, ylim;
  int W;
  int ok;                                      // 0: this map / direction cannot use the window (a >= 0, |os| > W, non-finite weights, lines too long for the bound)
};

// Host-side construction (the engine builds the table once per model geometry; tests/dt_window_host.cpp uses the same function).
inline WinParams make_params(float w_sq, float w_lin, int os, int N, int W) {
  WinParams P;
  std::memset(&P, 0, sizeof(P));
  P.W = W;
  const double a = (double)(-w_sq), b = (double)(-w_lin);
  const double A = -a;
  bool ok = W >= 1 && W <= kWMax && N >= 1 && a < 0.0 && std::isfinite(a) && std::isfinite(b) && (os < 0 ? -os : os) <= W;
  if (!ok) return P;
  double cmax = 0.0;
  for (int j = 0; j <= 2 * W; ++j) {
    const int x = W - j;                                             // env::table_E's expression; the volatiles keep a host compiler from fusing
    volatile double t1 = a * (double)(x * x);
    volatile double t2 = b * (double)x;
    P.ed[j] = t1 + t2;
    P.ef[j] = (float)P.ed[j];
    cmax = std::fmax(cmax, std::fabs(P.ed[j]));
  }
  int pmax = std::max(std::max(os < 0 ? -os : os, std::abs(os + N - 1)), 1);
  int e = 0;
  std::frexp((double)pmax, &e);                // pmax = m 2^e, m in [0.5, 1): floor(log2 pmax) = e - 1
  const double d32 = std::ldexp(1.0, e - 1 - 23);
  const double delta = 1.25 * d32;
  const double u = std::ldexp(1.0, -53);
  const double ylim = (0.25 * d32 / (1.01 * u) - 2.5 * std::fabs(b) / A - 4.0 * N) * A / 5.0;
  if (!(ylim > 1.0)) return P;
  P.ylim = (float)std::fmin(ylim * 0.999, 1e30);
  P.cmax = cmax;
  P.margin1 = 1.01 * 2.0 * A * delta;
  const double tau0 = 1.01 * 4.0 * A * W * delta + 1.01 * std::ldexp(cmax, -22) + 1e-37;
  P.tau0 = std::nextafterf((float)tau0, INFINITY);
  if (!std::isfinite(P.tau0) || !std::isfinite(cmax)) return P;
  P.ok = 1;
  return P;
}

// NaN-propagating maximum (a NaN sample must refuse the position: fmaxf would silently drop it) and the counting step of tier 1.
// On the device the count is a float accumulated with a PREDICATED FADD: the walk is bound by the integer/compare (ALU) pipe, which
// issues at half the rate of the FP32 pipe the FADD goes to (256 + j is exact in a float; eleven terms stay far below 2^24).
#if defined(__CUDA_ARCH__)
PBD_ENV_FN float max3_nan(float a, float b, float c) { float d; asm("max.NaN.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
template <int K> PBD_ENV_FN void count_ge(float& acc, float c, float thr) {
  asm("{ .reg .pred p; setp.ge.f32 p, %1, %2; @p add.rn.f32 %0, %0, %3; }" : "+f"(acc) : "f"(c), "f"(thr), "f"((float)K));
}
#else
PBD_ENV_FN float max3_nan(float a, float b, float c) {
  if (a != a || b != b || c != c) return NAN;
  return fmaxf(a, fmaxf(b, c));
}
template <int K> PBD_ENV_FN void count_ge(float& acc, float c, float thr) { if (c >= thr) acc += (float)K; }
#endif

// counting step that also carries the winner's SAMPLE: ysel starts at -0.0f and y is added under the same predicate, so that after
// a count of exactly one candidate ysel == y of that candidate, bit for bit (-0 + y = y for every y, zeros of either sign included)
#if defined(__CUDA_ARCH__)
template <int K> PBD_ENV_FN void count_ge_y(float& acc, float& ysel, float c, float thr, float y) {
  asm("{ .reg .pred p; setp.ge.f32 p, %2, %3; @p add.rn.f32 %0, %0, %4; @p add.rn.f32 %1, %1, %5; }"
      : "+f"(acc), "+f"(ysel) : "f"(c), "f"(thr), "f"((float)K), "f"(y));
}
#else
template <int K> PBD_ENV_FN void count_ge_y(float& acc, float& ysel, float c, float thr, float y) { if (c >= thr) { acc += (float)K; ysel += y; } }
#endif
template <int W, int LO, int J, int JEND>
PBD_ENV_FN void count_range(const float (&c)[2 * W + 1], const float (&y)[JEND - LO + 1], float thr, float& a0, float& a1, float& y0, float& y1) {
  if constexpr (J <= JEND) {
    if constexpr (J & 1) count_ge_y<0x100 + J>(a1, y1, c[J], thr, y[J - LO]); else count_ge_y<0x100 + J>(a0, y0, c[J], thr, y[J - LO]);
    count_range<W, LO, J + 1, JEND>(c, y, thr, a0, a1, y0, y1);
  }
}

// Tier 1.  c[j] = fl(y_j + ef[j]) (-inf where the sample does not exist), yin[k] = the sample of inner candidate j = W - RIN + k.
// Returns whether the position is certified, the certified candidate j and its sample yv (both meaningless otherwise).
// The candidates within RIN samples of the position (j in [W - RIN, W + RIN]; on score maps 99.9 % of the owners: measured offsets
// 0 / 1 / 2 / 3 = 75 % / 25 % / 0.2 % / 0.03 %) are counted one by one against the threshold; the candidates of the outer ring only have
// to lie BELOW it, so one NaN-propagating maximum and one comparison stand for their 2 (W - RIN) compare-and-count pairs.  The
// certificate is the same statement as before -- exactly one candidate of the whole window at or above best - tau -- restricted to
// owners of the inner ring; an owner in the outer ring leaves the position to tier 2 (which looks at the whole window again).
struct Pick { bool ok; int j; float yv; };
template <int W, int RIN = (W > 2 ? 2 : W)>
PBD_ENV_FN Pick pick_walk(const float (&c)[2 * W + 1], const float (&yin)[2 * RIN + 1], float tau0, float ylim) {
  static_assert(RIN >= 1 && RIN <= W, "inner ring inside the window");
  constexpr int LO = W - RIN, HI = W + RIN;
  float best = c[LO];
#pragma unroll
  for (int j = LO + 1; j + 1 <= HI; j += 2) best = max3_nan(best, c[j], c[j + 1]);     // 2 RIN + 1 is odd: pairs after c[LO]
  const float thr = env::fsub_r(env::fsub_r(best, tau0), env::fmul_r(fabsf(best), 4.76837158203125e-07f));   // 2^-21
  float a0 = 0.f, a1 = 0.f, y0 = -0.f, y1 = -0.f;
  count_range<W, LO, LO, HI>(c, yin, thr, a0, a1, y0, y1);
  bool outer_below = true;
  if constexpr (RIN < W) {
    // outer ring: j in [0, LO) and (HI, 2W], an even number of candidates
    float o[2 * (W - RIN)];
#pragma unroll
    for (int j = 0; j < LO; ++j) { o[j] = c[j]; o[LO + j] = c[HI + 1 + j]; }
    float mo = o[0];
    int i = 1;
#pragma unroll
    for (; i + 1 < 2 * (W - RIN); i += 2) mo = max3_nan(mo, o[i], o[i + 1]);
    mo = max3_nan(mo, o[2 * (W - RIN) - 1], o[2 * (W - RIN) - 1]);                     // the count is even: one candidate is left over
    outer_below = mo < thr;                                                            // false for a NaN candidate or a NaN threshold
  }
  // exactly one inner candidate at or above the threshold <=> the sum is 256 + j; none: 0; two or more: >= 512.  A NaN or too large
  // best fails the last test (and a NaN threshold counts nothing).
  Pick r;
  r.j = (int)env::fadd_r(a0, a1) - 0x100;
  r.ok = (unsigned)(r.j - LO) <= (unsigned)(2 * RIN) && outer_below && fabsf(best) <= ylim;
  r.yv = env::fadd_r(y0, y1);
  return r;
}
// the certified candidate or a negative number (host harness, tests)
template <int W, int RIN = (W > 2 ? 2 : W)>
PBD_ENV_FN int pick(const float (&c)[2 * W + 1], float tau0, float ylim) {
  float yin[2 * RIN + 1];
  for (int k = 0; k <= 2 * RIN; ++k) yin[k] = 0.f;
  const Pick r = pick_walk<W, RIN>(c, yin, tau0, ylim);
  return r.ok ? r.j : -1;
}

// Tier 2.  y[j] = the window's samples (-inf where none).  Returns the certified candidate or -1.
template <int W>
PBD_ENV_FN int pick_exact(const float (&y)[2 * W + 1], const double* ed, double margin1, double cmax, float ylim) {
  double f[2 * W + 1];
  double best = -INFINITY;
  int jb = -1;
#pragma unroll
  for (int j = 0; j <= 2 * W; ++j) {
    f[j] = env::dadd(ed[j], (double)y[j]);
    if (f[j] != f[j]) return -1;                                                       // a NaN sample in the window
    if (f[j] > best) { best = f[j]; jb = j; }
  }
  if (jb < 0 || !(fabs(best) <= (double)ylim)) return -1;
  const double slack = env::dmul(env::dadd(fabs(best), cmax), 8.8817841970012523e-16);                        // 2^-50
#pragma unroll
  for (int j = 0; j <= 2 * W; ++j) {
    if (j == jb) continue;
    const int dist = j > jb ? j - jb : jb - j;
    if (!(env::dsub(best, f[j]) > env::dadd(env::dmul(margin1, (double)dist), slack))) return -1;
  }
  return jb;
}

// Tier 3: LOCAL REPLAY of an isolated open position p0 (neither tier certifies it: a break point lies within delta of p0), or of a
// short run of them.  Its neighbours p0 - 1 and p0 + 1 are certified with owners uL <= uR, so (lemma) both are in the reference's final stack and are never
// popped.  Entries are only ever popped from the top, so whatever lies between uL and uR in the final stack is exactly what the
// reference's loop (:160-170) leaves there when it pushes the samples uL+1 .. uR on top of uL -- a computation that involves nothing
// but those samples and the reference's own break-point expression (env::isect_adjacent / env::isect_far): a pop test against uL
// itself is false by the lemma (at a line's start uL = sample 0, which the loop never pops), and nothing after uR can reach below uR.
// The scan (:171-176) stands at uL when it arrives at p0 (it chose uL for p0 - 1; at the first position it starts at sample 0) and
// advances while the next entry's break point is < p0.  At a line's end the replay simply runs to the last sample.
// The certified positions on either side keep their global margins: uL, uR and every sample between them must lie in the windows of
// BOTH p0 - 1 and p0 + 1 (checked by the caller: uR - (p0-1) <= W and (p0+1) - uL <= W), so the chain argument steps over p0.
constexpr int kLocalMax = 2 * kWMax + 8;
// A run of open positions pa .. pb (indices qa .. qb, usually a single one) between the certified positions pa - 1 (owner uL) and
// pb + 1 (owner uR): one local replay, then the scan advances through the run.  emit(p, owner, y_owner) once per position of the run.
template <typename LoadY, typename Emit>
PBD_ENV_FN void local_owners(const env::Quad& f, int pa, int pb, int uL, int uR, LoadY loady, Emit emit) {
  int v[kLocalMax];
  float z[kLocalMax], y[kLocalMax];
  int k = 0;
  v[0] = uL; y[0] = loady(uL); z[0] = -INFINITY;
  for (int q = uL + 1; q <= uR; ++q) {
    const float yq = loady(q);
    auto isect = [&](int kk) -> float {
      return q - v[kk] == 1 ? env::isect_adjacent(f, q, (double)y[kk], (double)yq) : env::isect_far(f, v[kk], q, (double)y[kk], (double)yq);
    };
    float s = isect(k);
    while (k > 0 && s <= z[k]) { --k; s = isect(k); }
    ++k;
    v[k] = q; y[k] = yq; z[k] = s;
  }
  int kk = 0;
  for (int p = pa; p <= pb; ++p) {
    while (kk < k && z[kk + 1] < (float)p) ++kk;
    emit(p, v[kk], y[kk]);
  }
}
// the caller's conditions for local_owners on the run of position indices qa .. qb of a line of N samples: uL / uR are the owners of
// the certified positions qa - 1 / qb + 1 (at the line's ends: uL = 0 / uR = N - 1, and the anchor must leave the window room)
PBD_ENV_FN bool local_ok(int W, int os, int N, int qa, int qb, int uL, int uR) {
  const int pa = qa + os, pb = qb + os;
  if (uL > uR || uR - uL > kLocalMax - 2) return false;
  if (qa > 0 ? uR - (pa - 1) > W : os > W - 1) return false;         // uL, uR and everything between them inside the windows of BOTH certified
  if (qb < N - 1 ? (pb + 1) - uL > W : -os > W - 1) return false;    // neighbours, so that the chain of margins steps over the run
  return true;
}

// The owner of position index q (0 <= q < N) must also lie in the windows of the neighbouring positions: the window's first candidate
// (j = 0, v = p - W) is acceptable only when there is no next position, the last (j = 2W) only when there is no previous one.
PBD_ENV_FN bool edge_ok(int j, int W, int q, int N) {
  if (j != 0 && j != 2 * W) return true;                             // the common case: one compare pair
  return j == 0 ? q == N - 1 : q == 0;
}

// the reference's value of position p for the certified sample: Quadratic::operator()(p - v, y_v), :102-104, rounded to float (:177)
PBD_ENV_FN float value_of(double ed_j, float y) { return (float)env::dadd(ed_j, (double)y); }

}  // namespace dtw
}  // namespace pbd
