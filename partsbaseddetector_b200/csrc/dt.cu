// dt.cu -- the max-sum tree DP of DynamicProgram<float>::min (reference src/DynamicProgram.cpp:67-173) over the
// Felzenszwalb generalised distance transform DistanceTransform<float> (reference include/DistanceTransform.hpp).
//
// 1-D transform (computeRow, :152-182): dst[q] = max_v src[v] + a (q+os-v)^2 + b (q+os-v), a = -w0 < 0.  The
// reference builds the upper envelope with a stack whose break points z[] are double quotients rounded to
// float; which parabola wins at a near-tie depends on those rounded values, so the literal sequential stack
// algorithm is reproduced, one lane per row (or column), with explicit IEEE round-to-nearest double
// intrinsics so nothing is contracted to FMA.
//
// Streaming formulation (bit-identical to build-then-scan).  In the reference's second loop parabola k is
// chosen for exactly the integer positions pos with z[k] < pos <= z[k+1].  On real score maps almost every
// sample stays on the envelope (measured: stack depth ~0.95 N, 5-8 % pops), so instead of materialising the
// whole stack and re-reading it, the range of the current top is evaluated and stored the moment the next
// sample is pushed (see envelope_stream); a pop just causes the affected positions to be stored again by
// their new owner, and the last store to every position is the one the reference's scan would produce.
// The newest 8 stack entries live in a per-lane shared-memory ring for pops; deeper pops (rare) go to a
// write-mostly backing store in local memory.
//
// 2-D transform (compute, :203-245) = dt_rows (x direction, anchor x) then dt_cols (y direction, anchor y);
// mix_max then forms, per cell and parent mixture,  max_mm(dt[mm] + bias[mm][pm])  (Math::reduceMax,
// include/Math.hpp:149-185), the best-mixture index Ik and parent.score += max (:134-156).  The reference's
// back-pointer composition Iy[y][x] <- Iy[y][Ix[y][x]] (:232-244) and the per-parent-mixture gather
// (Math::reducePickIndex) are NOT materialised: the raw row-pass / column-pass argmaxes are kept per child
// mixture (u16) and composed lazily by the backtrack (backtrack.cu).
#include <algorithm>
#include <cfloat>
#include <type_traits>
#include "kernels.cuh"
#include "dt_envelope.cuh"
#include "dt_lines.cuh"
#include "dt_window.cuh"

namespace pbd {
namespace {

using env::Ring;
using env::Quad;
static_assert(env::kRcp == kDtRcp && env::kTabPad == kDtTabPad, "table layout constants of kernels.cuh and dt_envelope.cuh differ");
constexpr int kPassWarps = 4;
#ifdef PBD_DT_WINDOWED_STORES
constexpr int kOutWindow = 8, kOutLag = 5;  // experiment: emissions parked per lane / write-back lag in samples
#endif
constexpr int kTileW = 16;        // samples per line staged per shared-memory tile

__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// base[idx] stores with a 32-bit index (the compiler otherwise re-derives the 64-bit base from its kernel-parameter
// components at every store)
__device__ __forceinline__ void st_f32(float* base, unsigned idx, float v) {
  asm volatile("{ .reg .u64 a; mad.wide.u32 a, %1, 4, %0; st.global.f32 [a], %2; }" ::"l"(base), "r"(idx), "f"(v) : "memory");
}
__device__ __forceinline__ void st_u16(unsigned short* base, unsigned idx, unsigned short v) {
  asm volatile("{ .reg .u64 a; mad.wide.u32 a, %1, 2, %0; st.global.u16 [a], %2; }" ::"l"(base), "r"(idx), "h"(v) : "memory");
}

// predicated forms (dt_pass_win keeps its walk free of branches)
// value + arg-max of position q, stored only if q < n (unsigned compare: q "negative" while the window fills); one predicate, no branch
__device__ __forceinline__ void st_pair_if_lt(float* vbase, unsigned short* pbase, unsigned idx, float v, unsigned short a, unsigned q, unsigned n) {
  asm volatile(
      "{ .reg .pred p; .reg .u64 a, b; setp.lt.u32 p, %5, %6; mad.wide.u32 a, %2, 4, %0; mad.wide.u32 b, %2, 2, %1;\n"
      "  @p st.global.f32 [a], %3; @p st.global.u16 [b], %4; }" ::"l"(vbase),
      "l"(pbase), "r"(idx), "f"(v), "h"(a), "r"(q), "r"(n)
      : "memory");
}

// ---------------------------------------------------------------------------------------------------
// One pass of the separable transform over every map of a wave: each lane owns one contiguous line of N samples
// of some (map, line); lines of all maps of a level are packed 32 per warp so that small pyramid levels still
// fill their warps.  Input lines are staged through a 32x32 shared-memory tile (coalesced 128-byte loads,
// conflict-free transposed reads); the output is written TRANSPOSED (out[i*nlines + line]), which makes the
// stores of adjacent lanes adjacent in memory and hands the next pass contiguous lines again:
//   rows pass:  in [y][x] (responses / working scores) -> tmp  [x][y], ixdt  [x][y]
//   cols pass:  in tmp [x][y]                           -> val  [y][x], iyraw [y][x]
// ---------------------------------------------------------------------------------------------------
#ifndef PBD_DT_MINBLOCKS
#define PBD_DT_MINBLOCKS 6        // 80 registers: 6 CTAs of 4 warps per SM (7 CTAs at 72 registers measured 9 % slower)
#endif
// SCAN = 0: eager emission, every break point through the reference's double expression -- the detector's kernel;
// SCAN = -1: eager emission with certified fp32 break points (env::envelope_stream_cert; option dt_variant 1).  Bit-identical, but
//            measured 12 % SLOWER on B200 (dt_rows + dt_cols 8.2 vs 7.3 ms per 64 frames): the fp32 path saves ~15 instructions per
//            step, the bound bookkeeping, the extra ring column and the larger loop body (the double path stays as fallback) cost more;
// SCAN > 0: lagged-scan emission with that lag (env::envelope_scan): one store per position unless a late pop rewinds the cursor --
// the variant for rough inputs (white-noise maps: 3.4 instead of 24.6 stores per position), standalone transform impl 3.
template <int MAXN, int SCAN>
__global__ void __launch_bounds__(kPassWarps * 32, PBD_DT_MINBLOCKS)
dt_pass(const PassGeom* __restrict__ pg, const PassMap* __restrict__ maps, int nmaps, const float* __restrict__ inA, size_t strideA,
        const float* __restrict__ inB, size_t strideB, float* __restrict__ out, size_t stride_out, unsigned short* __restrict__ ptr,
        size_t stride_ptr) {
  __shared__ typename std::conditional<SCAN < 0, env::RingE, Ring>::type rings[kPassWarps];
  __shared__ float tiles[kPassWarps][2][32][kTileW + 1];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int w = blockIdx.x * kPassWarps + wib;                          // warp index -> (level, first item)
  int l = 0, nlines = 0, items = 0;
  for (; l < pg->n_levels; ++l) {
    nlines = pg->nlines[l];
    items = nlines * nmaps;
    const int nw = (items + 31) >> 5;
    if (w < nw) break;
    w -= nw;
  }
  if (l >= pg->n_levels) return;                                  // warp-uniform
  const int frame = blockIdx.y;
  const int N = pg->N[l];
  const size_t cell_off = (size_t)pg->cell_off[l];
  const int t0 = w * 32;
  const bool active = t0 + lane < items;
  const int t = active ? t0 + lane : t0;                          // inactive lanes shadow the warp's first item
  const int mi = t / nlines, line = t - mi * nlines;
  const PassMap M = maps[mi];
  const float* src = (M.in_buf ? inB + (size_t)frame * strideB : inA + (size_t)frame * strideA) + M.in_off + cell_off + (size_t)line * N;
  float* dst = out + (size_t)frame * stride_out + M.out_off + cell_off + line;
  unsigned short* dp = ptr + (size_t)frame * stride_ptr + M.ptr_off + cell_off + line;
  const Quad f = env::make_quad(M.w_sq, M.w_lin, M.etab + M.tab_bias, M.etab + (M.tab_len - kDtRcp));
  float zb[MAXN];
  unsigned short pb[MAXN];
  // Input staging: the warp's 32 lines are read kTileW samples at a time into a double-buffered shared-memory tile
  // with cp.async (each lane copies 16 of the 32 x kTileW elements; a line's kTileW samples are one 64-byte run), the
  // next tile being in flight while the current one is consumed.  The main loop asks for q = 0,1,2,... in lock
  // step across the warp; deep-pop reloads of older samples go to global memory.
  const int c_col = lane & (kTileW - 1), c_row0 = lane / kTileW;  // this lane copies rows c_row0 + 2*i
  auto prefetch = [&](int q, int buf) {
#pragma unroll
    for (int i = 0; i < 32 * kTileW / 32; ++i) {
      const int r = c_row0 + i * (32 / kTileW);
      const float* p = (const float*)__shfl_sync(0xffffffffu, (unsigned long long)src, r);
      if (q + c_col < N) cp_async4(&tiles[wib][buf][r][c_col], p + q + c_col);
    }
    cp_async_commit();
  };
  // sequential reads: q = 0, 1, 2, ... in lock step across the warp; a new tile becomes current every kTileW samples
  float chk = 0.f;                                                // fma(y, 0, chk): NaN as soon as one sample is NaN or +-inf
  auto loady = [&](int q) -> float {
    if ((q & (kTileW - 1)) == 0) {
      if (q == 0) prefetch(0, 0);
      cp_async_wait_all();
      __syncwarp();
      if (q + kTileW < N) prefetch(q + kTileW, ((q / kTileW) + 1) & 1);
    }
    const float y = tiles[wib][(q / kTileW) & 1][lane][q & (kTileW - 1)];
    chk = __fmaf_rn(y, 0.f, chk);
    return y;
  };
  // deep-pop reloads of older samples: global memory (the tile that held them may already be refilled)
  auto reload = [&](int v) -> float { return __ldg(src + v); };
  // inactive lanes recompute the warp's first item and store the same values to the same addresses as lane 0
  asm volatile("" : "+l"(dst), "+l"(dp));                        // keep both bases as materialised 64-bit registers
  auto store = [&](int i, float val, unsigned short v) {
    const unsigned off = (unsigned)i * (unsigned)nlines;
    st_f32(dst, off, val); st_u16(dp, off, v);
  };
  const int os0 = M.os;
  // a line with a non-finite sample is redone with the reference's two loops, literally (env::envelope_literal): what its comparisons
  // make of a NaN is not what the streaming formulations make of it
  auto redo_if_not_finite = [&]() {
    if (chk != chk) env::envelope_literal(N, f, os0, zb, pb, reload, [&](int i, float val, int v) { store(i, val, (unsigned short)v); });
  };
  if constexpr (SCAN > 0) {
    env::envelope_scan<(SCAN > 0 ? SCAN : 1)>(N, f, os0, rings[wib], lane, zb, pb, loady, reload,
                                              [&](int i, float val, int v) { store(i, val, (unsigned short)v); });
    redo_if_not_finite();
    return;
  } else if constexpr (SCAN < 0) {
    env::envelope_stream_cert(N, f, os0, rings[wib], lane, zb, pb, loady, reload, [&](int i, float val, int v) { store(i, val, (unsigned short)v); });
    redo_if_not_finite();
    return;
  } else {
#if !defined(PBD_DT_WINDOWED_STORES)
  // every emission straight to global memory (the default, see below)
  env::envelope_stream(N, f, os0, rings[wib], lane, zb, pb, loady, reload,
                       [&](int i, float val, int v) { store(i, val, (unsigned short)v); }, [](int) {});
#else
  // A/B switch (tools/build_variant.sh ... -DPBD_DT_WINDOWED_STORES): emissions through a per-lane shared-memory write-back window
  // (env::OutWindow) so that a warp stores one position index per step, 32 consecutive floats.  Measured on B200 (round 1): the
  // stores do become coalesced, but the window's bookkeeping (~20 more instructions per step in an issue-bound kernel) costs more
  // than the transactions it saves: dt_pass 7.33 -> 9.45 ms per 64 frames.  Kept as an experiment; bit-identical results (GPU parity
  // suite and tests/test_dt_envelope_host.py).
  __shared__ float wval[kPassWarps][kOutWindow][32];
  __shared__ unsigned short wptr[kPassWarps][kOutWindow][32];
  env::OutWindow<kOutWindow, kOutLag> win;
  win.init(&wval[wib][0][lane], &wptr[wib][0][lane]);
  env::envelope_stream(N, f, os0, rings[wib], lane, zb, pb, loady, reload,
                       [&](int i, float val, int v) { win.put(i, val, v, store); }, [&](int q) { win.step(q, os0, store); });
  win.finish(store);
#endif
  redo_if_not_finite();
  }
}

// ---------------------------------------------------------------------------------------------------
// dt_variant 3: the same pass (same lane = line packing, same staging, same transposed output) with the WINDOWED CERTIFIED evaluation
// of dt_window.cuh instead of the stack: a lane walks along its line keeping the last 2W+1 samples in a 16-slot circular register
// window; the position whose window has just been completed, q = s - os - W, is decided by tier 1 (2W+1 adds, a NaN-propagating max
// tree over the five candidates within two samples of the position, their compare-and-count, one more max tree + one compare for the
// six outer candidates: dtw::pick_walk), its value formed with the reference's double add, and stored.  No data-dependent control
// flow except one warp-uniform branch for the rare open positions, all lanes busy.  A lane whose line cannot be certified somewhere
// (near-tie against a rounded break point, arg-max at the window's edge, NaN / inf sample, map not eligible) replays the line afterwards
// with the literal stack algorithm (env::envelope_stream, direct global loads) and overwrites its own stores -- same thread, so no
// ordering question arises (segmented walk: see below).  Real score maps: 0.07 % of the lines are replayed.
// ---------------------------------------------------------------------------------------------------
// The positions the walk left open (tier 1 undecided, or decided by the window's edge), resolved afterwards so that the walk itself is
// straight-line code: tier 2 on the window read again from the line in global memory, then tier 3, the local replay of what is still
// open between its certified neighbours (dt_window.cuh).  olist holds the nd open position indices in increasing order.  Returns false
// if the line has to be replayed as a whole.
constexpr int kOpenCap = 256;
template <int W>
__device__ __noinline__ bool win_resolve_open(const float* __restrict__ src, int N, int os, int nlines, const PassMap& M,
                                              const dtw::WinParams* __restrict__ wp, float* dst, unsigned short* dp, int* olist, int nd,
                                              int sqa, int sqb) {
  const Quad f = env::make_quad(M.w_sq, M.w_lin, M.etab + M.tab_bias, M.etab + (M.tab_len - kDtRcp));
  int left = 0;
  for (int k = 0; k < nd; ++k) {                                  // tier 2
    const int q0 = olist[k], p0 = q0 + os;
    float w[2 * W + 1];
#pragma unroll
    for (int j = 0; j <= 2 * W; ++j) { const int v = p0 - W + j; w[j] = (unsigned)v < (unsigned)N ? __ldg(src + v) : -INFINITY; }
    const int jb = dtw::pick_exact<W>(w, wp->ed, wp->margin1, wp->cmax, wp->ylim);
    if (jb < 0) { olist[left++] = q0; continue; }                 // stays open
    if (!dtw::edge_ok(jb, W, q0, N)) return false;
    float yv = 0.f;
#pragma unroll
    for (int j = 0; j <= 2 * W; ++j) if (j == jb) yv = w[j];
    const int v = p0 - W + jb;
    dst[(size_t)q0 * nlines] = dtw::value_of(env::ld_table(f.E, p0 - v), yv);
    dp[(size_t)q0 * nlines] = (unsigned short)v;
  }
  for (int k = 0; k < left;) {                                    // tier 3: runs of consecutive open positions
    int e = k;
    while (e + 1 < left && olist[e + 1] == olist[e] + 1) ++e;
    const int qa = olist[k], qb = olist[e];
    // the neighbours of the run are certified (walk or tier 2): their owners were stored by this thread -- unless the neighbour belongs
    // to another segment of the line (segmented walk: this thread owns the positions sqa .. sqb-1 only), in which case it is certified
    // here, by tier 2 on its own window
    auto owner_at = [&](int q, int& u) -> bool {
      if (q >= sqa && q < sqb) { u = (int)dp[(size_t)q * nlines]; return true; }
      const int pn = q + os;
      float w[2 * W + 1];
#pragma unroll
      for (int j = 0; j <= 2 * W; ++j) { const int v = pn - W + j; w[j] = (unsigned)v < (unsigned)N ? __ldg(src + v) : -INFINITY; }
      const int jb = dtw::pick_exact<W>(w, wp->ed, wp->margin1, wp->cmax, wp->ylim);
      if (jb < 0 || !dtw::edge_ok(jb, W, q, N)) return false;
      u = pn - W + jb;
      return true;
    };
    int uL = 0, uR = N - 1;
    if (qa > 0 && !owner_at(qa - 1, uL)) return false;
    if (qb < N - 1 && !owner_at(qb + 1, uR)) return false;
    if (!dtw::local_ok(W, os, N, qa, qb, uL, uR)) return false;
    dtw::local_owners(f, qa + os, qb + os, uL, uR, [&](int u) { return __ldg(src + u); }, [&](int p, int v, float yo) {
      dst[(size_t)(p - os) * nlines] = dtw::value_of(env::ld_table(f.E, p - v), yo);
      dp[(size_t)(p - os) * nlines] = (unsigned short)v;
    });
    k = e + 1;
  }
  return true;
}

// the walk's rare path: lists an open position (out of line on purpose: the walk pays one vote and one branch per step for it)
__device__ __noinline__ int win_note_open(int* olist, int nd, int q, bool open) {
  if (open) { olist[min(nd, kOpenCap - 1)] = q; ++nd; }
  return nd;
}

#ifndef PBD_DTW_RIN
#define PBD_DTW_RIN 2             // tier 1 counts the candidates within this many samples of the position one by one; the outer ring only has to lie below the threshold (dt_window.cuh)
#endif
#ifndef PBD_DTW_MINBLOCKS
#define PBD_DTW_MINBLOCKS 5       // 102 registers: the 16-slot window, the 2W+1 table values and the 2W+1 candidates stay in registers
#endif
// SEG = true: the SEGMENTED walk for launches that cannot fill the GPU (a single frame: the pass otherwise lasts as long as one lane needs
// for the longest line of the pyramid, 50 us at VGA).  The windowed decision of a position depends on nothing but its 2W+1 samples, so a
// line is cut into segments of seg_steps - 2W positions, each walked by its own lane over the samples [qa + os - W, qb - 1 + os + W]
// (seg_steps steps, the first 2W of which only fill the window).  Samples that do not exist (beyond either end of the line) are staged
// as kVirt, a finite value below every admissible sample.  Open positions are resolved as before; a neighbour that belongs to another
// segment is certified on the spot (win_resolve_open).  The certificate is a statement about the WHOLE line (the margins of the
// windows chain from position to position, dt_window.cuh), so a line is accepted only if all of its segments are: every lane adds its
// verdict to the line's counter (seg_ctr, one int per line: segments done in the low half, refusals in the high half), and the lane
// that finishes a line last replays the whole line with the stack algorithm if any segment refused it -- ordered after the other
// segments' stores by the fence / atomic pair -- and leaves the counter at zero for the next launch.
constexpr float kVirt = -3.0e38f;
template <int MAXN, int W, bool SEG>
__global__ void __launch_bounds__(kPassWarps * 32, PBD_DTW_MINBLOCKS)
dt_pass_win(const PassGeom* __restrict__ pg, const PassMap* __restrict__ maps, const dtw::WinParams* __restrict__ wps, int nmaps,
            const float* __restrict__ inA, size_t strideA, const float* __restrict__ inB, size_t strideB, float* __restrict__ out, size_t stride_out,
            unsigned short* __restrict__ ptr, size_t stride_ptr, int* __restrict__ counter, int seg_steps, int* __restrict__ seg_ctr, int seg_lines_total) {
  static_assert(2 * W + 1 <= 16, "the circular window has 16 slots");
  // per warp: the double-buffered input tile (the replay's stack ring reuses it)
  constexpr int kTileFloats = 2 * 32 * (kTileW + 1);
  static_assert(sizeof(Ring) <= kTileFloats * sizeof(float), "the replay ring must fit the tile buffers");
  __shared__ __align__(16) float tile_mem[kPassWarps][kTileFloats];
  constexpr int RIN = PBD_DTW_RIN < W ? PBD_DTW_RIN : W;
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float (*tiles)[32][kTileW + 1] = reinterpret_cast<float (*)[32][kTileW + 1]>(tile_mem[wib]);
  int w = blockIdx.x * kPassWarps + wib;                          // warp index -> (level, first item)
  const int seglen = SEG ? seg_steps - 2 * W : 1;                 // positions per segment
  int l = 0, nlines = 0, items = 0, nseg = 1;
  int line_base = 0, lines_total = 0;                             // SEG: index of the level's first line / lines of one frame in seg_ctr
  for (; l < pg->n_levels; ++l) {
    nlines = pg->nlines[l];
    if constexpr (SEG) nseg = (pg->N[l] + seglen - 1) / seglen;
    items = nlines * nmaps * nseg;
    const int nw = (items + 31) >> 5;
    if (w < nw) break;
    w -= nw;
    line_base += nlines * nmaps;
  }
  if (l >= pg->n_levels) return;                                  // warp-uniform
  if constexpr (SEG) lines_total = seg_lines_total;
  const int frame = blockIdx.y;
  const int N = pg->N[l];
  const size_t cell_off = (size_t)pg->cell_off[l];
  const int t0 = w * 32;
  const bool active = t0 + lane < items;
  const int t = active ? t0 + lane : t0;                          // inactive lanes shadow the warp's first item
  int mi, line, qa = 0, qb = N;
  if (SEG) {                                                      // item = (map, segment, line): the lanes of a warp mostly share a segment
    const int per_map = nlines * nseg;
    mi = t / per_map;
    const int r = t - mi * per_map, sg = r / nlines;
    line = r - sg * nlines;
    qa = sg * seglen; qb = min(N, qa + seglen);
  } else {
    mi = t / nlines; line = t - mi * nlines;
  }
  const PassMap M = maps[mi];
  const dtw::WinParams* wp = wps + mi;
  const float* src = (M.in_buf ? inB + (size_t)frame * strideB : inA + (size_t)frame * strideA) + M.in_off + cell_off + (size_t)line * N;
  float* dst = out + (size_t)frame * stride_out + M.out_off + cell_off + line;
  unsigned short* dp = ptr + (size_t)frame * stride_ptr + M.ptr_off + cell_off + line;
  const int os = M.os;
  // step i of the segmented walk reads sample sbase + i and completes the window of position qa + i - 2W
  const int sbase = SEG ? qa + os - W : 0;
  const long long srcw = (long long)src + (long long)sbase * 4;   // address of the (possibly virtual) sample of step 0
  const int lohi = SEG ? (max(0, -sbase) | (min(max(N - sbase, 0), 0xffff) << 8)) : 0;   // steps [lo, hi) have a sample (lo <= 2W < 256)
  const int c_col = lane & (kTileW - 1), c_row0 = lane / kTileW;
  auto prefetch = [&](int q, int buf) {
#pragma unroll
    for (int i = 0; i < 32 * kTileW / 32; ++i) {
      const int r = c_row0 + i * (32 / kTileW);
      if (SEG) {
        const float* p = (const float*)__shfl_sync(0xffffffffu, (unsigned long long)srcw, r);
        const int lh = __shfl_sync(0xffffffffu, lohi, r);
        const int k = q + c_col;
        if (k >= (lh & 0xff) && k < (lh >> 8)) cp_async4(&tiles[buf][r][c_col], p + k);
        else tiles[buf][r][c_col] = kVirt;
      } else {
        const float* p = (const float*)__shfl_sync(0xffffffffu, (unsigned long long)src, r);
        if (q + c_col < N) cp_async4(&tiles[buf][r][c_col], p + q + c_col);
      }
    }
    cp_async_commit();
  };
  const int steps = SEG ? seg_steps : (N + 2 * W + 15) & ~15;
  auto loady = [&](int q) -> float {                              // q = 0, 1, 2, ... in lock step across the warp
    if ((q & (kTileW - 1)) == 0) {
      if (q == 0) prefetch(0, 0);
      cp_async_wait_all();
      __syncwarp();
      if (q + kTileW < (SEG ? steps : N)) prefetch(q + kTileW, ((q / kTileW) + 1) & 1);
    }
    return tiles[(q / kTileW) & 1][lane][q & (kTileW - 1)];
  };
  asm volatile("" : "+l"(dst), "+l"(dp));
  auto store = [&](int i, float val, unsigned short v) {
    const unsigned off = (unsigned)i * (unsigned)nlines;
    st_f32(dst, off, val); st_u16(dp, off, v);
  };
  bool refused = wp->ok == 0;
  float ef[2 * W + 1];
#pragma unroll
  for (int j = 0; j <= 2 * W; ++j) ef[j] = wp->ef[j];
  const float tau0 = wp->tau0, ylim = wp->ylim;
  const double* ed = wp->ed;
  float buf[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) buf[k] = -INFINITY;
  float chk = 0.f;                                                // fma(y, 0, chk): NaN as soon as one sample is NaN or +-inf (they refuse the line)
  // sample index s = 0 .. N-1, then 2W virtual -inf samples flush the window; no branch inside a step but the rare-path vote
  int nd = 0;                                                     // positions left open by the walk, listed in olist (decided afterwards)
  int olist[kOpenCap];
  // q * nlines of the current step's position (wraps while q is outside the line / segment: never stored)
  unsigned off = (unsigned)(SEG ? qa - 2 * W : -os - W) * (unsigned)nlines;
  const int q0 = SEG ? qa - 2 * W : -os - W;                      // position of step 0
  // positions stored: qlo <= q < qlo + qn.  In the segmented walk the shadow lanes of a warp's last, partly filled batch store nothing
  // (the line's last finisher may rewrite it, and only the stores of lanes that vote are ordered before that)
  const unsigned qlo = SEG ? (unsigned)qa : 0u, qn = (SEG && !active) ? 0u : (unsigned)(qb - qa);
  const int vbase = SEG ? sbase - 2 * W : -2 * W;                 // owner = vbase + step + j
  // the walk needs N + 2W steps (segmented: seg_steps, a multiple of 16); the unroll of 16 comes from the register window's static
  // indexing, so a line may stop in the middle of the last round: one warp-uniform test per 16 steps
  const int steps_needed = SEG ? steps : N + 2 * W;
  for (int s0 = 0; s0 < steps; s0 += 16) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int s = s0 + u;
      float y = -INFINITY;
      if (SEG) { y = loady(s); chk = __fmaf_rn(y, 0.f, chk); }    // kVirt where the line has no sample: finite, passes the check
      else if (s < N) { y = loady(s); chk = __fmaf_rn(y, 0.f, chk); }  // warp-uniform
      buf[u] = y;
      const int q = q0 + s;                                       // the position whose last candidate is this step's sample
      const bool valid = (unsigned)q - qlo < qn;
      float c[2 * W + 1], yin[2 * RIN + 1];
#pragma unroll
      for (int j = 0; j <= 2 * W; ++j) c[j] = __fadd_rn(buf[(u + 16 - 2 * W + j) & 15], ef[j]);
#pragma unroll
      for (int k = 0; k <= 2 * RIN; ++k) yin[k] = buf[(u + 16 - W - RIN + k) & 15];
      // decided here only if certified inside the inner ring (tier 1 carries the winner's sample along); anything else is an open
      // position, listed and resolved after the walk.  Open positions are rare (7e-4 on score maps): one warp-uniform branch
      const dtw::Pick pk = dtw::pick_walk<W, RIN>(c, yin, tau0, ylim);
      const bool open = valid & !pk.ok;
      if (__any_sync(0xffffffffu, open)) nd = win_note_open(olist, nd, q, open);   // a call, so that the compiler keeps the branch
      const int j = pk.ok ? pk.j : W;
      const float val = dtw::value_of(__ldg(ed + j), pk.yv);
      st_pair_if_lt(dst, dp, off, val, (unsigned short)(vbase + s + j), (unsigned)q - qlo, qn);
      off += (unsigned)nlines;
    }
    if (s0 + 8 >= steps_needed) break;
#pragma unroll
    for (int u = 8; u < 16; ++u) {
      const int s = s0 + u;
      float y = -INFINITY;
      if (SEG) { y = loady(s); chk = __fmaf_rn(y, 0.f, chk); }    // kVirt where the line has no sample: finite, passes the check
      else if (s < N) { y = loady(s); chk = __fmaf_rn(y, 0.f, chk); }  // warp-uniform
      buf[u] = y;
      const int q = q0 + s;                                       // the position whose last candidate is this step's sample
      const bool valid = (unsigned)q - qlo < qn;
      float c[2 * W + 1], yin[2 * RIN + 1];
#pragma unroll
      for (int j = 0; j <= 2 * W; ++j) c[j] = __fadd_rn(buf[(u + 16 - 2 * W + j) & 15], ef[j]);
#pragma unroll
      for (int k = 0; k <= 2 * RIN; ++k) yin[k] = buf[(u + 16 - W - RIN + k) & 15];
      // decided here only if certified inside the inner ring (tier 1 carries the winner's sample along); anything else is an open
      // position, listed and resolved after the walk.  Open positions are rare (7e-4 on score maps): one warp-uniform branch
      const dtw::Pick pk = dtw::pick_walk<W, RIN>(c, yin, tau0, ylim);
      const bool open = valid & !pk.ok;
      if (__any_sync(0xffffffffu, open)) nd = win_note_open(olist, nd, q, open);   // a call, so that the compiler keeps the branch
      const int j = pk.ok ? pk.j : W;
      const float val = dtw::value_of(__ldg(ed + j), pk.yv);
      st_pair_if_lt(dst, dp, off, val, (unsigned short)(vbase + s + j), (unsigned)q - qlo, qn);
      off += (unsigned)nlines;
    }
  }
  if (nd > kOpenCap) refused = true;
  if (nd > 0 && !refused)                                         // ~2 % of the VGA lines: tier 2, then the local replay, for the open positions
    refused = !win_resolve_open<W>(src, N, os, nlines, M, wp, dst, dp, olist, nd, qa, qb);
  if (chk != chk) refused = true;
  if (SEG && nseg > 1) {
    if (!active) return;
    int* ctr = seg_ctr + (size_t)frame * lines_total + line_base + mi * nlines + line;
    __threadfence();                                              // this segment's stores before its verdict
    const int old = atomicAdd(ctr, 1 + (refused ? 0x10000 : 0));
    if ((old & 0xffff) != nseg - 1) return;                       // another segment of the line is still at work: its lane decides
    *ctr = 0;                                                     // last one: clean for the next launch
    if (!refused && (old >> 16) == 0) return;                     // every segment certified: the line stands
    __threadfence();
  } else {
    if (!(refused && active)) return;
  }
  // ---- replay: the reference's stack algorithm for this lane's line (all of it, also in the segmented walk) ----
  __syncwarp(__activemask());
  Ring& R = *reinterpret_cast<Ring*>(tile_mem[wib]);              // the tiles are dead (every lane of the warp has finished its walk)
  const Quad f = env::make_quad(M.w_sq, M.w_lin, M.etab + M.tab_bias, M.etab + (M.tab_len - kDtRcp));
  float zb[MAXN];
  unsigned short pb[MAXN];
  auto ld = [&](int q) -> float { return __ldg(src + q); };
  // a line with a non-finite sample goes through the reference's two loops literally (dt_envelope.cuh); a segment has only seen its
  // own samples, so the lane that replays a segmented line looks at the whole line first
  bool not_finite = chk != chk;
  if (SEG && !not_finite) {
    float c2 = 0.f;
    for (int q = 0; q < N; ++q) c2 = __fmaf_rn(ld(q), 0.f, c2);
    not_finite = c2 != c2;
  }
  if (not_finite) env::envelope_literal(N, f, os, zb, pb, ld, [&](int i, float val, int v) { store(i, val, (unsigned short)v); });
  else env::envelope_stream(N, f, os, R, lane, zb, pb, ld, ld, [&](int i, float val, int v) { store(i, val, (unsigned short)v); }, [](int) {});
  if (counter) atomicAdd(counter, 1);
}

// E[j] = a x^2 + b x for x = j - tab_bias, j in [0, tab_len - kDtRcp): the position-independent part of Quadratic::operator()(x, y);
// the last kDtRcp entries are the reciprocals 1/(2a*dd) used by the pop path
__global__ void __launch_bounds__(128) dt_build_tables(const PassMap* __restrict__ maps) {
  const PassMap M = maps[blockIdx.x];
  const double a = (double)(-M.w_sq), b = (double)(-M.w_lin);
  const int ne = M.tab_len - kDtRcp;
  for (int j = threadIdx.x; j < M.tab_len; j += blockDim.x)
    M.etab[j] = j < ne ? env::table_E(a, b, j - M.tab_bias) : env::table_rcp(a, j - ne);
}

// ---------------------------------------------------------------------------------------------------
// Parallel-in-q transform (dt_lines.cuh).  A warp owns a batch of b lines of one map: the batch is staged in the warp's
// shared-memory region (y[line][q], conflict-free for both the lane = sample phases and the lane = line phase), transformed by
// dtl::process_lines, and written back with the same access pattern it was read with -- all maps stay [y][x] in HBM:
//   rows pass  (COLS = false): a line is an image row (contiguous); values and arg-maxes go straight to global memory, coalesced;
//   cols pass  (COLS = true):  a line is an image column; the batch is a [N rows][b columns] tile, read and written as row
//                              segments of b elements (b = 8 floats = one 32-byte sector), transposed through shared memory.
// ---------------------------------------------------------------------------------------------------
constexpr int kLineWarps = 4;

struct LineTask { int level, unit, blk; };
// task index -> (level, unit, block of lines); units = maps (or jobs) per level
__device__ __forceinline__ bool find_task(const LineGeom* __restrict__ lg, int units, int t, LineTask& T) {
  const int nl = lg->n_levels;
  for (int l = 0; l < nl; ++l) {
    const int nblk = lg->nblk[l], nt = nblk * units;
    if (t < nt) { T.level = l; T.unit = t / nblk; T.blk = t - T.unit * nblk; return true; }
    t -= nt;
  }
  return false;
}

struct LineRegion { float* y; float* z; int* own; unsigned* bits; unsigned short* stash; };
// alias: own shares the z array (lines of at most dtl::kAliasMaxN samples); the u16 stash exists in the column kernels only
__device__ __forceinline__ LineRegion carve_region(unsigned char* base, int b, int N, bool alias) {
  const int LS = dtl::line_stride(N), NW = (N + 31) >> 5;
  LineRegion R;
  R.y = reinterpret_cast<float*>(base);
  R.z = R.y + b * LS;
  R.own = alias ? reinterpret_cast<int*>(R.z) : reinterpret_cast<int*>(R.z + b * LS);
  R.bits = reinterpret_cast<unsigned*>(R.own + b * LS);
  R.stash = reinterpret_cast<unsigned short*>(R.bits + b * NW);
  return R;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, m));
  return v;
}

// stage a [N][nb] column tile (columns x0 .. x0+nb-1 of a map with row pitch W) as y[c][r]; b = power of two; returns max |y|
__device__ __forceinline__ float load_col_tile(const float* __restrict__ src, int W, int x0, int nb, int b, int N, float* y, int LS, int lane) {
  const int c = lane & (b - 1), rstep = 32 / b;
  float ymax = 0.f;
  if (c < nb) {
    const float* p = src + x0 + c;
    float* d = y + c * LS;
#pragma unroll 4
    for (int r = lane / b; r < N; r += rstep) { const float v = __ldg(p + (size_t)r * W); d[r] = v; ymax = fmaxf(ymax, fabsf(v)); }
  }
  return warp_max(ymax);
}

template <bool COLS>
__global__ void __launch_bounds__(kLineWarps * 32)
dt_lines(const LineGeom* __restrict__ lg, const PassMap* __restrict__ maps, int nmaps, const float* __restrict__ inA, size_t strideA,
         const float* __restrict__ inB, size_t strideB, float* __restrict__ out, size_t stride_out, unsigned short* __restrict__ ptr,
         size_t stride_ptr) {
  extern __shared__ __align__(16) unsigned char smem_lines[];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  LineTask T;
  if (!find_task(lg, nmaps, blockIdx.x * kLineWarps + wib, T)) return;          // warp-uniform
  const int frame = blockIdx.y;
  const int N = lg->N[T.level], nlines = lg->nlines[T.level], b = lg->b[T.level];
  const int line0 = T.blk * b, nb = min(b, nlines - line0);
  const size_t cell_off = (size_t)lg->cell_off[T.level];
  const PassMap M = maps[T.unit];
  const float* src = (M.in_buf ? inB + (size_t)frame * strideB : inA + (size_t)frame * strideA) + M.in_off + cell_off;
  float* dst = out + (size_t)frame * stride_out + M.out_off + cell_off;
  unsigned short* dp = ptr + (size_t)frame * stride_ptr + M.ptr_off + cell_off;
  const Quad f = env::make_quad(M.w_sq, M.w_lin, M.etab + M.tab_bias, M.etab + (M.tab_len - kDtRcp));
  const LineRegion R = carve_region(smem_lines + (size_t)wib * lg->region_bytes, b, N, lg->alias != 0);
  const int LS = dtl::line_stride(N), LSP = (LS + 1) & ~1;
  const dtl::DevWarp w;
  float ymax = 0.f;
  if (!COLS) {
    // the batch's rows are one contiguous run of nb * N floats
    const float* s0 = src + (size_t)line0 * N;
    for (int l = 0; l < nb; ++l) {
      const float* sl = s0 + l * N;
      float* yl = R.y + l * LS;
#pragma unroll 4
      for (int q = lane; q < N; q += 32) { const float v = __ldg(sl + q); yl[q] = v; ymax = fmaxf(ymax, fabsf(v)); }
    }
    ymax = warp_max(ymax);
  } else {
    ymax = load_col_tile(src, nlines, line0, nb, b, N, R.y, LS, lane);
  }
  __syncwarp();
  if (!COLS) {
    float* d0 = dst + (size_t)line0 * N;
    unsigned short* p0 = dp + (size_t)line0 * N;
    dtl::process_lines_any(w, f, N, M.os, nb, ymax, R.y, R.z, R.own, R.bits,
                           [&](int l, int i, float val, int v) { d0[l * N + i] = val; p0[l * N + i] = (unsigned short)v; });
  } else {
    // values are parked in the position's own slot, arg-maxes in the u16 stash, then the tile is written as row segments
    dtl::process_lines_any(w, f, N, M.os, nb, ymax, R.y, R.z, R.own, R.bits,
                           [&](int l, int i, float val, int v) { R.own[l * LS + i] = __float_as_int(val); R.stash[l * LSP + i] = (unsigned short)v; });
    __syncwarp();
    const int c = lane & (b - 1), rstep = 32 / b;
    if (c < nb)
      for (int r = lane / b; r < N; r += rstep) {
        const size_t o = (size_t)r * nlines + line0 + c;
        dst[o] = __int_as_float(R.own[c * LS + r]);
        dp[o] = R.stash[c * LSP + r];
      }
  }
}

// ---------------------------------------------------------------------------------------------------
// Mixture maximum + parent accumulate (src/DynamicProgram.cpp:134-156), elementwise over all cells of all levels.
// ---------------------------------------------------------------------------------------------------
// V = cells per thread: 4 (128-bit loads / stores, one 32-bit store of four Ik bytes) when cells_total is a multiple of 4 so that
// every map base stays 16-byte aligned, else 1.  The job's bias matrix and slot tables are staged in shared memory once per block.
// PRE = true (launches too small to fill the GPU: single frames): the parent's maps are read before anything is written, so that the
// loads of all parent mixtures are in flight together -- such a launch is a chain of memory latencies otherwise; large launches are
// bandwidth-bound and keep the leaner register footprint (measured: PRE costs 2 % of the DP stage at 64 frames, saves 12 % of mix_max at 1).
template <int V, bool PRE>
__global__ void __launch_bounds__(256)
mix_max(const Geometry* __restrict__ g, const PartJob* __restrict__ jobs, const float* __restrict__ resp, float* __restrict__ work,
        const float* __restrict__ val, unsigned char* __restrict__ ik, int nfilters, int nwork, int npm, int tmp_maps) {
  __shared__ PartJob J;
  {
    const int* src = reinterpret_cast<const int*>(jobs + blockIdx.y);
    int* dst = reinterpret_cast<int*>(&J);
    for (int i = threadIdx.x; i < (int)(sizeof(PartJob) / sizeof(int)); i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const int idx = (blockIdx.x * blockDim.x + threadIdx.x) * V;
  if (idx >= g->cells_total) return;
  const int frame = blockIdx.z;
  const size_t ct = (size_t)g->cells_total;
  const int nmix = J.nmix, pnmix = J.pnmix;
  float v[kMaxMix][V];
#pragma unroll
  for (int mm = 0; mm < kMaxMix; ++mm) {
    if (mm < nmix) {
      const float* p = val + ((size_t)frame * tmp_maps + J.tmp_base + mm) * ct + idx;
      if (V == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
        v[mm][0] = t.x; v[mm][1 % V] = t.y; v[mm][2 % V] = t.z; v[mm][3 % V] = t.w;
      } else {
        v[mm][0] = __ldg(p);
      }
    }
  }
  // the parent's maps are read before anything is written (a map is only ever rewritten in place, so the loads of all parent
  // mixtures can be in flight together: a single frame's launch is a chain of memory latencies otherwise)
  float pb[PRE ? kMaxMix : 1][V];
#pragma unroll
  for (int pm = 0; pm < kMaxMix; ++pm) {
    if (PRE && pm < pnmix) {
      const float* bp = J.first_touch ? resp + ((size_t)frame * nfilters + J.out_resp_fid[pm]) * ct + idx
                                      : work + ((size_t)frame * nwork + J.out_work_slot[pm]) * ct + idx;
      if (V == 4) {
        const float4 t = *reinterpret_cast<const float4*>(bp);
        pb[pm][0] = t.x; pb[pm][1 % V] = t.y; pb[pm][2 % V] = t.z; pb[pm][3 % V] = t.w;
      } else {
        pb[pm][0] = *bp;
      }
    }
  }
  auto one_parent_mixture = [&](int pm, const float (&b)[V]) {
    float best[V];
    int bi[V];
#pragma unroll
    for (int c = 0; c < V; ++c) { best[c] = -INFINITY; bi[c] = 0; }
#pragma unroll
    for (int mm = 0; mm < kMaxMix; ++mm) {
      if (mm < nmix) {
        const float bs = J.bias[mm][pm];
#pragma unroll
        for (int c = 0; c < V; ++c) {
          const float wv = __fadd_rn(v[mm][c], bs);                 // scoresp[mm] + bias(mm)[m], :139
          if (wv > best[c]) { best[c] = wv; bi[c] = mm; }           // reduceMax: strict >, first wins
        }
      }
    }
    unsigned char* ikp = ik + ((size_t)frame * npm + J.pm_slot[pm]) * ct + idx;
    float* wp = work + ((size_t)frame * nwork + J.out_work_slot[pm]) * ct + idx;
    if (V == 4) {
      *reinterpret_cast<uchar4*>(ikp) = make_uchar4((unsigned char)bi[0], (unsigned char)bi[1 % V], (unsigned char)bi[2 % V], (unsigned char)bi[3 % V]);
      float4 o;                                                     // parent.score += maxv, :155-156
      o.x = __fadd_rn(b[0], best[0]); o.y = __fadd_rn(b[1 % V], best[1 % V]); o.z = __fadd_rn(b[2 % V], best[2 % V]); o.w = __fadd_rn(b[3 % V], best[3 % V]);
      *reinterpret_cast<float4*>(wp) = o;
    } else {
      *ikp = (unsigned char)bi[0];
      *wp = __fadd_rn(b[0], best[0]);
    }
  };
  if constexpr (PRE) {
#pragma unroll
    for (int pm = 0; pm < kMaxMix; ++pm)
      if (pm < pnmix) one_parent_mixture(pm, pb[pm]);
  } else {
    for (int pm = 0; pm < pnmix; ++pm) {
      const float* bp = J.first_touch ? resp + ((size_t)frame * nfilters + J.out_resp_fid[pm]) * ct + idx
                                      : work + ((size_t)frame * nwork + J.out_work_slot[pm]) * ct + idx;
      float b[V];
      if (V == 4) {
        const float4 t = *reinterpret_cast<const float4*>(bp);
        b[0] = t.x; b[1 % V] = t.y; b[2 % V] = t.z; b[3 % V] = t.w;
      } else {
        b[0] = *bp;
      }
      one_parent_mixture(pm, b);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Root: rootv = max_m(score[m] + bias), rooti = argmax (:163-171).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
root_select(const Geometry* __restrict__ g, const RootJob* __restrict__ roots, int ncomp, const float* __restrict__ resp,
            const float* __restrict__ work, int nfilters, int nwork, float* __restrict__ rootv, unsigned char* __restrict__ rooti) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= g->cells_total) return;
  const int comp = blockIdx.y, frame = blockIdx.z;
  const RootJob& R = roots[comp];
  const size_t ct = (size_t)g->cells_total;
  float best = -INFINITY;
  int bi = 0;
  for (int m = 0; m < R.nmix; ++m) {
    const float sc = R.in_is_work[m] ? work[((size_t)frame * nwork + R.in_slot[m]) * ct + idx]
                                     : resp[((size_t)frame * nfilters + R.in_slot[m]) * ct + idx];
    const float wv = __fadd_rn(sc, R.bias);
    if (wv > best) { best = wv; bi = m; }
  }
  rootv[((size_t)frame * ncomp + comp) * ct + idx] = best;
  rooti[((size_t)frame * ncomp + comp) * ct + idx] = (unsigned char)bi;
}

// Optional epilogue of min(): nonMaximaSuppression(rootv[n][c], sz, maxima) of reference src/nms.cpp:84-129 (the call the reference
// keeps commented out at src/PartsBasedDetector.cpp:86), one thread per root cell.  The map is cut into (sz+1)^2 blocks; a cell is
// kept iff it is its block's FIRST maximum in row-major order (cv::minMaxLoc: strict >) and strictly greater than every cell of the
// (2 sz + 1)^2 window centred on it that lies outside the block's rows x columns; an empty window compares against 0, as
// cv::minMaxLoc over an empty selection does.
__global__ void __launch_bounds__(256)
root_nms(const Geometry* __restrict__ g, int ncomp, const float* __restrict__ rootv, int sz, unsigned char* __restrict__ keep) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= g->cells_total) return;
  const int comp = blockIdx.y, frame = blockIdx.z;
  int l = 0;
  while (l + 1 < g->n_levels && idx >= g->lv[l + 1].cell_off) ++l;
  const int M = g->lv[l].oh, N = g->lv[l].ow;
  const int local = idx - g->lv[l].cell_off, y = local / N, x = local - y * N;
  const float* map = rootv + ((size_t)frame * ncomp + comp) * g->cells_total + g->lv[l].cell_off;
  const float v = map[local];
  const int m = y / (sz + 1) * (sz + 1), n = x / (sz + 1) * (sz + 1);                 // block origin
  const int i1 = min(m + sz + 1, M), j1 = min(n + sz + 1, N);
  bool cand = true;
  for (int yy = m; yy < i1 && cand; ++yy)
    for (int xx = n; xx < j1; ++xx) {
      const float u = map[yy * N + xx];
      const bool earlier = yy < y || (yy == y && xx < x);
      if (earlier ? !(v > u) : u > v) { cand = false; break; }                           // first maximum of the block
    }
  unsigned char out = 0;
  if (cand) {
    const int in0 = max(y - sz, 0), in1 = min(y + sz + 1, M), jn0 = max(x - sz, 0), jn1 = min(x + sz + 1, N);
    // the block's rows / columns as the reference clips them against the window (:111-113)
    const int by0 = m, by1 = in0 + min(m - in0 + sz + 1, in1 - in0), bx0 = n, bx1 = jn0 + min(n - jn0 + sz + 1, jn1 - jn0);
    float vn = 0.f;
    bool any = false;
    for (int yy = in0; yy < in1; ++yy)
      for (int xx = jn0; xx < jn1; ++xx) {
        if (yy >= by0 && yy < by1 && xx >= bx0 && xx < bx1) continue;
        const float u = map[yy * N + xx];
        if (!any || u > vn) { vn = u; any = true; }
      }
    out = v > vn ? 255 : 0;
  }
  keep[((size_t)frame * ncomp + comp) * g->cells_total + idx] = out;
}

// Hits of the computed rootv (DynamicProgram::argmin threshold + Math::find, :208-211); keep (optional): root-map NMS mask.
__global__ void __launch_bounds__(256)
hits_select(const Geometry* __restrict__ g, int ncomp, const float* __restrict__ rootv, const unsigned char* __restrict__ keep, float thresh,
            Hit* __restrict__ hits, int* __restrict__ nhits, int max_hits) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= g->cells_total) return;
  const int comp = blockIdx.y, frame = blockIdx.z;
  const float v = rootv[((size_t)frame * ncomp + comp) * g->cells_total + idx];
  if (!(v > thresh)) return;
  if (keep && !keep[((size_t)frame * ncomp + comp) * g->cells_total + idx]) return;
  int l = 0;
  while (l + 1 < g->n_levels && idx >= g->lv[l + 1].cell_off) ++l;
  const int local = idx - g->lv[l].cell_off;
  const int slot = atomicAdd(nhits, 1);
  if (slot < max_hits) {
    Hit h;
    h.frame = frame; h.level = l; h.comp = comp; h.y = local / g->lv[l].ow; h.x = local % g->lv[l].ow; h.score = v;
    hits[slot] = h;
  }
}

// ---------------------------------------------------------------------------------------------------
// Standalone 2-D DT (pbd_dt2d_f32 / config-5 microbenchmark) uses dt_pass twice plus this composition.
// Layouts: ixraw is [x][y] (row-pass output, transposed), iyraw is [y][x].
// ---------------------------------------------------------------------------------------------------
// mode 0 (reference, :232-244): Iy[y][x] <- Iyraw[y][Ix[y][x]];  mode 1: Ix[y][x] <- Ixraw[Iy[y][x]][x]
__global__ void __launch_bounds__(256)
dt2d_compose(int h, int w, const unsigned short* __restrict__ ixraw, const unsigned short* __restrict__ iyraw,
             unsigned short* __restrict__ ix, unsigned short* __restrict__ iy, int mode) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= w) return;
  const int y = blockIdx.y, m = blockIdx.z;
  const size_t base = (size_t)m * h * w;
  const size_t c = base + (size_t)y * w + x;
  if (mode == 0) {
    const int xi = ixraw[base + (size_t)x * h + y];
    ix[c] = (unsigned short)xi;
    iy[c] = iyraw[base + (size_t)y * w + xi];
  } else {
    const int yi = iyraw[c];
    iy[c] = (unsigned short)yi;
    ix[c] = ixraw[base + (size_t)x * h + yi];
  }
}

}  // namespace

template <int SCAN, typename... A>
static void launch_pass_v(int maxn, dim3 grid, cudaStream_t s, A... args) {
  if (maxn <= 160) dt_pass<160, SCAN><<<grid, kPassWarps * 32, 0, s>>>(args...);
  else if (maxn <= 512) dt_pass<512, SCAN><<<grid, kPassWarps * 32, 0, s>>>(args...);
  else if (maxn <= 1024) dt_pass<1024, SCAN><<<grid, kPassWarps * 32, 0, s>>>(args...);
  else dt_pass<4096, SCAN><<<grid, kPassWarps * 32, 0, s>>>(args...);
}
template <typename... A>
static void launch_pass(int maxn, dim3 grid, cudaStream_t s, A... args) { launch_pass_v<0>(maxn, grid, s, args...); }
template <bool SEG, typename... A>
static void launch_pass_win_v(int maxn, dim3 grid, cudaStream_t s, A... args) {
  constexpr int W = kDtWindowW;
  if (maxn <= 160) dt_pass_win<160, W, SEG><<<grid, kPassWarps * 32, 0, s>>>(args...);
  else if (maxn <= 512) dt_pass_win<512, W, SEG><<<grid, kPassWarps * 32, 0, s>>>(args...);
  else if (maxn <= 1024) dt_pass_win<1024, W, SEG><<<grid, kPassWarps * 32, 0, s>>>(args...);
  else dt_pass_win<4096, W, SEG><<<grid, kPassWarps * 32, 0, s>>>(args...);
}
// seg_steps = 0: one lane per line; otherwise lines are cut into segments of seg_steps - 2W positions (dt_pass_win<.., SEG = true>)
template <typename... A>
static void launch_pass_win(int maxn, int seg_steps, int* d_seg_ctr, int lines_total, dim3 grid, cudaStream_t s, A... args) {
  if (seg_steps > 0 && d_seg_ctr) launch_pass_win_v<true>(maxn, grid, s, args..., seg_steps, d_seg_ctr, lines_total);
  else launch_pass_win_v<false>(maxn, grid, s, args..., 0, (int*)nullptr, 0);
}
static int pass_lines(const PassGeom& pg, int nmaps) {
  int n = 0;
  for (int l = 0; l < pg.n_levels; ++l) n += pg.nlines[l] * nmaps;
  return n;
}

// number of warps a pass needs for `nmaps` maps; seg_steps > 0: lines cut into segments of seg_steps - 2W positions (dt_pass_win<.., true>)
static int pass_warps(const PassGeom& pg, int nmaps, int seg_steps = 0) {
  int w = 0;
  const int seglen = seg_steps - 2 * kDtWindowW;
  for (int l = 0; l < pg.n_levels; ++l) {
    const int nseg = seg_steps > 0 ? (pg.N[l] + seglen - 1) / seglen : 1;
    w += (pg.nlines[l] * nmaps * nseg + 31) / 32;
  }
  return w;
}
// Segment length of the windowed walk for one pass.  A lane walks its line sequentially (about 0.3 us per sample), so a launch that
// cannot fill the GPU lasts as long as the longest line of the pyramid; such launches (single frames, small batches) cut their lines
// into the shortest segments that still fit the GPU's resident warps at once.  Launches that fill the GPU anyway keep one lane per line
// (a segment re-reads 2W samples).  mode: -1 automatic, 0 never, otherwise the forced number of steps per segment (multiple of 16, >= 32).
#ifndef PBD_DT_SEG_MIN
#define PBD_DT_SEG_MIN 32
#endif
constexpr int kSegMinSteps = PBD_DT_SEG_MIN;      // shortest segment the automatic choice considers (steps; 2W of them fill the window)
static int choose_seg_steps(const PassGeom& pg, int nmaps, int nframes, int warp_slots, int mode, const int* d_seg_ctr) {
  int maxn = 1;
  long long lines = 0;
  for (int l = 0; l < pg.n_levels; ++l) { maxn = std::max(maxn, pg.N[l]); lines += (long long)pg.nlines[l] * nmaps; }
  const int full = (maxn + 2 * kDtWindowW + 15) & ~15;              // steps of the unsegmented walk of the longest line
  if (mode == 0 || !d_seg_ctr || lines * nframes > (long long)warp_slots * 32) return 0;   // one counter per line: warp_slots * 32 of them
  if (mode > 0) return mode < full ? mode : 0;
  if ((long long)pass_warps(pg, nmaps) * nframes * 2 > warp_slots) return 0;
  for (int st = kSegMinSteps; st < full; st += 16)
    if ((long long)pass_warps(pg, nmaps, st) * nframes <= warp_slots) return st;
  return 0;
}

// ---- parallel-in-q launch plan ------------------------------------------------------------------------------------------
void plan_line_geom(LineGeom& lg, int budget_bytes, bool stash) {
  int maxn = 1;
  for (int l = 0; l < lg.n_levels; ++l) maxn = std::max(maxn, lg.N[l]);
  lg.alias = maxn <= dtl::kAliasMaxN ? 1 : 0;
  int region = 0;
  for (int l = 0; l < lg.n_levels; ++l) {
    const int lb = dtl::line_bytes(lg.N[l], lg.alias != 0, stash);
    int b = 32;
    while (b > 1 && (b * lb > budget_bytes || b / 2 >= lg.nlines[l])) b >>= 1;   // no wider than needed for the level's lines
    lg.b[l] = b;
    lg.nblk[l] = (lg.nlines[l] + b - 1) / b;
    region = std::max(region, b * lb);
  }
  lg.region_bytes = (region + 15) / 16 * 16;
}
long long line_tasks(const LineGeom& lg, int units) {
  long long t = 0;
  for (int l = 0; l < lg.n_levels; ++l) t += (long long)lg.nblk[l] * units;
  return t;
}

template <typename K>
static void set_smem(K kernel, int bytes) {
  // opt in to more than 48 KB of dynamic shared memory (idempotent; the attribute is per kernel and device)
  if (bytes > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

template <bool COLS, typename... A>
static void launch_lines(const LineGeom& lg, dim3 grid, cudaStream_t s, A... args) {
  const int smem = kLineWarps * lg.region_bytes;
  set_smem(dt_lines<COLS>, smem);
  dt_lines<COLS><<<grid, kLineWarps * 32, smem, s>>>(args...);
}

int launch_dt_wave(const Geometry& g, const Geometry* d_g, const DeviceBuffers& b, const PassGeom& pg_rows, const PassGeom* d_pg_rows,
                   const PassGeom& pg_cols, const PassGeom* d_pg_cols, const PassMap* d_maps_rows, const PassMap* d_maps_cols, int nmaps,
                   int max_ow, int max_oh, const PartJob* d_jobs, int njobs, int nfilters, int nwork, int ncm, int npm, int tmp_maps,
                   cudaStream_t s, void (*mark)(void*, int), void* mark_ctx, int scan, const dtw::WinParams* d_wp_rows,
                   const dtw::WinParams* d_wp_cols, int* d_replayed, int warp_slots, int seg_mode, int* d_seg_ctr) {
  if (nmaps <= 0 || njobs <= 0 || g.cells_total <= 0) return 0;
  const size_t ct = (size_t)g.cells_total;
  const int seg_r = scan == 3 ? choose_seg_steps(pg_rows, nmaps, g.n_frames, warp_slots, seg_mode, d_seg_ctr) : 0;
  const int seg_c = scan == 3 ? choose_seg_steps(pg_cols, nmaps, g.n_frames, warp_slots, seg_mode, d_seg_ctr) : 0;
  dim3 gr((pass_warps(pg_rows, nmaps, seg_r) + kPassWarps - 1) / kPassWarps, g.n_frames);
  if (scan == 3) launch_pass_win(max_ow, seg_r, d_seg_ctr, pass_lines(pg_rows, nmaps), gr, s, d_pg_rows, d_maps_rows, d_wp_rows, nmaps, (const float*)b.resp, ct * nfilters, (const float*)b.work, ct * nwork,
                                 b.tmp, ct * tmp_maps, b.ixdt, ct * ncm, d_replayed);
  else if (scan == 2) launch_pass_v<4>(max_ow, gr, s, d_pg_rows, d_maps_rows, nmaps, (const float*)b.resp, ct * nfilters, (const float*)b.work, ct * nwork, b.tmp,
                                  ct * tmp_maps, b.ixdt, ct * ncm);
  else if (scan == 1) launch_pass_v<-1>(max_ow, gr, s, d_pg_rows, d_maps_rows, nmaps, (const float*)b.resp, ct * nfilters, (const float*)b.work, ct * nwork, b.tmp,
                                       ct * tmp_maps, b.ixdt, ct * ncm);
  else launch_pass(max_ow, gr, s, d_pg_rows, d_maps_rows, nmaps, (const float*)b.resp, ct * nfilters, (const float*)b.work, ct * nwork, b.tmp,
                   ct * tmp_maps, b.ixdt, ct * ncm);
  if (mark) mark(mark_ctx, 2);
  dim3 gc((pass_warps(pg_cols, nmaps, seg_c) + kPassWarps - 1) / kPassWarps, g.n_frames);
  if (scan == 3) launch_pass_win(max_oh, seg_c, d_seg_ctr, pass_lines(pg_cols, nmaps), gc, s, d_pg_cols, d_maps_cols, d_wp_cols, nmaps, (const float*)b.tmp, ct * tmp_maps, (const float*)b.tmp, ct * tmp_maps,
                                 b.val, ct * tmp_maps, b.iyraw, ct * ncm, d_replayed);
  else if (scan == 2) launch_pass_v<4>(max_oh, gc, s, d_pg_cols, d_maps_cols, nmaps, (const float*)b.tmp, ct * tmp_maps, (const float*)b.tmp, ct * tmp_maps, b.val,
                                  ct * tmp_maps, b.iyraw, ct * ncm);
  else if (scan == 1) launch_pass_v<-1>(max_oh, gc, s, d_pg_cols, d_maps_cols, nmaps, (const float*)b.tmp, ct * tmp_maps, (const float*)b.tmp, ct * tmp_maps, b.val,
                                       ct * tmp_maps, b.iyraw, ct * ncm);
  else launch_pass(max_oh, gc, s, d_pg_cols, d_maps_cols, nmaps, (const float*)b.tmp, ct * tmp_maps, (const float*)b.tmp, ct * tmp_maps, b.val,
                   ct * tmp_maps, b.iyraw, ct * ncm);
  if (mark) mark(mark_ctx, 3);
  const bool pre = (long long)g.cells_total * njobs * g.n_frames < (long long)warp_slots * 32 * 16;   // under ~16 cells per resident thread
  if (g.cells_total % 4 == 0) {
    dim3 gm((g.cells_total / 4 + 255) / 256, njobs, g.n_frames);
    if (pre) mix_max<4, true><<<gm, 256, 0, s>>>(d_g, d_jobs, b.resp, b.work, b.val, b.ik, nfilters, nwork, npm, tmp_maps);
    else mix_max<4, false><<<gm, 256, 0, s>>>(d_g, d_jobs, b.resp, b.work, b.val, b.ik, nfilters, nwork, npm, tmp_maps);
  } else {
    dim3 gm((g.cells_total + 255) / 256, njobs, g.n_frames);
    if (pre) mix_max<1, true><<<gm, 256, 0, s>>>(d_g, d_jobs, b.resp, b.work, b.val, b.ik, nfilters, nwork, npm, tmp_maps);
    else mix_max<1, false><<<gm, 256, 0, s>>>(d_g, d_jobs, b.resp, b.work, b.val, b.ik, nfilters, nwork, npm, tmp_maps);
  }
  if (mark) mark(mark_ctx, 4);
  return 3;
}

int launch_dt_tables(const PassMap* d_maps, int nmaps, cudaStream_t s) {
  if (nmaps <= 0) return 0;
  dt_build_tables<<<nmaps, 128, 0, s>>>(d_maps);
  return 1;
}

int launch_root(const Geometry& g, const Geometry* d_g, const DeviceBuffers& b, const RootJob* d_roots, int ncomp, int nfilters, int nwork,
                cudaStream_t s) {
  if (g.cells_total <= 0) return 0;
  dim3 grid((g.cells_total + 255) / 256, ncomp, g.n_frames);
  root_select<<<grid, 256, 0, s>>>(d_g, d_roots, ncomp, b.resp, b.work, nfilters, nwork, b.rootv, b.rooti);
  return 1;
}

int launch_hits(const Geometry& g, const Geometry* d_g, const DeviceBuffers& b, int ncomp, float thresh, Hit* d_hits, int* d_nhits,
                int max_hits, cudaStream_t s, int root_nms_sz, unsigned char* d_keep) {
  if (g.cells_total <= 0) return 0;
  dim3 grid((g.cells_total + 255) / 256, ncomp, g.n_frames);
  int n = 1;
  if (root_nms_sz > 0 && d_keep) { root_nms<<<grid, 256, 0, s>>>(d_g, ncomp, b.rootv, root_nms_sz, d_keep); ++n; }
  hits_select<<<grid, 256, 0, s>>>(d_g, ncomp, b.rootv, root_nms_sz > 0 ? d_keep : nullptr, thresh, d_hits, d_nhits, max_hits);
  return n;
}

int launch_dt2d_standalone(const float* d_in, int n_maps, int h, int w, const PassGeom* d_pg2 /* [rows, cols] */, const PassMap* d_maps2,
                           float* d_tmp, float* d_out, uint16_t* d_ix, uint16_t* d_iy, uint16_t* d_ixraw, uint16_t* d_iyraw, int backptr_mode,
                           cudaStream_t s, int scan, const dtw::WinParams* d_wp2, int* d_replayed, int seg_steps, int* d_seg_ctr) {
  if (n_maps <= 0 || h <= 0 || w <= 0) return 0;
  // maps are launched in chunks so that the warp index stays small; every map is h*w cells
  PassGeom pr{}, pc{};
  pr.n_levels = pc.n_levels = 1;
  pr.nlines[0] = h; pr.N[0] = w; pc.nlines[0] = w; pc.N[0] = h;
  // seg_steps > 0 (tests): the windowed transform with its lines cut into segments, as the detector does for launches that cannot fill the GPU
  const int seg = (scan == 3 && d_wp2 && d_seg_ctr && seg_steps > 2 * kDtWindowW) ? seg_steps : 0;
  const int seg_r = seg && seg < ((w + 2 * kDtWindowW + 15) & ~15) ? seg : 0, seg_c = seg && seg < ((h + 2 * kDtWindowW + 15) & ~15) ? seg : 0;
  dim3 gr((pass_warps(pr, n_maps, seg_r) + kPassWarps - 1) / kPassWarps, 1), gc((pass_warps(pc, n_maps, seg_c) + kPassWarps - 1) / kPassWarps, 1);
  if (scan == 3 && d_wp2) {                                         // windowed certified evaluation (dt_pass_win): rows, then columns
    launch_pass_win(w, seg_r, d_seg_ctr, pass_lines(pr, n_maps), gr, s, d_pg2, d_maps2, d_wp2, n_maps, d_in, (size_t)0, d_in, (size_t)0, d_tmp, (size_t)0, d_ixraw, (size_t)0, d_replayed);
    launch_pass_win(h, seg_c, d_seg_ctr, pass_lines(pc, n_maps), gc, s, d_pg2 + 1, d_maps2 + n_maps, d_wp2 + n_maps, n_maps, (const float*)d_tmp, (size_t)0, (const float*)d_tmp, (size_t)0, d_out,
                    (size_t)0, d_iyraw, (size_t)0, d_replayed);
  } else if (scan) {
    launch_pass_v<4>(w, gr, s, d_pg2, d_maps2, n_maps, d_in, (size_t)0, d_in, (size_t)0, d_tmp, (size_t)0, d_ixraw, (size_t)0);
    launch_pass_v<4>(h, gc, s, d_pg2 + 1, d_maps2 + n_maps, n_maps, (const float*)d_tmp, (size_t)0, (const float*)d_tmp, (size_t)0, d_out, (size_t)0,
                     d_iyraw, (size_t)0);
  } else {
    launch_pass(w, gr, s, d_pg2, d_maps2, n_maps, d_in, (size_t)0, d_in, (size_t)0, d_tmp, (size_t)0, d_ixraw, (size_t)0);
    launch_pass(h, gc, s, d_pg2 + 1, d_maps2 + n_maps, n_maps, (const float*)d_tmp, (size_t)0, (const float*)d_tmp, (size_t)0, d_out, (size_t)0,
                d_iyraw, (size_t)0);
  }
  for (int m0 = 0; m0 < n_maps; m0 += 65535) {                     // gridDim.z limit
    const int nm = std::min(65535, n_maps - m0);
    const size_t o = (size_t)m0 * h * w;
    dim3 gx((w + 255) / 256, h, nm);
    dt2d_compose<<<gx, 256, 0, s>>>(h, w, d_ixraw + o, d_iyraw + o, d_ix + o, d_iy + o, backptr_mode);
  }
  return 2 + (n_maps + 65534) / 65535;
}


namespace {
// all maps [y][x]: mode 0 (reference, :232-244): Iy[y][x] <- Iyraw[y][Ix[y][x]];  mode 1: Ix[y][x] <- Ixraw[Iy[y][x]][x]
__global__ void __launch_bounds__(256)
dt2d_compose_yx(int h, int w, const unsigned short* __restrict__ ixraw, const unsigned short* __restrict__ iyraw,
                unsigned short* __restrict__ ix, unsigned short* __restrict__ iy, int mode) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= w) return;
  const int y = blockIdx.y;
  const size_t base = (size_t)blockIdx.z * h * w;
  const size_t c = base + (size_t)y * w + x;
  if (mode == 0) {
    const int xi = ixraw[c];
    ix[c] = (unsigned short)xi;
    iy[c] = iyraw[base + (size_t)y * w + xi];
  } else {
    const int yi = iyraw[c];
    iy[c] = (unsigned short)yi;
    ix[c] = ixraw[base + (size_t)yi * w + x];
  }
}
}  // namespace

int launch_dt2d_lines(const float* d_in, int n_maps, int h, int w, const LineGeom& lg_rows, const LineGeom& lg_cols, const LineGeom* d_lg2,
                      const PassMap* d_maps2, float* d_tmp, float* d_out, uint16_t* d_ix, uint16_t* d_iy, uint16_t* d_ixraw, uint16_t* d_iyraw,
                      int backptr_mode, cudaStream_t s) {
  if (n_maps <= 0 || h <= 0 || w <= 0) return 0;
  dim3 gr((unsigned)((line_tasks(lg_rows, n_maps) + kLineWarps - 1) / kLineWarps), 1), gc((unsigned)((line_tasks(lg_cols, n_maps) + kLineWarps - 1) / kLineWarps), 1);
  launch_lines<false>(lg_rows, gr, s, d_lg2, d_maps2, n_maps, d_in, (size_t)0, d_in, (size_t)0, d_tmp, (size_t)0, d_ixraw, (size_t)0);
  launch_lines<true>(lg_cols, gc, s, d_lg2 + 1, d_maps2 + n_maps, n_maps, (const float*)d_tmp, (size_t)0, (const float*)d_tmp, (size_t)0, d_out, (size_t)0,
                     d_iyraw, (size_t)0);
  for (int m0 = 0; m0 < n_maps; m0 += 65535) {                     // gridDim.z limit
    const int nm = std::min(65535, n_maps - m0);
    const size_t o = (size_t)m0 * h * w;
    dim3 gx((w + 255) / 256, h, nm);
    dt2d_compose_yx<<<gx, 256, 0, s>>>(h, w, d_ixraw + o, d_iyraw + o, d_ix + o, d_iy + o, backptr_mode);
  }
  return 2 + (n_maps + 65534) / 65535;
}

}  // namespace pbd
