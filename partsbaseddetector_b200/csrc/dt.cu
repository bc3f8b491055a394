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
#include "kernels.cuh"

namespace pbd {
namespace {

constexpr int kRing = 8;          // ring window (stack entries kept in shared memory per lane)
constexpr int kPassWarps = 4;
constexpr int kTileW = 16;        // samples per line staged per shared-memory tile

__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

struct Quad {
  double a, b, a2;   // a2 = 2*a (exact), reference evaluates 2*a*(x1-x0) left to right
  double r1;         // correctly rounded 1/(2a): reciprocal of the divisor for adjacent samples (x1-x0 = 1)
  const double* E;   // E[x] = a*x^2 + b*x (both products and the sum rounded as the reference does), x = pos - v
};
__device__ __forceinline__ Quad make_quad(const PassMap& M) {
  Quad f;
  f.a = (double)(-M.w_sq);          // Quadratic(-w[0], -w[1]), src/DynamicProgram.cpp:126-127
  f.b = (double)(-M.w_lin);
  f.a2 = __dmul_rn(2.0, f.a);
  f.r1 = __drcp_rn(f.a2);
  f.E = M.etab + M.tab_bias;
  return f;
}
// the rare exact quotient: kept out of line so that the common path does not carry the division's instructions
__device__ __noinline__ float quotient_exact(double num, double den) { return (float)__ddiv_rn(num, den); }

// Quadratic::operator()(x0,x1,y0,y1), include/DistanceTransform.hpp:98-100, result rounded to float as `T s = f(...)`,
// for adjacent samples x1 = x0 + 1 (the first intersection of every step: the previous sample is always the top).
// b*1 = b and a*(x1^2-x0^2) = a*(2*x1-1) are the reference's own values; the quotient num/(2a) is formed with a Markstein
// correction step from the precomputed reciprocal.  It is within 1 ulp(double) of the correctly rounded quotient, so its
// float rounding equals the reference's double-division-then-float unless it lies within 2 ulp of a float rounding
// boundary (probability ~2^-27): those cases, and anything outside 2^-100..2^100, take the exact division.
__device__ __forceinline__ float isect_adjacent(const Quad& f, int x1, double y0, double y1) {
  const double num = __dadd_rn(__dsub_rn(__dsub_rn(y1, y0), f.b), __dmul_rn(f.a, (double)(2 * x1 - 1)));
  const double q0 = __dmul_rn(num, f.r1);
  const double q1 = __fma_rn(__fma_rn(-f.a2, q0, num), f.r1, q0);
  const int lo = __double2loint(q1) & 0x1FFFFFFF;
  const unsigned ex = ((unsigned)__double2hiint(q1) & 0x7ff00000u) - ((1023u - 100u) << 20);
  if (ex <= (200u << 20) && abs(lo - 0x10000000) > 2) return (float)q1;
  return quotient_exact(num, f.a2);
}
// the general case (after a pop x1 - x0 >= 2)
__device__ __forceinline__ float isect_far(const Quad& f, int x0, int x1, double y0, double y1) {
  const double dd = (double)(x1 - x0);
  const double t = __dsub_rn(__dsub_rn(y1, y0), __dmul_rn(f.b, dd));
  const double num = __dadd_rn(t, __dmul_rn(f.a, (double)(x1 * x1 - x0 * x0)));
  return (float)__ddiv_rn(num, __dmul_rn(f.a2, dd));
}

// base[idx] accesses with a 32-bit index: one IMAD.WIDE forms the address (the compiler otherwise re-derives the 64-bit
// base from its kernel-parameter components at every store)
__device__ __forceinline__ double ld_table(const double* base, int idx) {
  double v;
  asm("{ .reg .u64 a; mad.wide.s32 a, %2, 8, %1; ld.global.nc.f64 %0, [a]; }" : "=d"(v) : "l"(base), "r"(idx));
  return v;
}
__device__ __forceinline__ void st_f32(float* base, unsigned idx, float v) {
  asm volatile("{ .reg .u64 a; mad.wide.u32 a, %1, 4, %0; st.global.f32 [a], %2; }" ::"l"(base), "r"(idx), "f"(v) : "memory");
}
__device__ __forceinline__ void st_u16(unsigned short* base, unsigned idx, unsigned short v) {
  asm volatile("{ .reg .u64 a; mad.wide.u32 a, %1, 2, %0; st.global.u16 [a], %2; }" ::"l"(base), "r"(idx), "h"(v) : "memory");
}

// per-warp shared-memory ring: [slot][lane]; vp = v | (v of the entry below << 16), 0xFFFF = none
struct Ring {
  float z[kRing][32];
  float y[kRing][32];
  unsigned int vp[kRing][32];
};

// One lane's 1-D transform.  loady(q) = src[q] (called for q = 0..N-1 in order, and again for deep-pop reloads);
// emit(i, val, v) stores dst[i] = val, ptr[i] = v (may be called more than once for an i; the last call wins).
//
// Eager emission: when sample q is pushed with break point s, the previous top P (still in registers) owns exactly the
// integer positions z_P < pos <= s, and they are evaluated and stored at once.  If q (or P) is popped later, the positions
// are simply stored again by their new owner: every position's FINAL owner E_i emits when its final successor E_{i+1} is
// pushed, and since break points increase up the stack no later emission (all above E_{i+1}) can touch a position
// <= z_{E_{i+1}}, so the last store to every position is the reference's scan result.
//
// The ring holds the newest 8 stack entries for pops (5-8 % of the steps); zb/pb are the backing store of the whole
// envelope as a linked list threaded through the sample index (zb[q] = break point of the parabola pushed at q,
// pb[q] = the sample below it), written in lock step across lanes (coalesced) and read only by pops deeper than the ring.
template <typename LoadY, typename Reload, typename Emit>
__device__ __forceinline__ void envelope_stream(int N, const Quad& f, int os0, unsigned stride, Ring& R, int lane, float* zb, unsigned short* pb,
                                                LoadY loady, Reload reload, Emit emit) {
  const int pos_last = os0 + N - 1;
  auto emit_range = [&](float zlo, float zhi, int v, double yd) {
    // integer positions with zlo < pos <= zhi, clipped to [os0, os0+N-1] (the float -> int conversions saturate, so the
    // -inf / +inf break points of the bottom and the top need no special case); value = Quadratic::operator()(pos - v, y),
    // :102-104, = (a x^2 + b x) + y with the parenthesis taken from the map's table
    const int lo = max(min(__float2int_rd(zlo), pos_last) + 1, os0);
    const int hi = min(__float2int_rd(zhi), pos_last);
    int x = lo - v;
    unsigned off = (unsigned)(lo - os0) * stride;
#pragma unroll 1
    for (int pos = lo; pos <= hi; ++pos, ++x, off += stride) emit(off, (float)__dadd_rn(ld_table(f.E, x), yd), v);
  };
  int k = 0, base = 0;                                            // stack depth of the top; lowest depth still valid in the ring
  int vt = 0, pt = 0xFFFF;
  float ytf = loady(0), zt = -INFINITY;
  double yt = (double)ytf;
  R.z[0][lane] = zt; R.y[0][lane] = ytf; R.vp[0][lane] = 0xFFFF0000u;
  zb[0] = zt; pb[0] = 0xFFFF;
  for (int q = 1; q < N; ++q) {                                   // :160-170
    const float yqf = loady(q);
    const double yq = (double)yqf;
    float s = isect_adjacent(f, q, yt, yq);                       // the top is sample q - 1
    while (s <= zt && k > 0) {
      --k;
      const int slot = k & (kRing - 1);
      if (k < base) {                                             // popped below the ring: reload from the backing store
        base = k;
        const int vv = pt;                                        // the entry below the one just popped
        R.vp[slot][lane] = (unsigned)vv | ((unsigned)pb[vv] << 16); R.z[slot][lane] = zb[vv]; R.y[slot][lane] = reload(vv);
      }
      const unsigned vp = R.vp[slot][lane];
      vt = vp & 0xFFFF; pt = vp >> 16; ytf = R.y[slot][lane]; yt = (double)ytf; zt = R.z[slot][lane];
      s = isect_far(f, vt, q, yt, yq);
    }
    emit_range(zt, s, vt, yt);                                   // the top's positions up to the new break point
    ++k;
    base = max(base, k - (kRing - 1));                            // the slot of depth k - kRing is overwritten
    const int slot = k & (kRing - 1);
    R.vp[slot][lane] = (unsigned)q | ((unsigned)vt << 16); R.y[slot][lane] = yqf; R.z[slot][lane] = s;
    zb[q] = s; pb[q] = (unsigned short)vt;
    pt = vt; vt = q; ytf = yqf; yt = yq; zt = s;
  }
  emit_range(zt, INFINITY, vt, yt);
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
template <int MAXN>
__global__ void __launch_bounds__(kPassWarps * 32)
dt_pass(const PassGeom* __restrict__ pg, const PassMap* __restrict__ maps, int nmaps, const float* __restrict__ inA, size_t strideA,
        const float* __restrict__ inB, size_t strideB, float* __restrict__ out, size_t stride_out, unsigned short* __restrict__ ptr,
        size_t stride_ptr) {
  __shared__ Ring rings[kPassWarps];
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
  const Quad f = make_quad(M);
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
  auto loady = [&](int q) -> float {
    if ((q & (kTileW - 1)) == 0) {
      if (q == 0) prefetch(0, 0);
      cp_async_wait_all();
      __syncwarp();
      if (q + kTileW < N) prefetch(q + kTileW, ((q / kTileW) + 1) & 1);
    }
    return tiles[wib][(q / kTileW) & 1][lane][q & (kTileW - 1)];
  };
  // deep-pop reloads of older samples: global memory (the tile that held them may already be refilled)
  auto reload = [&](int v) -> float { return __ldg(src + v); };
  // inactive lanes recompute the warp's first item and store the same values to the same addresses as lane 0
  asm volatile("" : "+l"(dst), "+l"(dp));                        // keep both bases as materialised 64-bit registers
  envelope_stream(N, f, M.os, (unsigned)nlines, rings[wib], lane, zb, pb, loady, reload, [&](unsigned off, float val, int v) {
    st_f32(dst, off, val); st_u16(dp, off, (unsigned short)v);
  });
}

// E[j] = a x^2 + b x for x = j - tab_bias, j in [0, tab_len): the position-independent part of Quadratic::operator()(x, y)
__global__ void __launch_bounds__(128) dt_build_tables(const PassMap* __restrict__ maps) {
  const PassMap M = maps[blockIdx.x];
  const double a = (double)(-M.w_sq), b = (double)(-M.w_lin);
  for (int j = threadIdx.x; j < M.tab_len; j += blockDim.x) {
    const int x = j - M.tab_bias;
    M.etab[j] = __dadd_rn(__dmul_rn(a, (double)(x * x)), __dmul_rn(b, (double)x));
  }
}

// ---------------------------------------------------------------------------------------------------
// Mixture maximum + parent accumulate (src/DynamicProgram.cpp:134-156), elementwise over all cells of all levels.
// ---------------------------------------------------------------------------------------------------
// V = cells per thread: 4 (128-bit loads / stores, one 32-bit store of four Ik bytes) when cells_total is a multiple of 4 so that
// every map base stays 16-byte aligned, else 1.  The job's bias matrix and slot tables are staged in shared memory once per block.
template <int V>
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
  for (int pm = 0; pm < pnmix; ++pm) {
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
    const float* bp = J.first_touch ? resp + ((size_t)frame * nfilters + J.out_resp_fid[pm]) * ct + idx : wp;
    if (V == 4) {
      *reinterpret_cast<uchar4*>(ikp) = make_uchar4((unsigned char)bi[0], (unsigned char)bi[1 % V], (unsigned char)bi[2 % V], (unsigned char)bi[3 % V]);
      const float4 b4 = *reinterpret_cast<const float4*>(bp);
      float4 o;                                                     // parent.score += maxv, :155-156
      o.x = __fadd_rn(b4.x, best[0]); o.y = __fadd_rn(b4.y, best[1 % V]); o.z = __fadd_rn(b4.z, best[2 % V]); o.w = __fadd_rn(b4.w, best[3 % V]);
      *reinterpret_cast<float4*>(wp) = o;
    } else {
      *ikp = (unsigned char)bi[0];
      *wp = __fadd_rn(*bp, best[0]);
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

// Hits of the computed rootv (DynamicProgram::argmin threshold + Math::find, :208-211).
__global__ void __launch_bounds__(256)
hits_select(const Geometry* __restrict__ g, int ncomp, const float* __restrict__ rootv, float thresh, Hit* __restrict__ hits,
            int* __restrict__ nhits, int max_hits) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= g->cells_total) return;
  const int comp = blockIdx.y, frame = blockIdx.z;
  const float v = rootv[((size_t)frame * ncomp + comp) * g->cells_total + idx];
  if (!(v > thresh)) return;
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

template <typename... A>
static void launch_pass(int maxn, dim3 grid, cudaStream_t s, A... args) {
  if (maxn <= 160) dt_pass<160><<<grid, kPassWarps * 32, 0, s>>>(args...);
  else if (maxn <= 512) dt_pass<512><<<grid, kPassWarps * 32, 0, s>>>(args...);
  else if (maxn <= 1024) dt_pass<1024><<<grid, kPassWarps * 32, 0, s>>>(args...);
  else dt_pass<4096><<<grid, kPassWarps * 32, 0, s>>>(args...);
}

// number of warps a pass needs for `nmaps` maps
static int pass_warps(const PassGeom& pg, int nmaps) {
  int w = 0;
  for (int l = 0; l < pg.n_levels; ++l) w += (pg.nlines[l] * nmaps + 31) / 32;
  return w;
}

int launch_dt_wave(const Geometry& g, const Geometry* d_g, const DeviceBuffers& b, const PassGeom& pg_rows, const PassGeom* d_pg_rows,
                   const PassGeom& pg_cols, const PassGeom* d_pg_cols, const PassMap* d_maps_rows, const PassMap* d_maps_cols, int nmaps,
                   int max_ow, int max_oh, const PartJob* d_jobs, int njobs, int nfilters, int nwork, int ncm, int npm, int tmp_maps,
                   cudaStream_t s, void (*mark)(void*, int), void* mark_ctx) {
  if (nmaps <= 0 || njobs <= 0 || g.cells_total <= 0) return 0;
  const size_t ct = (size_t)g.cells_total;
  dim3 gr((pass_warps(pg_rows, nmaps) + kPassWarps - 1) / kPassWarps, g.n_frames);
  launch_pass(max_ow, gr, s, d_pg_rows, d_maps_rows, nmaps, (const float*)b.resp, ct * nfilters, (const float*)b.work, ct * nwork, b.tmp,
              ct * tmp_maps, b.ixdt, ct * ncm);
  if (mark) mark(mark_ctx, 2);
  dim3 gc((pass_warps(pg_cols, nmaps) + kPassWarps - 1) / kPassWarps, g.n_frames);
  launch_pass(max_oh, gc, s, d_pg_cols, d_maps_cols, nmaps, (const float*)b.tmp, ct * tmp_maps, (const float*)b.tmp, ct * tmp_maps, b.val,
              ct * tmp_maps, b.iyraw, ct * ncm);
  if (mark) mark(mark_ctx, 3);
  if (g.cells_total % 4 == 0) {
    dim3 gm((g.cells_total / 4 + 255) / 256, njobs, g.n_frames);
    mix_max<4><<<gm, 256, 0, s>>>(d_g, d_jobs, b.resp, b.work, b.val, b.ik, nfilters, nwork, npm, tmp_maps);
  } else {
    dim3 gm((g.cells_total + 255) / 256, njobs, g.n_frames);
    mix_max<1><<<gm, 256, 0, s>>>(d_g, d_jobs, b.resp, b.work, b.val, b.ik, nfilters, nwork, npm, tmp_maps);
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
                int max_hits, cudaStream_t s) {
  if (g.cells_total <= 0) return 0;
  dim3 grid((g.cells_total + 255) / 256, ncomp, g.n_frames);
  hits_select<<<grid, 256, 0, s>>>(d_g, ncomp, b.rootv, thresh, d_hits, d_nhits, max_hits);
  return 1;
}

int launch_dt2d_standalone(const float* d_in, int n_maps, int h, int w, const PassGeom* d_pg2 /* [rows, cols] */, const PassMap* d_maps2,
                           float* d_tmp, float* d_out, uint16_t* d_ix, uint16_t* d_iy, uint16_t* d_ixraw, uint16_t* d_iyraw, int backptr_mode,
                           cudaStream_t s) {
  if (n_maps <= 0 || h <= 0 || w <= 0) return 0;
  // maps are launched in chunks so that the warp index stays small; every map is h*w cells
  PassGeom pr{}, pc{};
  pr.n_levels = pc.n_levels = 1;
  pr.nlines[0] = h; pr.N[0] = w; pc.nlines[0] = w; pc.N[0] = h;
  dim3 gr((pass_warps(pr, n_maps) + kPassWarps - 1) / kPassWarps, 1), gc((pass_warps(pc, n_maps) + kPassWarps - 1) / kPassWarps, 1);
  launch_pass(w, gr, s, d_pg2, d_maps2, n_maps, d_in, (size_t)0, d_in, (size_t)0, d_tmp, (size_t)0, d_ixraw, (size_t)0);
  launch_pass(h, gc, s, d_pg2 + 1, d_maps2 + n_maps, n_maps, (const float*)d_tmp, (size_t)0, (const float*)d_tmp, (size_t)0, d_out, (size_t)0,
              d_iyraw, (size_t)0);
  dim3 gx((w + 255) / 256, h, n_maps);
  dt2d_compose<<<gx, 256, 0, s>>>(h, w, d_ixraw, d_iyraw, d_ix, d_iy, backptr_mode);
  return 3;
}

}  // namespace pbd
