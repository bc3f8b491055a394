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
#include "dt_envelope.cuh"

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
template <int MAXN>
__global__ void __launch_bounds__(kPassWarps * 32, PBD_DT_MINBLOCKS)
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
  auto store = [&](int i, float val, unsigned short v) {
    const unsigned off = (unsigned)i * (unsigned)nlines;
    st_f32(dst, off, val); st_u16(dp, off, v);
  };
  const int os0 = M.os;
#if defined(PBD_DT_SCAN)
  // A/B switch (tools/build_variant.sh ... -DPBD_DT_SCAN=4): lagged-scan emission (env::envelope_scan<LAG>), one position per lane and
  // step, the same index in all lanes.  Bit-identical on the host (tests/test_dt_envelope_host.py); not yet measured on a GPU.
  env::envelope_scan<PBD_DT_SCAN>(N, f, os0, rings[wib], lane, zb, pb, loady, reload,
                                  [&](int i, float val, int v) { store(i, val, (unsigned short)v); });
#elif !defined(PBD_DT_WINDOWED_STORES)
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
