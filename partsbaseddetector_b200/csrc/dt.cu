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
// whole stack and re-reading it, the top W entries live in a per-lane shared-memory ring; when a push
// overflows the ring the oldest entry is *retired*: its position range is evaluated and written at once.
// A pop that digs below the ring (rare) reloads the entry from a write-mostly backing store in local
// memory and marks it un-retired, so its (now longer) range is simply written again later; the last write
// to every position is therefore the one the reference's scan would produce.
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

struct Quad {
  double a, b, a2;   // a2 = 2*a (exact), reference evaluates 2*a*(x1-x0) left to right
};
__device__ __forceinline__ Quad make_quad(float w_sq, float w_lin) {
  Quad f;
  f.a = (double)(-w_sq);          // Quadratic(-w[0], -w[1]), src/DynamicProgram.cpp:126-127
  f.b = (double)(-w_lin);
  f.a2 = __dmul_rn(2.0, f.a);
  return f;
}
// Quadratic::operator()(x0,x1,y0,y1), include/DistanceTransform.hpp:98-100
__device__ __forceinline__ float isect(const Quad& f, int x0, int x1, float y0, float y1) {
  const int d = x1 - x0;
  const double dd = (double)d;
  const double t = __dsub_rn(__dsub_rn((double)y1, (double)y0), __dmul_rn(f.b, dd));
  const double num = __dadd_rn(t, __dmul_rn(f.a, (double)(x1 * x1 - x0 * x0)));
  return (float)__ddiv_rn(num, __dmul_rn(f.a2, dd));
}
// Quadratic::operator()(x,y), :102-104
__device__ __forceinline__ float envelope(const Quad& f, int x, float y) {
  return (float)__dadd_rn(__dadd_rn(__dmul_rn(f.a, (double)(x * x)), __dmul_rn(f.b, (double)x)), (double)y);
}

// per-warp shared-memory ring: [slot][lane]
struct Ring {
  float z[kRing][32];
  float y[kRing][32];
  unsigned short v[kRing][32];
};

// One lane's 1-D transform.  loady(q) = src[q] (called for q = 0..N-1 in order, and again for deep-pop reloads);
// emit(i, val, v) stores dst[i] = val, ptr[i] = v (may be called more than once for an i; the last call wins).
template <int MAXN, typename LoadY, typename Emit>
__device__ __forceinline__ void envelope_stream(int N, const Quad& f, int os0, Ring& R, int lane, float* zb, unsigned short* vb,
                                                LoadY loady, Emit emit) {
  const float pos_lo = (float)os0, pos_hi = (float)(os0 + N - 1);
  auto retire = [&](int j, float znext) {
    const int slot = j & (kRing - 1);
    const float zj = R.z[slot][lane];
    // integer positions with zj < pos <= znext, clipped to [os0, os0+N-1]
    const int lo = (zj < pos_lo) ? os0 : (int)floorf(fminf(zj, pos_hi + 1.f)) + 1;
    const int hi = (znext >= pos_hi) ? os0 + N - 1 : (int)floorf(fmaxf(znext, pos_lo - 1.f));
    if (lo > hi) return;
    const int v = R.v[slot][lane];
    const float y = R.y[slot][lane];
    for (int pos = lo; pos <= hi; ++pos) emit(pos - os0, envelope(f, pos - v, y), v);
  };
  int k = 0, ret = 0;
  int vt = 0;
  float yt = loady(0), zt = -INFINITY;
  R.z[0][lane] = zt; R.y[0][lane] = yt; R.v[0][lane] = 0;
  zb[0] = zt; vb[0] = 0;
  for (int q = 1; q < N; ++q) {                                   // :160-170
    const float yq = loady(q);
    float s = isect(f, vt, q, yt, yq);
    while (s <= zt && k > 0) {
      --k;
      const int slot = k & (kRing - 1);
      if (k < ret) {                                              // popped below the ring: reload from the backing store
        ret = k;
        const int vv = vb[k];
        R.v[slot][lane] = (unsigned short)vv; R.z[slot][lane] = zb[k]; R.y[slot][lane] = loady(vv);
      }
      vt = R.v[slot][lane]; yt = R.y[slot][lane]; zt = R.z[slot][lane];
      s = isect(f, vt, q, yt, yq);
    }
    ++k;
    if (k - ret == kRing) {                                       // ring full: retire the oldest entry
      retire(ret, R.z[(ret + 1) & (kRing - 1)][lane]);
      ++ret;
    }
    const int slot = k & (kRing - 1);
    R.v[slot][lane] = (unsigned short)q; R.y[slot][lane] = yq; R.z[slot][lane] = s;
    zb[k] = s; vb[k] = (unsigned short)q;
    vt = q; yt = yq; zt = s;
  }
  for (int j = ret; j <= k; ++j) retire(j, j < k ? R.z[(j + 1) & (kRing - 1)][lane] : INFINITY);
}

// Warp-cooperative source of one row per lane: the 32 rows are staged through a 32x32 shared-memory tile
// (coalesced 128-byte loads, conflict-free transposed reads).  The main loop asks for q = 0,1,2,... in lock
// step across the warp, so the tile is refilled every 32 columns; deep-pop reloads of older samples go to
// global memory.  All 32 lanes must call get() for every q of the main sequence (inactive lanes included).
struct RowTile {
  float (*tile)[33];
  const float* src0;     // first row of the group
  const float* mine;     // this lane's row (a valid row for inactive lanes)
  int nrows, N, lane, myrow, q0;
  __device__ __forceinline__ float get(int q) {
    if (q >= q0 + 32 && (q & 31) == 0) {
      __syncwarp();
      for (int i = 0; i < nrows; ++i)
        if (q + lane < N) tile[i][lane] = __ldg(src0 + (size_t)i * N + q + lane);
      __syncwarp();
      q0 = q;
    }
    if (q >= q0) return tile[myrow][q - q0];
    return __ldg(mine + q);
  }
};

// ---------------------------------------------------------------------------------------------------
// Row pass: one lane per row of one (frame, job, child mixture, level) map; 32 rows per warp.
// ---------------------------------------------------------------------------------------------------
constexpr int kRowWarps = 4;

template <int MAXN>
__global__ void __launch_bounds__(kRowWarps * 32)
dt_rows(const Geometry* __restrict__ g, const int* __restrict__ rg_level, const int* __restrict__ rg_row0, int nrg,
        const PartJob* __restrict__ jobs, const float* __restrict__ resp, const float* __restrict__ work,
        float* __restrict__ tmp, unsigned short* __restrict__ ixdt, int nfilters, int nwork, int ncm, int tmp_maps) {
  __shared__ Ring rings[kRowWarps];
  __shared__ float tiles[kRowWarps][32][33];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rg = blockIdx.x * kRowWarps + wib;
  const PartJob& J = jobs[blockIdx.y / kMaxMix];
  const int mm = blockIdx.y % kMaxMix;
  if (rg >= nrg || mm >= J.nmix) return;                           // warp-uniform
  const int frame = blockIdx.z;
  const LevelDesc& L = g->lv[rg_level[rg]];
  const int row0 = rg_row0[rg];
  const int N = L.ow;
  const int nrows = min(32, L.oh - row0);
  const bool active = lane < nrows;
  const int myrow = active ? lane : 0;                             // inactive lanes shadow row 0 and discard their results
  const size_t ct = (size_t)g->cells_total;
  const float* src0 = (J.in_is_work[mm] ? work + ((size_t)frame * nwork + J.in_slot[mm]) * ct
                                        : resp + ((size_t)frame * nfilters + J.in_slot[mm]) * ct) + L.cell_off + (size_t)row0 * N;
  float* dst = tmp + ((size_t)frame * tmp_maps + J.tmp_base + mm) * ct + L.cell_off + (size_t)(row0 + myrow) * N;
  unsigned short* ptr = ixdt + ((size_t)frame * ncm + J.cm_slot[mm]) * ct + L.cell_off + (size_t)(row0 + myrow) * N;
  const Quad f = make_quad(J.w[mm][0], J.w[mm][1]);
  float zb[MAXN];
  unsigned short vb[MAXN];
  RowTile T{tiles[wib], src0, src0 + (size_t)myrow * N, nrows, N, lane, myrow, -64};
  envelope_stream<MAXN>(N, f, J.ax[mm], rings[wib], lane, zb, vb, [&](int q) { return T.get(q); },
                        [&](int i, float val, int v) { if (active) { dst[i] = val; ptr[i] = (unsigned short)v; } });
}

// ---------------------------------------------------------------------------------------------------
// Column pass: one lane per column (coalesced loads and stores), one warp per 32 columns of one
// (frame, job, child mixture, level) map.  Writes the transformed map (val) and the raw y back-pointer.
// ---------------------------------------------------------------------------------------------------
constexpr int kColWarps = 4;

// software prefetch of a strided column: the next 4 samples are kept in registers so that the load latency
// does not sit on the sequential critical path
struct ColStream {
  const float* src;
  size_t stride;
  int N, next_q;
  float p0, p1, p2, p3;
  __device__ __forceinline__ void init() {
    next_q = 0;
    p0 = __ldg(src);
    p1 = N > 1 ? __ldg(src + stride) : 0.f;
    p2 = N > 2 ? __ldg(src + 2 * stride) : 0.f;
    p3 = N > 3 ? __ldg(src + 3 * stride) : 0.f;
  }
  __device__ __forceinline__ float get(int q) {
    if (q == next_q) {
      const float r = p0;
      p0 = p1; p1 = p2; p2 = p3;
      p3 = (q + 4 < N) ? __ldg(src + (size_t)(q + 4) * stride) : 0.f;
      ++next_q;
      return r;
    }
    return __ldg(src + (size_t)q * stride);                        // deep-pop reload
  }
};

template <int MAXN>
__global__ void __launch_bounds__(kColWarps * 32)
dt_cols(const Geometry* __restrict__ g, const int* __restrict__ cg_level, const int* __restrict__ cg_col0, int ncg,
        const PartJob* __restrict__ jobs, const float* __restrict__ tmp, float* __restrict__ val,
        unsigned short* __restrict__ iyraw, int ncm, int tmp_maps) {
  __shared__ Ring rings[kColWarps];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cg = blockIdx.x * kColWarps + wib;
  const PartJob& J = jobs[blockIdx.y / kMaxMix];
  const int mm = blockIdx.y % kMaxMix;
  if (cg >= ncg || mm >= J.nmix) return;
  const int frame = blockIdx.z;
  const LevelDesc& L = g->lv[cg_level[cg]];
  const int col = cg_col0[cg] + lane;
  if (col >= L.ow) return;
  const int N = L.oh, ow = L.ow;
  const size_t ct = (size_t)g->cells_total;
  const size_t mapoff = ((size_t)frame * tmp_maps + J.tmp_base + mm) * ct + L.cell_off + col;
  float* dst = val + mapoff;
  unsigned short* ptr = iyraw + ((size_t)frame * ncm + J.cm_slot[mm]) * ct + L.cell_off + col;
  const Quad f = make_quad(J.w[mm][2], J.w[mm][3]);
  float zb[MAXN];
  unsigned short vb[MAXN];
  ColStream S{tmp + mapoff, (size_t)ow, N};
  S.init();
  envelope_stream<MAXN>(N, f, J.ay[mm], rings[wib], lane, zb, vb, [&](int q) { return S.get(q); },
                        [&](int i, float v_, int v) { dst[(size_t)i * ow] = v_; ptr[(size_t)i * ow] = (unsigned short)v; });
}

// ---------------------------------------------------------------------------------------------------
// Mixture maximum + parent accumulate (src/DynamicProgram.cpp:134-156), elementwise over all cells of all levels.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mix_max(const Geometry* __restrict__ g, const PartJob* __restrict__ jobs, const float* __restrict__ resp, float* __restrict__ work,
        const float* __restrict__ val, unsigned char* __restrict__ ik, int nfilters, int nwork, int npm, int tmp_maps) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= g->cells_total) return;
  const PartJob& J = jobs[blockIdx.y];
  const int frame = blockIdx.z;
  const size_t ct = (size_t)g->cells_total;
  const int nmix = J.nmix, pnmix = J.pnmix;
  float v[kMaxMix];
#pragma unroll
  for (int mm = 0; mm < kMaxMix; ++mm)
    v[mm] = mm < nmix ? __ldg(val + ((size_t)frame * tmp_maps + J.tmp_base + mm) * ct + idx) : 0.f;
  for (int pm = 0; pm < pnmix; ++pm) {
    float best = -INFINITY;
    int bi = 0;
#pragma unroll
    for (int mm = 0; mm < kMaxMix; ++mm) {
      if (mm < nmix) {
        const float wv = __fadd_rn(v[mm], J.bias[mm][pm]);         // scoresp[mm] + bias(mm)[m], :139
        if (wv > best) { best = wv; bi = mm; }                      // reduceMax: strict >, first wins
      }
    }
    ik[((size_t)frame * npm + J.pm_slot[pm]) * ct + idx] = (unsigned char)bi;
    float* wp = work + ((size_t)frame * nwork + J.out_work_slot[pm]) * ct + idx;
    const float base = J.first_touch ? __ldg(resp + ((size_t)frame * nfilters + J.out_resp_fid[pm]) * ct + idx) : *wp;
    *wp = __fadd_rn(base, best);                                    // parent.score += maxv, :155-156
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
// Standalone 2-D DT (pbd_dt2d_f32 / config-5 microbenchmark): rows, columns, composition.
// ---------------------------------------------------------------------------------------------------
template <int MAXN>
__global__ void __launch_bounds__(kRowWarps * 32)
dt2d_rows(const float* __restrict__ in, int h, int w, const float* __restrict__ defw4, const int* __restrict__ anchor,
          float* __restrict__ tmp, unsigned short* __restrict__ ix) {
  __shared__ Ring rings[kRowWarps];
  __shared__ float tiles[kRowWarps][32][33];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = (blockIdx.x * kRowWarps + wib) * 32;
  if (row0 >= h) return;
  const int m = blockIdx.y;
  const int nrows = min(32, h - row0);
  const bool active = lane < nrows;
  const int myrow = active ? lane : 0;
  const float* src0 = in + ((size_t)m * h + row0) * w;
  float* dst = tmp + ((size_t)m * h + row0 + myrow) * w;
  unsigned short* ptr = ix + ((size_t)m * h + row0 + myrow) * w;
  const Quad f = make_quad(defw4[m * 4 + 0], defw4[m * 4 + 1]);
  float zb[MAXN];
  unsigned short vb[MAXN];
  RowTile T{tiles[wib], src0, src0 + (size_t)myrow * w, nrows, w, lane, myrow, -64};
  envelope_stream<MAXN>(w, f, anchor[m * 2 + 0], rings[wib], lane, zb, vb, [&](int q) { return T.get(q); },
                        [&](int i, float val, int v) { if (active) { dst[i] = val; ptr[i] = (unsigned short)v; } });
}
template <int MAXN>
__global__ void __launch_bounds__(kColWarps * 32)
dt2d_cols(const float* __restrict__ tmp, int h, int w, const float* __restrict__ defw4, const int* __restrict__ anchor,
          float* __restrict__ out, unsigned short* __restrict__ iy) {
  __shared__ Ring rings[kColWarps];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col = (blockIdx.x * kColWarps + wib) * 32 + lane;
  if (col >= w) return;
  const int m = blockIdx.y;
  const size_t base = (size_t)m * h * w + col;
  const Quad f = make_quad(defw4[m * 4 + 2], defw4[m * 4 + 3]);
  float zb[MAXN];
  unsigned short vb[MAXN];
  ColStream S{tmp + base, (size_t)w, h};
  S.init();
  envelope_stream<MAXN>(h, f, anchor[m * 2 + 1], rings[wib], lane, zb, vb, [&](int q) { return S.get(q); },
                        [&](int i, float val, int v) { out[base + (size_t)i * w] = val; iy[base + (size_t)i * w] = (unsigned short)v; });
}
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

int launch_dt_rows_tab(const Geometry& g, const Geometry* d_g, const DeviceBuffers& b, const int* d_rg_level, const int* d_rg_row0, int nrg,
                       int max_ow, const PartJob* d_jobs, int njobs, int nfilters, int nwork, int ncm, int tmp_maps, cudaStream_t s) {
  if (nrg <= 0 || njobs <= 0) return 0;
  dim3 grid((nrg + kRowWarps - 1) / kRowWarps, njobs * kMaxMix, g.n_frames);
#define PBD_ROWS(M) dt_rows<M><<<grid, kRowWarps * 32, 0, s>>>(d_g, d_rg_level, d_rg_row0, nrg, d_jobs, b.resp, b.work, b.tmp, b.ixdt, nfilters, nwork, ncm, tmp_maps)
  if (max_ow <= 160) PBD_ROWS(160);
  else if (max_ow <= 512) PBD_ROWS(512);
  else PBD_ROWS(1024);
#undef PBD_ROWS
  return 1;
}

int launch_dt_cols_tab(const Geometry& g, const Geometry* d_g, const DeviceBuffers& b, const int* d_cg_level, const int* d_cg_col0, int ncg,
                       int max_oh, const PartJob* d_jobs, int njobs, int max_mix, int nfilters, int nwork, int ncm, int npm, int tmp_maps,
                       cudaStream_t s) {
  if (ncg <= 0 || njobs <= 0) return 0;
  (void)max_mix;
  dim3 grid((ncg + kColWarps - 1) / kColWarps, njobs * kMaxMix, g.n_frames);
#define PBD_COLS(M) dt_cols<M><<<grid, kColWarps * 32, 0, s>>>(d_g, d_cg_level, d_cg_col0, ncg, d_jobs, b.tmp, b.val, b.iyraw, ncm, tmp_maps)
  if (max_oh <= 160) PBD_COLS(160);
  else if (max_oh <= 512) PBD_COLS(512);
  else PBD_COLS(1024);
#undef PBD_COLS
  dim3 gm((g.cells_total + 255) / 256, njobs, g.n_frames);
  mix_max<<<gm, 256, 0, s>>>(d_g, d_jobs, b.resp, b.work, b.val, b.ik, nfilters, nwork, npm, tmp_maps);
  return 2;
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

int launch_dt2d_standalone(const float* d_in, int n_maps, int h, int w, const float* d_defw4, const int* d_anchor, float* d_tmp,
                           float* d_out, uint16_t* d_ix, uint16_t* d_iy, uint16_t* d_ixraw, uint16_t* d_iyraw, int backptr_mode,
                           cudaStream_t s) {
  if (n_maps <= 0 || h <= 0 || w <= 0) return 0;
  dim3 gr((h + 32 * kRowWarps - 1) / (32 * kRowWarps), n_maps), gc((w + 32 * kColWarps - 1) / (32 * kColWarps), n_maps);
  if (w <= 160) dt2d_rows<160><<<gr, kRowWarps * 32, 0, s>>>(d_in, h, w, d_defw4, d_anchor, d_tmp, d_ixraw);
  else if (w <= 512) dt2d_rows<512><<<gr, kRowWarps * 32, 0, s>>>(d_in, h, w, d_defw4, d_anchor, d_tmp, d_ixraw);
  else dt2d_rows<4096><<<gr, kRowWarps * 32, 0, s>>>(d_in, h, w, d_defw4, d_anchor, d_tmp, d_ixraw);
  if (h <= 160) dt2d_cols<160><<<gc, kColWarps * 32, 0, s>>>(d_tmp, h, w, d_defw4, d_anchor, d_out, d_iyraw);
  else if (h <= 512) dt2d_cols<512><<<gc, kColWarps * 32, 0, s>>>(d_tmp, h, w, d_defw4, d_anchor, d_out, d_iyraw);
  else dt2d_cols<4096><<<gc, kColWarps * 32, 0, s>>>(d_tmp, h, w, d_defw4, d_anchor, d_out, d_iyraw);
  dim3 gx((w + 255) / 256, h, n_maps);
  dt2d_compose<<<gx, 256, 0, s>>>(h, w, d_ixraw, d_iyraw, d_ix, d_iy, backptr_mode);
  return 3;
}

}  // namespace pbd
