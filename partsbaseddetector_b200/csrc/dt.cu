// dt.cu -- the max-sum tree DP of DynamicProgram<float>::min (reference src/DynamicProgram.cpp:67-173) over the
// Felzenszwalb generalised distance transform DistanceTransform<float> (reference include/DistanceTransform.hpp).
//
// 1-D transform (computeRow, :152-182): dst[q] = max_v src[v] + a (q+os-v)^2 + b (q+os-v), a = -w0 < 0.  The
// reference builds the upper envelope with a stack whose break points z[] are double quotients rounded to
// float; which parabola wins at a near-tie depends on those rounded values, so the literal sequential
// algorithm is reproduced (one lane = one row or one column, stack in thread-local memory, all
// arithmetic with explicit IEEE round-to-nearest double intrinsics so nothing is contracted to FMA).
//
// 2-D transform (compute, :203-245) = row pass (dt_rows: x direction, anchor x) then column pass
// (dt_cols: y direction, anchor y).  dt_cols handles all child mixtures of one part for a column and fuses,
// per cell, the parent-mixture maximum  max_mm(dt[mm] + bias[mm][pm])  (Math::reduceMax, include/Math.hpp:149-185),
// the best-mixture index Ik and the accumulation into the parent's working score (:134-156) into its
// write, so the transformed maps never round-trip through HBM.  The reference's back-pointer composition
// Iy[y][x] <- Iy[y][Ix[y][x]] (:232-244) and the per-parent-mixture gather (Math::reducePickIndex) are NOT
// materialised: the raw row-pass / column-pass argmaxes are kept per child mixture (u16) and composed lazily
// by the backtrack (backtrack.cu), which is what makes the DP write 5 B instead of 12 B per (cell, map).
#include <algorithm>
#include <cfloat>
#include "kernels.cuh"

namespace pbd {
namespace {

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

// Lane-private envelope stack.  v = parabola positions, z = break points (float, as the reference's T z[]),
// y = src[v] (kept so that pops never re-read the source).
template <int MAXN>
struct Stack {
  unsigned short v[MAXN];
  float z[MAXN + 1];
  float y[MAXN];
};

// Build phase of computeRow (:154-170).  load(q) returns src[q].  Returns k (top index).
template <int MAXN, typename Load>
__device__ __forceinline__ int build_envelope(Stack<MAXN>& st, int N, const Quad& f, Load load) {
  int k = 0;
  int vt = 0;
  float yt = load(0);
  float zt = -INFINITY;
  st.v[0] = 0; st.y[0] = yt; st.z[0] = zt;
  for (int q = 1; q < N; ++q) {
    const float yq = load(q);
    float s = isect(f, vt, q, yt, yq);
    while (s <= zt && k > 0) {
      --k;
      vt = st.v[k]; yt = st.y[k]; zt = st.z[k];
      s = isect(f, vt, q, yt, yq);
    }
    ++k;
    st.v[k] = (unsigned short)q; st.y[k] = yq; st.z[k] = s;
    vt = q; yt = yq; zt = s;
  }
  st.z[k + 1] = INFINITY;
  return k;
}

// ---------------------------------------------------------------------------------------------------
// Row pass: one lane per row of one (frame, job, child mixture, level) map.
// ---------------------------------------------------------------------------------------------------
template <int MAXN>
__global__ void __launch_bounds__(128)
dt_rows(const Geometry* __restrict__ g, const int* __restrict__ rg_level, const int* __restrict__ rg_row0, int nrg,
        const PartJob* __restrict__ jobs, const float* __restrict__ resp, const float* __restrict__ work,
        float* __restrict__ tmp, unsigned short* __restrict__ ixdt, int nfilters, int nwork, int ncm, int tmp_maps) {
  const int rg = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (rg >= nrg) return;
  const int lane = threadIdx.x & 31;
  const PartJob& J = jobs[blockIdx.y / kMaxMix];
  const int mm = blockIdx.y % kMaxMix;
  if (mm >= J.nmix) return;
  const int frame = blockIdx.z;
  const LevelDesc& L = g->lv[rg_level[rg]];
  const int row = rg_row0[rg] + lane;
  if (row >= L.oh) return;
  const int N = L.ow;
  const size_t ct = (size_t)g->cells_total;
  const float* src = (J.in_is_work[mm] ? work + ((size_t)frame * nwork + J.in_slot[mm]) * ct
                                       : resp + ((size_t)frame * nfilters + J.in_slot[mm]) * ct) + L.cell_off + (size_t)row * N;
  float* dst = tmp + ((size_t)frame * tmp_maps + J.tmp_base + mm) * ct + L.cell_off + (size_t)row * N;
  unsigned short* ptr = ixdt + ((size_t)frame * ncm + J.cm_slot[mm]) * ct + L.cell_off + (size_t)row * N;
  const Quad f = make_quad(J.w[mm][0], J.w[mm][1]);
  Stack<MAXN> st;
  build_envelope<MAXN>(st, N, f, [&](int q) { return __ldg(src + q); });
  int k = 0;
  int os = J.ax[mm];
  for (int q = 0; q < N; ++q, ++os) {                       // :172-178
    while (st.z[k + 1] < (float)os) ++k;
    const int v = st.v[k];
    dst[q] = envelope(f, os - v, st.y[k]);
    ptr[q] = (unsigned short)v;
  }
}

// ---------------------------------------------------------------------------------------------------
// Column pass + mixture maximum + parent accumulate.  One CTA per (frame, job, level, 32 columns); warp w runs
// the column transform of child mixture w (lane = column), so all mixtures of a part advance concurrently.
// The evaluation phase proceeds in chunks of CH rows: every warp deposits its CH x 32 transformed values in
// shared memory, then the warps split the parent mixtures and compute, per cell,
//   max_mm(dt[mm] + bias[mm][pm])  (Math::reduceMax, strict >, first wins), Ik, and parent.score += max.
// ---------------------------------------------------------------------------------------------------
template <int MAXN, int CH>
__global__ void __launch_bounds__(kMaxMix * 32)
dt_cols(const Geometry* __restrict__ g, const int* __restrict__ cg_level, const int* __restrict__ cg_col0, int ncg,
        const PartJob* __restrict__ jobs, const float* __restrict__ resp, float* __restrict__ work,
        const float* __restrict__ tmp, unsigned short* __restrict__ iyraw, unsigned char* __restrict__ ik,
        int nfilters, int nwork, int ncm, int npm, int tmp_maps) {
  __shared__ float sval[2][kMaxMix][CH][32];
  const int cg = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const PartJob& J = jobs[blockIdx.y];
  const int frame = blockIdx.z;
  const LevelDesc& L = g->lv[cg_level[cg]];
  const int col = cg_col0[cg] + lane;
  const int N = L.oh, ow = L.ow;
  const size_t ct = (size_t)g->cells_total;
  const int nmix = J.nmix, pnmix = J.pnmix;
  const bool col_ok = col < ow;
  const bool mine = col_ok && warp < nmix;      // this thread owns the column transform of mixture `warp`

  Stack<MAXN> st;
  Quad f = make_quad(1.f, 0.f);
  int os = 0;
  unsigned short* iyp = nullptr;
  if (mine) {
    const float* src = tmp + ((size_t)frame * tmp_maps + J.tmp_base + warp) * ct + L.cell_off + col;
    f = make_quad(J.w[warp][2], J.w[warp][3]);
    build_envelope<MAXN>(st, N, f, [&](int q) { return __ldg(src + (size_t)q * ow); });
    os = J.ay[warp];
    iyp = iyraw + ((size_t)frame * ncm + J.cm_slot[warp]) * ct + L.cell_off + col;
  }
  int k = 0;
  for (int q0 = 0; q0 < N; q0 += CH) {
    const int buf = (q0 / CH) & 1;
    if (mine) {
#pragma unroll 1
      for (int qq = 0; qq < CH && q0 + qq < N; ++qq) {          // :172-178
        const int pos = os + q0 + qq;
        while (st.z[k + 1] < (float)pos) ++k;
        const int v = st.v[k];
        sval[buf][warp][qq][lane] = envelope(f, pos - v, st.y[k]);
        iyp[(size_t)(q0 + qq) * ow] = (unsigned short)v;
      }
    }
    __syncthreads();
    if (col_ok) {
      for (int pm = warp; pm < pnmix; pm += nwarps) {           // src/DynamicProgram.cpp:134-156
        unsigned char* ikp = ik + ((size_t)frame * npm + J.pm_slot[pm]) * ct + L.cell_off + col;
        float* wp = work + ((size_t)frame * nwork + J.out_work_slot[pm]) * ct + L.cell_off + col;
        const float* rp = resp + ((size_t)frame * nfilters + J.out_resp_fid[pm]) * ct + L.cell_off + col;
        float bias[kMaxMix];
#pragma unroll
        for (int mm = 0; mm < kMaxMix; ++mm) bias[mm] = mm < nmix ? J.bias[mm][pm] : 0.f;
        for (int qq = 0; qq < CH && q0 + qq < N; ++qq) {
          float best = -INFINITY;
          int bi = 0;
#pragma unroll
          for (int mm = 0; mm < kMaxMix; ++mm) {
            if (mm < nmix) {
              const float wv = __fadd_rn(sval[buf][mm][qq][lane], bias[mm]);   // scoresp[mm] + bias(mm)[m], :139
              if (wv > best) { best = wv; bi = mm; }                            // reduceMax: strict >, first wins
            }
          }
          const size_t o = (size_t)(q0 + qq) * ow;
          ikp[o] = (unsigned char)bi;
          const float base = J.first_touch ? __ldg(rp + o) : wp[o];
          wp[o] = __fadd_rn(base, best);                                         // parent.score += maxv, :155-156
        }
      }
    }
    // double-buffered sval: the next chunk writes the other buffer, and the chunk after that is separated
    // from these reads by the next __syncthreads
  }
}

// ---------------------------------------------------------------------------------------------------
// Root: rootv = max_m(score[m] + bias), rooti = argmax (:163-171); hits rootv > thresh appended (:208-211).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
root_select(const Geometry* __restrict__ g, const RootJob* __restrict__ roots, int ncomp, const float* __restrict__ resp,
            const float* __restrict__ work, int nfilters, int nwork, float thresh, float* __restrict__ rootv,
            unsigned char* __restrict__ rooti, Hit* __restrict__ hits, int* __restrict__ nhits, int max_hits) {
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
  if (hits != nullptr && best > thresh) {
    int l = 0;
    while (l + 1 < g->n_levels && idx >= g->lv[l + 1].cell_off) ++l;
    const int local = idx - g->lv[l].cell_off;
    const int slot = atomicAdd(nhits, 1);
    if (slot < max_hits) {
      Hit h;
      h.frame = frame; h.level = l; h.comp = comp; h.y = local / g->lv[l].ow; h.x = local % g->lv[l].ow; h.score = best;
      hits[slot] = h;
    }
  }
}

// Hits of an already computed rootv (DynamicProgram::argmin threshold + Math::find, :208-211).
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
__global__ void __launch_bounds__(128)
dt2d_rows(const float* __restrict__ in, int h, int w, const float* __restrict__ defw4, const int* __restrict__ anchor,
          float* __restrict__ tmp, unsigned short* __restrict__ ix) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= h) return;
  const int m = blockIdx.y;
  const size_t base = ((size_t)m * h + row) * w;
  const Quad f = make_quad(defw4[m * 4 + 0], defw4[m * 4 + 1]);
  Stack<MAXN> st;
  build_envelope<MAXN>(st, w, f, [&](int q) { return __ldg(in + base + q); });
  int k = 0, os = anchor[m * 2 + 0];
  for (int q = 0; q < w; ++q, ++os) {
    while (st.z[k + 1] < (float)os) ++k;
    const int v = st.v[k];
    tmp[base + q] = envelope(f, os - v, st.y[k]);
    ix[base + q] = (unsigned short)v;
  }
}
template <int MAXN>
__global__ void __launch_bounds__(128)
dt2d_cols(const float* __restrict__ tmp, int h, int w, const float* __restrict__ defw4, const int* __restrict__ anchor,
          float* __restrict__ out, unsigned short* __restrict__ iy) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= w) return;
  const int m = blockIdx.y;
  const size_t base = (size_t)m * h * w + col;
  const Quad f = make_quad(defw4[m * 4 + 2], defw4[m * 4 + 3]);
  Stack<MAXN> st;
  build_envelope<MAXN>(st, h, f, [&](int q) { return __ldg(tmp + base + (size_t)q * w); });
  int k = 0, os = anchor[m * 2 + 1];
  for (int q = 0; q < h; ++q, ++os) {
    while (st.z[k + 1] < (float)os) ++k;
    const int v = st.v[k];
    out[base + (size_t)q * w] = envelope(f, os - v, st.y[k]);
    iy[base + (size_t)q * w] = (unsigned short)v;
  }
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
  dim3 grid((nrg + 3) / 4, njobs * kMaxMix, g.n_frames);
#define PBD_ROWS(M) dt_rows<M><<<grid, 128, 0, s>>>(d_g, d_rg_level, d_rg_row0, nrg, d_jobs, b.resp, b.work, b.tmp, b.ixdt, nfilters, nwork, ncm, tmp_maps)
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
  dim3 grid(ncg, njobs, g.n_frames);
  const int threads = 32 * std::min(std::max(max_mix, 1), kMaxMix);
#define PBD_COLS(M) dt_cols<M, 8><<<grid, threads, 0, s>>>(d_g, d_cg_level, d_cg_col0, ncg, d_jobs, b.resp, b.work, b.tmp, b.iyraw, b.ik, nfilters, nwork, ncm, npm, tmp_maps)
  if (max_oh <= 160) PBD_COLS(160);
  else if (max_oh <= 512) PBD_COLS(512);
  else PBD_COLS(1024);
#undef PBD_COLS
  return 1;
}

int launch_root(const Geometry& g, const Geometry* d_g, const DeviceBuffers& b, const RootJob* d_roots, int ncomp, int nfilters, int nwork,
                float thresh, Hit* d_hits, int* d_nhits, int max_hits, cudaStream_t s) {
  if (g.cells_total <= 0) return 0;
  dim3 grid((g.cells_total + 255) / 256, ncomp, g.n_frames);
  root_select<<<grid, 256, 0, s>>>(d_g, d_roots, ncomp, b.resp, b.work, nfilters, nwork, thresh, b.rootv, b.rooti, d_hits, d_nhits, max_hits);
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
  dim3 gr((h + 127) / 128, n_maps), gc((w + 127) / 128, n_maps);
  if (w <= 160) dt2d_rows<160><<<gr, 128, 0, s>>>(d_in, h, w, d_defw4, d_anchor, d_tmp, d_ixraw);
  else if (w <= 512) dt2d_rows<512><<<gr, 128, 0, s>>>(d_in, h, w, d_defw4, d_anchor, d_tmp, d_ixraw);
  else dt2d_rows<4096><<<gr, 128, 0, s>>>(d_in, h, w, d_defw4, d_anchor, d_tmp, d_ixraw);
  if (h <= 160) dt2d_cols<160><<<gc, 128, 0, s>>>(d_tmp, h, w, d_defw4, d_anchor, d_out, d_iyraw);
  else if (h <= 512) dt2d_cols<512><<<gc, 128, 0, s>>>(d_tmp, h, w, d_defw4, d_anchor, d_out, d_iyraw);
  else dt2d_cols<4096><<<gc, 128, 0, s>>>(d_tmp, h, w, d_defw4, d_anchor, d_out, d_iyraw);
  dim3 gx((w + 255) / 256, h, n_maps);
  dt2d_compose<<<gx, 256, 0, s>>>(h, w, d_ixraw, d_iyraw, d_ix, d_iy, backptr_mode);
  return 3;
}

}  // namespace pbd
