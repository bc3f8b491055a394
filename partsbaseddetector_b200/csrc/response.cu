// response.cu -- SpatialConvolutionEngine::pdf (reference src/SpatialConvolutionEngine.cpp:70-124) for all
// filters x all levels x all frames of a batch in one launch.
//
//   resp[f](y,x) = sum_{c<32} R_c,   R_c = sum_{ky,kx} w_f[ky][kx][c] * F_c(y+ky-ay, x+kx-ax),  ay = kh/2, ax = kw/2
// with F_c outside the map = 0 for c < 31 and 1 for c = 31 (the BORDER_CONSTANT engines built at
// src/SpatialConvolutionEngine.cpp:147-156).  In EXACT mode every product and every sum is rounded
// separately (__fmul_rn/__fadd_rn) and accumulated in the reference's order -- taps row-major inside a
// channel (cv::Filter2D, src/filter.cpp:3898-3922), then channels 0..31 (:85-93) -- so the scores are
// bit-identical to the CPU path.  In fast mode the taps are fused multiply-adds into one accumulator.
//
// This stage is a dense contraction (M = cells, N = filters, K = kh*kw*32 = 800; 325 FLOP per algorithmic
// byte), so it is bounded by the FP32 issue rate, not by HBM.  Blocking (sm_100a, measured issue rates in
// DESIGN.md): a CTA owns a TY x TX tile of cells of one level; the HOG tile plus halo is staged once into
// shared memory, transposed from HWC to channel-planar so that a thread's P=4 adjacent cells are one
// 128-bit load; filters are pre-packed on the host as [group of 8][channel][tap][8] and streamed through a
// double-buffered TMA bulk-copy stage (cp.async.bulk + mbarrier, one 6.4 KB copy per filter group and channel chunk),
// so a thread's 8 filter taps are two broadcast 128-bit loads.  Each
// thread keeps 4 cells x 8 filters in registers: 64 FMUL+FADD (or 32..64 FFMA) per 6 shared-memory loads.
#include <vector>

#include "kernels.cuh"

namespace pbd {
namespace {

constexpr int P = 4, Q = 8;          // register tile: cells (along x) x filters
// Cells per CTA tile = one warp of quads (P adjacent cells per lane).  Two shapes, chosen PER LEVEL by response_plan_tiles: 8 rows x 4
// quads (32 lanes) and 6 rows x 5 quads (30 lanes; lanes 30, 31 idle) -- whichever needs fewer tiles for the level's size: the VGA
// pyramid takes 420 tiles instead of 428 (8x16 alone pads it to 90 % real cells; 8x32 only 84 %).  Both stage 240 floats per channel
// (12 x 20 / 10 x 24), and each shape is its own kernel instance with compile-time strides, launched over its own tile list.
template <int SHAPE> struct TileShape;
template <> struct TileShape<0> { static constexpr int TY = 8, LX = 4; };
template <> struct TileShape<1> { static constexpr int TY = 6, LX = 5; };
constexpr int POSW = 1;              // warps covering the tile's cells
constexpr int WF = 6;                // filter groups (of Q) processed per pass by different warps
constexpr int NT = POSW * WF * 32;   // threads per CTA (192)
constexpr int CCH = 8;               // channels per weight stage

// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier: one elected lane moves a whole weight slab ----
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  unsigned done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(a), "r"(parity) : "memory");
  } while (!done);
}

template <int KH, int KW, bool EXACT, int SHAPE>
__global__ void __launch_bounds__(NT, 2)
part_response(const Geometry* __restrict__ g, const int* __restrict__ tile_level, const int* __restrict__ tile_first,
              const float* __restrict__ feat, const float* __restrict__ wbank, int nfilters, int ngroups, int trunc_zero,
              float* __restrict__ resp) {
  constexpr int TY = TileShape<SHAPE>::TY, LX = TileShape<SHAPE>::LX, TX = LX * P;
  constexpr int TAPS = KH * KW;
  constexpr int HY = TY + KH - 1;                       // tile rows incl. halo
  constexpr int ROWP = ((TX + KW - 1 + 3) / 4) * 4;     // padded row pitch (words), 16-byte aligned rows
  constexpr int PLANE = HY * ROWP;
  constexpr int WSLAB = CCH * TAPS * Q;                 // floats per (group, channel-chunk)
  extern __shared__ __align__(16) float smem[];
  float* sfeat = smem;                                  // [32][HY][ROWP]
  float* sw = smem + 32 * PLANE;                        // [2][WF][CCH][TAPS][Q]

  const int tile = blockIdx.x, frame = blockIdx.y;
  const int l = tile_level[tile];
  const LevelDesc& L = g->lv[l];
  const int tiles_x = (L.ow + TX - 1) / TX;
  const int tl = tile - tile_first[l];                  // tile_first: first tile of the level within this shape's list
  const int y0 = (tl / tiles_x) * TY, x0 = (tl % tiles_x) * TX;
  const int ow = L.ow, oh = L.oh;
  constexpr int AY = KH / 2, AX = KW / 2;               // anchor, include/filterengine.hpp:310-318
  const int tid = threadIdx.x;

  // ---- stage the HOG tile (+halo), HWC -> channel-planar, border 0 (c<31) / 1 (c=31) ----
  const float* fbase = feat + ((size_t)frame * g->cells_total + L.cell_off) * 32;
  for (int i = tid; i < HY * ROWP * 8; i += NT) {
    const int cq = i / (HY * ROWP);                     // channel quad 0..7
    const int pos = i % (HY * ROWP);
    const int ty = pos / ROWP, tx = pos % ROWP;
    const int gy = y0 + ty - AY, gx = x0 + tx - AX;
    float4 v;
    if (gy >= 0 && gy < oh && gx >= 0 && gx < ow) v = __ldg(reinterpret_cast<const float4*>(fbase + ((size_t)gy * ow + gx) * 32 + cq * 4));
    else v = make_float4(0.f, 0.f, 0.f, cq == 7 ? 1.f : 0.f);
    sfeat[(cq * 4 + 0) * PLANE + pos] = v.x;
    sfeat[(cq * 4 + 1) * PLANE + pos] = v.y;
    sfeat[(cq * 4 + 2) * PLANE + pos] = v.z;
    sfeat[(cq * 4 + 3) * PLANE + pos] = v.w;
  }

  // Channel 31 (truncation feature) is identically zero inside a map produced by the HOG stage (src/HOGFeatures.cpp:338;
  // trunc_zero = 0 for features injected through pbd_set_features), so for a tile whose halo
  // lies entirely inside the map its 25 products are all +-0 and acc + (+-0) == acc: the channel is skipped exactly.
  const bool interior = trunc_zero && (y0 - AY >= 0) && (x0 - AX >= 0) && (y0 + TY - 1 - AY + KH - 1 < oh) && (x0 + ROWP - 1 - AX < ow);
  const int nch_last = interior ? CCH - 1 : CCH;
  const int warp = tid >> 5, lane = tid & 31;
  const int wf = warp;                                  // which filter group of the pass
  const int lx = lane % LX, lyr = lane / LX;            // lanes along x / y inside the warp
  const bool lane_on = lyr < TY;                        // 6 x 5: lanes 30 and 31 shadow the last row and store nothing
  const int cy = lane_on ? lyr : TY - 1;                // tile-local cell row
  const int cx = lx * P;                                // tile-local first cell column
  const int npass = (ngroups + WF - 1) / WF;

  // Weight slabs are private to a filter group: the POSW warps that share group `wf` stage and consume the slab
  // themselves and synchronise on their own named barrier, so the groups of a CTA never wait for each other.
  const int gtid = tid - wf * (POSW * 32);               // thread index inside the group (warps wf*POSW .. +POSW-1)
  // The slab of (group, channel chunk) is one contiguous WSLAB*4-byte run of the packed bank: a single TMA bulk copy issued
  // by the group's first lane, completion signalled on the group's mbarrier of that buffer.
  __shared__ __align__(8) unsigned long long wbar[2][WF];
  auto stage_weights = [&](int pass, int chunk, int buf) {
    if (gtid != 0) return;
    int gq = pass * WF + wf;
    if (gq >= ngroups) gq = ngroups - 1;                  // idle groups read a valid slab, results discarded
    const float* srcw = wbank + ((size_t)gq * 32 + chunk * CCH) * TAPS * Q;   // contiguous WSLAB floats
    float* dstw = sw + ((size_t)buf * WF + wf) * WSLAB;
    mbar_expect_tx(&wbar[buf][wf], WSLAB * 4);
    tma_bulk_g2s(dstw, srcw, WSLAB * 4, &wbar[buf][wf]);
  };
  auto group_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(wf + 1), "n"(POSW * 32) : "memory"); };

  constexpr int NCH = 32 / CCH;
  if (tid < 2 * WF) mbar_init(&wbar[tid / WF][tid % WF], 1);
  mbar_fence_init();
  __syncthreads();                                      // HOG tile staged, barriers initialised
  stage_weights(0, 0, 0);
  int it = 0;                                           // global stage counter over (pass, chunk)
  for (int pass = 0; pass < npass; ++pass) {
    float acc[P][Q];
#pragma unroll
    for (int i = 0; i < P; ++i)
#pragma unroll
      for (int j = 0; j < Q; ++j) acc[i][j] = 0.f;

    for (int chunk = 0; chunk < NCH; ++chunk, ++it) {
      const int buf = it & 1;
      // prefetch the next stage, then wait for the current one
      const int nchunk = chunk + 1 == NCH ? 0 : chunk + 1;
      const int npassi = chunk + 1 == NCH ? pass + 1 : pass;
      if (npassi < npass) stage_weights(npassi, nchunk, buf ^ 1);
      mbar_wait(&wbar[buf][wf], (it >> 1) & 1);         // this group's slab has landed (n-th use of the buffer: parity n & 1)
      const float* wsl = sw + ((size_t)buf * WF + wf) * WSLAB;
      const int ncl = (chunk == 32 / CCH - 1) ? nch_last : CCH;
#pragma unroll 1
      for (int cl = 0; cl < ncl; ++cl) {
        const int c = chunk * CCH + cl;
        const float* fp = sfeat + c * PLANE + cy * ROWP + cx;
        const float* wp = wsl + cl * TAPS * Q;
        float r[P][Q];
#pragma unroll
        for (int ky = 0; ky < KH; ++ky) {
          float xs[P + KW - 1];
#pragma unroll
          for (int i = 0; i < (P + KW - 1 + 3) / 4; ++i) {
            const float4 v = *reinterpret_cast<const float4*>(fp + ky * ROWP + 4 * i);
            if (4 * i + 0 < P + KW - 1) xs[4 * i + 0] = v.x;
            if (4 * i + 1 < P + KW - 1) xs[4 * i + 1] = v.y;
            if (4 * i + 2 < P + KW - 1) xs[4 * i + 2] = v.z;
            if (4 * i + 3 < P + KW - 1) xs[4 * i + 3] = v.w;
          }
#pragma unroll
          for (int kx = 0; kx < KW; ++kx) {
            const float4 w0 = *reinterpret_cast<const float4*>(wp + (ky * KW + kx) * Q);
            const float4 w1 = *reinterpret_cast<const float4*>(wp + (ky * KW + kx) * Q + 4);
            const float w[Q] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
            for (int i = 0; i < P; ++i)
#pragma unroll
              for (int j = 0; j < Q; ++j) {
                if (EXACT) {
                  const float pr = __fmul_rn(w[j], xs[i + kx]);
                  r[i][j] = (ky == 0 && kx == 0) ? pr : __fadd_rn(r[i][j], pr);
                } else {
                  acc[i][j] = __fmaf_rn(w[j], xs[i + kx], acc[i][j]);
                }
              }
          }
        }
        if (EXACT) {
#pragma unroll
          for (int i = 0; i < P; ++i)
#pragma unroll
            for (int j = 0; j < Q; ++j) acc[i][j] = __fadd_rn(acc[i][j], r[i][j]);
        }
      }
      group_sync();                                     // the group is done with `buf` before it is refilled
    }
    // ---- write the pass's responses: resp[frame][f][cell] ----
    const int gq = pass * WF + wf;
    const int gy = y0 + cy;
    if (gq < ngroups && gy < oh && lane_on) {
#pragma unroll
      for (int j = 0; j < Q; ++j) {
        const int f = gq * Q + j;
        if (f >= nfilters) break;
        float* dst = resp + ((size_t)frame * nfilters + f) * g->cells_total + L.cell_off + (size_t)gy * ow + x0 + cx;
#pragma unroll
        for (int i = 0; i < P; ++i)
          if (x0 + cx + i < ow) dst[i] = acc[i][j];
      }
    }
  }
}

// Generic fallback for models whose filters are not all the same size (e.g. 15x5 roots with 6x6 parts):
// one thread per cell, loops over all filters; same arithmetic order, weights read through L1 (uniform).
template <bool EXACT>
__global__ void __launch_bounds__(128)
part_response_generic(const Geometry* __restrict__ g, const int* __restrict__ tile_level, const int* __restrict__ tile_first,
                      const float* __restrict__ feat, const float* __restrict__ wg, const int* __restrict__ foff,
                      const int* __restrict__ fkh, const int* __restrict__ fkw, int nfilters, int khm, int kwm,
                      float* __restrict__ resp) {
  constexpr int GX = 16, GY = 8;
  extern __shared__ __align__(16) float smem[];
  const int HYm = GY + khm - 1, ROWm = GX + kwm - 1, PLANEm = HYm * ROWm + 1;
  const int tile = blockIdx.x, frame = blockIdx.y;
  const int l = tile_level[tile];
  const LevelDesc& L = g->lv[l];
  const int tiles_x = (L.ow + GX - 1) / GX;
  const int tl = tile - tile_first[l];
  const int y0 = (tl / tiles_x) * GY, x0 = (tl % tiles_x) * GX;
  const int ow = L.ow, oh = L.oh;
  const int aym = khm / 2, axm = kwm / 2;               // tile halo is laid out for the largest anchor
  const float* fbase = feat + ((size_t)frame * g->cells_total + L.cell_off) * 32;
  for (int i = threadIdx.x; i < HYm * ROWm * 32; i += blockDim.x) {
    const int c = i % 32, pos = i / 32;
    const int ty = pos / ROWm, tx = pos % ROWm;
    const int gy = y0 + ty - aym, gx = x0 + tx - axm;
    float v = (c == 31) ? 1.f : 0.f;
    if (gy >= 0 && gy < oh && gx >= 0 && gx < ow) v = fbase[((size_t)gy * ow + gx) * 32 + c];
    smem[c * PLANEm + pos] = v;
  }
  __syncthreads();
  const int cx = threadIdx.x % GX, cy = threadIdx.x / GX;
  const int gx = x0 + cx, gy = y0 + cy;
  if (gx >= ow || gy >= oh) return;
  for (int f = 0; f < nfilters; ++f) {
    const int kh = fkh[f], kw = fkw[f];
    const int ay = kh / 2, ax = kw / 2;
    const float* w = wg + foff[f];                      // [c][ky][kx]
    float acc = 0.f;
    for (int c = 0; c < 32; ++c) {
      const float* fp = smem + c * PLANEm + (cy + aym - ay) * ROWm + (cx + axm - ax);
      float r = 0.f;
      for (int ky = 0; ky < kh; ++ky)
        for (int kx = 0; kx < kw; ++kx) {
          const float wv = __ldg(w + (c * kh + ky) * kw + kx);
          if (EXACT) r = __fadd_rn(r, __fmul_rn(wv, fp[ky * ROWm + kx]));
          else acc = __fmaf_rn(wv, fp[ky * ROWm + kx], acc);
        }
      if (EXACT) acc = __fadd_rn(acc, r);
    }
    resp[((size_t)frame * nfilters + f) * g->cells_total + L.cell_off + (size_t)gy * ow + gx] = acc;
  }
}

template <int KH, int KW, int SHAPE>
size_t fast_smem_bytes() {
  constexpr int HY = TileShape<SHAPE>::TY + KH - 1, ROWP = ((TileShape<SHAPE>::LX * P + KW - 1 + 3) / 4) * 4;
  return (size_t)(32 * HY * ROWP + 2 * WF * CCH * KH * KW * Q) * sizeof(float);
}

template <int KH, int KW, int SHAPE>
void launch_fast_shape(const Geometry* d_g, const int* d_tl, const int* d_tf, int ntiles, int nframes, const DeviceBuffers& b,
                       const FilterBank& fb, int exact, int trunc_zero, cudaStream_t s) {
  if (ntiles <= 0) return;
  const size_t smem = fast_smem_bytes<KH, KW, SHAPE>();
  // per launch: the attribute is per device and a process may drive several devices
  if (exact) cudaFuncSetAttribute(part_response<KH, KW, true, SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  else cudaFuncSetAttribute(part_response<KH, KW, false, SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid(ntiles, nframes);
  if (exact) part_response<KH, KW, true, SHAPE><<<grid, NT, smem, s>>>(d_g, d_tl, d_tf, b.feat, fb.w, fb.nfilters, fb.ngroups, trunc_zero, b.resp);
  else part_response<KH, KW, false, SHAPE><<<grid, NT, smem, s>>>(d_g, d_tl, d_tf, b.feat, fb.w, fb.nfilters, fb.ngroups, trunc_zero, b.resp);
}
// tiles [0, ntiles0) of the list are 8 x 4-quad tiles, the rest 6 x 5; d_tf holds the first-tile table of either shape (n_levels ints each)
template <int KH, int KW>
int launch_fast(const Geometry* d_g, const int* d_tl, const int* d_tf, int n_levels, int ntiles0, int ntiles, int nframes, const DeviceBuffers& b,
                const FilterBank& fb, int exact, int trunc_zero, cudaStream_t s) {
  launch_fast_shape<KH, KW, 0>(d_g, d_tl, d_tf, ntiles0, nframes, b, fb, exact, trunc_zero, s);
  launch_fast_shape<KH, KW, 1>(d_g, d_tl + ntiles0, d_tf + n_levels, ntiles - ntiles0, nframes, b, fb, exact, trunc_zero, s);
  return (ntiles0 > 0) + (ntiles > ntiles0);
}

}  // namespace

}  // namespace pbd

namespace pbd {

bool response_has_fast_path(const FilterBank& fb) {
  return fb.uniform && fb.flen == 32 && fb.kh == fb.kw && (fb.kh == 4 || fb.kh == 5 || fb.kh == 6);
}

// Tiles of the response kernels.  tile_level = [tiles of shape 0 ..., tiles of shape 1 ...] (level of every tile), level_first =
// [first tile of level l within the shape-0 list ... | ... within the shape-1 list] (2 n_levels ints).  A level uses the shape that
// needs fewer tiles (ties: 8 x 4); the generic kernel (filters of different sizes) keeps 8 x 16 cells for every level.
int response_plan_tiles(const Geometry& g, const FilterBank& fb, std::vector<int>& tile_level, std::vector<int>& level_first) {
  const bool fast = response_has_fast_path(fb);
  std::vector<int> t0, t1;
  level_first.assign((size_t)2 * g.n_levels, 0);
  // the second shape is a second launch: worth it only when the launches are many waves of CTAs long (a single VGA frame is 1.4 waves:
  // measured 0.66 ms with two launches, 0.48 ms with one)
  // ... and when it saves at least 1 % of the tiles (1080p / 10 levels: 4 of 2 755 tiles -- the second launch's tail costs more)
  long long all0 = 0, best = 0;
  for (int l = 0; l < g.n_levels; ++l) {
    const long long n0 = (long long)((g.lv[l].oh + 7) / 8) * ((g.lv[l].ow + 15) / 16);
    const long long n1 = (long long)((g.lv[l].oh + 5) / 6) * (((g.lv[l].ow + P - 1) / P + 4) / 5);
    all0 += n0; best += n1 < n0 ? n1 : n0;
  }
  const bool two_shapes = fast && all0 * g.n_frames >= 2400 && (all0 - best) * 100 >= all0;
  for (int l = 0; l < g.n_levels; ++l) {
    const LevelDesc& L = g.lv[l];
    const int qw = (L.ow + P - 1) / P;
    const int n0 = fast ? ((L.oh + 7) / 8) * ((qw + 3) / 4) : ((L.oh + 7) / 8) * ((L.ow + 15) / 16);
    const int n1 = ((L.oh + 5) / 6) * ((qw + 4) / 5);
    level_first[l] = (int)t0.size(); level_first[g.n_levels + l] = (int)t1.size();
    if (two_shapes && n1 < n0) t1.insert(t1.end(), n1, l); else t0.insert(t0.end(), n0, l);
  }
  tile_level = t0;
  tile_level.insert(tile_level.end(), t1.begin(), t1.end());
  return (int)t0.size();
}

int launch_response_tiles(const Geometry& g, const Geometry* d_g, const DeviceBuffers& b, const FilterBank& fb, const int* d_tile_level,
                          const int* d_tile_first, int ntiles0, int ntiles, int exact, int trunc_zero, cudaStream_t s) {
  if (ntiles <= 0 || g.n_frames <= 0) return 0;
  const int khm = fb.khm, kwm = fb.kwm;
  if (response_has_fast_path(fb)) {
    if (fb.kh == 5) return launch_fast<5, 5>(d_g, d_tile_level, d_tile_first, g.n_levels, ntiles0, ntiles, g.n_frames, b, fb, exact, trunc_zero, s);
    if (fb.kh == 4) return launch_fast<4, 4>(d_g, d_tile_level, d_tile_first, g.n_levels, ntiles0, ntiles, g.n_frames, b, fb, exact, trunc_zero, s);
    return launch_fast<6, 6>(d_g, d_tile_level, d_tile_first, g.n_levels, ntiles0, ntiles, g.n_frames, b, fb, exact, trunc_zero, s);
  }
  const int HYm = 8 + khm - 1, ROWm = 16 + kwm - 1;
  const size_t smem = (size_t)32 * (HYm * ROWm + 1) * sizeof(float);
  if (exact) cudaFuncSetAttribute(part_response_generic<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  else cudaFuncSetAttribute(part_response_generic<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid(ntiles, g.n_frames);
  if (exact) part_response_generic<true><<<grid, 128, smem, s>>>(d_g, d_tile_level, d_tile_first, b.feat, fb.wg, fb.foff, fb.fkh, fb.fkw, fb.nfilters, khm, kwm, b.resp);
  else part_response_generic<false><<<grid, 128, smem, s>>>(d_g, d_tile_level, d_tile_first, b.feat, fb.wg, fb.foff, fb.fkh, fb.fkw, fb.nfilters, khm, kwm, b.resp);
  return 1;
}

}  // namespace pbd
