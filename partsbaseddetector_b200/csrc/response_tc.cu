// response_tc.cu -- SpatialConvolutionEngine::pdf (reference src/SpatialConvolutionEngine.cpp:70-124) on the 5th-generation
// tensor cores: response mode 2 ("tf32x3").
//
// The stage is the implicit GEMM  resp[cell][f] = sum_{tap, c} F[cell + tap][c] * w_f[tap][c]   (M = cells, N = filters,
// K = kh*kw*32 = 800 for the person model).  Every fp32 operand is split as v = hi + lo with hi = tf32(v) and
// lo = tf32(v - hi) and the three products hi*hi + lo*hi + hi*lo are formed by tcgen05.mma.kind::tf32 with fp32 accumulators
// in tensor memory (the dropped lo*lo term is 2^-22 of a product).  The tensor core's accumulation TRUNCATES (rounds toward
// zero at every MMA: measured -50 ulp on a 300-MMA chain), so the chains are kept short: the hi*hi products are summed on
// the tensor core over `taps_per_partial` taps only (default one filter row) into ping-pong partial accumulators that the
// epilogue warps add up in fp32 round-to-nearest registers, and the small lo terms go to an accumulator of their own
// (2^-11 of the score, so its truncation is invisible).  Scores then agree with the reference to a few ulp (~3e-7).  The
// mode is NOT bit-identical to the reference's separately rounded multiply/add chain -- that is response mode 0
// (response.cu) -- and is held to the north-star tolerance (root scores 1e-4 relative, integer outputs identical) by the
// parity tests.
//
// Layout.  feat_split writes the HOG cells of a level into a padded, flattened strip of 128-byte rows (one row = the 32
// channels of one cell): real cell (y, x) of level l sits at row R_l + y*Wp + x, Wp = ow + ax, so that consecutive image
// rows are separated by ax border cells (value 0, channel 31 = 1: the BORDER_CONSTANT engines of
// src/SpatialConvolutionEngine.cpp:147-156) and ay / kh-1-ay border rows lie above / below.  In this layout the cells a tap
// (ky, kx) needs for 128 CONSECUTIVE output rows are again 128 consecutive rows, shifted by (ky-ay)*Wp + (kx-ax).  The
// rows are stored pre-swizzled (16-byte chunk c of row P at chunk c ^ (P & 7)) so that a plain 1-D TMA bulk copy
// (cp.async.bulk, no tensor map) of an 8-row aligned run lands in shared memory in exactly the K-major SWIZZLE_128B
// layout tcgen05.mma reads, and the kx shift is a descriptor START-ADDRESS offset of kx rows into the same staged strip
// (the hardware swizzle is a function of the absolute shared-memory address: verified by tools/umma_probe.cu).
//
// Kernel (persistent, one CTA per SM, 8 warps; work item = one tile of 128 strip rows): warp 0 streams the A strips (per
// filter row ky: 144 rows x {hi, lo}, double buffered), warp 1 streams the per-tap weight slabs ({hi, lo} x NP filters x
// 128 B, 4 stages), one lane of warp 2 issues the MMAs (M = 128, N = NP, K = 8: 12 per tap), warps 4-7 read finished
// partial accumulators (tcgen05.ld), sum them and store the planar response maps while the next partial / tile is being
// multiplied.  All hand-offs are mbarriers (TMA complete_tx / tcgen05.commit / epilogue arrivals).
#include "kernels.cuh"

#include <cuda_fp16.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>

namespace pbd {
namespace {

constexpr int TM = 128;                       // output rows (cells) per tile = UMMA M
constexpr int STRIP_ROWS = 144;               // 128 + 7 (8-row alignment of the copy) + kw-1 <= 9
constexpr int STRIP_BYTES = STRIP_ROWS * 128;
constexpr int A_BUFS = 2, B_STAGES = 4;
constexpr int A_BUFS_F16 = 4;                 // fp16 flavour: a strip is 18 KB, so four filter rows fit (a strip load takes longer than one
                                              // row's MMAs: with two buffers the issuing warp waits for every strip)
constexpr int B_STAGES_F16 = 8;               // fp16 flavour: one 128-byte-row slab per tap instead of two => twice the taps in flight
constexpr int NP_MAX = 160;                   // 3 accumulators of NP columns must fit the 512 TMEM columns
constexpr int NTHREADS = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint32_t bar, unsigned parity) {
  unsigned done;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done != 0;
}
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, unsigned parity) {
  for (long long spins = 0; !mbar_try(bar, parity); ++spins)
    if (spins > (1ll << 26)) asm volatile("trap;\n");      // a broken hand-off must fail loudly, not hang the device
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, unsigned parity) {
  if (!mbar_try(bar, parity)) mbar_wait_slow(bar, parity);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, unsigned bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// K-major SWIZZLE_128B operand descriptor: 128-byte rows, 8-row groups 1024 B apart.  Low word = (start address >> 4) | LBO
// (1 << 16, unused for swizzled K-major), high word = SBO (1024 >> 4) | version 1 (bit 46) | SWIZZLE_128B (2 << 61) = constant.
// The start address may be ANY row of a staged strip (absolute-address swizzle, tools/umma_probe.cu).
constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint32_t da_lo, uint32_t db_lo, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %4, 0;\nmov.b64 da, {%1, %5};\nmov.b64 db, {%2, %5};\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %3, p;\n}\n" ::"r"(tmem_d), "r"(da_lo), "r"(db_lo), "r"(idesc), "r"(acc), "r"(kDescHi)
      : "memory");
}
// the same issue for fp16 operands (K = 16 per instruction, again 32 bytes of a 128-byte row)
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint32_t da_lo, uint32_t db_lo, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %4, 0;\nmov.b64 da, {%1, %5};\nmov.b64 db, {%2, %5};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n}\n" ::"r"(tmem_d), "r"(da_lo), "r"(db_lo), "r"(idesc), "r"(acc), "r"(kDescHi)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n.reg .b32 rx;\n.reg .pred px;\nelect.sync rx|px, 0xFFFFFFFF;\n@px mov.s32 %0, 1;\n}\n" : "+r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

__device__ __forceinline__ float tf32_rna(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

// ---- HOG cells [n][cells][32] (HWC) -> padded, pre-swizzled hi / lo strips ----
__global__ void __launch_bounds__(256)
feat_split(const Geometry* __restrict__ g, const TcLevel* __restrict__ lv, const float* __restrict__ feat, float* __restrict__ fhi,
           float* __restrict__ flo, long long frame_rows) {
  const int frame = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;        // (cell of the frame, 16-byte chunk)
  const int gcell = i >> 3, c = i & 7;
  if (gcell >= g->cells_total) return;
  int l = 0;
  while (l + 1 < g->n_levels && gcell >= g->lv[l + 1].cell_off) ++l;
  const LevelDesc& L = g->lv[l];
  const int cell = gcell - L.cell_off;
  const int y = cell / L.ow, x = cell - y * L.ow;
  const float4 v = __ldg(reinterpret_cast<const float4*>(feat + ((size_t)frame * g->cells_total + gcell) * 32) + c);
  const long long P = (long long)frame * frame_rows + lv[l].R + (long long)y * lv[l].Wp + x;
  float4 h, o;
  h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
  o.x = tf32_rna(v.x - h.x); o.y = tf32_rna(v.y - h.y); o.z = tf32_rna(v.z - h.z); o.w = tf32_rna(v.w - h.w);
  const size_t off = (size_t)P * 32 + ((c ^ (int)(P & 7)) << 2);
  *reinterpret_cast<float4*>(fhi + off) = h;
  *reinterpret_cast<float4*>(flo + off) = o;
}

// every row that is not a real cell is a border cell: channels 0..30 = 0, channel 31 (chunk 7, word 3) = 1
__global__ void __launch_bounds__(256) tc_border_init(float* __restrict__ fhi, long long rows) {
  const long long P = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (P >= rows) return;
  fhi[(size_t)P * 32 + ((7 ^ (int)(P & 7)) << 2) + 3] = 1.f;
}

// ---- fp16 flavour (response mode 3): one 128-byte strip row per cell = [hi(32 ch) | lo(32 ch)] in fp16 with
// F' = F * 2^kFeatExp, hi = fp16(F'), lo = fp16(F' - hi): the same 11 + 11 significand bits as the tf32 split.  fp16 has only 5
// exponent bits, hence the power-of-two pre-scaling (exact, undone in the epilogue): with F' up to 2^12 the residual lo ~ 2^-12 F'
// stays a normal fp16 number for every feature above 6e-5, and below that its absolute error (2^-25) is < 2^-23 of F' anyway ----
constexpr int kFeatExp = 12;                  // |F| < 16 stays finite in fp16; HOG features are <= 1
__global__ void __launch_bounds__(256)
feat_split_f16(const Geometry* __restrict__ g, const TcLevel* __restrict__ lv, const float* __restrict__ feat, uint4* __restrict__ strips,
               long long frame_rows) {
  const int frame = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;        // (cell of the frame, group of 8 channels)
  const int gcell = i >> 2, c = i & 3;
  if (gcell >= g->cells_total) return;
  int l = 0;
  while (l + 1 < g->n_levels && gcell >= g->lv[l + 1].cell_off) ++l;
  const LevelDesc& L = g->lv[l];
  const int cell = gcell - L.cell_off;
  const int y = cell / L.ow, x = cell - y * L.ow;
  const float4* src = reinterpret_cast<const float4*>(feat + ((size_t)frame * g->cells_total + gcell) * 32) + 2 * c;
  const float4 v0 = __ldg(src), v1 = __ldg(src + 1);
  const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
  unsigned short h[8], o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float s = __fmul_rn(v[j], (float)(1 << kFeatExp));
    const __half hh = __float2half_rn(s);
    h[j] = __half_as_ushort(hh);
    o[j] = __half_as_ushort(__float2half_rn(__fsub_rn(s, __half2float(hh))));
  }
  const long long P = (long long)frame * frame_rows + lv[l].R + (long long)y * lv[l].Wp + x;
  const int sw = (int)(P & 7);
  uint4 a, b;
  a.x = h[0] | ((unsigned)h[1] << 16); a.y = h[2] | ((unsigned)h[3] << 16); a.z = h[4] | ((unsigned)h[5] << 16); a.w = h[6] | ((unsigned)h[7] << 16);
  b.x = o[0] | ((unsigned)o[1] << 16); b.y = o[2] | ((unsigned)o[3] << 16); b.z = o[4] | ((unsigned)o[5] << 16); b.w = o[6] | ((unsigned)o[7] << 16);
  strips[(size_t)P * 8 + (c ^ sw)] = a;                       // chunks 0..3: hi channels 8c..8c+7
  strips[(size_t)P * 8 + ((4 + c) ^ sw)] = b;                 // chunks 4..7: lo
}
// border cells: channels 0..30 = 0, channel 31 = 1 (hi = fp16(2^kFeatExp) in chunk 3, element 7; lo = 0)
__global__ void __launch_bounds__(256) tc_border_init_f16(unsigned short* __restrict__ strips, long long rows) {
  const long long P = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (P >= rows) return;
  strips[(size_t)P * 64 + ((3 ^ (int)(P & 7)) << 3) + 7] = __half_as_ushort(__float2half_rn((float)(1 << kFeatExp)));
}

struct TcParams {
  const float* fhi;
  const float* flo;
  const float* wpk;            // [taps][2][NP][32] pre-swizzled weight slabs
  float* resp;
  const TcLevel* levels;
  const TcTile* tiles;
  int n_tiles, n_frames;
  long long frame_rows;
  int cells_total, nfilters, NP, kh, kw;
  float out_scale[NP_MAX];     // fp16 flavour: per filter 2^-(kFeatExp + weight exponent of the filter); in the parameter (constant) bank so
                               // that the unrolled epilogue reads it as an instruction operand
};

// F16 = false: tf32 operands, fhi / flo strips, three split products per tap (4 + 8 MMAs).
// F16 = true:  fp16 operands, one strip of [F_hi | F_lo] rows and ONE weight slab of [W_hi | W_lo] rows per tap; the three products
//              F_hi W_hi, F_lo W_hi, F_hi W_lo are K-halves of those rows (2 MMAs of K = 16 each): 6 MMAs per tap instead of 12 and half
//              the operand bytes through L2 and shared memory.
template <bool PER_TAP, bool F16>
__global__ void __launch_bounds__(NTHREADS, 1) part_response_tc(const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int NP = p.NP;
  const uint32_t slab_bytes = (uint32_t)NP * 128u;             // one part (hi or lo) of one tap; fp16: the whole tap
  constexpr uint32_t AP = F16 ? 1u : 2u;                       // parts per A buffer / B stage
  constexpr int BS = F16 ? B_STAGES_F16 : B_STAGES;
  constexpr int AB = F16 ? A_BUFS_F16 : A_BUFS;
  // A: [AB][AP parts][STRIP_BYTES], B: [BS][AP parts][slab_bytes]
  const uint32_t sA = sbase, sB = sbase + (uint32_t)AB * AP * STRIP_BYTES;
  __shared__ __align__(8) unsigned long long bars[2 * A_BUFS_F16 + 2 * B_STAGES_F16 + 6];
  __shared__ uint32_t tmem_base_s;
  const uint32_t fullA = smem_u32(&bars[0]), emptyA = fullA + 8 * AB, fullB = emptyA + 8 * AB, emptyB = fullB + 8 * BS;
  const uint32_t hFull = emptyB + 8 * BS, hEmpty = hFull + 16, cFull = hEmpty + 16, cEmpty = cFull + 8;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    constexpr unsigned releasers = 2;                         // both MMA warps release
    for (int i = 0; i < AB; ++i) { mbar_init(fullA + 8 * i, 1); mbar_init(emptyA + 8 * i, releasers); }
    for (int i = 0; i < BS; ++i) { mbar_init(fullB + 8 * i, 1); mbar_init(emptyB + 8 * i, releasers); }
    mbar_init(hFull, 1); mbar_init(hFull + 8, 1); mbar_init(hEmpty, 4); mbar_init(hEmpty + 8, 4);
    mbar_init(cFull, 1); mbar_init(cEmpty, 4);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  const int kh = p.kh, kw = p.kw, ay = kh / 2, ax = kw / 2;   // anchor = centre, include/filterengine.hpp:310-318
  const int taps = kh * kw, G = PER_TAP ? 1 : kw;
  const int total = p.n_tiles * p.n_frames;

  if (warp == 0) {
    // ===== A producer: per tile and filter row ky, one 144-row strip per part =====
    if (lane == 0) {
      int it = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int frame = w / p.n_tiles;
        const TcTile S = p.tiles[w - frame * p.n_tiles];
        const int Wp = p.levels[S.level].Wp;
        for (int ky = 0; ky < kh; ++ky, ++it) {
          const int buf = it % AB;
          mbar_wait(emptyA + 8 * buf, ((it / AB) & 1) ^ 1);
          mbar_expect_tx(fullA + 8 * buf, F16 ? (uint32_t)STRIP_BYTES : 2u * STRIP_BYTES);
          const long long pstart = (long long)S.q0 + (long long)(ky - ay) * Wp - ax;
          const long long pa = (long long)frame * p.frame_rows + (pstart & ~7ll);
          const uint32_t dst = sA + (uint32_t)buf * AP * STRIP_BYTES;
          bulk_g2s(dst, p.fhi + (size_t)pa * 32, STRIP_BYTES, fullA + 8 * buf);        // rows are 128 bytes in both flavours
          if (!F16) bulk_g2s(dst + STRIP_BYTES, p.flo + (size_t)pa * 32, STRIP_BYTES, fullA + 8 * buf);
        }
      }
    }
  } else if (warp == 1) {
    // ===== B producer: one {hi, lo} weight slab per tap =====
    if (lane == 0) {
      int it = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x) {
        for (int tap = 0; tap < taps; ++tap, ++it) {
          const int st = it % BS;
          mbar_wait(emptyB + 8 * st, ((it / BS) & 1) ^ 1);
          mbar_expect_tx(fullB + 8 * st, AP * slab_bytes);
          bulk_g2s(sB + (uint32_t)st * AP * slab_bytes, p.wpk + (size_t)tap * AP * NP * 32, AP * slab_bytes, fullB + 8 * st);
        }
      }
    }
  } else if (warp == 2 || warp == 3) {
    // ===== MMA issuers.  Two warps, because a single warp's instruction latency (not the tensor pipe) bounds the kernel when
    // one warp issues all 12 MMAs of a tap: warp 2 issues the hi*hi products (4 per tap) into the ping-pong partial
    // accumulators [0,NP) / [NP,2NP) -- chains of one tap (PER_TAP) or one filter row -- and warp 3 the lo*hi + hi*lo
    // corrections (8 per tap) into [2NP,3NP) for the whole tile; the two chains are independent, so the warps need no
    // ordering between them.  Each warp runs its loop convergently (operands stay in uniform registers), one elected lane
    // issues.  Short hi*hi chains keep the tensor core's truncating accumulation below one ulp of the final score; the
    // epilogue adds the partials in fp32 RN. =====
    const bool isH = warp == 2;
    // instruction descriptor: D = f32 (bit 4), A / B format (bits 7.. / 10..: 2 = tf32, 0 = f16), N >> 3, M >> 4
    const uint32_t idesc = (1u << 4) | (F16 ? 0u : ((2u << 7) | (2u << 10))) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
    const uint32_t dC = tmem + (uint32_t)(2 * NP);
    const uint32_t a_lo_off = STRIP_BYTES >> 4, b_lo_off = slab_bytes >> 4;
    uint32_t bufA = 0, phA = 0, stB = 0, phB = 0, hs = 0, phH = 1, phC = 1;       // ring positions and wait parities
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
      const int frame = w / p.n_tiles;
      const TcTile S = p.tiles[w - frame * p.n_tiles];
      const int Wp = p.levels[S.level].Wp;
      int prow = S.q0 - ay * Wp - ax;                          // first strip row of filter row ky (>= 0 by construction)
      if (!isH) { mbar_wait(cEmpty, phC); phC ^= 1; }          // the correction accumulator of the previous tile has been read
      for (int ky = 0; ky < kh; ++ky, prow += Wp) {
        mbar_wait(fullA + 8 * bufA, phA);
        // descriptor low words = (start address >> 4) | LBO; one strip row = 128 B adds 8, a k-step of 8 tf32 = 32 B adds 2
        uint32_t ahi = ((sA + bufA * AP * STRIP_BYTES + (uint32_t)(prow & 7) * 128u) >> 4) | 0x10000u;
        for (int kx = 0; kx < kw; ++kx, ahi += 8) {
          mbar_wait(fullB + 8 * stB, phB);
          const uint32_t bhi = ((sB + stB * AP * slab_bytes) >> 4) | 0x10000u;
          if (isH) {
            const bool hfirst = PER_TAP || kx == 0, hlast = PER_TAP || kx == kw - 1;
            if (hfirst) mbar_wait(hEmpty + 8 * hs, phH);
            tc_fence_after();
            const uint32_t dH = tmem + hs * (uint32_t)NP;
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (!F16) umma_tf32(dH, ahi + 2 * k, bhi + 2 * k, idesc, (hfirst && k == 0) ? 0u : 1u);
              }
              if (F16) {
                // k-steps 0, 1 = hi half, 2, 3 = lo half of the 128-byte rows.  F_hi W_hi goes to the ping-pong partial accumulator
                // (2 MMAs per tap); warp 3 issues the two small products (2^-11 of it) into the correction accumulator of the tile,
                // where the tensor core's truncating accumulation costs nothing
#pragma unroll
                for (int k = 0; k < 2; ++k) umma_f16(dH, ahi + 2 * k, bhi + 2 * k, idesc, (hfirst && k == 0) ? 0u : 1u);
              }
              umma_commit(emptyB + 8 * stB);                 // weight slab consumed once these MMAs retire (and warp 3's)
              if (hlast) umma_commit(hFull + 8 * hs);
              if (kx == kw - 1) umma_commit(emptyA + 8 * bufA);
            }
            __syncwarp();
            if (hlast) { phH ^= hs; hs ^= 1; }               // parity flips every second partial
          } else {
            tc_fence_after();
            const bool first = (ky | kx) == 0;
            if (elect_one()) {
              if (F16) {
#pragma unroll
                for (int k = 0; k < 2; ++k) umma_f16(dC, ahi + 4 + 2 * k, bhi + 2 * k, idesc, (first && k == 0) ? 0u : 1u);      // F_lo W_hi
#pragma unroll
                for (int k = 0; k < 2; ++k) umma_f16(dC, ahi + 2 * k, bhi + 4 + 2 * k, idesc, 1u);                               // F_hi W_lo
              } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_tf32(dC, ahi + a_lo_off + 2 * k, bhi + 2 * k, idesc, (first && k == 0) ? 0u : 1u);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_tf32(dC, ahi + 2 * k, bhi + b_lo_off + 2 * k, idesc, 1u);
              }
              umma_commit(emptyB + 8 * stB);
              if (kx == kw - 1) umma_commit(emptyA + 8 * bufA);
              if (ky == kh - 1 && kx == kw - 1) umma_commit(cFull);
            }
            __syncwarp();
          }
          if (++stB == BS) { stB = 0; phB ^= 1; }
        }
        if (++bufA == AB) { bufA = 0; phA ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: TMEM lanes 32*(warp%4).. ; sums the partial accumulators in registers, then -> resp[frame][f][cell] =====
    const int q = warp & 3;
    const int nparts = (taps + G - 1) / G;
    int hcount = 0, tcount = 0;
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    for (int w = blockIdx.x; w < total; w += gridDim.x, ++tcount) {
      const int frame = w / p.n_tiles;
      const TcTile S = p.tiles[w - frame * p.n_tiles];
      const TcLevel L = p.levels[S.level];
      float acc[NP_MAX];
#pragma unroll
      for (int j = 0; j < NP_MAX; ++j) acc[j] = 0.f;
      // the partial sums and the correction accumulator.  The last partial and the correction accumulator complete together; the
      // correction accumulator is drained first because the next tile's first tap already needs it (the partial's slot only two
      // filter rows later)
      for (int step = 0; step <= nparts; ++step) {
        const int part = step == nparts - 1 ? nparts : (step == nparts ? nparts - 1 : step);
        uint32_t taddr;
        int hs = 0;
        if (part < nparts) {
          hs = hcount & 1;
          mbar_wait(hFull + 8 * hs, (hcount >> 1) & 1);
          taddr = lane_base + (uint32_t)(hs * NP);
        } else {
          mbar_wait(cFull, tcount & 1);
          taddr = lane_base + (uint32_t)(2 * NP);
        }
        tc_fence_after();
        // three 16-column loads in flight per wait (one TMEM round trip per 48 columns instead of per 16)
#pragma unroll
        for (int c0 = 0; c0 < NP_MAX; c0 += 48) {
          if (c0 < NP) {
            uint32_t v[48];
#pragma unroll
            for (int b = 0; b < 3; ++b) {
              if (c0 + 16 * b < NP) {
                uint32_t* w = v + 16 * b;
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                             : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]), "=r"(w[8]),
                               "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15])
                             : "r"(taddr + c0 + 16 * b));
              }
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
            for (int j = 0; j < 48; ++j)
              if (c0 + j < NP_MAX) acc[c0 + j] = __fadd_rn(acc[c0 + j], __uint_as_float(v[j]));
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(part < nparts ? hEmpty + 8 * hs : cEmpty);
        if (part < nparts) ++hcount;
      }
      const int r = S.q0 - L.R + q * 32 + lane;             // row relative to real cell (0, 0)
      const int y = r / L.Wp, x = r - y * L.Wp;
      if (y < L.oh && x < L.ow) {
        // one running pointer (a 64-bit add per plane) and constant-bank scales: the stores issue back to back.  (A first version
        // recomputed the address and fetched the scale from shared memory per plane: 80 dependent cycles per store, 1 ms per 64 frames.)
        float* dst = p.resp + (size_t)frame * p.nfilters * p.cells_total + L.cell_off + (size_t)y * L.ow + x;
        const size_t plane = (size_t)p.cells_total;
#pragma unroll
        for (int j = 0; j < NP_MAX; ++j)
          if (j < p.nfilters) { *dst = F16 ? __fmul_rn(acc[j], p.out_scale[j]) : acc[j]; dst += plane; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem));
}

inline int round_up(int v, int a) { return (v + a - 1) / a * a; }

}  // namespace

bool response_tc_supported(const FilterBank& fb) {
  return fb.uniform && fb.flen == 32 && fb.kw >= 1 && fb.kw <= 10 && fb.kh >= 1 && round_up(fb.nfilters, 16) <= NP_MAX;
}
int response_tc_np(int nfilters) { return round_up(nfilters, 16); }

// round to tf32 (10 explicit mantissa bits), ties away from zero like cvt.rna.tf32.f32
static float host_tf32(float v) {
  uint32_t u;
  memcpy(&u, &v, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return v;
  u = (u + 0x1000u) & ~0x1FFFu;
  memcpy(&v, &u, 4);
  return v;
}

// filters[f] = [tap][c] (HWC taps) -> [tap][part][NP][32] slabs, 16-byte chunk c of filter row n stored at chunk c ^ (n & 7)
void response_tc_pack_weights(const std::vector<std::vector<float>>& filters, int taps, std::vector<float>& out) {
  const int nf = (int)filters.size(), NP = response_tc_np(nf);
  out.assign((size_t)taps * 2 * NP * 32, 0.f);
  for (int t = 0; t < taps; ++t)
    for (int n = 0; n < nf; ++n)
      for (int c = 0; c < 32; ++c) {
        const float v = filters[n][(size_t)t * 32 + c];
        const float hi = host_tf32(v), lo = host_tf32(v - hi);
        const size_t col = (size_t)(((c >> 2) ^ (n & 7)) << 2) + (c & 3);
        out[(((size_t)t * 2 + 0) * NP + n) * 32 + col] = hi;
        out[(((size_t)t * 2 + 1) * NP + n) * 32 + col] = lo;
      }
}

// fp16 flavour: [tap][NP][64 halves], row n = [W_hi | W_lo] with W' = W * 2^wexp[n], W_hi = fp16(W'), W_lo = fp16(W' - W_hi) and a PER-FILTER
// exponent chosen so that the filter's largest |W'| lies in [2^13, 2^14) (shipped models contain filters that are ~1e-19 throughout: a
// common scale would flush them to zero; W_lo ~ 2^-12 W' stays a normal fp16 number for all but the smallest weights of a filter, whose
// absolute error 2^-25 is then < 2^-38 of the filter's scale); same chunk swizzle (16-byte chunks of 8 halves).
// out_scales[n] = 2^-(kFeatExp + wexp[n]).
void response_tc_pack_weights_f16(const std::vector<std::vector<float>>& filters, int taps, std::vector<uint16_t>& out, std::vector<float>& out_scales) {
  const int nf = (int)filters.size(), NP = response_tc_np(nf);
  out.assign((size_t)taps * NP * 64, 0);
  out_scales.assign(NP, 0.f);
  for (int n = 0; n < nf; ++n) {
    float wmax = 0.f;
    for (float v : filters[n]) if (std::isfinite(v)) wmax = std::max(wmax, std::fabs(v));
    int wexp = 0;
    if (wmax > 0.f) { int e; std::frexp(wmax, &e); wexp = 14 - e; }        // wmax = m * 2^e, m in [0.5, 1)  =>  wmax * 2^wexp in [2^13, 2^14)
    wexp = std::max(-100, std::min(100, wexp));
    out_scales[n] = std::ldexp(1.f, -(kFeatExp + wexp));
    const float ws = std::ldexp(1.f, wexp);
    for (int t = 0; t < taps; ++t)
      for (int c = 0; c < 32; ++c) {
        const float v = filters[n][(size_t)t * 32 + c] * ws;
        const __half hi = __float2half_rn(v);
        const size_t row = ((size_t)t * NP + n) * 64;
        auto col = [&](int ch) { return (size_t)((((ch >> 3)) ^ (n & 7)) << 3) + (ch & 7); };   // ch = position in the 64-half row
        out[row + col(c)] = __half_as_ushort(hi);
        out[row + col(32 + c)] = __half_as_ushort(__float2half_rn(v - __half2float(hi)));
      }
  }
}

// padded strip layout of one frame + the work list of one frame; returns rows per frame (multiple of 8)
long long response_tc_plan(const Geometry& g, int kh, int kw, std::vector<TcLevel>& levels, std::vector<TcTile>& tiles, long long* slack_rows) {
  const int ay = kh / 2, ax = kw / 2;
  levels.assign(g.n_levels, TcLevel{});
  tiles.clear();
  long long base = 0;
  int wp_max = 0;
  for (int l = 0; l < g.n_levels; ++l) {
    const LevelDesc& L = g.lv[l];
    TcLevel T{};
    T.Wp = L.ow + ax; T.ow = L.ow; T.oh = L.oh; T.cell_off = L.cell_off;
    T.R = (int)base + ay * T.Wp + ax;
    levels[l] = T;
    wp_max = std::max(wp_max, T.Wp);
    const int n_out = (L.oh - 1) * T.Wp + L.ow;                   // first to last real cell
    const int ntiles = (n_out + TM - 1) / TM;
    for (int k = 0; k < ntiles; ++k) tiles.push_back(TcTile{l, T.R + k * TM});
    base += round_up((L.oh + kh - 1) * T.Wp + ax, 8);
  }
  *slack_rows = (long long)TM + (long long)kh * wp_max + 32;   // phantom rows of a level's last tile read past its block
  return base;
}

int launch_feat_split(const Geometry& g, const Geometry* d_g, const TcLevel* d_levels, const float* feat, float* fhi, float* flo, long long frame_rows,
                      cudaStream_t s) {
  if (g.cells_total <= 0) return 0;
  dim3 grid((g.cells_total * 8 + 255) / 256, g.n_frames);
  feat_split<<<grid, 256, 0, s>>>(d_g, d_levels, feat, fhi, flo, frame_rows);
  return 1;
}

int launch_tc_border_init(float* fhi, float* flo, long long rows, cudaStream_t s) {
  cudaMemsetAsync(fhi, 0, (size_t)rows * 128, s);
  cudaMemsetAsync(flo, 0, (size_t)rows * 128, s);
  tc_border_init<<<(unsigned)((rows + 255) / 256), 256, 0, s>>>(fhi, rows);
  return 1;
}

int launch_feat_split_f16(const Geometry& g, const Geometry* d_g, const TcLevel* d_levels, const float* feat, uint16_t* strips, long long frame_rows,
                          cudaStream_t s) {
  if (g.cells_total <= 0) return 0;
  dim3 grid((g.cells_total * 4 + 255) / 256, g.n_frames);
  feat_split_f16<<<grid, 256, 0, s>>>(d_g, d_levels, feat, reinterpret_cast<uint4*>(strips), frame_rows);
  return 1;
}

int launch_tc_border_init_f16(uint16_t* strips, long long rows, cudaStream_t s) {
  cudaMemsetAsync(strips, 0, (size_t)rows * 128, s);
  tc_border_init_f16<<<(unsigned)((rows + 255) / 256), 256, 0, s>>>(strips, rows);
  return 1;
}

// f16_scales != nullptr (HOST array of NP floats) selects the fp16 flavour: fhi = the [hi | lo] strips, wpk = the fp16 slabs
int launch_response_tc(const Geometry& g, const DeviceBuffers& b, const FilterBank& fb, const float* fhi, const float* flo, const float* wpk,
                       const TcLevel* d_levels, const TcTile* d_tiles, int n_tiles, long long frame_rows, int num_sms, int taps_per_partial,
                       cudaStream_t s, const float* f16_scales) {
  if (n_tiles <= 0 || g.n_frames <= 0) return 0;
  TcParams p{};
  p.fhi = fhi; p.flo = flo; p.wpk = wpk; p.resp = b.resp; p.levels = d_levels; p.tiles = d_tiles;
  p.n_tiles = n_tiles; p.n_frames = g.n_frames; p.frame_rows = frame_rows; p.cells_total = g.cells_total;
  p.nfilters = fb.nfilters; p.NP = response_tc_np(fb.nfilters); p.kh = fb.kh; p.kw = fb.kw;
  const bool per_tap = taps_per_partial == 1;      // hi*hi chains of one tap (most accurate) or of one filter row (default)
  const bool f16 = f16_scales != nullptr;
  const size_t smem = f16 ? 1024 + (size_t)A_BUFS_F16 * STRIP_BYTES + (size_t)B_STAGES_F16 * p.NP * 128
                          : 1024 + (size_t)A_BUFS * 2 * STRIP_BYTES + (size_t)B_STAGES * 2 * p.NP * 128;
  if (f16) for (int j = 0; j < p.NP; ++j) p.out_scale[j] = f16_scales[j];
  const long long total = (long long)n_tiles * g.n_frames;
  const int grid = (int)std::min<long long>(total, num_sms);
  auto go = [&](auto kernel) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kernel<<<grid, NTHREADS, smem, s>>>(p);
  };
  if (f16) { if (per_tap) go(part_response_tc<true, true>); else go(part_response_tc<false, true>); }
  else { if (per_tap) go(part_response_tc<true, false>); else go(part_response_tc<false, false>); }
  return 1;
}

}  // namespace pbd
