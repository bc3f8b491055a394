// hog.cu -- HOGFeatures<float>::features<uint8_t> (reference src/HOGFeatures.cpp:168-341) as two kernels.
//
// hog_hist   CTA = 16x8 HOG blocks.  Phase 1 computes gradient magnitude and snapped orientation of every pixel of the
//            tile's footprint exactly once into shared memory.  Phase 2: one thread per block (by,bx) gathers the
//            (2*sbin)^2 pixel window that scatters into this block in the reference (bilinear scatter, :252-265),
//            visiting pixels in raster order and adding the term into its shared-memory histogram with a separately
//            rounded multiply and add.  Because every bin receives its terms in the same order as the reference's
//            sequential scatter, the float sums are bit-identical (no atomics, deterministic).  Also emits the
//            block energy (:270-283).
// hog_feat   one thread per output cell: 4 normalisers with double sqrt/divide (:292-299), 18 contrast-
//            sensitive + 9 insensitive + 4 texture + 1 truncation features (:304-338), written HWC.
// Both are HBM/L2-bound streaming kernels (48 B of image read and 128 B written per cell).
#include "kernels.cuh"

namespace pbd {
namespace {

__device__ __forceinline__ int find_level_by_cell(const Geometry* g, int idx) {
  int l = 0;
  while (l + 1 < g->n_levels && idx >= g->lv[l + 1].cell_off) ++l;
  return l;
}

// (T)(((T)p + 0.5) / (T)sbin - 0.5), reference :252-253 (double intermediates).  For power-of-two bin sizes the
// division is an exact scaling, so it is replaced by the (equally exact) multiplication with 1/sbin.
__device__ __forceinline__ float bin_coord(int p, int sbin, double dsb, double inv_sb, bool pow2) {
  const double t = (double)(float)p + 0.5;
  return (float)((pow2 ? __dmul_rn(t, inv_sb) : __ddiv_rn(t, dsb)) - 0.5);
}

constexpr int HB_X = 16, HB_Y = 8;                 // HOG blocks per CTA (one thread per block)

// Orientation snap (reference :243-249): best of 9 directions and their opposites by the separately rounded dot product
// uu[o]*dx + vv[o]*dy.  dx, dy are differences of 8-bit samples, so there are only 511 x 511 distinct inputs: the snapped
// orientation is tabulated once per detector BY THIS VERY CODE (bit-identical decisions, ties included) and hog_hist looks it up
// instead of evaluating 9 dot products per pixel.
__device__ __forceinline__ int snap_orientation(float dx, float dy) {
  const float uu[9] = {(float)1.000, (float)0.9397, (float)0.7660, (float)0.5000, (float)0.1736,
                       (float)-0.1736, (float)-0.5000, (float)-0.7660, (float)-0.9397};
  const float vv[9] = {(float)0.000, (float)0.3420, (float)0.6428, (float)0.8660, (float)0.9848,
                       (float)0.9848, (float)0.8660, (float)0.6428, (float)0.3420};
  float best_dot = 0.f;
  int best_o = 0;
#pragma unroll
  for (int o = 0; o < 9; ++o) {
    const float dot = __fadd_rn(__fmul_rn(uu[o], dx), __fmul_rn(vv[o], dy));
    if (dot > best_dot) { best_dot = dot; best_o = o; }
    else if (-dot > best_dot) { best_dot = -dot; best_o = o + 9; }
  }
  return best_o;
}
constexpr int kGradSpan = 511;                     // dx, dy in [-255, 255]
__global__ void __launch_bounds__(256) hog_orient_lut(unsigned char* __restrict__ lut) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kGradSpan * kGradSpan) return;
  lut[i] = (unsigned char)snap_orientation((float)(i % kGradSpan - 255), (float)(i / kGradSpan - 255));
}

// SB = the bin size for the specialised gather (even bin sizes 4 and 8: every shipped model), 0 = any bin size.
// For an even bin size 2h the pixels that scatter into block b are exactly the 2 sbin pixels [sbin b - h, sbin b + 3h - 1]
// (floor((x + 0.5) / sbin - 0.5) = floor((x - h) / sbin): (2x + 1 - 2h) / (2 sbin) is never an integer, so no rounding of the
// reference's expression can move a pixel across a bin boundary), the first sbin of them with weight vx0, the others with vx1.  The
// gather of a block is then a fixed 2 sbin x 2 sbin loop without tests: pixels the reference does not visit (outside [1, visible - 2])
// are staged with magnitude +0, and adding +0 to a bin (never -0: bins start at +0 and receive non-negative terms) changes nothing.
template <int CN, int SB>
__global__ void __launch_bounds__(HB_X * HB_Y)
hog_hist(const Geometry* __restrict__ g, const uint8_t* __restrict__ frames, const uint8_t* __restrict__ pyr, const unsigned char* __restrict__ orient_lut,
         float* __restrict__ hist, float* __restrict__ norm, int sbin, int frame0) {
  extern __shared__ __align__(16) unsigned char hsm[];
  // ---- which tile of which level ----
  int tile = blockIdx.x, l = 0, tiles_x = 0;
  for (; l < g->n_levels; ++l) {
    tiles_x = (g->lv[l].bw + HB_X - 1) / HB_X;
    const int nt = tiles_x * ((g->lv[l].bh + HB_Y - 1) / HB_Y);
    if (tile < nt) break;
    tile -= nt;
  }
  if (l >= g->n_levels) return;
  const int frame = frame0 + blockIdx.y;
  const LevelDesc& L = g->lv[l];
  const int bx0 = (tile % tiles_x) * HB_X, by0 = (tile / tiles_x) * HB_Y;
  const int cols = L.img_w, rows = L.img_h;
  const int vis_w = L.bw * sbin, vis_h = L.bh * sbin;
  const uint8_t* im = L.identity ? frames + (size_t)frame * g->in_h * g->in_w * CN : pyr + (size_t)frame * g->img_bytes + L.img_off;
  const size_t stride = (size_t)cols * CN;
  // pixel region that can scatter into this tile's blocks: floor((p+0.5)/sbin - 0.5) in {b-1, b}
  // SB: the region is exactly the union of the blocks' windows, rows 16-byte aligned per block column (PW = 17 sbin, a multiple of 4)
  const int marg = SB ? SB / 2 : (sbin + 1) / 2 + 1;
  const int px0 = bx0 * sbin - marg, py0 = by0 * sbin - marg;
  const int PW = SB ? (HB_X + 1) * SB : HB_X * sbin + sbin + 2 * marg, PH = SB ? (HB_Y + 1) * SB : HB_Y * sbin + sbin + 2 * marg;
  float* smag = reinterpret_cast<float*>(hsm);                       // [PH][PW] gradient magnitude (sqrt(v)); < 0 = pixel not visited
  float* shist = smag + PH * PW;                                     // [18][HB_X*HB_Y] histogram of each thread's block
  float* sfx = shist + 18 * HB_X * HB_Y;                             // [PW] fractional bin coordinate vx0 of every column of the region
  float* sfy = sfx + PW;                                             // [PH] ... vy0 of every row
  float* sgx = sfy + PH;                                             // [PW] vx1 = (float)(1.0 - vx0), :258-259
  float* sgy = sgx + PW;                                             // [PH] vy1
  int* sbx = reinterpret_cast<int*>(sgy + PH);                       // [PW] floor(xp) of every column, [PH] floor(yp) of every row
  int* sby = sbx + PW;
  unsigned char* sbo = reinterpret_cast<unsigned char*>(sby + PH);   // [PH][PW] snapped orientation

  // ---- phase 1: per pixel gradient, channel pick, orientation snap (each pixel of the region exactly once) ----
  int rx = threadIdx.x % PW, ry = threadIdx.x / PW;                  // position inside the region, advanced without divisions
  const int stepx = (HB_X * HB_Y) % PW, stepy = (HB_X * HB_Y) / PW;
  // The loop is bound by the latency of its byte loads (twelve per colour pixel), and the slow-path branch inside the correctly rounded
  // square root keeps the compiler from overlapping iterations: so the loads of KP pixels are issued first, then the arithmetic.
  // Pixels the reference does not visit (:202-203) read a clamped, valid address and are masked afterwards.
#ifndef PBD_HOG_KP
#define PBD_HOG_KP 4
#endif
  constexpr int KP = PBD_HOG_KP, NB = CN == 1 ? 4 : 12;
  for (int i0 = threadIdx.x; i0 < PW * PH; i0 += KP * HB_X * HB_Y) {
    int raw[KP][NB];
    bool visited[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      const int x = px0 + rx, y = py0 + ry;
      rx += stepx; ry += stepy;
      if (rx >= PW) { rx -= PW; ++ry; }
      visited[k] = x >= 1 && x <= vis_w - 2 && y >= 1 && y <= vis_h - 2 && i0 + k * (HB_X * HB_Y) < PW * PH;
      const int sx = min(max(x, 1), cols - 2), sy = min(max(y, 1), rows - 2);   // = min(x, cols - 2), min(y, rows - 2) for a visited pixel
      const uint8_t* s = im + (size_t)sy * stride + sx * CN;
      if (CN == 1) {
        raw[k][0] = s[stride]; raw[k][1] = *(s - stride); raw[k][2] = s[1]; raw[k][3] = *(s - 1);
      } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          raw[k][4 * c + 0] = s[stride + c]; raw[k][4 * c + 1] = *(s - stride + c);
          raw[k][4 * c + 2] = s[3 + c]; raw[k][4 * c + 3] = *(s - 3 + c);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      // dx, dy are integers of magnitude <= 255: dx*dx + dy*dy <= 130050 is exact in float (and in int), so the reference's float
      // comparisons of the squared magnitudes (:221-239) are integer comparisons
      int dx, dy, v;
      if (CN == 1) {                                                   // :207-212
        dy = raw[k][0] - raw[k][1];
        dx = raw[k][2] - raw[k][3];
        v = dx * dx + dy * dy;
      } else {                                                         // :217-240: blue, green, then red is the starting point
        const int dyb = raw[k][0] - raw[k][1], dxb = raw[k][2] - raw[k][3];
        const int vb = dxb * dxb + dyb * dyb;
        const int dyg = raw[k][4] - raw[k][5], dxg = raw[k][6] - raw[k][7];
        const int vg = dxg * dxg + dyg * dyg;
        dy = raw[k][8] - raw[k][9];
        dx = raw[k][10] - raw[k][11];
        v = dx * dx + dy * dy;
        if (vg > v) { v = vg; dx = dxg; dy = dyg; }
        if (vb > v) { v = vb; dx = dxb; dy = dyb; }
      }
      const int lut_o = __ldg(orient_lut + (dy + 255) * kGradSpan + (dx + 255));   // :243-249, tabulated
      const float root = __fsqrt_rn((float)v);                          // :260
      const int i = i0 + k * (HB_X * HB_Y);
      if (i < PW * PH) {
        smag[i] = visited[k] ? root : (SB ? 0.f : -1.f);
        sbo[i] = (unsigned char)(visited[k] ? lut_o : 0);
      }
    }
  }
#pragma unroll
  for (int o = 0; o < 18; ++o) shist[o * (HB_X * HB_Y) + threadIdx.x] = 0.f;
  {
    // bilinear bin coordinates (:252-257) depend on the pixel column / row only: once per CTA instead of once per (block, pixel)
    const double dsb = (double)(float)sbin, inv_sb = 1.0 / dsb;
    const bool pow2 = (sbin & (sbin - 1)) == 0;
    for (int i = threadIdx.x; i < PW + PH; i += HB_X * HB_Y) {
      const bool isx = i < PW;
      const int p = isx ? px0 + i : py0 + (i - PW);
      const float c = bin_coord(p, sbin, dsb, inv_sb, pow2);
      const int ic = (int)floorf(c);
      const float f = __fsub_rn(c, (float)ic);
      const float fc = (float)(1.0 - (double)f);                       // vx1 / vy1, :258-259
      if (isx) { sfx[i] = f; sgx[i] = fc; sbx[i] = ic; } else { sfy[i - PW] = f; sgy[i - PW] = fc; sby[i - PW] = ic; }
    }
  }
  __syncthreads();

  // ---- phase 2: one thread per block gathers its window in raster order (bit-identical to the sequential scatter) ----
  const int bx = bx0 + threadIdx.x % HB_X, by = by0 + threadIdx.x / HB_X;
  if (bx >= L.bw || by >= L.bh) return;
  float* myh = shist + threadIdx.x;
  if constexpr (SB > 0) {
    const int lx0 = SB * (threadIdx.x % HB_X), ly0 = SB * (threadIdx.x / HB_X);   // window origin inside the region
    float wx[2 * SB];
#pragma unroll
    for (int i = 0; i < 2 * SB; ++i) wx[i] = i < SB ? sfx[lx0 + i] : sgx[lx0 + i];   // vx0 for ixp = bx - 1, vx1 for ixp = bx
#pragma unroll 2
    for (int yy = 0; yy < 2 * SB; ++yy) {
      const float wy = yy < SB ? sfy[ly0 + yy] : sgy[ly0 + yy];
      const float4* mrow = reinterpret_cast<const float4*>(smag + (ly0 + yy) * PW + lx0);
      const unsigned* brow = reinterpret_cast<const unsigned*>(sbo + (ly0 + yy) * PW + lx0);
#pragma unroll
      for (int k = 0; k < 2 * SB / 4; ++k) {
        const float4 m4 = mrow[k];
        const unsigned o4 = brow[k];
        const float m[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float term = __fmul_rn(__fmul_rn(wy, wx[4 * k + i]), m[i]);      // :262-265
          float* hb = myh + ((o4 >> (8 * i)) & 0xffu) * (HB_X * HB_Y);
          *hb = __fadd_rn(*hb, term);
        }
      }
    }
  } else {
  const int y_lo = max(1, by * sbin - marg), y_hi = min(vis_h - 2, by * sbin + sbin + marg - 1);
  const int x_lo = max(1, bx * sbin - marg), x_hi = min(vis_w - 2, bx * sbin + sbin + marg - 1);
  for (int y = y_lo; y <= y_hi; ++y) {
    const int iyp = sby[y - py0];                                      // :252
    float wy;
    if (iyp == by) wy = sgy[y - py0];                                  // vy1, :258
    else if (iyp == by - 1) wy = sfy[y - py0];
    else continue;
    const float* mrow = smag + (y - py0) * PW - px0;
    const unsigned char* brow = sbo + (y - py0) * PW - px0;
    for (int x = x_lo; x <= x_hi; ++x) {
      const int ixp = sbx[x - px0];
      float wx;
      if (ixp == bx) wx = sgx[x - px0];
      else if (ixp == bx - 1) wx = sfx[x - px0];
      else continue;
      const float term = __fmul_rn(__fmul_rn(wy, wx), mrow[x]);         // :262-265
      float* hb = myh + brow[x] * (HB_X * HB_Y);
      *hb = __fadd_rn(*hb, term);
    }
  }
  }
  const int idx = L.block_off + by * L.bw + bx;
  float* hp = hist + ((size_t)frame * g->blocks_total + idx) * 18;
  float h[18];
#pragma unroll
  for (int o = 0; o < 18; ++o) { h[o] = myh[o * (HB_X * HB_Y)]; hp[o] = h[o]; }
  float e = 0.f;                                                       // :270-283
#pragma unroll
  for (int o = 0; o < 9; ++o) { const float t = __fadd_rn(h[o], h[o + 9]); e = __fadd_rn(e, __fmul_rn(t, t)); }
  norm[(size_t)frame * g->blocks_total + idx] = e;
}

__device__ __forceinline__ float normaliser(const float* p, int stride) {  // :292-299
  const float s = __fadd_rn(__fadd_rn(__fadd_rn(p[0], p[1]), p[stride]), p[stride + 1]);
  return (float)__ddiv_rn(1.0, __dsqrt_rn(__dadd_rn((double)s, 0.0001)));
}

__global__ void __launch_bounds__(128) hog_feat(const Geometry* __restrict__ g, const float* __restrict__ hist,
                                                const float* __restrict__ norm, float* __restrict__ feat, int frame0) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= g->cells_total) return;
  const int frame = frame0 + blockIdx.y;
  const int l = find_level_by_cell(g, idx);
  const LevelDesc& L = g->lv[l];
  const int local = idx - L.cell_off;
  const int x = local % L.ow, y = local / L.ow;
  const float* nb = norm + (size_t)frame * g->blocks_total + L.block_off;
  const int ns = L.bw;
  const float n1 = normaliser(nb + (y + 1) * ns + (x + 1), ns);
  const float n2 = normaliser(nb + y * ns + (x + 1), ns);
  const float n3 = normaliser(nb + (y + 1) * ns + x, ns);
  const float n4 = normaliser(nb + y * ns + x, ns);
  const float* src = hist + ((size_t)frame * g->blocks_total + L.block_off + (size_t)(y + 1) * L.bw + (x + 1)) * 18;
  float hv[18];
#pragma unroll
  for (int o = 0; o < 18; ++o) hv[o] = src[o];
  float out[32];
  float t1 = 0.f, t2 = 0.f, t3 = 0.f, t4 = 0.f;
  const float clip = (float)0.2;
#pragma unroll
  for (int o = 0; o < 18; ++o) {                                             // :305-317
    const float h1 = fminf(__fmul_rn(hv[o], n1), clip), h2 = fminf(__fmul_rn(hv[o], n2), clip);
    const float h3 = fminf(__fmul_rn(hv[o], n3), clip), h4 = fminf(__fmul_rn(hv[o], n4), clip);
    out[o] = __fmul_rn(0.5f, __fadd_rn(__fadd_rn(__fadd_rn(h1, h2), h3), h4));   // 0.5*sum is exact in either precision
    t1 = __fadd_rn(t1, h1); t2 = __fadd_rn(t2, h2); t3 = __fadd_rn(t3, h3); t4 = __fadd_rn(t4, h4);
  }
#pragma unroll
  for (int o = 0; o < 9; ++o) {                                              // :321-329
    const float sum = __fadd_rn(hv[o], hv[o + 9]);
    const float h1 = fminf(__fmul_rn(sum, n1), clip), h2 = fminf(__fmul_rn(sum, n2), clip);
    const float h3 = fminf(__fmul_rn(sum, n3), clip), h4 = fminf(__fmul_rn(sum, n4), clip);
    out[18 + o] = __fmul_rn(0.5f, __fadd_rn(__fadd_rn(__fadd_rn(h1, h2), h3), h4));
  }
  out[27] = (float)__dmul_rn(0.2357, (double)t1);                            // :332-335
  out[28] = (float)__dmul_rn(0.2357, (double)t2);
  out[29] = (float)__dmul_rn(0.2357, (double)t3);
  out[30] = (float)__dmul_rn(0.2357, (double)t4);
  out[31] = 0.f;                                                             // :338
  float4* dst = reinterpret_cast<float4*>(feat + ((size_t)frame * g->cells_total + idx) * 32);
#pragma unroll
  for (int i = 0; i < 8; ++i) dst[i] = make_float4(out[4 * i], out[4 * i + 1], out[4 * i + 2], out[4 * i + 3]);
}

}  // namespace

size_t hog_orient_lut_bytes() { return (size_t)kGradSpan * kGradSpan; }
int launch_hog_orient_lut(unsigned char* d_lut, cudaStream_t s) {
  hog_orient_lut<<<(kGradSpan * kGradSpan + 255) / 256, 256, 0, s>>>(d_lut);
  return 1;
}

int launch_hog(const Geometry& g, const Geometry* d_g, const DeviceBuffers& b, const unsigned char* d_orient_lut, int sbin, int frame0, int nframes, cudaStream_t s) {
  if (g.blocks_total <= 0) return 0;
  int ntiles = 0;
  for (int l = 0; l < g.n_levels; ++l) ntiles += ((g.lv[l].bw + HB_X - 1) / HB_X) * ((g.lv[l].bh + HB_Y - 1) / HB_Y);
  const int SB = (sbin == 4 || sbin == 8) ? sbin : 0;
  const int marg = SB ? SB / 2 : (sbin + 1) / 2 + 1;
  const int PW = SB ? (HB_X + 1) * SB : HB_X * sbin + sbin + 2 * marg, PH = SB ? (HB_Y + 1) * SB : HB_Y * sbin + sbin + 2 * marg;
  const size_t smem = (size_t)PW * PH * 5 + (size_t)18 * HB_X * HB_Y * 4 + (size_t)(PW + PH) * 16 + 16;   // + the coordinate tables
  dim3 gh(ntiles, nframes);
  auto go = [&](auto kernel) {
    // per launch: the attribute is per device and a process may drive several devices
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kernel<<<gh, HB_X * HB_Y, smem, s>>>(d_g, b.frames, b.pyr, d_orient_lut, b.hist, b.norm, sbin, frame0);
  };
  if (g.in_c == 1) { if (SB == 4) go(hog_hist<1, 4>); else if (SB == 8) go(hog_hist<1, 8>); else go(hog_hist<1, 0>); }
  else { if (SB == 4) go(hog_hist<3, 4>); else if (SB == 8) go(hog_hist<3, 8>); else go(hog_hist<3, 0>); }
  int n = 1;
  if (g.cells_total > 0) {
    dim3 gf((g.cells_total + 127) / 128, nframes);
    hog_feat<<<gf, 128, 0, s>>>(d_g, b.hist, b.norm, b.feat, frame0);
    ++n;
  }
  return n;
}

}  // namespace pbd
