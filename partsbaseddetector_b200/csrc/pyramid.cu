// pyramid.cu -- image pyramid of HOGFeatures<T>::pyramid (reference src/HOGFeatures.cpp:109-127):
// `interval` bilinear resizes of the input frame (cv::resize, INTER_LINEAR, 8U fixed point) followed by
// chains of cv::pyrDown (5x5 binomial, BORDER_REFLECT_101, 8U).  Pure integer arithmetic => bit-exact.
// HBM-bound: every output pixel is written once, sources are read through L1/L2 (a pyrDown tap
// footprint is re-read by 6.25 outputs on average, all hits).
#include <algorithm>
#include <vector>

#include "kernels.cuh"

namespace pbd {
namespace {

__device__ __forceinline__ int reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
  return p;
}

// One thread per destination pixel of one resized level.  Coefficient tables are built on the host
// exactly as OpenCV's resize() does (float coordinate, cvFloor, cvRound(f*2048) as int16).
__global__ void __launch_bounds__(256) pyr_resize_u8(const Geometry* __restrict__ g, const uint8_t* __restrict__ frames,
                                                     uint8_t* __restrict__ pyr, const int* __restrict__ xofs,
                                                     const short* __restrict__ xalpha, const int* __restrict__ yofs,
                                                     const short* __restrict__ ybeta, int4 levels, int frame0) {
  const int level = blockIdx.z == 0 ? levels.x : blockIdx.z == 1 ? levels.y : blockIdx.z == 2 ? levels.z : levels.w;   // independent levels share a launch
  const LevelDesc& L = g->lv[level];
  const int dw = L.img_w, dh = L.img_h, cn = g->in_c, sw = g->in_w, sh = g->in_h;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= dw * dh) return;
  const int dx = idx % dw, dy = idx / dw;
  const int frame = frame0 + blockIdx.y;
  const uint8_t* S = frames + (size_t)frame * sh * sw * cn;
  uint8_t* D = pyr + (size_t)frame * g->img_bytes + L.img_off + (size_t)idx * cn;
  const int sx = xofs[L.xofs_off + dx], sx1 = min(sx + 1, sw - 1);
  const int a0 = xalpha[2 * (L.xofs_off + dx)], a1 = xalpha[2 * (L.xofs_off + dx) + 1];
  const int sy = yofs[L.yofs_off + dy];
  const int sy0 = min(max(sy, 0), sh - 1), sy1 = min(max(sy + 1, 0), sh - 1);
  const int b0 = ybeta[2 * (L.yofs_off + dy)], b1 = ybeta[2 * (L.yofs_off + dy) + 1];
  const uint8_t* r0p = S + (size_t)sy0 * sw * cn;
  const uint8_t* r1p = S + (size_t)sy1 * sw * cn;
  for (int c = 0; c < cn; ++c) {
    const int r0 = r0p[sx * cn + c] * a0 + r0p[sx1 * cn + c] * a1;
    const int r1 = r1p[sx * cn + c] * a0 + r1p[sx1 * cn + c] * a1;
    const int v = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
    D[c] = (uint8_t)min(max(v, 0), 255);
  }
}

// out = (sum_{i,j} k[i]k[j] src[2y+i-2][2x+j-2] + 128) >> 8, k = [1 4 6 4 1] (integer arithmetic: any summation order gives cv::pyrDown's
// bits).  One thread per VY = 4 vertically adjacent destination pixels: their windows span 2 VY + 3 = 11 source rows, so the horizontal
// 5-tap sum of a source row is formed once and enters up to three of the four outputs -- 41 byte loads per destination pixel instead
// of 75 (the kernel is bound by instruction issue).  Threads of a warp still walk along x: loads and stores stay coalesced.
constexpr int kPyrVY = 4;
__global__ void __launch_bounds__(256) pyr_down_u8(const Geometry* __restrict__ g, const uint8_t* __restrict__ frames, uint8_t* __restrict__ pyr, int4 levels,
                                                   int frame0) {
  const int level = blockIdx.z == 0 ? levels.x : blockIdx.z == 1 ? levels.y : blockIdx.z == 2 ? levels.z : levels.w;
  const LevelDesc& L = g->lv[level];
  const LevelDesc& P = g->lv[L.src_level];
  const int dw = L.img_w, dh = L.img_h, cn = g->in_c, sw = P.img_w, sh = P.img_h;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int nyq = (dh + kPyrVY - 1) / kPyrVY;                            // groups of VY rows
  if (idx >= dw * nyq) return;
  const int x = idx % dw, y0 = (idx / dw) * kPyrVY;
  const int frame = frame0 + blockIdx.y;
  const uint8_t* S = P.identity ? frames + (size_t)frame * sh * sw * cn : pyr + (size_t)frame * g->img_bytes + P.img_off;   // level 0 is the frame itself
  uint8_t* D = pyr + (size_t)frame * g->img_bytes + L.img_off;
  const int k[5] = {1, 4, 6, 4, 1};
  int xs[5];
#pragma unroll
  for (int j = 0; j < 5; ++j) xs[j] = reflect101(2 * x + j - 2, sw) * cn;
  int acc[kPyrVY][3];
#pragma unroll
  for (int v = 0; v < kPyrVY; ++v) acc[v][0] = acc[v][1] = acc[v][2] = 0;
#pragma unroll
  for (int r = 0; r < 2 * kPyrVY + 3; ++r) {                             // source row 2 y0 - 2 + r
    const uint8_t* row = S + (size_t)reflect101(2 * y0 - 2 + r, sh) * sw * cn;
    int h[3] = {0, 0, 0};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (c < cn) {
#pragma unroll
        for (int j = 0; j < 5; ++j) h[c] += k[j] * row[xs[j] + c];
      }
    }
#pragma unroll
    for (int v = 0; v < kPyrVY; ++v) {
      const int i = r - 2 * v;                                           // tap of output y0 + v that this row feeds
      if (i >= 0 && i < 5) {
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[v][c] += k[i] * h[c];
      }
    }
  }
#pragma unroll
  for (int v = 0; v < kPyrVY; ++v) {
    if (y0 + v >= dh) break;
    uint8_t* d = D + ((size_t)(y0 + v) * dw + x) * cn;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      if (c < cn) d[c] = (uint8_t)((acc[v][c] + 128) >> 8);
  }
}

}  // namespace

int launch_pyramid(const Geometry& g, const Geometry* d_g, const DeviceBuffers& b, const int* d_xofs, const short* d_xalpha,
                   const int* d_yofs, const short* d_ybeta, int frame0, int nframes, cudaStream_t s) {
  int launches = 0;
  // Resized levels depend only on the frame; pyrDown level l depends on level l - interval.  Levels of the same "generation"
  // (all resizes; then the pyrDowns whose sources exist) are independent and share one launch (blockIdx.z picks the level, up to 4
  // per launch), generations follow each other on the stream.
  std::vector<int> gen(g.n_levels, 0);
  int ngen = 0;
  for (int l = 0; l < g.n_levels; ++l) { gen[l] = g.lv[l].src_level < 0 ? 0 : gen[g.lv[l].src_level] + 1; ngen = std::max(ngen, gen[l] + 1); }
  for (int k = 0; k < ngen; ++k) {
    std::vector<int> ls;
    for (int l = 0; l < g.n_levels; ++l)
      if (gen[l] == k && g.lv[l].img_w * g.lv[l].img_h > 0 && !g.lv[l].identity) ls.push_back(l);   // a level of the frame's own size is read in place
    for (size_t i = 0; i < ls.size(); i += 4) {
      const int n = (int)std::min<size_t>(4, ls.size() - i);
      int npx = 0;
      int4 lv = make_int4(ls[i], ls[i], ls[i], ls[i]);
      for (int j = 0; j < n; ++j) {
        const LevelDesc& Lj = g.lv[ls[i + j]];
        npx = std::max(npx, k == 0 ? Lj.img_w * Lj.img_h : Lj.img_w * ((Lj.img_h + kPyrVY - 1) / kPyrVY));   // pyr_down: a thread per VY rows
        (j == 0 ? lv.x : j == 1 ? lv.y : j == 2 ? lv.z : lv.w) = ls[i + j];
      }
      dim3 grid((npx + 255) / 256, nframes, n);
      if (k == 0) pyr_resize_u8<<<grid, 256, 0, s>>>(d_g, b.frames, b.pyr, d_xofs, d_xalpha, d_yofs, d_ybeta, lv, frame0);
      else pyr_down_u8<<<grid, 256, 0, s>>>(d_g, b.frames, b.pyr, lv, frame0);
      ++launches;
    }
  }
  return launches;
}

}  // namespace pbd
