// nms.cu -- Candidate::sort + Candidate::nonMaximaSuppression (reference include/Candidate.hpp:97-99, 277-304) on the device,
// as the callers of detect() run them (ros/Node.cpp:192-196, cells/detect.cpp:237-238): per frame, candidates in
// descending score order, each kept iff the painted fraction of its bounding box (the hull of its part rectangles clipped
// to the image) is <= overlap, then painted.  The greedy painting is order dependent, so it stays sequential per frame
// (one CTA per frame walks its sorted candidates; the CTA's threads count / paint the box cooperatively on a 1-bit-per-pixel
// scratch image); frames are independent.  Only the kept candidates are compacted for the download.
//
// Ties in score are broken by the canonical order of the raw candidate list (level, component, row-major hit), i.e. the
// result equals a stable sort of detect()'s output followed by the reference's loop -- what pbd_candidates_sort +
// pbd_candidates_nms compute on the host (abi.cpp), against which tests/test_gpu_parity.py checks this path.
#include <algorithm>

#include "kernels.cuh"

namespace pbd {
namespace {

struct IRect { int x, y, w, h; };
__device__ __forceinline__ bool rect_empty(const IRect& r) { return r.w <= 0 || r.h <= 0; }
__device__ __forceinline__ IRect rect_or(IRect a, const IRect& b) {            // cv::Rect operator|
  if (rect_empty(a)) return b;
  if (rect_empty(b)) return a;
  const int x1 = min(a.x, b.x), y1 = min(a.y, b.y);
  a.w = max(a.x + a.w, b.x + b.w) - x1; a.h = max(a.y + a.h, b.y + b.h) - y1; a.x = x1; a.y = y1;
  return a;
}
__device__ __forceinline__ IRect rect_and(IRect a, const IRect& b) {           // cv::Rect operator&
  const int x1 = max(a.x, b.x), y1 = max(a.y, b.y);
  a.w = min(a.x + a.w, b.x + b.w) - x1; a.h = min(a.y + a.h, b.y + b.h) - y1; a.x = x1; a.y = y1;
  if (a.w <= 0 || a.h <= 0) a = IRect{0, 0, 0, 0};
  return a;
}

// per hit: bounding box (Candidate::boundingBox, :104-110, clipped to the image as nonMaximaSuppression does) and the sort key;
// per frame: hit count
__global__ void __launch_bounds__(128)
nms_prepare(const Geometry* __restrict__ g, const Hit* __restrict__ hits, const int* __restrict__ nhits, int max_hits, const int* __restrict__ xym,
            int out_parts, const int* __restrict__ nparts, const int* __restrict__ ksize /* [ncomp][kMaxParts][kMaxMix] filter rows */,
            int4* __restrict__ boxes, unsigned long long* __restrict__ keys, int* __restrict__ frame_count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = min(*nhits, max_hits);
  if (i >= n) return;
  const Hit h = hits[i];
  const float scale = g->lv[h.level].scale;
  const int* xs = xym + (size_t)i * 3 * out_parts;
  const int np = nparts[h.comp];
  IRect hull{0, 0, 0, 0};
  for (int p = 0; p < np; ++p) {                                              // part rectangles: src/DynamicProgram.cpp:238-244
    const int x = xs[p], y = xs[out_parts + p], mix = xs[2 * out_parts + p];
    const int ks = ksize[(h.comp * kMaxParts + p) * kMaxMix + mix];
    const int x1 = __float2int_rn(__fmul_rn((float)(x - 1), scale)), y1 = __float2int_rn(__fmul_rn((float)(y - 1), scale));
    const int sz = __float2int_rn(__fmul_rn((float)ks, scale));
    const int x2 = x1 + sz - 1, y2 = y1 + sz - 1;
    const IRect r{min(x1, x2), min(y1, y2), max(x1, x2) - min(x1, x2), max(y1, y2) - min(y1, y2)};
    hull = p == 0 ? r : rect_or(hull, r);
  }
  const IRect b = rect_and(hull, IRect{0, 0, g->in_w, g->in_h});
  boxes[i] = make_int4(b.x, b.y, b.w, b.h);
  unsigned int u = __float_as_uint(h.score);
  u ^= (u >> 31) ? 0xFFFFFFFFu : 0x80000000u;                                  // ascending-orderable float
  const unsigned int canon = ((unsigned)h.level << 25) | ((unsigned)h.comp << 20) | ((unsigned)h.y << 10) | (unsigned)h.x;
  keys[i] = ((unsigned long long)(~u) << 32) | canon;                          // ascending key = score descending, then canonical order
  atomicAdd(&frame_count[h.frame], 1);
}

// segment offsets: every frame's candidates are sorted in a power-of-two sized segment
__global__ void nms_offsets(const int* __restrict__ frame_count, int n_frames, int* __restrict__ seg_off /* [n_frames + 1] */,
                            int* __restrict__ fill /* [n_frames] */) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int off = 0;
  for (int f = 0; f < n_frames; ++f) {
    seg_off[f] = off;
    fill[f] = 0;
    int c = frame_count[f], p2 = 1;
    while (p2 < c) p2 <<= 1;
    off += c ? p2 : 0;
  }
  seg_off[n_frames] = off;
}

__global__ void __launch_bounds__(128)
nms_scatter(const Hit* __restrict__ hits, const int* __restrict__ nhits, int max_hits, const unsigned long long* __restrict__ keys,
            const int* __restrict__ seg_off, int* __restrict__ fill, unsigned long long* __restrict__ skeys, int* __restrict__ sidx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = min(*nhits, max_hits);
  if (i >= n) return;
  const int f = hits[i].frame;
  const int pos = seg_off[f] + atomicAdd(&fill[f], 1);
  skeys[pos] = keys[i];
  sidx[pos] = i;
}

// one CTA per frame: bitonic sort of the frame's segment (keys are unique within a frame, so the result does not depend on the
// scatter order), then the sequential painting loop
constexpr int kNmsThreads = 512;
__global__ void __launch_bounds__(kNmsThreads)
nms_frames(const int* __restrict__ frame_count, const int* __restrict__ seg_off, unsigned long long* __restrict__ skeys, int* __restrict__ sidx,
           const int4* __restrict__ boxes, unsigned int* __restrict__ scratch, int words_per_row, int im_h, float overlap,
           int* __restrict__ kept_count, int* __restrict__ kept_idx /* [seg] compacted per frame, in order */) {
  const int f = blockIdx.x;
  const int cnt = frame_count[f];
  const int base = seg_off[f];
  int p2 = cnt ? 1 : 0;                                        // a frame without candidates owns no segment
  while (p2 < cnt) p2 <<= 1;
  const int tid = threadIdx.x;
  for (int i = cnt + tid; i < p2; i += kNmsThreads) { skeys[base + i] = ~0ull; sidx[base + i] = -1; }   // padding sorts last
  __syncthreads();
  for (int k = 2; k <= p2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < p2; i += kNmsThreads) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = skeys[base + i], b = skeys[base + ixj];
          const bool up = (i & k) == 0;
          if ((a > b) == up) {
            skeys[base + i] = b; skeys[base + ixj] = a;
            const int t = sidx[base + i]; sidx[base + i] = sidx[base + ixj]; sidx[base + ixj] = t;
          }
        }
      }
      __syncthreads();
    }
  unsigned int* img = scratch + (size_t)f * im_h * words_per_row;
  for (int i = tid; i < im_h * words_per_row; i += kNmsThreads) img[i] = 0u;
  __shared__ int warp_sum[kNmsThreads / 32];
  __shared__ int s_keep;
  int kept = 0;
  __syncthreads();
  for (int c = 0; c < cnt; ++c) {
    const int hit = sidx[base + c];
    const int4 b = boxes[hit];
    const int w0 = b.x >> 5, w1 = (b.x + b.z - 1) >> 5;                        // word range of the box's columns
    const int nw = b.z > 0 && b.w > 0 ? w1 - w0 + 1 : 0;
    const unsigned int first_mask = 0xFFFFFFFFu << (b.x & 31);
    const unsigned int last_mask = 0xFFFFFFFFu >> (31 - ((b.x + b.z - 1) & 31));
    int local = 0;
    for (int i = tid; i < nw * b.w; i += kNmsThreads) {
      const int r = i / nw, wi = i - r * nw;
      unsigned int m = 0xFFFFFFFFu;
      if (wi == 0) m &= first_mask;
      if (wi == nw - 1) m &= last_mask;
      local += __popc(img[(size_t)(b.y + r) * words_per_row + w0 + wi] & m);
    }
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((tid & 31) == 0) warp_sum[tid >> 5] = local;
    __syncthreads();
    if (tid == 0) {
      long long sum = 0;
      for (int w = 0; w < kNmsThreads / 32; ++w) sum += warp_sum[w];
      // `if (boxsum[0] / box.area() > overlap) continue;` -- 0/0 = NaN for an empty box, which is kept
      const double ratio = (double)sum / (double)(b.z * b.w);
      s_keep = !(ratio > (double)overlap);
    }
    __syncthreads();
    if (s_keep) {
      for (int i = tid; i < nw * b.w; i += kNmsThreads) {
        const int r = i / nw, wi = i - r * nw;
        unsigned int m = 0xFFFFFFFFu;
        if (wi == 0) m &= first_mask;
        if (wi == nw - 1) m &= last_mask;
        img[(size_t)(b.y + r) * words_per_row + w0 + wi] |= m;
      }
      if (tid == 0) kept_idx[base + kept] = hit;
      ++kept;
    }
    __syncthreads();
  }
  if (tid == 0) kept_count[f] = kept;
}

// compaction: kept candidates of all frames, frame by frame, into dense hit / part arrays
__global__ void nms_out_offsets(const int* __restrict__ kept_count, int n_frames, int* __restrict__ out_off, int* __restrict__ total) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int off = 0;
  for (int f = 0; f < n_frames; ++f) { out_off[f] = off; off += kept_count[f]; }
  *total = off;
}
__global__ void __launch_bounds__(128)
nms_gather(const int* __restrict__ total, int n_frames, const int* __restrict__ seg_off, const int* __restrict__ out_off,
           const int* __restrict__ kept_idx, const Hit* __restrict__ hits, const int* __restrict__ xym, int row_ints, Hit* __restrict__ hits_out,
           int* __restrict__ xym_out) {
  const int dst = blockIdx.x * blockDim.y + threadIdx.y;
  if (dst >= *total) return;
  int lo = 0, hi = n_frames - 1;                               // the frame whose output range holds dst
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (out_off[mid] <= dst) lo = mid; else hi = mid - 1;
  }
  const int src = kept_idx[seg_off[lo] + (dst - out_off[lo])];
  if (threadIdx.x == 0) hits_out[dst] = hits[src];
  for (int i = threadIdx.x; i < row_ints; i += blockDim.x) xym_out[(size_t)dst * row_ints + i] = xym[(size_t)src * row_ints + i];
}

}  // namespace

size_t nms_scratch_words(const Geometry& g) { return (size_t)g.n_frames * g.in_h * ((g.in_w + 31) / 32); }

int launch_device_nms(const Geometry& g, const Geometry* d_g, const NmsBuffers& nb, const Hit* d_hits, const int* d_nhits, int max_hits,
                      const int* d_xym, int out_parts, const int* d_nparts, const int* d_ksize, float overlap, cudaStream_t s) {
  if (max_hits <= 0 || g.n_frames <= 0) return 0;
  cudaMemsetAsync(nb.frame_count, 0, sizeof(int) * g.n_frames, s);
  nms_prepare<<<(max_hits + 127) / 128, 128, 0, s>>>(d_g, d_hits, d_nhits, max_hits, d_xym, out_parts, d_nparts, d_ksize, nb.boxes, nb.keys,
                                                     nb.frame_count);
  nms_offsets<<<1, 32, 0, s>>>(nb.frame_count, g.n_frames, nb.seg_off, nb.fill);
  nms_scatter<<<(max_hits + 127) / 128, 128, 0, s>>>(d_hits, d_nhits, max_hits, nb.keys, nb.seg_off, nb.fill, nb.skeys, nb.sidx);
  nms_frames<<<g.n_frames, kNmsThreads, 0, s>>>(nb.frame_count, nb.seg_off, nb.skeys, nb.sidx, nb.boxes, nb.scratch, (g.in_w + 31) / 32, g.in_h,
                                                overlap, nb.kept_count, nb.kept_idx);
  nms_out_offsets<<<1, 32, 0, s>>>(nb.kept_count, g.n_frames, nb.out_off, nb.total);
  const int per_block = 4;                                     // launched for the capacity; surplus blocks read `total` and exit
  nms_gather<<<(max_hits + per_block - 1) / per_block, dim3(32, per_block), 0, s>>>(nb.total, g.n_frames, nb.seg_off, nb.out_off, nb.kept_idx,
                                                                                   d_hits, d_xym, 3 * out_parts, nb.hits_out, nb.xym_out);
  return 6;
}

}  // namespace pbd
