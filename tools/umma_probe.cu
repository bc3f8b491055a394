// umma_probe.cu -- hardware probe for the tensor-core response kernel (sm_100a): checks that a K-major
// SWIZZLE_128B shared-memory descriptor of tcgen05.mma.kind::tf32 (a) multiplies correctly for M=128, N=144,
// (b) may START at any 128-byte row of a swizzled tile (row-shifted start address, base_offset 0), which is what
// lets one staged strip of HOG cells serve the kx taps of a filter row; and measures the MMA issue pace.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_probe tools/umma_probe.cu && tools/umma_probe
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int M = 128, N = 144, KROW = 32;   // one 128-byte row = 32 tf32 = 4 MMA k-steps of 8
constexpr int AROWS = 160;                   // rows staged for A (so the start can shift by up to 32 rows)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t base_offset) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);        // start address
  d |= (uint64_t)1 << 16;                         // LBO (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;               // SBO: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                         // version 1 (sm_100)
  d |= (uint64_t)(base_offset & 7) << 49;
  d |= (uint64_t)2 << 61;                         // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

__global__ void __launch_bounds__(128, 1)
probe(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int shift, int base_offset, int reps,
      long long* cycles, int* err) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                      // AROWS x 128 B, swizzled on absolute address bits
  uint8_t* sB = smem + AROWS * 128;        // N x 128 B (AROWS*128 = 20480 = 20 x 1024: still 1024-aligned)
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;

  for (int i = tid; i < AROWS * 8; i += 128) {      // 16-byte chunks
    const int r = i >> 3, c = i & 7;
    const float4 v = *reinterpret_cast<const float4*>(A + r * KROW + c * 4);
    *reinterpret_cast<float4*>(sA + r * 128 + ((c ^ (r & 7)) << 4)) = v;
  }
  for (int i = tid; i < N * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    const float4 v = *reinterpret_cast<const float4*>(B + r * KROW + c * 4);
    *reinterpret_cast<float4*>(sB + r * 128 + ((c ^ (r & 7)) << 4)) = v;
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;\n" ::"r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    const uint32_t a0 = smem_u32(sA) + shift * 128, b0 = smem_u32(sB);
    t0 = clock64();
    for (int rep = 0; rep < reps; ++rep)
      for (int k = 0; k < 4; ++k)
        mma_tf32(tmem, make_desc(a0 + k * 32, base_offset), make_desc(b0 + k * 32, 0), idesc, (rep | k) ? 1u : 0u);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)) : "memory");
  }
  // bounded wait: a wrong descriptor must not hang the box
  {
    const uint32_t a = smem_u32(&bar);
    uint32_t done = 0;
    long long spins = 0;
    while (!done) {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(a), "r"(0u) : "memory");
      if (++spins > 200000000ll) { if (tid == 0) *err = 1; break; }
    }
  }
  if (tid == 0) { t1 = clock64(); *cycles = t1 - t0; }
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  // epilogue: warp w reads TMEM lanes 32w..32w+31 (= rows of D), 16 columns at a time
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    for (int j = 0; j < 16; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;\n" ::"r"(tmem));
}

// Issue pattern of part_response_tc without any data movement: per "tap" 4 MMAs into a ping-pong accumulator and 8 into a third
// one, operands taken from varying rows / slabs of shared memory, optionally one tcgen05.commit per tap.  Measures the tensor
// pipe's sustained pace for that pattern (cycles per MMA).
__global__ void __launch_bounds__(128, 1) pattern(int taps, int mode, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) unsigned long long bar[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 200 * 1024 / 4; i += 128) reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0.f;
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  __shared__ volatile int done_flag;
  if (tid == 0) done_flag = 0;
  __syncthreads();
  if (warp != 0 && (mode & 8)) {         // other warps poll an mbarrier that never completes, like idle pipeline roles do
    unsigned done = 0;
    while (!done_flag)
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_u32(&bar[1])), "r"(0u) : "memory");
  }
  if (warp != 0 && (mode & 16)) {        // other warps read tensor memory continuously, like the epilogue does
    unsigned sink = 0;
    while (!done_flag) {
      uint32_t v[16];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                     "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                   : "r"(tmem_base_s + ((uint32_t)(warp * 32) << 16) + 300u));
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      sink += v[3];
    }
    if (sink == 77u) cycles[3] = 1;
  }
  if (tid == 0) {
    long long commit_cycles = 0;
    unsigned busy = (unsigned)taps;
    const long long t0 = clock64();
    for (int tap = 0; tap < taps; ++tap) {
      const uint32_t a_hi = base + (uint32_t)((tap / 5) & 1) * 36864u + (uint32_t)(3 + tap % 5) * 128u, a_lo = a_hi + 18432u;
      const uint32_t b_hi = base + 73728u + (uint32_t)(tap & 3) * 36864u, b_lo = b_hi + 18432u;
      const uint32_t dH = (mode & 2) ? tmem : tmem + (uint32_t)(((tap / 5) & 1) * N), dC = (mode & 2) ? tmem : tmem + 2 * N;
      for (int k = 0; k < 4; ++k) mma_tf32(dH, make_desc(a_hi + k * 32, 0), make_desc(b_hi + k * 32, 0), idesc, 1u);
      for (int k = 0; k < 4; ++k) mma_tf32(dC, make_desc(a_lo + k * 32, 0), make_desc(b_hi + k * 32, 0), idesc, 1u);
      for (int k = 0; k < 4; ++k) mma_tf32(dC, make_desc(a_hi + k * 32, 0), make_desc(b_lo + k * 32, 0), idesc, 1u);
      long long tc0 = clock64();
      if (mode & 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar[1])) : "memory");
      commit_cycles += clock64() - tc0;
      if (mode & 4) {                      // ~500 cycles of dependent busy work per tap (stands for waits / bookkeeping)
        for (int i = 0; i < 100; ++i) busy = busy * 1664525u + 1013904223u;
      }
    }
    if (busy == 12345u) cycles[2] = 1;
    cycles[1] = commit_cycles;
    done_flag = 1;
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar[0])) : "memory");
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_u32(&bar[0])), "r"(0u) : "memory");
    *cycles = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem));
}

int main() {
  std::vector<float> hA(AROWS * KROW), hB(N * KROW), hD(M * N);
  srand(7);
  for (auto& v : hA) v = (float)(rand() % 9 - 4);
  for (auto& v : hB) v = (float)(rand() % 9 - 4);
  float *dA, *dB, *dD;
  long long* dcyc;
  int* derr;
  CK(cudaMalloc(&dA, hA.size() * 4)); CK(cudaMalloc(&dB, hB.size() * 4)); CK(cudaMalloc(&dD, hD.size() * 4));
  CK(cudaMalloc(&dcyc, 64)); CK(cudaMalloc(&derr, 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));
  const int smem = (AROWS + N) * 128 + 1024;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  struct Case { int shift, bo, reps; } cases[] = {{0, 0, 1}, {8, 0, 1}, {1, 0, 1}, {1, 1, 1}, {3, 0, 1}, {3, 3, 1}, {5, 0, 1}, {13, 0, 1}, {13, 5, 1},
                                                 {0, 0, 1000}, {3, 0, 1000}};
  for (auto c : cases) {
    CK(cudaMemset(dD, 0, hD.size() * 4)); CK(cudaMemset(derr, 0, 4));
    probe<<<1, 128, smem>>>(dA, dB, dD, c.shift, c.bo, c.reps, dcyc, derr);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("shift %d bo %d: launch failed: %s\n", c.shift, c.bo, cudaGetErrorString(e)); return 1; }
    long long cyc; int err;
    CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&err, derr, 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int i = 0; i < M; ++i)
      for (int n = 0; n < N; ++n) {
        double s = 0;
        for (int k = 0; k < KROW; ++k) s += (double)hA[(i + c.shift) * KROW + k] * hB[n * KROW + k];
        if ((double)hD[i * N + n] != s * c.reps) ++bad;
      }
    printf("shift %2d base_offset %d reps %4d: %s mismatches %d / %d, timeout %d, %lld cycles (%.1f per MMA)\n", c.shift, c.bo, c.reps,
           bad ? "FAIL" : "ok  ", bad, M * N, err, cyc, (double)cyc / (4.0 * c.reps));
  }
  const int psmem = 222 * 1024;
  CK(cudaFuncSetAttribute(pattern, cudaFuncAttributeMaxDynamicSharedMemorySize, psmem));
  for (int mode : {1, 5, 9, 13, 17, 29}) {
    for (int grid : {1}) {
      pattern<<<grid, 128, psmem>>>(500, mode, dcyc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("pattern launch failed: %s\n", cudaGetErrorString(e)); return 1; }
      long long cyc[2];
      CK(cudaMemcpy(cyc, dcyc, 16, cudaMemcpyDeviceToHost));
      printf("pattern mode %2d (commit per tap %d, busy work per tap %d, polling warps %d, LDTM warps %d) grid %3d: %.1f cycles per tap (MMA floor 864), of which in commit %.1f\n",
             mode, mode & 1, (mode >> 2) & 1, (mode >> 3) & 1, (mode >> 4) & 1, grid, (double)cyc[0] / 500.0, (double)cyc[1] / 500.0);
    }
  }
  return 0;
}
