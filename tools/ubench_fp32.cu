// Micro-benchmark: FP32 issue rates on sm_100a for the instruction mixes the
// part-response kernel can use (scalar vs packed f32x2, fused vs separately
// rounded).  Prints instr/clk/SM and "mul-add pairs"/clk/SM.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float2 up(u64 v) { float2 r; asm("mov.b64 {%0,%1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

#define NACC 16
#define ITERS 16384

template <int MODE>
__global__ void __launch_bounds__(256) bench(float* out, const float* in, u64 zero64, long long* cyc) {
  float w0 = in[threadIdx.x & 31], w1 = in[32 + (threadIdx.x & 31)];
  float acc[NACC], acd[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { acc[i] = in[i]; acd[i] = in[i + 7]; }
  u64 a2[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) a2[i] = pk(acc[i], acd[i]);
  u64 w2 = pk(w0, w1), x2 = pk(w1, w0);
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
    if (MODE == 0) {           // scalar FFMA: NACC*2 FFMA
#pragma unroll
      for (int i = 0; i < NACC; ++i) { acc[i] = __fmaf_rn(w0, acd[i], acc[i]); acd[i] = __fmaf_rn(w1, acc[i], acd[i]); }
    } else if (MODE == 1) {    // scalar FMUL + FADD (exact order)
#pragma unroll
      for (int i = 0; i < NACC; ++i) { acc[i] = __fadd_rn(acc[i], __fmul_rn(w0, acd[i])); acd[i] = __fadd_rn(acd[i], __fmul_rn(w1, acc[i])); }
    } else if (MODE == 2) {    // FFMA2
#pragma unroll
      for (int i = 0; i < NACC; ++i) a2[i] = fma2(w2, a2[(i + 1) % NACC], a2[i]);
#pragma unroll
      for (int i = 0; i < NACC; ++i) a2[i] = fma2(x2, a2[(i + 3) % NACC], a2[i]);
    } else if (MODE == 3) {    // FMUL2 + 2 scalar FADD
#pragma unroll
      for (int i = 0; i < NACC; ++i) {
        float2 p = up(mul2(w2, a2[(i + 1) % NACC])); float2 a = up(a2[i]);
        a.x = __fadd_rn(a.x, p.x); a.y = __fadd_rn(a.y, p.y); a2[i] = pk(a.x, a.y);
      }
    } else if (MODE == 4) {    // 2 scalar FMUL + FADD2
#pragma unroll
      for (int i = 0; i < NACC; ++i) {
        float2 s = up(a2[(i + 1) % NACC]);
        a2[i] = add2(a2[i], pk(__fmul_rn(w0, s.x), __fmul_rn(w1, s.y)));
      }
    } else if (MODE == 5) {    // FMUL2 + opaque 64-bit integer add + FADD2
#pragma unroll
      for (int i = 0; i < NACC; ++i) {
        u64 p = mul2(w2, a2[(i + 1) % NACC]);
        asm volatile("add.u64 %0, %0, %1;" : "+l"(p) : "l"(zero64));
        a2[i] = add2(a2[i], p);
      }
    } else if (MODE == 6) {    // FMUL2 only
#pragma unroll
      for (int i = 0; i < NACC; ++i) a2[i] = mul2(w2, a2[i]);
    } else if (MODE == 7) {    // FADD2 only
#pragma unroll
      for (int i = 0; i < NACC; ++i) a2[i] = add2(w2, a2[i]);
    } else if (MODE == 8) {    // scalar FMUL only
#pragma unroll
      for (int i = 0; i < NACC; ++i) { acc[i] = __fmul_rn(w0, acc[i]); acd[i] = __fmul_rn(w1, acd[i]); }
    } else if (MODE == 9) {    // scalar FADD only
#pragma unroll
      for (int i = 0; i < NACC; ++i) { acc[i] = __fadd_rn(w0, acc[i]); acd[i] = __fadd_rn(w1, acd[i]); }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) { float2 v = up(a2[i]); s += acc[i] + acd[i] + v.x + v.y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// FP64: DFMA chain and the DT intersection expression with a true division
template <int MODE>
__global__ void __launch_bounds__(256) bench64(double* out, const double* in, long long* cyc) {
  double a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
  double acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = in[i];
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < 1024; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = __fma_rn(a, acc[(i + 1) & 7], acc[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = __ddiv_rn(__dadd_rn(acc[i], a), __dmul_rn(b, acc[(i + 1) & 7]));
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}


__global__ void calib(long long* out) {
  long long c0 = clock64(); unsigned long long g0, g1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
  float x = 1.0f; for (int i = 0; i < 2000000; ++i) x = __fmaf_rn(x, 1.0000001f, 1e-9f);
  long long c1 = clock64();
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
  out[0] = c1 - c0; out[1] = (long long)(g1 - g0); out[2] = (long long)x;
}
static double g_mhz = 0;
template <int MODE>
void run(const char* name, int instr_per_iter, int pairs_per_iter, float* out, float* in, long long* cyc, int nsm) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int wps = 8; wps <= 32; wps *= 2) {
    int blocks = nsm * (wps / 8);
    bench<MODE><<<blocks, 256>>>(out, in, 0ull, cyc);
    cudaEventRecord(e0);
    bench<MODE><<<blocks, 256>>>(out, in, 0ull, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double cycles = ms * 1e-3 * g_mhz * 1e6;
    double ipc = (double)wps * ITERS * instr_per_iter / cycles;
    double pairs = (double)wps * 32 * ITERS * pairs_per_iter / cycles;
    printf("%-28s warps/SM=%2d ms=%7.3f warp-instr/clk/SM=%5.2f muladd/clk/SM=%6.1f\n", name, wps, ms, ipc, pairs);
  }
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("device %s SMs=%d clock=%d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  float *out, *in; long long* cyc; double *o64, *i64;
  cudaMalloc(&out, 4 << 20); cudaMalloc(&in, 4096); cudaMalloc(&cyc, 8 * 4096); cudaMalloc(&o64, 8 << 20); cudaMalloc(&i64, 4096);
  float hin[1024]; double hin64[512];
  for (int i = 0; i < 1024; ++i) hin[i] = 1.0f + 1e-3f * (i % 13);
  for (int i = 0; i < 512; ++i) hin64[i] = 1.0 + 1e-3 * (i % 13);
  cudaMemcpy(in, hin, 4096, cudaMemcpyHostToDevice); cudaMemcpy(i64, hin64, 4096, cudaMemcpyHostToDevice);
  int nsm = p.multiProcessorCount;
  for (int r = 0; r < 3; ++r) { calib<<<1, 1>>>(cyc); cudaDeviceSynchronize(); long long h[3]; cudaMemcpy(h, cyc, 24, cudaMemcpyDeviceToHost);
    g_mhz = (double)h[0] / (double)h[1] * 1e3; printf("calib: %lld cycles in %lld ns => %.1f MHz\n", h[0], h[1], g_mhz); }
  // warm the clocks with a saturating kernel
  for (int r = 0; r < 20; ++r) bench<0><<<nsm * 4, 256>>>(out, in, 0ull, cyc);
  cudaDeviceSynchronize();
  { calib<<<1, 1>>>(cyc); cudaDeviceSynchronize(); long long h[3]; cudaMemcpy(h, cyc, 24, cudaMemcpyDeviceToHost);
    g_mhz = (double)h[0] / (double)h[1] * 1e3; printf("calib(after warm): %.1f MHz\n", g_mhz); }
  run<0>("FFMA scalar", 2 * NACC, 2 * NACC, out, in, cyc, nsm);
  run<1>("FMUL+FADD scalar (exact)", 4 * NACC, 2 * NACC, out, in, cyc, nsm);
  run<2>("FFMA2", 2 * NACC, 4 * NACC, out, in, cyc, nsm);
  run<3>("FMUL2 + 2 FADD (exact)", 3 * NACC, 2 * NACC, out, in, cyc, nsm);
  run<4>("2 FMUL + FADD2 (exact)", 3 * NACC, 2 * NACC, out, in, cyc, nsm);
  run<5>("FMUL2+IADD64+FADD2 (exact)", 4 * NACC, 2 * NACC, out, in, cyc, nsm);
  run<6>("FMUL2 only", NACC, 2 * NACC, out, in, cyc, nsm);
  run<7>("FADD2 only", NACC, 2 * NACC, out, in, cyc, nsm);
  run<8>("FMUL scalar only", 2 * NACC, 2 * NACC, out, in, cyc, nsm);
  run<9>("FADD scalar only", 2 * NACC, 2 * NACC, out, in, cyc, nsm);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 2; ++mode) {
    for (int wps = 8; wps <= 32; wps *= 2) {
      int blocks = nsm * (wps / 8);
      float ms = 0;
      for (int r = 0; r < 2; ++r) { cudaEventRecord(e0); if (mode == 0) bench64<0><<<blocks, 256>>>(o64, i64, cyc); else bench64<1><<<blocks, 256>>>(o64, i64, cyc); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); }
      double cycles = ms * 1e-3 * g_mhz * 1e6;
      printf("%-28s warps/SM=%2d ms=%7.3f ops/clk/SM=%6.2f\n", mode == 0 ? "DFMA" : "DADD+DMUL+DDIV", wps, ms, (double)wps * 32 * 1024 * 8 / cycles);
    }
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
