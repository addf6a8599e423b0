// Cost of skipping a DMMA with a real warp-uniform branch (forced by wrapping the instruction in a never-repeating loop, which
// ptxas cannot if-convert) versus predicating it off.  32 slots per iteration, mask density varied.  1 CTA (8 warps) per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
template <bool LOOP> __global__ void k(double* out, const uint32_t* masks, int iters, int never) {
  extern __shared__ double sm[];
  for (int t = threadIdx.x; t < 4096; t += blockDim.x) sm[t] = 1e-3 * t;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  double c[32][2];
#pragma unroll
  for (int i = 0; i < 32; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
  const int itU = __reduce_max_sync(0xffffffffu, iters);
  for (int it = 0; it < itU; ++it) {
    const uint32_t m = __reduce_or_sync(0xffffffffu, masks[it & 1023]);
    const double a = sm[(it & 63) * 32 + lane], b = sm[2048 + (it & 63) * 32 + lane];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      if ((m >> i) & 1u) {
        if (LOOP) { do { dmma(c[i], a, b); } while (never); }
        else dmma(c[i], a, b);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount, iters = 4000, smem = 150 * 1024;
  double* out; cudaMalloc(&out, sizeof(double) * sms * 1024);
  uint32_t* masks; cudaMalloc(&masks, 4096);
  cudaFuncSetAttribute(k<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int dens : {32, 21, 16, 8, 0}) {
    uint32_t h[1024]; uint32_t seed = 12345; long on = 0;
    for (int i = 0; i < 1024; ++i) { uint32_t v = 0; for (int b = 0; b < 32; ++b) { seed = seed * 1664525u + 1013904223u; if ((int)((seed >> 16) % 32) < dens) v |= 1u << b; } h[i] = v; on += __builtin_popcount(v); }
    cudaMemcpy(masks, h, 4096, cudaMemcpyHostToDevice);
    for (int loop = 0; loop < 2; ++loop) for (int warps : {4, 8}) {
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      auto run = [&] { if (loop) k<true><<<sms, warps * 32, smem>>>(out, masks, iters, 0); else k<false><<<sms, warps * 32, smem>>>(out, masks, iters, 0); };
      run(); cudaDeviceSynchronize(); cudaEventRecord(e0); run(); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double cyc = ms * 1e-3 * 1.965e9, slots = 32.0 * iters * warps / 4, live = slots * (on / (1024.0 * 32));
      printf("density %2d/32 %-10s warps/SM %d : %.1f cycles per slot per SMSP, %.1f per live DMMA\n", dens, loop ? "branch" : "predicated", warps, cyc / slots, live > 0 ? cyc / live : 0.0);
    }
  }
  return 0;
}
