// DMMA throughput vs resident warps per SM (1 CTA/SM forced by dynamic shared memory), 32 independent accumulators per warp,
// operands from registers (R) or re-loaded from shared memory before every instruction (S).
#include <cstdio>
#include <cuda_runtime.h>
template <int NACC, bool SMEM, bool PRED> __global__ void k(double* out, int iters, unsigned bits) {
  extern __shared__ double sm[];
  for (int t = threadIdx.x; t < 4096; t += blockDim.x) sm[t] = 1e-3 * t;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
  double a = sm[lane], b = sm[lane + 32];
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      if (SMEM) { a = sm[((it + i) & 63) * 32 + lane]; b = sm[2048 + ((it + i) & 63) * 32 + lane]; }
      if (PRED)
        asm volatile("{\n.reg .pred p;\n.reg .b32 t;\nand.b32 t, %4, %5;\nsetp.ne.u32 p, t, 0;\n@p mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n}\n"
                     : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b), "r"(bits), "r"(1u << (i % 8)));
      else
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F> static float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
template <int NACC, bool SMEM, bool PRED> void run(const char* name, double* out, int sms, unsigned bits = 0xffu) {
  const int iters = 4000, smem = 150 * 1024;
  cudaFuncSetAttribute(k<NACC, SMEM, PRED>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int warps : {4, 8, 12, 16, 32}) {
    float ms = timeit([&] { k<NACC, SMEM, PRED><<<sms, warps * 32, smem>>>(out, iters, bits); });
    printf("%-28s warps/SM %2d : %6.2f TFLOP/s  (%.1f cycles per DMMA per SMSP at 1.965 GHz)\n", name, warps, 2.0 * 256 * NACC * iters * (double)sms * warps / ms * 1e-9,
           ms * 1e-3 * 1.965e9 / ((double)NACC * iters * warps / 4));
  }
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double* out; cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 1024);
  run<32, false, false>("R 32acc", out, p.multiProcessorCount);
  run<32, true, false>("S 32acc", out, p.multiProcessorCount);
  run<32, true, true>("S 32acc predicated", out, p.multiProcessorCount);
  run<8, false, false>("R 8acc", out, p.multiProcessorCount);
  run<32, true, true>("S 32acc predicated half on", out, p.multiProcessorCount, 0x55u);
  run<32, true, true>("S 32acc predicated all off", out, p.multiProcessorCount, 0x0u);
  return 0;
}
