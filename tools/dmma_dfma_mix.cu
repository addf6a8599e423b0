// Do DMMA (tensor pipe) and DFMA (FP64 pipe) overlap on this GPU?  Each warp issues R dfma per dmma.
#include <cstdio>
#include <cuda_runtime.h>
template <int R> __global__ void k(double* out, int iters, double a, double b) {
  double c[8][2], x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; x[i] = i + threadIdx.x; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
#pragma unroll
      for (int r = 0; r < R; ++r) x[(i + r) & 7] = fma(x[(i + r) & 7], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int R> void run(double* out, int sms) {
  const int iters = 20000, nt = 256, nb = sms * 2;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<R><<<nb, nt>>>(out, iters, 1.0000001, 1e-9); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<R><<<nb, nt>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double warps = (double)nb * nt / 32, dm = 8.0 * iters * warps, df = dm * R;
  printf("dfma per dmma %d : DMMA %.2f TF + DFMA %.2f TF = %.2f TF\n", R, 2 * 256 * dm / ms * 1e-9, 2 * 32 * df / ms * 1e-9, (2 * 256 * dm + 2 * 32 * df) / ms * 1e-9);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double* out; cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 2 * 256);
  run<0>(out, p.multiProcessorCount); run<1>(out, p.multiProcessorCount); run<2>(out, p.multiProcessorCount); run<4>(out, p.multiProcessorCount); run<8>(out, p.multiProcessorCount);
  return 0;
}
