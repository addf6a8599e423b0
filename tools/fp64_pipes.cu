// Microbenchmark: FP64 FMA pipe vs FP64 tensor (DMMA) throughput on the device it runs on.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_pipes fp64_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC> __global__ void dmma884_kernel(double* out, int iters, double a, double b) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC> __global__ void dmma16816_kernel(double* out, int iters, double a, double b) {
  double c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%4,%4,%4,%4,%4,%4,%4}, {%5,%5,%5,%5}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA fed from shared memory: every k-step loads MI A-fragments and NJ B-fragments (LDS.64 per lane) for MI*NJ mma
template <int MI, int NJ> __global__ void dmma_smem_kernel(double* out, int iters) {
  __shared__ double sA[8][MI * 32 + 8], sB[8][NJ * 32 + 8];
  for (int t = threadIdx.x; t < 8 * (MI * 32 + 8); t += blockDim.x) (&sA[0][0])[t] = 1e-3 * t;
  for (int t = threadIdx.x; t < 8 * (NJ * 32 + 8); t += blockDim.x) (&sB[0][0])[t] = 1e-3 * t;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  double c[MI][NJ][2];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) { c[i][j][0] = 0; c[i][j][1] = 0; }
  for (int it = 0; it < iters; ++it) {
    const int s = it & 7;
    double a[MI], b[NJ];
#pragma unroll
    for (int i = 0; i < MI; ++i) a[i] = sA[s][i * 32 + lane];
#pragma unroll
    for (int j = 0; j < NJ; ++j) b[j] = sB[s][j * 32 + lane];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
      for (int j = 0; j < NJ; ++j)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][j][0]), "+d"(c[i][j][1]) : "d"(a[i]), "d"(b[j]));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) s += c[i][j][0] + c[i][j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F> static float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  const int iters = 20000;
  for (int warps : {4, 8, 16, 32}) {
    const int nt = warps * 32, nb = sms * (warps >= 16 ? 2 : 4) / (warps >= 16 ? 2 : 1);
    float ms = timeit([&] { dfma_kernel<<<nb, nt>>>(out, iters, 1.0000001, 1e-9); });
    printf("DFMA      warps/CTA %2d CTAs %4d : %.2f TFLOP/s\n", warps, nb, 2.0 * 16 * iters * (double)nb * nt / ms * 1e-9);
    ms = timeit([&] { dmma884_kernel<8><<<nb, nt>>>(out, iters, 1.0000001, 1e-9); });
    printf("DMMA 8x8x4   (8 acc) warps/CTA %2d : %.2f TFLOP/s\n", warps, 2.0 * 256 * 8 * iters * (double)nb * warps / ms * 1e-9);
    ms = timeit([&] { dmma884_kernel<16><<<nb, nt>>>(out, iters, 1.0000001, 1e-9); });
    printf("DMMA 8x8x4  (16 acc) warps/CTA %2d : %.2f TFLOP/s\n", warps, 2.0 * 256 * 16 * iters * (double)nb * warps / ms * 1e-9);
    ms = timeit([&] { dmma16816_kernel<4><<<nb, nt>>>(out, iters / 4, 1.0000001, 1e-9); });
    printf("DMMA 16x8x16 (4 acc) warps/CTA %2d : %.2f TFLOP/s\n", warps, 2.0 * 2048 * 4 * (iters / 4) * (double)nb * warps / ms * 1e-9);
    ms = timeit([&] { dmma_smem_kernel<3, 3><<<nb, nt>>>(out, iters); });
    printf("DMMA 8x8x4 smem 3x3  warps/CTA %2d : %.2f TFLOP/s\n", warps, 2.0 * 256 * 9 * iters * (double)nb * warps / ms * 1e-9);
    ms = timeit([&] { dmma_smem_kernel<2, 4><<<nb, nt>>>(out, iters); });
    printf("DMMA 8x8x4 smem 2x4  warps/CTA %2d : %.2f TFLOP/s\n", warps, 2.0 * 256 * 8 * iters * (double)nb * warps / ms * 1e-9);
    ms = timeit([&] { dmma_smem_kernel<1, 1><<<nb, nt>>>(out, iters); });
    printf("DMMA 8x8x4 smem 1x1  warps/CTA %2d : %.2f TFLOP/s\n", warps, 2.0 * 256 * 1 * iters * (double)nb * warps / ms * 1e-9);
  }
  printf("SMs %d, clock %d kHz\n", sms, p.clockRate);
  return 0;
}
