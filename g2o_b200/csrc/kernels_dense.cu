// Dense FP64 Cholesky path for small reduced camera systems: the device-side counterpart of LinearSolverDense
// (g2o/solvers/dense/linear_solver_dense.h:65-115): copy the upper blocks of the sparse block matrix into a dense symmetric matrix,
// factorise, solve.  The reference uses Eigen::LDLT and refuses non-positive matrices (`isPositive()`); here a blocked right-looking
// LL^T is used (identical solution for SPD systems up to rounding) and a non-positive / non-finite pivot raises the info flag, which
// api.cu maps to "solve() returned false" exactly like the reference.
//   potrf_diag_kernel   64 x 64 diagonal block, one CTA, in shared memory
//   trsm_panel_kernel   L21 = A21 L11^-T, thread per row, L11 broadcast from shared memory
//   syrk_dmma_kernel    A22 -= L21 L21^T on the FP64 tensor pipe (mma.sync.m8n8k4.f64): 128 x 128 tile per CTA, 8 warps x (64 x 32),
//                       panel staged [k][row] with a row stride of 132 doubles (conflict-free 64-bit fragment loads)
//   trsv_*              blocked forward / backward substitution
// H is column-major n x n with leading dimension n; only the lower triangle is referenced after assembly.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "kernels.hpp"

namespace g2ocu {

namespace {

constexpr int NB = 64;                    // panel width
constexpr int TS = 128;                   // trailing-update tile
constexpr int LDS_ = TS + 4;              // shared row stride of a staged panel column

template <int P> __global__ void dense_assemble_kernel(PcgDev p, double* __restrict__ H, int n) {
  constexpr int PP = P * P;
  const int row = blockIdx.x;             // block row of the upper CSR
  for (int k = p.rowPtr[row] + (threadIdx.x / PP); k < p.rowPtr[row + 1]; k += blockDim.x / PP) {
    const int el = threadIdx.x % PP, r = el % P, c = el / P, col = p.colIdx[k];
    double v = p.A[(size_t)k * PP + el];
    if (col == row && r == c) v += p.lambda;
    const size_t gi = (size_t)row * P + r, gj = (size_t)col * P + c;
    H[gi + gj * n] = v;
    if (col != row) H[gj + gi * n] = v;
  }
}

__global__ void __launch_bounds__(256) potrf_diag_kernel(double* __restrict__ H, int n, int k0, int nb, int* info) {
  __shared__ double sA[NB][NB + 1];
  const int tid = threadIdx.x;
  for (int t = tid; t < nb * nb; t += 256) { const int r = t % nb, c = t / nb; sA[r][c] = H[(size_t)(k0 + r) + (size_t)(k0 + c) * n]; }
  __syncthreads();
  for (int j = 0; j < nb; ++j) {
    const double d = sA[j][j];
    if (!(d > 0.0) || !isfinite(d)) { if (tid == 0) atomicExch(info, 1); }
    const double sd = (d > 0.0) ? sqrt(d) : 1.0;
    __syncthreads();
    if (tid == 0) sA[j][j] = sd;
    for (int r = j + 1 + tid; r < nb; r += 256) sA[r][j] /= sd;
    __syncthreads();
    const int m = nb - j - 1;             // trailing update of the lower triangle
    for (int t = tid; t < m * m; t += 256) { const int r = j + 1 + t % m, c = j + 1 + t / m; if (r >= c) sA[r][c] -= sA[r][j] * sA[c][j]; }
    __syncthreads();
  }
  for (int t = tid; t < nb * nb; t += 256) { const int r = t % nb, c = t / nb; if (r >= c) H[(size_t)(k0 + r) + (size_t)(k0 + c) * n] = sA[r][c]; }
}

// rows k0+nb .. n-1 of the panel: X L11^T = A21  (forward substitution along the row; a short last panel is padded with the identity)
__global__ void __launch_bounds__(128) trsm_panel_kernel(double* __restrict__ H, int n, int k0, int nb) {
  __shared__ double sL[NB][NB + 1];
  for (int t = threadIdx.x; t < NB * NB; t += 128) {
    const int r = t % NB, c = t / NB;
    sL[r][c] = (r < nb && c < nb) ? ((r >= c) ? H[(size_t)(k0 + r) + (size_t)(k0 + c) * n] : 0.0) : (r == c ? 1.0 : 0.0);
  }
  __syncthreads();
  const int row = k0 + nb + blockIdx.x * 128 + threadIdx.x;
  if (row >= n) return;
  double x[NB];
#pragma unroll
  for (int j = 0; j < NB; ++j) x[j] = j < nb ? H[(size_t)row + (size_t)(k0 + j) * n] : 0.0;
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    x[j] /= sL[j][j];
#pragma unroll
    for (int q = j + 1; q < NB; ++q) x[q] -= x[j] * sL[q][j];
  }
#pragma unroll
  for (int j = 0; j < NB; ++j) if (j < nb) H[(size_t)row + (size_t)(k0 + j) * n] = x[j];
}

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// A22(tile) -= L21(rows of the tile) L21(columns of the tile)^T, lower tiles only (blockIdx.y >= blockIdx.x)
__global__ void __launch_bounds__(256) syrk_dmma_kernel(double* __restrict__ H, int n, int k0, int nb) {
  if (blockIdx.y < blockIdx.x) return;
  extern __shared__ double smem[];
  double* sA = smem;                      // [nb][LDS_] rows of the tile
  double* sB = smem + NB * LDS_;          // [nb][LDS_] columns of the tile
  const int t0 = k0 + nb;
  const int r0 = t0 + blockIdx.y * TS, c0 = t0 + blockIdx.x * TS;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int t = tid; t < nb * TS; t += 256) {
    const int r = t % TS, k = t / TS;
    sA[k * LDS_ + r] = (r0 + r < n) ? H[(size_t)(r0 + r) + (size_t)(k0 + k) * n] : 0.0;
    sB[k * LDS_ + r] = (c0 + r < n) ? H[(size_t)(c0 + r) + (size_t)(k0 + k) * n] : 0.0;
  }
  __syncthreads();
  const int wr = (warp >> 2) * 64, wc = (warp & 3) * 32;      // warp sub-tile: 64 rows x 32 columns
  const int m = lane >> 2, kq = lane & 3;
  double C[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { C[i][j][0] = 0; C[i][j][1] = 0; }
  for (int kk = 0; kk < nb; kk += 4) {
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = sA[(kk + kq) * LDS_ + wr + 8 * i + m];
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = sB[(kk + kq) * LDS_ + wc + 8 * j + m];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dmma(C[i][j], a[i], b[j]);
  }
  const bool diagTile = blockIdx.x == blockIdx.y;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gr = r0 + wr + 8 * i + m;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int gc = c0 + wc + 8 * j + 2 * kq + h;
        if (gr < n && gc < n && (!diagTile || gr >= gc)) H[(size_t)gr + (size_t)gc * n] -= C[i][j][h];
      }
    }
}

// forward: y_k = L11^-1 b_k (one CTA), then b[below] -= L21 y_k (thread per row)
__global__ void __launch_bounds__(NB) trsv_diag_fwd_kernel(const double* __restrict__ H, int n, int k0, int nb, double* __restrict__ x) {
  __shared__ double sx[NB];
  const int t = threadIdx.x;
  if (t < nb) sx[t] = x[k0 + t];
  __syncthreads();
  for (int j = 0; j < nb; ++j) {
    if (t == j) sx[j] /= H[(size_t)(k0 + j) + (size_t)(k0 + j) * n];
    __syncthreads();
    if (t > j && t < nb) sx[t] -= H[(size_t)(k0 + t) + (size_t)(k0 + j) * n] * sx[j];
    __syncthreads();
  }
  if (t < nb) x[k0 + t] = sx[t];
}
__global__ void __launch_bounds__(128) trsv_update_fwd_kernel(const double* __restrict__ H, int n, int k0, int nb, double* __restrict__ x) {
  __shared__ double sx[NB];
  if (threadIdx.x < nb) sx[threadIdx.x] = x[k0 + threadIdx.x];
  __syncthreads();
  const int row = k0 + nb + blockIdx.x * 128 + threadIdx.x;
  if (row >= n) return;
  double v = 0;
  for (int j = 0; j < nb; ++j) v += H[(size_t)row + (size_t)(k0 + j) * n] * sx[j];
  x[row] -= v;
}
// backward: x_k = L11^-T (y_k - L21^T x_below): column dot products (CTA per panel column), then the transposed diagonal solve
__global__ void __launch_bounds__(256) trsv_update_bwd_kernel(const double* __restrict__ H, int n, int k0, int nb, double* __restrict__ x) {
  __shared__ double sm[8];
  const int j = blockIdx.x;
  double v = 0;
  for (int row = k0 + nb + threadIdx.x; row < n; row += 256) v += H[(size_t)row + (size_t)(k0 + j) * n] * x[row];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) { double r = 0; for (int q = 0; q < 8; ++q) r += sm[q]; x[k0 + j] -= r; }
}
__global__ void __launch_bounds__(NB) trsv_diag_bwd_kernel(const double* __restrict__ H, int n, int k0, int nb, double* __restrict__ x) {
  __shared__ double sx[NB];
  const int t = threadIdx.x;
  if (t < nb) sx[t] = x[k0 + t];
  __syncthreads();
  for (int j = nb - 1; j >= 0; --j) {
    if (t == j) sx[j] /= H[(size_t)(k0 + j) + (size_t)(k0 + j) * n];
    __syncthreads();
    if (t < j) sx[t] -= H[(size_t)(k0 + j) + (size_t)(k0 + t) * n] * sx[j];     // L^T(t, j) = L(j, t)
    __syncthreads();
  }
  if (t < nb) x[k0 + t] = sx[t];
}

}  // namespace

void launchDenseAssemble(const PcgDev& p, double* H, cudaStream_t st, int64_t* launches) {
  const int n = p.n;
  cudaMemsetAsync(H, 0, sizeof(double) * (size_t)n * n, st);
  switch (p.P) {
    case 3: dense_assemble_kernel<3><<<p.nb, 9 * 28, 0, st>>>(p, H, n); break;
    case 6: dense_assemble_kernel<6><<<p.nb, 36 * 7, 0, st>>>(p, H, n); break;
    case 9: dense_assemble_kernel<9><<<p.nb, 81 * 3, 0, st>>>(p, H, n); break;
    default: break;
  }
  *launches += 1;
}

// factorise H = L L^T in place (lower), then x = H^-1 b.  *info (device int, zeroed here) becomes non-zero when a pivot is not positive.
int launchDenseCholeskySolve(double* H, int n, const double* b, double* x, int* info, cudaStream_t st, int64_t* launches) {
  static bool configured = false;
  constexpr int kSyrkSmem = 2 * NB * LDS_ * (int)sizeof(double);
  if (!configured) { cudaFuncSetAttribute(syrk_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSyrkSmem); configured = true; }
  cudaMemsetAsync(info, 0, sizeof(int), st);
  for (int k0 = 0; k0 < n; k0 += NB) {
    const int nb = n - k0 < NB ? n - k0 : NB, rem = n - k0 - nb;
    potrf_diag_kernel<<<1, 256, 0, st>>>(H, n, k0, nb, info); *launches += 1;
    if (rem > 0) {
      trsm_panel_kernel<<<(rem + 127) / 128, 128, 0, st>>>(H, n, k0, nb);
      const int tiles = (rem + TS - 1) / TS;
      syrk_dmma_kernel<<<dim3(tiles, tiles), 256, kSyrkSmem, st>>>(H, n, k0, nb);
      *launches += 2;
    }
  }
  if (x != b) cudaMemcpyAsync(x, b, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, st);
  for (int k0 = 0; k0 < n; k0 += NB) {
    const int nb = n - k0 < NB ? n - k0 : NB, rem = n - k0 - nb;
    trsv_diag_fwd_kernel<<<1, NB, 0, st>>>(H, n, k0, nb, x); *launches += 1;
    if (rem > 0) { trsv_update_fwd_kernel<<<(rem + 127) / 128, 128, 0, st>>>(H, n, k0, nb, x); *launches += 1; }
  }
  for (int k0 = ((n - 1) / NB) * NB; k0 >= 0; k0 -= NB) {
    const int nb = n - k0 < NB ? n - k0 : NB, rem = n - k0 - nb;
    if (rem > 0) { trsv_update_bwd_kernel<<<nb, 256, 0, st>>>(H, n, k0, nb, x); *launches += 1; }
    trsv_diag_bwd_kernel<<<1, NB, 0, st>>>(H, n, k0, nb, x); *launches += 1;
  }
  return 0;
}

}  // namespace g2ocu
