// Dense FP64 Cholesky path for small reduced camera systems: the device-side counterpart of LinearSolverDense
// (g2o/solvers/dense/linear_solver_dense.h:65-115): copy the upper blocks of the sparse block matrix into a dense symmetric matrix,
// factorise, solve.  The reference uses Eigen::LDLT and refuses non-positive matrices (`isPositive()`); here a blocked right-looking
// LL^T is used (identical solution for SPD systems up to rounding) and a non-positive / non-finite pivot raises the info flag, which
// api.cu maps to "solve() returned false" exactly like the reference.
//   potrf_diag_kernel   64 x 64 diagonal block, one CTA, in shared memory
//   trsm_panel_kernel   L21 = A21 L11^-T, thread per row, L11 broadcast from shared memory
//   syrk_dmma_kernel    C -= L L^T on the FP64 tensor pipe (mma.sync.m8n8k4.f64): 128 x 128 tile per CTA, 8 warps x (64 x 32), the panels
//                       streamed through a 2-stage cp.async ring, staged [k][row] with a row stride of 132 doubles (conflict-free 64-bit
//                       fragment loads).  Two-level blocking: 64-wide panels inside 512-wide outer panels, so that the trailing matrix is
//                       read and written once per 512 columns
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
    // trailing update of the lower triangle, 16 x 16 thread grid striding over rows / columns (no integer division in the loop)
    for (int c = j + 1 + (tid >> 4); c < nb; c += 16) {
      const double lc = sA[c][j];
      for (int r = c + (tid & 15); r < nb; r += 16) sA[r][c] -= sA[r][j] * lc;
    }
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

// H[r, c] -= sum_k L[r, kBegin + k] L[c, kBegin + k] for the lower tiles of rows >= colBegin, columns in [colBegin, colEnd):
// 128 x 128 tile per CTA, the two 128 x kLen panels streamed in chunks of KC columns through a 2-stage cp.async ring.
// Used twice per outer panel: K = 64 inside the panel, K = up to 512 for the trailing matrix (one read-modify-write of C per 512 columns).
constexpr int KC = 32;
__device__ __forceinline__ void cpAsync8(double* dstSmem, const double* src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dstSmem);
  const int sz = valid ? 8 : 0;                       // src-size 0: zero fill
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__global__ void __launch_bounds__(256) syrk_dmma_kernel(double* __restrict__ H, int n, int kBegin, int kLen, int colBegin, int colEnd) {
  if (blockIdx.y < blockIdx.x) return;
  extern __shared__ double smem[];                    // [2 stages][A | B][KC][LDS_]
  const int r0 = colBegin + blockIdx.y * TS, c0 = colBegin + blockIdx.x * TS;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nChunks = (kLen + KC - 1) / KC;
  auto load = [&](int c, int stage) {
    double* sA = smem + (size_t)stage * 2 * KC * LDS_; double* sB = sA + KC * LDS_;
    for (int t = tid; t < KC * TS; t += 256) {
      const int r = t % TS, k = t / TS, kk = c * KC + k;
      const bool kOk = kk < kLen;
      const size_t col = (size_t)(kBegin + (kOk ? kk : 0)) * n;
      const bool aOk = kOk && r0 + r < n, bOk = kOk && c0 + r < n;
      cpAsync8(sA + k * LDS_ + r, H + (aOk ? (size_t)(r0 + r) : 0) + col, aOk);
      cpAsync8(sB + k * LDS_ + r, H + (bOk ? (size_t)(c0 + r) : 0) + col, bOk);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  const int wr = (warp >> 2) * 64, wc = (warp & 3) * 32;      // warp sub-tile: 64 rows x 32 columns
  const int m = lane >> 2, kq = lane & 3;
  double C[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { C[i][j][0] = 0; C[i][j][1] = 0; }
  load(0, 0);
  for (int c = 0; c < nChunks; ++c) {
    if (c + 1 < nChunks) { load(c + 1, (c + 1) & 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const double* sA = smem + (size_t)(c & 1) * 2 * KC * LDS_; const double* sB = sA + KC * LDS_;
#pragma unroll 2
    for (int kk = 0; kk < KC; kk += 4) {
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
    __syncthreads();                                  // the stage is refilled two iterations later
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
        if (gr < n && gc < colEnd && (!diagTile || gr >= gc)) H[(size_t)gr + (size_t)gc * n] -= C[i][j][h];
      }
    }
}

// forward: y_k = L11^-1 b_k (one CTA), then b[below] -= L21 y_k (thread per row)
__global__ void __launch_bounds__(256) trsv_diag_fwd_kernel(const double* __restrict__ H, int n, int k0, int nb, double* __restrict__ x, size_t ldx) {
  __shared__ double sL[NB][NB + 1];
  __shared__ double sx[NB];
  const int t = threadIdx.x;
  x += (size_t)blockIdx.x * ldx;                       // one CTA per right-hand side
  for (int q = t; q < nb * nb; q += 256) { const int r = q % nb, c = q / nb; if (r >= c) sL[r][c] = H[(size_t)(k0 + r) + (size_t)(k0 + c) * n]; }
  if (t < nb) sx[t] = x[k0 + t];
  __syncthreads();
  if (t < 32) {                                        // one warp: lane owns rows lane and lane + 32
    for (int j = 0; j < nb; ++j) {
      const double xj = sx[j] / sL[j][j];
      __syncwarp();
      if (t == 0) sx[j] = xj;
      for (int r = j + 1 + t; r < nb; r += 32) sx[r] -= sL[r][j] * xj;
      __syncwarp();
    }
  }
  __syncthreads();
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
__global__ void __launch_bounds__(256) trsv_diag_bwd_kernel(const double* __restrict__ H, int n, int k0, int nb, double* __restrict__ x, size_t ldx) {
  __shared__ double sL[NB][NB + 1];
  __shared__ double sx[NB];
  const int t = threadIdx.x;
  x += (size_t)blockIdx.x * ldx;
  for (int q = t; q < nb * nb; q += 256) { const int r = q % nb, c = q / nb; if (r >= c) sL[r][c] = H[(size_t)(k0 + r) + (size_t)(k0 + c) * n]; }
  if (t < nb) sx[t] = x[k0 + t];
  __syncthreads();
  if (t < 32) {
    for (int j = nb - 1; j >= 0; --j) {
      const double xj = sx[j] / sL[j][j];
      __syncwarp();
      if (t == 0) sx[j] = xj;
      for (int r = t; r < j; r += 32) sx[r] -= sL[j][r] * xj;      // L^T(r, j) = L(j, r)
      __syncwarp();
    }
  }
  __syncthreads();
  if (t < nb) x[k0 + t] = sx[t];
}

// ---- several right-hand sides at once (blocks of the inverse, computeMarginals) ----
// X is n x nrhs, column-major with leading dimension ldx.  A CTA of the update kernels serves RC right-hand sides, so that the panel of
// L is read once per RC columns: forward, a thread keeps its row of the panel in registers and walks the RC columns staged in shared
// memory; backward, a CTA owns one panel column and RC dot products over the rows below the panel.
constexpr int RC = 16;
__global__ void __launch_bounds__(128) trsm_update_fwd_kernel(const double* __restrict__ H, int n, int k0, int nb, double* __restrict__ X, size_t ldx, int nrhs) {
  __shared__ double sx[RC][NB];
  const int c0 = blockIdx.y * RC, nc = min(RC, nrhs - c0);
  for (int t = threadIdx.x; t < RC * NB; t += 128) { const int c = t / NB, j = t % NB; sx[c][j] = (c < nc && j < nb) ? X[(size_t)(c0 + c) * ldx + k0 + j] : 0.0; }
  __syncthreads();
  const int row = k0 + nb + blockIdx.x * 128 + threadIdx.x;
  if (row >= n) return;
  double l[NB];
#pragma unroll
  for (int j = 0; j < NB; ++j) l[j] = j < nb ? H[(size_t)row + (size_t)(k0 + j) * n] : 0.0;
  for (int c = 0; c < nc; ++c) {
    double v = 0;
#pragma unroll
    for (int j = 0; j < NB; ++j) v += l[j] * sx[c][j];
    X[(size_t)(c0 + c) * ldx + row] -= v;
  }
}
__global__ void __launch_bounds__(256) trsm_update_bwd_kernel(const double* __restrict__ H, int n, int k0, int nb, double* __restrict__ X, size_t ldx, int nrhs) {
  __shared__ double sm[RC][8];
  const int j = blockIdx.x, c0 = blockIdx.y * RC, nc = min(RC, nrhs - c0);
  double acc[RC];
#pragma unroll
  for (int c = 0; c < RC; ++c) acc[c] = 0;
  for (int row = k0 + nb + threadIdx.x; row < n; row += 256) {
    const double l = H[(size_t)row + (size_t)(k0 + j) * n];
#pragma unroll
    for (int c = 0; c < RC; ++c) if (c < nc) acc[c] += l * X[(size_t)(c0 + c) * ldx + row];
  }
#pragma unroll
  for (int c = 0; c < RC; ++c) {
    double v = acc[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sm[c][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < nc) { double r = 0; for (int q = 0; q < 8; ++q) r += sm[threadIdx.x][q]; X[(size_t)(c0 + threadIdx.x) * ldx + k0 + j] -= r; }
}
// column c of X = unit vector of scalar index colScalar[c]
__global__ void unit_columns_kernel(double* __restrict__ X, size_t ldx, const int32_t* __restrict__ colScalar, int nCols) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < nCols) X[(size_t)c * ldx + (size_t)colScalar[c]] = 1.0;
}
// out block i (rowDim x colDim, column-major, at out + outOff[i]) = rows [rowScalar[i], + rowDim[i]) of the columns [colStart[i], + colDim[i]) of X
__global__ void gather_blocks_kernel(const double* __restrict__ X, size_t ldx, const int32_t* __restrict__ rowScalar, const int32_t* __restrict__ rowDim,
                                     const int32_t* __restrict__ colStart, const int32_t* __restrict__ colDim, const int64_t* __restrict__ outOff, double* __restrict__ out) {
  const int i = blockIdx.x, nr = rowDim[i], nc = colDim[i];
  for (int el = threadIdx.x; el < nr * nc; el += blockDim.x) {
    const int r = el % nr, c = el / nr;
    out[outOff[i] + el] = X[(size_t)(colStart[i] + c) * ldx + (size_t)rowScalar[i] + r];
  }
}
// the point blocks and the pose-point blocks of a whole system [Hpp Hpl; Hpl^T Hll] + lambda I into the dense matrix (internal order: poses, then points)
__global__ void dense_assemble_hll_kernel(const double* __restrict__ Hll, int numLandmarks, int L, int np, double lambda, double* __restrict__ H, int n) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= numLandmarks * L * L) return;
  const int l = t / (L * L), el = t - l * L * L, r = el % L, c = el / L;
  H[(size_t)(np + l * L + r) + (size_t)(np + l * L + c) * n] = Hll[t] + (r == c ? lambda : 0.0);
}
__global__ void dense_assemble_hpl_kernel(const double* __restrict__ Hpl, const int32_t* __restrict__ hplRow, const int32_t* __restrict__ hplLm, int nBlocks, int P, int L, int np,
                                          double* __restrict__ H, int n) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)nBlocks * P * L) return;
  const int k = (int)(t / (P * L)), el = (int)(t - (int64_t)k * P * L), r = el % P, c = el / P;
  const size_t gi = (size_t)hplRow[k] * P + r, gj = (size_t)np + (size_t)hplLm[k] * L + c;
  const double v = Hpl[t];
  H[gi + gj * n] = v; H[gj + gi * n] = v;
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

// factorise H = L L^T in place (lower).  *info (device int, zeroed here) becomes non-zero when a pivot is not positive.
int launchDenseCholeskyFactor(double* H, int n, int* info, cudaStream_t st, int64_t* launches) {
  constexpr int kSyrkSmem = 2 * 2 * KC * LDS_ * (int)sizeof(double);
  constexpr int NO = 512;                             // outer panel: the trailing matrix is updated once per NO columns
  cudaFuncSetAttribute(syrk_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSyrkSmem);   // per device: set on every call (solvers may live on several GPUs of one process)
  cudaMemsetAsync(info, 0, sizeof(int), st);
  for (int K0 = 0; K0 < n; K0 += NO) {
    const int pend = K0 + NO < n ? K0 + NO : n;
    for (int k0 = K0; k0 < pend; k0 += NB) {
      const int nb = pend - k0 < NB ? pend - k0 : NB, rem = n - k0 - nb;
      potrf_diag_kernel<<<1, 256, 0, st>>>(H, n, k0, nb, info); *launches += 1;
      if (rem > 0) {
        trsm_panel_kernel<<<(rem + 127) / 128, 128, 0, st>>>(H, n, k0, nb); *launches += 1;
        const int cb = k0 + nb;                       // inside the outer panel: only its remaining columns are updated now
        if (cb < pend) {
          syrk_dmma_kernel<<<dim3((pend - cb + TS - 1) / TS, (n - cb + TS - 1) / TS), 256, kSyrkSmem, st>>>(H, n, k0, nb, cb, pend);
          *launches += 1;
        }
      }
    }
    if (pend < n) {                                   // trailing matrix -= L[pend:, K0:pend] L[pend:, K0:pend]^T
      const int tiles = (n - pend + TS - 1) / TS;
      syrk_dmma_kernel<<<dim3(tiles, tiles), 256, kSyrkSmem, st>>>(H, n, K0, pend - K0, pend, n);
      *launches += 1;
    }
  }
  return 0;
}

// ... then x = H^-1 b
int launchDenseCholeskySolve(double* H, int n, const double* b, double* x, int* info, cudaStream_t st, int64_t* launches) {
  launchDenseCholeskyFactor(H, n, info, st, launches);
  if (x != b) cudaMemcpyAsync(x, b, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, st);
  for (int k0 = 0; k0 < n; k0 += NB) {
    const int nb = n - k0 < NB ? n - k0 : NB, rem = n - k0 - nb;
    trsv_diag_fwd_kernel<<<1, 256, 0, st>>>(H, n, k0, nb, x, 0); *launches += 1;
    if (rem > 0) { trsv_update_fwd_kernel<<<(rem + 127) / 128, 128, 0, st>>>(H, n, k0, nb, x); *launches += 1; }
  }
  for (int k0 = ((n - 1) / NB) * NB; k0 >= 0; k0 -= NB) {
    const int nb = n - k0 < NB ? n - k0 : NB, rem = n - k0 - nb;
    if (rem > 0) { trsv_update_bwd_kernel<<<nb, 256, 0, st>>>(H, n, k0, nb, x); *launches += 1; }
    trsv_diag_bwd_kernel<<<1, 256, 0, st>>>(H, n, k0, nb, x, 0); *launches += 1;
  }
  return 0;
}

// X <- (L L^T)^-1 X for nrhs columns.  Rows above `firstNonZero` are zero in every column on entry (the forward sweep starts at that
// panel); rows above `firstNeeded` of the result are not needed (the backward sweep stops there).
void launchDenseSolveMany(const double* H, int n, double* X, size_t ldx, int nrhs, int firstNonZero, int firstNeeded, cudaStream_t st, int64_t* launches) {
  const int yb = (nrhs + RC - 1) / RC;
  for (int k0 = (firstNonZero / NB) * NB; k0 < n; k0 += NB) {
    const int nb = n - k0 < NB ? n - k0 : NB, rem = n - k0 - nb;
    trsv_diag_fwd_kernel<<<nrhs, 256, 0, st>>>(H, n, k0, nb, X, ldx); *launches += 1;
    if (rem > 0) { trsm_update_fwd_kernel<<<dim3((rem + 127) / 128, yb), 128, 0, st>>>(H, n, k0, nb, X, ldx, nrhs); *launches += 1; }
  }
  for (int k0 = ((n - 1) / NB) * NB; k0 >= (firstNeeded / NB) * NB; k0 -= NB) {
    const int nb = n - k0 < NB ? n - k0 : NB, rem = n - k0 - nb;
    if (rem > 0) { trsm_update_bwd_kernel<<<dim3(nb, yb), 256, 0, st>>>(H, n, k0, nb, X, ldx, nrhs); *launches += 1; }
    trsv_diag_bwd_kernel<<<nrhs, 256, 0, st>>>(H, n, k0, nb, X, ldx); *launches += 1;
  }
}
void launchUnitColumns(double* X, size_t ldx, const int32_t* colScalar, int nCols, cudaStream_t st, int64_t* launches) {
  cudaMemsetAsync(X, 0, sizeof(double) * ldx * (size_t)nCols, st);
  unit_columns_kernel<<<(nCols + 127) / 128, 128, 0, st>>>(X, ldx, colScalar, nCols); *launches += 1;
}
void launchGatherBlocks(const double* X, size_t ldx, const int32_t* rowScalar, const int32_t* rowDim, const int32_t* colStart, const int32_t* colDim, const int64_t* outOff, int nPairs,
                        double* out, cudaStream_t st, int64_t* launches) {
  if (nPairs <= 0) return;
  gather_blocks_kernel<<<nPairs, 96, 0, st>>>(X, ldx, rowScalar, rowDim, colStart, colDim, outOff, out); *launches += 1;
}
// dense copy of the whole system of a graph whose points are not marginalized: hpp = the PCG view of Hpp (its lambda goes on every diagonal)
void launchDenseAssembleFull(const PcgDev& hpp, const double* Hll, const double* Hpl, const int32_t* hplRow, const int32_t* hplLm, int nHplBlocks, int numLandmarks, int L, double* H, int n,
                             cudaStream_t st, int64_t* launches) {
  cudaMemsetAsync(H, 0, sizeof(double) * (size_t)n * n, st);
  switch (hpp.P) {
    case 3: dense_assemble_kernel<3><<<hpp.nb, 9 * 28, 0, st>>>(hpp, H, n); break;
    case 6: dense_assemble_kernel<6><<<hpp.nb, 36 * 7, 0, st>>>(hpp, H, n); break;
    case 9: dense_assemble_kernel<9><<<hpp.nb, 81 * 3, 0, st>>>(hpp, H, n); break;
    default: break;
  }
  if (numLandmarks > 0) dense_assemble_hll_kernel<<<(numLandmarks * L * L + 255) / 256, 256, 0, st>>>(Hll, numLandmarks, L, hpp.n, hpp.lambda, H, n);
  if (nHplBlocks > 0) dense_assemble_hpl_kernel<<<(unsigned)(((int64_t)nHplBlocks * hpp.P * L + 255) / 256), 256, 0, st>>>(Hpl, hplRow, hplLm, nHplBlocks, hpp.P, L, hpp.n, H, n);
  *launches += 3;
}

}  // namespace g2ocu
