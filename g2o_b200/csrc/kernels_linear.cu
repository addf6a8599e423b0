// sm_100a kernels for the linear-algebra half of the LM inner loop:
//   schur_*            BlockSolver::solve Schur part: Dinv = (Hll+λI)^-1, S = Hpp+λI - Σ_l B Dinv B^T, b_s      (block_solver.hpp:333-405)
//   backsub_kernel     x_l = Dinv (b_l - Hpl^T x_p)                                                            (block_solver.hpp:420-444)
//   block_inverse      block-Jacobi preconditioner M^-1 = diag-block inverses                                    (linear_solver_pcg.hpp:93-95)
//   spmv_sym_kernel    q = A d using the upper blocks twice                                                      (linear_solver_pcg.hpp:179-197)
//   pcg_* kernels      the CG recurrences with device-resident scalars                                           (linear_solver_pcg.hpp:112-150)
//   maxdiag / scale    computeLambdaInit / computeScale reductions                                               (optimization_algorithm_levenberg.cpp:152-184)
// λ is never written into the stored diagonals: every consumer adds it on the fly, so setLambda/restoreDiagonal
// (block_solver.hpp:525-565) cost nothing and need no backup copies.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>

#include "kernels.hpp"

namespace g2ocu {

#define G2D __device__ __forceinline__

G2D double warpSumL(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}
template <int NT> G2D double blockSumL(double v, double* sm) {
  v = warpSumL(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sm[wid] = v;
  __syncthreads();
  double r = 0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NT / 32; ++k) r += sm[k];
  }
  return r;
}
// every thread of the CTA gets the sum of partial[0..n) (fixed order => identical in all CTAs)
template <int NT> G2D double sumPartialsAll(const double* partial, int n, double* sm) {
  double v = 0;
  for (int k = threadIdx.x; k < n; k += NT) v += partial[k];
  v = warpSumL(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sm[wid] = v;
  __syncthreads();
  double r = 0;
#pragma unroll
  for (int k = 0; k < NT / 32; ++k) r += sm[k];
  __syncthreads();
  return r;
}

// closed-form inverses, as Eigen does for fixed sizes <= 4 (block_solver.hpp:350 Hll^-1)
template <int L> G2D void invSmall(const double* A, double* X);
template <> G2D void invSmall<2>(const double* A, double* X) {
  const double id = 1.0 / (A[0] * A[3] - A[2] * A[1]);
  X[0] = A[3] * id; X[1] = -A[1] * id; X[2] = -A[2] * id; X[3] = A[0] * id;
}
template <> G2D void invSmall<3>(const double* A, double* X) {
  const double c00 = A[4] * A[8] - A[7] * A[5], c10 = A[7] * A[2] - A[1] * A[8], c20 = A[1] * A[5] - A[4] * A[2];
  const double id = 1.0 / (c00 * A[0] + c10 * A[3] + c20 * A[6]);
  X[0] = c00 * id; X[1] = c10 * id; X[2] = c20 * id;
  X[3] = (A[6] * A[5] - A[3] * A[8]) * id; X[4] = (A[0] * A[8] - A[6] * A[2]) * id; X[5] = (A[3] * A[2] - A[0] * A[5]) * id;
  X[6] = (A[3] * A[7] - A[6] * A[4]) * id; X[7] = (A[6] * A[1] - A[0] * A[7]) * id; X[8] = (A[0] * A[4] - A[3] * A[1]) * id;
}


template <int P> __global__ void schur_init_kernel(SchurDev d, const double* Hpp, const double* b, double lambda) {
  constexpr int PP = P * P;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < (int64_t)d.nnzHpp * PP) {
    const int k = (int)(t / PP), el = (int)(t - (int64_t)k * PP);
    d.S[(size_t)d.hppToS[k] * PP + el] = Hpp[t];
  }
  if (t < (int64_t)d.numPoses * P) d.bschur[t] = b[t];
}
template <int P> __global__ void add_lambda_diag_kernel(double* A, const int32_t* diag, int nb, double lambda) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nb * P) return;
  const int i = t / P, k = t - i * P;
  A[(size_t)diag[i] * P * P + k * (P + 1)] += lambda;
}

// Dinv_l = (Hll_l + lambda I)^-1 and db_l = Dinv_l b_l, thread per landmark (block_solver.hpp:347-356)
template <int P, int L> __global__ void dinv_kernel(SchurDev d, const double* __restrict__ Hll, const double* __restrict__ b, double lambda) {
  constexpr int LL = L * L;
  const int lm = d.lmBegin + blockIdx.x * blockDim.x + threadIdx.x;
  if (lm >= d.lmEnd) return;
  double H[LL], X[LL], bl[L];
#pragma unroll
  for (int q = 0; q < LL; ++q) H[q] = Hll[(size_t)lm * LL + q];
#pragma unroll
  for (int q = 0; q < L; ++q) { H[q * (L + 1)] += lambda; bl[q] = b[(size_t)d.numPoses * P + (size_t)lm * L + q]; }
  invSmall<L>(H, X);
#pragma unroll
  for (int q = 0; q < LL; ++q) d.Dinv[(size_t)lm * LL + q] = X[q];
#pragma unroll
  for (int r = 0; r < L; ++r) { double v = 0;
#pragma unroll
    for (int c = 0; c < L; ++c) v += X[r + L * c] * bl[c];
    d.db[(size_t)lm * L + r] = v; }
}

// b_schur -= B_i (Dinv b_l) for every Hpl block (block_solver.hpp:366-374): thread per block, blocks staged through shared memory.
// The blocks of short tracks also leave as W = B Dinv (block_solver.hpp:366, BDinv) for the pair kernel: formed in place in the stage and
// written with coalesced stores at the block's compact index (d.hplShortIdx; 2.6 M of the 5.0 M blocks on C3).
template <int P, int L> __global__ void __launch_bounds__(128) coeff_kernel(SchurDev d, const double* __restrict__ Hpl, const int32_t* __restrict__ hplLm, int nBlocks) {
  constexpr int PLn = P * L, LL = L * L;
  __shared__ double sB[128 * PLn];
  __shared__ int32_t sW[128];
  const int tid = threadIdx.x, k0 = d.blockBegin + blockIdx.x * 128;
  const int nb = min(128, d.blockBegin + nBlocks - k0);
  const double* src = Hpl + (size_t)k0 * PLn;
  // this thread's block: camera, landmark, Dinv b_l, compact index and (short tracks) Dinv first - dependent global loads that then overlap the
  // load of the tile
  int wi = -1, ci = 0, lm = 0; double dbv[L], Di[LL];
  if (tid < nb) {
    const int k = k0 + tid; ci = d.hplRowIdx[k]; lm = hplLm[k];
    if (d.hplShortIdx) wi = d.hplShortIdx[k];
#pragma unroll
    for (int a = 0; a < L; ++a) dbv[a] = d.db[(size_t)lm * L + a];
    if (wi >= 0) {
#pragma unroll
      for (int a = 0; a < LL; ++a) Di[a] = d.Dinv[(size_t)lm * LL + a];
    }
  }
  for (int t = tid; t < nb * PLn; t += 128) sB[t] = src[t];
  __syncthreads();
  if (tid < nb) {
    double* blk = sB + tid * PLn;
#pragma unroll
    for (int r = 0; r < P; ++r) { double v = 0;
#pragma unroll
      for (int a = 0; a < L; ++a) v += blk[r + P * a] * dbv[a];
      atomicAdd(d.bschur + (size_t)ci * P + r, -v); }
    if (wi >= 0) {
#pragma unroll
      for (int r = 0; r < P; ++r) {
        double bv[L];
#pragma unroll
        for (int a = 0; a < L; ++a) bv[a] = blk[r + P * a];
#pragma unroll
        for (int a = 0; a < L; ++a) { double w = 0;
#pragma unroll
          for (int a2 = 0; a2 < L; ++a2) w += bv[a2] * Di[a2 + L * a];
          blk[r + P * a] = w; }
      }
    }
  }
  sW[tid] = wi;
  __syncthreads();
  if (!d.hplShortIdx) return;
  for (int t = tid; t < nb * PLn; t += 128) { const int bk = t / PLn, w = sW[bk]; if (w >= 0) d.Wshort[(size_t)w * PLn + (t - bk * PLn)] = sB[t]; }
}

__device__ __forceinline__ void pairDmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// Short tracks (fewer than kTileMinTrack observations): the pairs (block i, block j), i <= j, of every such landmark are sorted by target
// Hschur block on the host (the short tracks of a ring of cameras hit a narrow band of Hschur: 5.8 M pairs fall on 30 k blocks on C3) and cut
// into segments of <= kPairSegment pairs of one block.  Earlier forms of this kernel: a warp per pair with one RED per element per pair
// (round 1: 470 M REDs, bound by the L2 reduction rate, 1.55 ms on C3), then a warp per segment summing scalar products in registers
// (1.25 ms, 86 % of the L1 pipe: 18 loads + 12 FMAs + 9 shuffles per pair).
// Short tracks on the FP64 tensor pipe: the pairs of a segment all add into one block, Hschur(i,j) -= sum_p W_p B_p^T, which is one small GEMM
// with the pairs stacked along K (L scalars each, <= kPairSegment pairs: K <= 48).  One warp per segment; per K step of 4 a lane fetches one
// scalar of W (row m of the pair its K slot belongs to, from the coefficient pass) and one of B_j and issues the DMMA of rows / columns
// 0..7; for P = 9 row 8 and column 8 are three plain FMAs per lane on the ninth scalars of its K slot, summed over the 4 K lanes at the
// end.  Against the scalar kernel above: 4 loads + 1 DMMA + 3 FMAs per 4/3 pairs instead of 18 loads + 12 FMAs + 9 shuffles per pair (that kernel sat at 86 % of the L1
// pipe).  (One pair per K step - L scalars + zero padding, so that a step's lanes read one W and one B block instead of two - was measured:
// 5 % slower, the extra K steps cost more than the fewer cache lines save.)  wIdx = index of the row-side W block per pair (d.pairW into d.Wshort, or d.pairEdgeI into the full W of the older tile path).
template <int P, int L> __global__ void __launch_bounds__(256) schur_pairs_dmma_kernel(SchurDev d, const double* __restrict__ Hpl, const double* __restrict__ W, const int32_t* __restrict__ wIdx) {
  constexpr int PP = P * P, PLn = P * L;
  constexpr bool kFringe = P > 8;
  const int lane = threadIdx.x & 31, m = lane >> 2, kq = lane & 3;
  const int nWarps = gridDim.x * (blockDim.x >> 5);
  for (int sgm = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); sgm < d.nPairSegs; sgm += nWarps) {
    const int pb = d.pairSegBegin[sgm], np = d.pairSegBegin[sgm + 1] - pb;
    int myW = 0, myJ = 0;                               // lane t < np: the blocks of pair t
    if (lane < np) { myW = wIdx[pb + lane]; myJ = d.pairEdgeJ[pb + lane]; }
    double C00[2] = {0, 0}, c01 = 0, c10 = 0, c11 = 0;   // fringe partial sums over this lane's K slots: (row m, col 8), (row 8, col m), (8, 8)
    const int K = np * L;
    for (int k0 = 0; k0 < K; k0 += 16) {                 // four K steps at a time: their 16 loads are in flight together
      double a0[4], b0[4], w8[4], b8[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int kg = k0 + 4 * u + kq, pi = kg / L, a = kg - pi * L;
        const bool on = kg < K;
        const int wi = __shfl_sync(0xffffffffu, myW, pi & 31), ej = __shfl_sync(0xffffffffu, myJ, pi & 31);
        const double* wp = W + (size_t)wi * PLn + P * a; const double* bp = Hpl + (size_t)ej * PLn + P * a;
        const bool in = on && m < P;
        a0[u] = in ? wp[m] : 0.0; b0[u] = in ? bp[m] : 0.0;
        if (kFringe) { w8[u] = on ? wp[8] : 0.0; b8[u] = on ? bp[8] : 0.0; }   // row 8 / column 8: every lane reads the two ninth scalars of its K slot
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (k0 + 4 * u < K) pairDmma(C00, a0[u], b0[u]);
        if (kFringe) { c01 += a0[u] * b8[u]; c10 += w8[u] * b0[u]; c11 += w8[u] * b8[u]; }   // plain FMAs (three more DMMAs would be 15/16 padding)
      }
    }
    double* Sb = d.S + (size_t)d.pairSegSlot[sgm] * PP;   // column-major block: element (r, c) at r + P c
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = 2 * kq + h;
      if (m < P && c < P) atomicAdd(Sb + m + P * c, -C00[h]);
    }
    if (kFringe) {
#pragma unroll
      for (int o = 1; o < 4; o <<= 1) { c01 += __shfl_xor_sync(0xffffffffu, c01, o); c10 += __shfl_xor_sync(0xffffffffu, c10, o); c11 += __shfl_xor_sync(0xffffffffu, c11, o); }
      if (kq == 0) atomicAdd(Sb + m + P * 8, -c01);
      if (kq == 1) atomicAdd(Sb + 8 + P * m, -c10);
      if (lane == 2) atomicAdd(Sb + 8 + P * 8, -c11);
    }
  }
}

// (Measured alternative: one cp.reduce.async.bulk .add.f64 of the whole block per pair instead of 81 REDs is slower on B200,
// 1.80 ms vs 1.54 ms on C3 - the L2 reduction rate, not the SM-side RED issue, is the limit.)

// Long tracks: output-stationary accumulation.  A CTA owns the Hschur blocks (rows = kTileRows consecutive cameras, columns = a strip of
// 32 consecutive cameras): warp w <-> camera row, lane <-> camera column, every thread keeps its whole P x P block in registers and
// walks the list of (landmark, cameras present in the rows, cameras present in the strip) entries of its tile.  The B blocks of an
// entry are staged in shared memory once (they are contiguous in Hpl because a landmark's blocks are sorted by camera); W_i = B_i Dinv
// is formed once per row.  Nothing is written to global memory until the end: one atomic add per block element per chunk, instead of one
// per landmark pair (the per-pair atomics ran at ~1.4 elements/clk/SM and made this phase 20 ms on the Venice-shaped problem).
constexpr int kTileBatch = 12;   // entries staged in shared memory per barrier pair
constexpr int kTileParts = 3;    // the P x P block of one (row camera, column camera) pair is split by columns over 3 warps
constexpr int kTileWarps = kTileRows * kTileParts, kTileThreads = kTileWarps * 32;
template <int P, int L> __global__ void __launch_bounds__(kTileThreads, 2) schur_tile_kernel(SchurDev d, const double* __restrict__ Hpl) {
  constexpr int PP = P * P, PLn = P * L, LL = L * L, SB = PLn | 1, CW = P / kTileParts, NA = P * CW;   // odd row stride: conflict-free 64-bit shared loads
  static_assert(P % kTileParts == 0, "block size must split into column parts");
  constexpr int ENT = kTileCols * SB + kTileRows * PLn;      // doubles of shared memory per staged entry
  extern __shared__ double smem[];
  __shared__ int sHdr[kTileBatch][5];                        // lm, baseI, baseJ, maskJ, maskI
  const int chunk = blockIdx.x, tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  const int bi = wid / kTileParts, part = wid - bi * kTileParts, c0 = part * CW;
  const int ci = d.chunkI[chunk] * kTileRows + bi, cj = d.chunkJ[chunk] * kTileCols + lane;
  const bool mine = ci < d.numPoses && cj < d.numPoses && cj >= ci;
  double acc[NA];                                            // columns c0 .. c0+CW-1 of the block, all P rows
#pragma unroll
  for (int q = 0; q < NA; ++q) acc[q] = 0;
  bool touched = false;
  const int eBegin = d.chunkBegin[chunk], eEnd = d.chunkEnd[chunk];
  for (int e0 = eBegin; e0 < eEnd; e0 += kTileBatch) {
    const int nb = min(kTileBatch, eEnd - e0);
    __syncthreads();
    // stage: warp w copies entries w, w + kTileWarps, ... of the batch (B_j blocks of the strip, W_i = B_i Dinv of the rows)
    for (int eb = wid; eb < nb; eb += kTileWarps) {
      const int e = e0 + eb;
      const int lm = d.entLm[e], baseI = d.entBaseI[e], baseJ = d.entBaseJ[e];
      const unsigned maskJ = d.entMaskJ[e], maskI = d.entMaskI[e];
      if (lane == 0) { sHdr[eb][0] = lm; sHdr[eb][1] = baseI; sHdr[eb][2] = baseJ; sHdr[eb][3] = (int)maskJ; sHdr[eb][4] = (int)maskI; }
      double* sB = smem + (size_t)eb * ENT; double* sW = sB + kTileCols * SB;
      const int nJ = __popc(maskJ), nI = __popc(maskI);
      const double* src = Hpl + (size_t)baseJ * PLn;
      for (int q = lane; q < nJ * PLn; q += 32) { const int jb = q / PLn, el = q - jb * PLn; sB[jb * SB + el] = src[q]; }
      for (int q = lane; q < nI * PLn; q += 32) {
        const int ii = q / PLn, el = q - ii * PLn, r = el % P, bc = el / P;
        const double* Bi = Hpl + (size_t)(baseI + ii) * PLn;
        double v = 0;
#pragma unroll
        for (int a = 0; a < L; ++a) v += Bi[r + P * a] * d.Dinv[(size_t)lm * LL + a + L * bc];
        sW[q] = v;
      }
    }
    __syncthreads();
    if (mine) {
      for (int eb = 0; eb < nb; ++eb) {
        const unsigned maskJ = (unsigned)sHdr[eb][3], maskI = (unsigned)sHdr[eb][4];
        if (!((maskI >> bi) & 1u)) continue;                 // warp-uniform
        if (!((maskJ >> lane) & 1u)) continue;
        touched = true;
        const double* sB = smem + (size_t)eb * ENT;
        const double* Wi = sB + kTileCols * SB + __popc(maskI & ((1u << bi) - 1u)) * PLn;
        const double* Bj = sB + __popc(maskJ & ((1u << lane) - 1u)) * SB + c0;
#pragma unroll
        for (int a = 0; a < L; ++a) {
          double wv[P], bv[CW];
#pragma unroll
          for (int r = 0; r < P; ++r) wv[r] = Wi[r + P * a];
#pragma unroll
          for (int c = 0; c < CW; ++c) bv[c] = Bj[c + P * a];
#pragma unroll
          for (int c = 0; c < CW; ++c)
#pragma unroll
            for (int r = 0; r < P; ++r) acc[r + P * c] += wv[r] * bv[c];
        }
      }
    }
  }
  if (touched) {
    int lo = d.sRowPtr[ci], hi = d.sRowPtr[ci + 1];
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (d.sColIdx[mid] < cj) lo = mid + 1; else hi = mid; }
    double* Sb = d.S + (size_t)lo * PP + (size_t)c0 * P;
#pragma unroll
    for (int q = 0; q < NA; ++q) atomicAdd(Sb + q, -acc[q]);
  }
}

// x_l = Dinv (b_l - Σ_i B_i^T x_p[c_i])   (block_solver.hpp:420-444)
// pass 1: thread per Hpl block; the CTA stages its 128 contiguous blocks in shared memory with coalesced loads, every thread forms
// B^T x_p for its block, runs of equal landmark are summed by their first thread and added to the accumulator (xl, zeroed before).
template <int P, int L> __global__ void __launch_bounds__(128) backsub_accum_kernel(SchurDev d, const double* __restrict__ Hpl, const int32_t* __restrict__ hplLm, int nBlocks,
                                                                                    const double* __restrict__ xp, double* __restrict__ xl) {
  constexpr int PLn = P * L;
  __shared__ double sB[128 * PLn];
  __shared__ double sV[128 * L];
  __shared__ int sLm[128];
  const int tid = threadIdx.x, k0 = d.blockBegin + blockIdx.x * 128;
  const int nb = min(128, d.blockBegin + nBlocks - k0);
  const double* src = Hpl + (size_t)k0 * PLn;
  // this thread's block: camera, landmark and the camera's step first - two dependent global loads that then overlap the load of the tile
  int lm = -1; double x[P];
  if (tid < nb) {
    const int k = k0 + tid, ci = d.hplRowIdx[k];
    lm = hplLm[k];
#pragma unroll
    for (int r = 0; r < P; ++r) x[r] = xp[(size_t)ci * P + r];
  }
  for (int t = tid; t < nb * PLn; t += 128) sB[t] = src[t];
  __syncthreads();
  if (tid < nb) {
#pragma unroll
    for (int q = 0; q < L; ++q) { double v = 0;
#pragma unroll
      for (int r = 0; r < P; ++r) v += sB[tid * PLn + r + P * q] * x[r];
      sV[tid * L + q] = v; }
  }
  sLm[tid] = lm;
  __syncthreads();
  if (lm >= 0 && (tid == 0 || sLm[tid - 1] != lm)) {
    double acc[L];
#pragma unroll
    for (int q = 0; q < L; ++q) acc[q] = sV[tid * L + q];
    for (int t = tid + 1; t < nb && sLm[t] == lm; ++t) {
#pragma unroll
      for (int q = 0; q < L; ++q) acc[q] += sV[t * L + q];
    }
#pragma unroll
    for (int q = 0; q < L; ++q) atomicAdd(xl + (size_t)lm * L + q, acc[q]);
  }
}
// pass 2: x_l = Dinv (b_l - acc_l)
template <int P, int L> __global__ void backsub_finish_kernel(SchurDev d, const double* __restrict__ b, double* __restrict__ xl) {
  constexpr int LL = L * L;
  const int lm = d.lmBegin + blockIdx.x * blockDim.x + threadIdx.x;
  if (lm >= d.lmEnd) return;
  double c[L];
#pragma unroll
  for (int q = 0; q < L; ++q) c[q] = b[(size_t)d.numPoses * P + (size_t)lm * L + q] - xl[(size_t)lm * L + q];
#pragma unroll
  for (int r = 0; r < L; ++r) { double v = 0;
#pragma unroll
    for (int q = 0; q < L; ++q) v += d.Dinv[(size_t)lm * LL + r + L * q] * c[q];
    xl[(size_t)lm * L + r] = v; }
}

// ------------------------------------------------------------------------------------------------
// block-Jacobi preconditioner: Gauss-Jordan with partial pivoting per diagonal block (Eigen inverse() is LU based for P > 4)
template <int P> __global__ void block_inverse_kernel(PcgDev p) {
  constexpr int PP = P * P;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.nb) return;
  double A[PP], X[PP];
  if (p.diag[i] < p.ownLo || p.diag[i] >= p.ownHi) {   // slab PCG: another rank owns (and inverts) this diagonal block; zero so that the sum over ranks is the inverse
    double* dst = p.Minv + (size_t)i * PP;
    for (int q = 0; q < PP; ++q) dst[q] = 0.0;
    return;
  }
  const double* src = p.A + (size_t)p.diag[i] * PP;
#pragma unroll
  for (int q = 0; q < PP; ++q) { A[q] = src[q]; X[q] = 0; }
#pragma unroll
  for (int q = 0; q < P; ++q) { A[q * (P + 1)] += p.lambda; X[q * (P + 1)] = 1; }
  for (int k = 0; k < P; ++k) {
    int piv = k; double best = fabs(A[k + P * k]);
    for (int r = k + 1; r < P; ++r) { const double v = fabs(A[r + P * k]); if (v > best) { best = v; piv = r; } }
    if (piv != k) for (int c = 0; c < P; ++c) { double t = A[k + P * c]; A[k + P * c] = A[piv + P * c]; A[piv + P * c] = t; t = X[k + P * c]; X[k + P * c] = X[piv + P * c]; X[piv + P * c] = t; }
    const double ip = 1.0 / A[k + P * k];
    for (int c = 0; c < P; ++c) { A[k + P * c] *= ip; X[k + P * c] *= ip; }
    for (int r = 0; r < P; ++r) {
      if (r == k) continue;
      const double f = A[r + P * k];
      for (int c = 0; c < P; ++c) { A[r + P * c] -= f * A[k + P * c]; X[r + P * c] -= f * X[k + P * c]; }
    }
  }
  double* dst = p.Minv + (size_t)i * PP;
  for (int q = 0; q < PP; ++q) dst[q] = X[q];
}

// symmetric block SpMV over work items (row, block range); q must be zero on entry.
// Each warp stages G = 32/P consecutive blocks in shared memory with coalesced loads; lane (g, c) then owns COLUMN c of block g
// (9 consecutive doubles, odd stride => conflict-free 64-bit shared loads): the transposed product (A_ij^T d_i)[c] is complete
// inside the lane (one RED into q_j), and the lane's contributions A_ij[:, c] d_j[c] to q_i are kept in P registers for the
// whole item and reduced across the warp once at its end.  Upper blocks are read once.
template <int P> __global__ void __launch_bounds__(128) spmv_sym_kernel(PcgDev p, const int32_t* __restrict__ itemRow, const int32_t* __restrict__ itemBegin,
                                                                         const int32_t* __restrict__ itemEnd, int nItems, const double* __restrict__ src, double* __restrict__ dst) {
  constexpr int PP = P * P, G = 32 / P;
  __shared__ double sA[4][G * PP];
  if (p.scal && p.scal[6] != 0.0) return;     // PCG already converged: remaining launches are no-ops
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * 4 + w;
  if (item >= nItems) return;
  const int row = itemRow[item], kb = itemBegin[item], ke = itemEnd[item];
  const int g = lane / P, c = lane - g * P;
  const bool act = g < G;
  double di[P], yacc[P];
#pragma unroll
  for (int r = 0; r < P; ++r) { di[r] = src[(size_t)row * P + r]; yacc[r] = 0; }
  for (int k0 = kb; k0 < ke; k0 += G) {
    const int nblk = min(G, ke - k0);
    const double* Ab = p.A + (size_t)k0 * PP;
    __syncwarp();
    for (int t = lane; t < nblk * PP; t += 32) sA[w][t] = Ab[t];
    __syncwarp();
    if (act && g < nblk) {
      const int j = p.colIdx[k0 + g];
      const double* a = &sA[w][g * PP + c * P];
      const double djc = src[(size_t)j * P + c];
      double z = 0;
#pragma unroll
      for (int r = 0; r < P; ++r) { const double v = a[r]; yacc[r] += v * djc; z += v * di[r]; }
      if (j != row) atomicAdd(dst + (size_t)j * P + c, z);
    }
  }
  // q_i: sum the per-lane partial columns over the warp (idle lanes hold zeros), lanes 0..P-1 publish
#pragma unroll
  for (int r = 0; r < P; ++r) {
    double v = yacc[r];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    yacc[r] = v;
  }
  if (lane < P) {
    double v = 0;
#pragma unroll
    for (int r = 0; r < P; ++r) if (r == lane) v = yacc[r];
    if (kb == p.rowPtr[row]) v += p.lambda * src[(size_t)row * P + lane];   // the item holding the diagonal block adds lambda d_i
    atomicAdd(dst + (size_t)row * P + lane, v);
  }
}

// Same product with the blocks streamed by the bulk-copy engine: every warp owns a 3-stage ring of shared-memory buffers (2 G blocks each),
// lane 0 issues one cp.async.bulk per stage two stages ahead and an mbarrier per stage reports the bytes (no LSU work, no registers for data
// in flight; 16 warps x 2 x 3.9 KB in flight per SM instead of one 1.9 KB round per warp).  Block k starts at byte 8 P^2 k, which is only
// 8-byte aligned for odd P: the copy starts at the 16-byte boundary below and the compute side skips the extra double.
namespace {
__device__ __forceinline__ uint32_t smemAddrL(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarWaitL(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "SPMV_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra SPMV_DONE;\n"
      "bra SPMV_WAIT;\n"
      "SPMV_DONE:\n"
      "}\n" ::"r"(smemAddrL(bar)), "r"(parity) : "memory");
}
}  // namespace
template <int P> __global__ void __launch_bounds__(128) spmv_tma_kernel(PcgDev p, const int32_t* __restrict__ itemRow, const int32_t* __restrict__ itemBegin,
                                                                         const int32_t* __restrict__ itemEnd, int nItems, const double* __restrict__ src, double* __restrict__ dst) {
  constexpr int PP = P * P, G = 32 / P, HV = 2, SB = HV * G, ST = 3;   // (HV = 5 for 3 x 3 blocks was measured: 137 vs 79 ms per C5 iteration - the small-block product is bound by issue slots and REDs, not by bytes in flight)
  constexpr int STAGE = (SB * PP + 2 + 1) & ~1;              // doubles per stage (room for the alignment slack), even => 16-byte aligned stages
  __shared__ __align__(16) double sA[4][ST][STAGE];
  __shared__ uint64_t sBar[4][ST];
  if (p.scal && p.scal[6] != 0.0) return;     // PCG already converged: remaining launches are no-ops
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * 4 + w;
  if (item >= nItems) return;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < ST; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smemAddrL(&sBar[w][s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const int row = itemRow[item], kb = itemBegin[item], ke = itemEnd[item];
  const int nChunks = (ke - kb + SB - 1) / SB;
  auto issue = [&](int c) {                   // lane 0 only
    const int k0 = kb + c * SB, nblk = min(SB, ke - k0), st = c % ST;
    const int64_t start = (int64_t)k0 * PP, s0 = start & ~(int64_t)1;
    const uint32_t bytes = (uint32_t)(((start + (int64_t)nblk * PP + 1) & ~(int64_t)1) - s0) * 8u;
    const uint32_t bar = smemAddrL(&sBar[w][st]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smemAddrL(&sA[w][st][0])), "l"(p.A + s0), "r"(bytes), "r"(bar) : "memory");
  };
  if (lane == 0) { for (int c = 0; c < ST - 1 && c < nChunks; ++c) issue(c); }
  const int g = lane / P, cc = lane - g * P;
  const bool act = g < G;
  double di[P], yacc[P];
#pragma unroll
  for (int r = 0; r < P; ++r) { di[r] = src[(size_t)row * P + r]; yacc[r] = 0; }
  // The column index of a block and the entry of d it selects are two dependent global loads; they run ahead of the blocks through
  // registers - indices two chunks ahead, the gathered d one chunk ahead - so that a chunk's products wait for its mbarrier only.
  // (On 3 x 3 blocks, where a chunk is only 1.4 KB, that chain was the top stall: 4.1 TB/s on C5 before.)
  int jCur[HV], jNext[HV]; double dCur[HV], dNext[HV];
  auto loadIdx = [&](int c, int (&jv)[HV]) {
    const int k0 = kb + c * SB, nblk = min(SB, ke - k0);
#pragma unroll
    for (int h = 0; h < HV; ++h) { const int gb = h * G + g; jv[h] = (c < nChunks && act && gb < nblk) ? p.colIdx[k0 + gb] : -1; }
  };
  auto gather = [&](const int (&jv)[HV], double (&dv)[HV]) {
#pragma unroll
    for (int h = 0; h < HV; ++h) dv[h] = jv[h] >= 0 ? src[(size_t)jv[h] * P + cc] : 0.0;
  };
  loadIdx(0, jCur); loadIdx(1, jNext); gather(jCur, dCur);
  for (int c = 0; c < nChunks; ++c) {
    __syncwarp();                             // every lane is done with the stage that is refilled now
    if (lane == 0 && c + ST - 1 < nChunks) issue(c + ST - 1);
    int jAfter[HV];
    loadIdx(c + 2, jAfter); gather(jNext, dNext);
    const int st = c % ST, k0 = kb + c * SB;
    mbarWaitL(&sBar[w][st], (uint32_t)(c / ST) & 1u);
    const double* base = &sA[w][st][0] + (((int64_t)k0 * PP) & 1);
#pragma unroll
    for (int h = 0; h < HV; ++h) {
      const int gb = h * G + g, j = jCur[h];
      if (j >= 0) {
        const double* a = base + gb * PP + cc * P;
        const double djc = dCur[h];
        double z = 0;
#pragma unroll
        for (int r = 0; r < P; ++r) { const double v = a[r]; yacc[r] += v * djc; z += v * di[r]; }
        if (j != row) atomicAdd(dst + (size_t)j * P + cc, z);
      }
    }
#pragma unroll
    for (int h = 0; h < HV; ++h) { jCur[h] = jNext[h]; dCur[h] = dNext[h]; jNext[h] = jAfter[h]; }
  }
#pragma unroll
  for (int r = 0; r < P; ++r) {
    double v = yacc[r];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    yacc[r] = v;
  }
  if (lane < P) {
    double v = 0;
#pragma unroll
    for (int r = 0; r < P; ++r) if (r == lane) v = yacc[r];
    if (kb == p.rowPtr[row]) v += p.lambda * src[(size_t)row * P + lane];   // the item holding the diagonal block adds lambda d_i
    atomicAdd(dst + (size_t)row * P + lane, v);
  }
}

// d.q partial sums
__global__ void __launch_bounds__(256) dot_partial_kernel(const double* scal, const double* a, const double* b, int n, double* partial) {
  __shared__ double sm[8];
  if (scal && scal[6] != 0.0) return;
  double v = 0;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) v += a[i] * b[i];
  const double r = blockSumL<256>(v, sm);
  if (threadIdx.x == 0) partial[blockIdx.x] = r;
}

// ---- slab PCG over NVLink peer memory ----------------------------------------------------------------------------------------
// Exchange k (a device-side sequence number, so that launches skipped after convergence do not count): every rank stores its partial
// product into slot [k & 1][rank] of EVERY rank's buffer (plain stores through the peer mappings), fences at system scope, and the last
// CTA publishes k in flags[rank] of every rank.  The second kernel waits until all flags have reached k, sums the slots in rank order
// and forms the d.q partial sums in the same pass.  Two buffers suffice: a rank can only start exchange k + 2 after it has received the
// flags of k + 1, which its peers set after they finished reading exchange k.
namespace {
__device__ __forceinline__ uint64_t* p2pFlags(double* base, int world, int64_t cap) { return reinterpret_cast<uint64_t*>(base + 2 * (size_t)world * cap); }
__device__ __forceinline__ uint64_t ldAcquireSys(const uint64_t* p) { uint64_t v; asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
// Bounded wait for a peer's flag: a rank that died (CUDA error, host exception) must not hang the others for ever.  After 2^32 cycles
// (about 2 s) the wait gives up and returns false; the caller marks the solve as failed (scal[6] = 3), the host returns G2OCU_E_COMM.
__device__ __forceinline__ bool waitFlagAtLeast(const uint64_t* p, uint64_t k) {
  const long long t0 = clock64();
  while (ldAcquireSys(p) < k) { if (clock64() - t0 > (1LL << 32)) return false; }
  return true;
}
__device__ __forceinline__ void stReleaseSys(uint64_t* p, uint64_t v) { asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
}  // namespace
__global__ void __launch_bounds__(256) p2p_push_kernel(PcgDev p, P2pDev x) {
  if (p.scal[6] != 0.0) return;
  double* mine = x.peer[x.rank];
  uint64_t* fl = p2pFlags(mine, x.world, x.cap);
  const uint64_t k = fl[x.world] + 1;                 // seq lives right after the flags
  const size_t slot = ((size_t)(k & 1) * x.world + x.rank) * x.cap;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < p.n; i += gridDim.x * 256) {
    const double v = p.q[i];
    for (int r = 0; r < x.world; ++r) x.peer[r][slot + i] = v;
  }
  __threadfence_system();
  __syncthreads();
  unsigned int* ticket = reinterpret_cast<unsigned int*>(fl + x.world + 1);
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (last) {
    __threadfence_system();
    if (threadIdx.x == 0) *ticket = 0u;
    if (threadIdx.x < x.world) stReleaseSys(p2pFlags(x.peer[threadIdx.x], x.world, x.cap) + x.rank, k);
  }
}
__global__ void __launch_bounds__(256) p2p_sum_dot_kernel(PcgDev p, P2pDev x) {
  __shared__ double sm[8];
  if (p.scal[6] != 0.0) return;
  double* mine = x.peer[x.rank];
  uint64_t* fl = p2pFlags(mine, x.world, x.cap);
  const uint64_t k = fl[x.world] + 1;
  if (threadIdx.x < x.world && !waitFlagAtLeast(fl + threadIdx.x, k)) p.scal[6] = 3.0;   // peer lost: the solve fails (host: G2OCU_E_COMM)
  __syncthreads();
  const double* slots = mine + (size_t)(k & 1) * x.world * x.cap;
  double v = 0;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < p.n; i += gridDim.x * 256) {
    double s = 0;
    for (int r = 0; r < x.world; ++r) s += __ldcv(slots + (size_t)r * x.cap + i);
    p.q[i] = s;
    v += p.d[i] * s;
  }
  const double r = blockSumL<256>(v, sm);
  if (threadIdx.x == 0) {
    p.partialDq[blockIdx.x] = r;
    unsigned int* ticket = reinterpret_cast<unsigned int*>(fl + x.world + 1) + 1;
    __threadfence();
    if (atomicAdd(ticket, 1u) == gridDim.x - 1) { *ticket = 0u; fl[x.world] = k; }   // every CTA has read seq: advance it
  }
}
// Reduction of the reduced camera system over NVLink peer memory: every rank holds its partial Hschur (all blocks) in a buffer its peers
// have mapped; rank r sums the slab it solves with - blocks [begin, begin + count) - over all ranks in rank order by reading the peers'
// partial sums directly (16-byte loads through the peer mappings).  Replaces ncclReduceScatter of the whole 0.76 GB buffer.
__global__ void __launch_bounds__(256) slab_reduce_kernel(P2pDev x, size_t begin, size_t count) {
  // only [begin, begin + count) of this rank's buffer is written: the neighbours read everything else of it
  const size_t head = begin & 1, body = (count - head) >> 1, tail = (count - head) & 1;
  double* base = x.peer[x.rank];
  double2* mine = reinterpret_cast<double2*>(base + begin + head);
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < body; i += (size_t)gridDim.x * 256) {
    double2 v = make_double2(0.0, 0.0);
    for (int r = 0; r < x.world; ++r) {
      const double2 u = r == x.rank ? mine[i] : __ldcv(reinterpret_cast<const double2*>(x.peer[r] + begin + head) + i);
      v.x += u.x; v.y += u.y;
    }
    mine[i] = v;
  }
  if (blockIdx.x == 0 && threadIdx.x < 2) {
    const bool doIt = threadIdx.x == 0 ? head != 0 : tail != 0;
    const size_t i = threadIdx.x == 0 ? begin : begin + count - 1;
    if (doIt) {
      double v = 0.0;
      for (int r = 0; r < x.world; ++r) v += r == x.rank ? base[i] : __ldcv(x.peer[r] + i);
      base[i] = v;
    }
  }
}
void launchSlabReduce(const P2pDev& x, size_t begin, size_t count, cudaStream_t st, int64_t* launches) {
  if (count == 0) return;
  slab_reduce_kernel<<<148 * 8, 256, 0, st>>>(x, begin, count);
  *launches += 1;
}
void launchP2pExchangeDot(const PcgDev& p, const P2pDev& x, cudaStream_t st, int64_t* launches) {
  p2p_push_kernel<<<p.nPartialDq, 256, 0, st>>>(p, x);
  p2p_sum_dot_kernel<<<p.nPartialDq, 256, 0, st>>>(p, x);
  *launches += 2;
}

// x = 0, r = b, d = M^-1 r, partial r.d   (thread per scalar row: the P threads of a block row read the same r and one row of M^-1 each,
// so every load of the warp is contiguous)
template <int P> __global__ void __launch_bounds__(256) pcg_init_kernel(PcgDev p, const double* __restrict__ b) {
  constexpr int PP = P * P;
  __shared__ double sm[8];
  const int t = blockIdx.x * 256 + threadIdx.x;
  double acc = 0;
  if (t < p.n) {
    const int i = t / P, r = t - i * P;
    const double* M = p.Minv + (size_t)i * PP + r;
    double v = 0, mine = 0;
#pragma unroll
    for (int c = 0; c < P; ++c) { const double rc = b[(size_t)i * P + c]; v += M[P * c] * rc; if (c == r) mine = rc; }
    p.r[t] = mine; p.x[t] = 0; p.d[t] = v; acc = mine * v;
  }
  const double s = blockSumL<256>(acc, sm);
  if (threadIdx.x == 0) p.partial[blockIdx.x] = s;
}
__global__ void __launch_bounds__(256) pcg_init_finish_kernel(PcgDev p, double tolerance, double residual, int absoluteTolerance) {
  __shared__ double sm[8];
  const double dn = sumPartialsAll<256>(p.partial, p.nPartial, sm);
  if (threadIdx.x == 0) {
    double d0 = tolerance * dn;
    if (absoluteTolerance && residual > 0.0 && residual > d0) d0 = residual;
    p.scal[0] = dn; p.scal[2] = dn; p.scal[5] = d0; p.scal[6] = (dn <= d0) ? 1.0 : 0.0; p.scal[7] = 0.0;
  }
}
// alpha = dn / (d.q);  x += alpha d;  s = M^-1 (r - alpha q);  partial (r - alpha q).s   (thread per scalar row, see pcg_init_kernel).
// r itself is left untouched here - the P threads of a block row all read its old values and a block row may straddle two CTAs -
// and is updated by pcg_update2_commit_kernel, which runs after this kernel has finished.
template <int P> __global__ void __launch_bounds__(256) pcg_update1_kernel(PcgDev p, const double* dqPartial, int nDq) {
  constexpr int PP = P * P;
  __shared__ double sm[8];
  if (p.scal[6] != 0.0) return;
  const double dq = sumPartialsAll<256>(dqPartial, nDq, sm);
  const double alpha = p.scal[0] / dq;
  const int t = blockIdx.x * 256 + threadIdx.x;
  double acc = 0;
  if (t < p.n) {
    const int i = t / P, r = t - i * P;
    const double* M = p.Minv + (size_t)i * PP + r;
    double v = 0, mine = 0;
#pragma unroll
    for (int c = 0; c < P; ++c) {
      const size_t o = (size_t)i * P + c;
      const double rc = p.r[o] - alpha * p.q[o];
      v += M[P * c] * rc;
      if (c == r) mine = rc;
    }
    p.x[t] += alpha * p.d[t];
    p.s[t] = v; acc = mine * v;
  }
  const double s = blockSumL<256>(acc, sm);
  if (threadIdx.x == 0) p.partial[blockIdx.x] = s;
}
// r -= alpha q;  beta = dn_new / dn;  d = s + beta d;  q = 0 for the next product;  the CTA that finishes last commits the scalars (dn <- dn_new, iteration
// count, convergence flag): every CTA has read scal[0] / scal[6] before it takes its ticket, so the commit cannot race with them.
__global__ void __launch_bounds__(256) pcg_update2_commit_kernel(PcgDev p, unsigned int* ticket) {
  __shared__ double sm[8];
  if (p.scal[6] != 0.0) return;
  const double dnNew = sumPartialsAll<256>(p.partial, p.nPartial, sm);
  const double dq = sumPartialsAll<256>(p.partialDq, p.nPartialDq, sm);
  const double alpha = p.scal[0] / dq, beta = dnNew / p.scal[0];     // the same alpha pcg_update1_kernel used (same partials, same order)
  for (int i = blockIdx.x * 256 + threadIdx.x; i < p.n; i += gridDim.x * 256) { p.r[i] -= alpha * p.q[i]; p.d[i] = p.s[i] + beta * p.d[i]; p.q[i] = 0.0; }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
      *ticket = 0u;
      p.scal[2] = dnNew; p.scal[0] = dnNew; p.scal[7] += 1.0;
      if (dnNew <= p.scal[5]) p.scal[6] = 1.0;
    }
  }
}
// ---- whole-system PCG (poses and points in one matrix, nothing marginalized): the same recurrences with a block-Jacobi preconditioner of TWO
// block sizes - unknowns [0, np) are pose blocks of P (inverses in p.Minv), the rest point blocks of L (inverses in Dinv).  Thread per scalar row;
// dot_partial_kernel and pcg_update2_commit_kernel above serve this path unchanged.
__device__ __forceinline__ void precondRow2(const PcgDev& p, int np, const double* __restrict__ Dinv, int L, int t, const double* __restrict__ rv, const double* __restrict__ qv,
                                            double alpha, double& v, double& mine) {
  int B, r; const double* M; size_t o;
  if (t < np) { B = p.P; const int i = t / B; r = t - i * B; M = p.Minv + (size_t)i * B * B + r; o = (size_t)i * B; }
  else { B = L; const int u = t - np, i = u / B; r = u - i * B; M = Dinv + (size_t)i * B * B + r; o = (size_t)np + (size_t)i * B; }
  v = 0; mine = 0;
  for (int c = 0; c < B; ++c) { double rc = rv[o + c]; if (qv) rc -= alpha * qv[o + c]; v += M[B * c] * rc; if (c == r) mine = rc; }
}
__global__ void __launch_bounds__(256) pcg_full_init_kernel(PcgDev p, int np, const double* __restrict__ Dinv, int L, const double* __restrict__ b) {
  __shared__ double sm[8];
  const int t = blockIdx.x * 256 + threadIdx.x;
  double acc = 0;
  if (t < p.n) { double v, mine; precondRow2(p, np, Dinv, L, t, b, nullptr, 0.0, v, mine); p.r[t] = mine; p.x[t] = 0; p.d[t] = v; acc = mine * v; }
  const double s = blockSumL<256>(acc, sm);
  if (threadIdx.x == 0) p.partial[blockIdx.x] = s;
}
__global__ void __launch_bounds__(256) pcg_full_update1_kernel(PcgDev p, int np, const double* __restrict__ Dinv, int L, const double* dqPartial, int nDq) {
  __shared__ double sm[8];
  if (p.scal[6] != 0.0) return;
  const double dq = sumPartialsAll<256>(dqPartial, nDq, sm);
  const double alpha = p.scal[0] / dq;
  const int t = blockIdx.x * 256 + threadIdx.x;
  double acc = 0;
  if (t < p.n) { double v, mine; precondRow2(p, np, Dinv, L, t, p.r, p.q, alpha, v, mine); p.x[t] += alpha * p.d[t]; p.s[t] = v; acc = mine * v; }
  const double s = blockSumL<256>(acc, sm);
  if (threadIdx.x == 0) p.partial[blockIdx.x] = s;
}

// The three kernels above (dot_partial, pcg_update1, pcg_update2_commit) as ONE launch for systems of up to 65 536 unknowns: a thread-block
// cluster of 8 CTAs x 1024 threads; the two global reductions of a CG iteration (d.q and r.s) go through distributed shared memory and two
// cluster barriers instead of two kernel boundaries.  Every sum is formed in exactly the order of the three-kernel path - 256 consecutive
// unknowns per partial (8 warp trees, then in warp order), partials summed the way sumPartialsAll<256> does - so that both paths give
// bit-identical iterates and the iteration counts compared with the reference do not depend on the path taken.
namespace cg = cooperative_groups;
constexpr int kFusedCtas = 8, kFusedThreads = 1024, kFusedRoundsMax = 8;   // 8 x 4 groups of 256 threads per round
constexpr int kFusedTailMaxUnknowns = 256 * 32 * kFusedRoundsMax;           // what the kernel can hold; see pcgFusedTail for what it is used for
template <int P> __global__ void __cluster_dims__(kFusedCtas, 1, 1) __launch_bounds__(kFusedThreads) pcg_tail_fused_kernel(PcgDev p, int dotDone, P2pDev x) {
  constexpr int PP = P * P;
  __shared__ double sWarp[32];                       // one sum per warp of this CTA
  __shared__ double sPart[2][256];                   // all partials of the cluster: [0] d.q, [1] r.s (every CTA holds a full copy)
  __shared__ double sFinal[8];
  if (p.scal[6] != 0.0) return;                      // converged: the same value in every CTA, nobody reaches a cluster barrier
  cg::cluster_group cluster = cg::this_cluster();
  const int cta = (int)cluster.block_rank(), tid = threadIdx.x, grp = tid >> 8, t256 = tid & 255, lane = tid & 31, wid = tid >> 5;
  const int nPart = (p.n + 255) >> 8, rounds = (nPart + 31) >> 5;
  // partial k of a quantity = sum over the unknowns [256 k, 256 k + 256): warp trees, then the 8 warps in order (blockSumL<256>)
  auto reducePartials = [&](const double (&acc)[kFusedRoundsMax], int which) {
#pragma unroll
    for (int ro = 0; ro < kFusedRoundsMax; ++ro) {
      if (ro < rounds) {                               // uniform over the cluster
        const double v = warpSumL(acc[ro]);
        __syncthreads();
        if (lane == 0) sWarp[wid] = v;
        __syncthreads();
        const int vb = cta * 4 + grp + 32 * ro;
        if (t256 == 0 && vb < nPart) {
          double r = 0;
#pragma unroll
          for (int k = 0; k < 8; ++k) r += sWarp[8 * grp + k];
          for (int c = 0; c < kFusedCtas; ++c) cluster.map_shared_rank(&sPart[which][0], c)[vb] = r;
        }
      }
    }
    cluster.sync();
  };
  // sum of partial[0 .. n) the way sumPartialsAll<256> forms it
  auto sumPartials = [&](const double* part, int n) -> double {
    double v = 0;
    if (tid < 256) { for (int k = tid; k < n; k += 256) v += part[k]; v = warpSumL(v); }
    __syncthreads();
    if (tid < 256 && lane == 0) sFinal[wid] = v;
    __syncthreads();
    double r = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) r += sFinal[k];
    __syncthreads();
    return r;
  };
  double acc[kFusedRoundsMax];
  double dq;
  if (dotDone == 2) {
    // slab PCG over NVLink peer memory: wait until every rank has published its partial product of this exchange (p2p_push_kernel), sum the
    // slots in rank order (bit-identical on all ranks) into q and form d.q in the same pass - what p2p_sum_dot_kernel does as a launch of its own
    double* mine = x.peer[x.rank];
    uint64_t* fl = p2pFlags(mine, x.world, x.cap);
    const uint64_t k = fl[x.world] + 1;
    if (tid < x.world && !waitFlagAtLeast(fl + tid, k)) p.scal[6] = 3.0;   // peer lost: the solve fails (host: G2OCU_E_COMM); this launch still runs to its end
    __syncthreads();
    const double* slots = mine + (size_t)(k & 1) * x.world * x.cap;
#pragma unroll
    for (int ro = 0; ro < kFusedRoundsMax; ++ro) {
      acc[ro] = 0;
      const int t = (cta * 4 + grp + 32 * ro) * 256 + t256;
      if (ro < rounds && t < p.n) {
        double sum = 0;
        for (int r = 0; r < x.world; ++r) sum += __ldcv(slots + (size_t)r * x.cap + t);
        p.q[t] = sum;
        acc[ro] = 0.0 + p.d[t] * sum;
      }
    }
    reducePartials(acc, 0);                          // its cluster barrier also makes q visible to the whole cluster
    if (cta == 0 && tid == 0) fl[x.world] = k;       // every CTA has read the sequence number and its slots: advance it
    dq = sumPartials(sPart[0], nPart);
  } else if (!dotDone) {
#pragma unroll
    for (int ro = 0; ro < kFusedRoundsMax; ++ro) {
      acc[ro] = 0;
      const int t = (cta * 4 + grp + 32 * ro) * 256 + t256;
      if (ro < rounds && t < p.n) acc[ro] = 0.0 + p.d[t] * p.q[t];
    }
    reducePartials(acc, 0);
    dq = sumPartials(sPart[0], nPart);
  } else {
    dq = sumPartials(p.partialDq, p.nPartialDq);     // formed by the peer-memory exchange kernel
  }
  const double dn = p.scal[0], alpha = dn / dq;
  double sv[kFusedRoundsMax];
#pragma unroll
  for (int ro = 0; ro < kFusedRoundsMax; ++ro) {
    acc[ro] = 0; sv[ro] = 0;
    const int t = (cta * 4 + grp + 32 * ro) * 256 + t256;
    if (ro < rounds && t < p.n) {
      const int i = t / P, r = t - i * P;
      const double* M = p.Minv + (size_t)i * PP + r;
      double v = 0, mine = 0;
#pragma unroll
      for (int c = 0; c < P; ++c) {
        const size_t o = (size_t)i * P + c;
        const double rc = p.r[o] - alpha * p.q[o];
        v += M[P * c] * rc;
        if (c == r) mine = rc;
      }
      p.x[t] += alpha * p.d[t];
      sv[ro] = v; acc[ro] = mine * v;
    }
  }
  reducePartials(acc, 1);                            // also orders every read of r / q above before the writes below, cluster wide
  const double dnNew = sumPartials(sPart[1], nPart);
  const double beta = dnNew / dn;
#pragma unroll
  for (int ro = 0; ro < kFusedRoundsMax; ++ro) {
    const int t = (cta * 4 + grp + 32 * ro) * 256 + t256;
    if (ro < rounds && t < p.n) { p.r[t] -= alpha * p.q[t]; p.d[t] = sv[ro] + beta * p.d[t]; p.q[t] = 0.0; }
  }
  if (cta == 0 && tid == 0) {
    p.scal[2] = dnNew; p.scal[0] = dnNew; p.scal[7] += 1.0;
    if (dnNew <= p.scal[5] && p.scal[6] == 0.0) p.scal[6] = 1.0;
  }
}
// computeLambdaInit: max |H_vv(j,j)| over pose and landmark diagonal blocks (levenberg.cpp:152-175)
// poseDiag (optional): the pose diagonals already summed over all ranks; landmarks: the owned range only
__global__ void extract_pose_diag_kernel(SystemDev sys, double* out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x, P = sys.P;
  if (t >= sys.numPoses * P) return;
  const int i = t / P, k = t - i * P;
  out[t] = sys.Hpp[(size_t)sys.hppDiag[i] * P * P + k * (P + 1)];
}
__global__ void __launch_bounds__(1024) maxdiag_kernel(SystemDev sys, const double* poseDiag, int lmBegin, int lmEnd, double* partial) {
  __shared__ double sm[32];
  double m = 0;
  const int P = sys.P, L = sys.L;
  const int64_t tid = (int64_t)blockIdx.x * 1024 + threadIdx.x, nth = (int64_t)gridDim.x * 1024;
  for (int64_t t = tid; t < (int64_t)sys.numPoses * P; t += nth) { const int i = (int)(t / P), k = (int)(t - (int64_t)i * P); m = fmax(m, fabs(poseDiag ? poseDiag[t] : sys.Hpp[(size_t)sys.hppDiag[i] * P * P + k * (P + 1)])); }
  for (int64_t t = (int64_t)lmBegin * L + tid; t < (int64_t)lmEnd * L; t += nth) { const int64_t i = t / L; const int k = (int)(t - i * L); m = fmax(m, fabs(sys.Hll[(size_t)i * L * L + k * (L + 1)])); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_down_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) { double r = 0; for (int k = 0; k < 32; ++k) r = fmax(r, sm[k]); partial[blockIdx.x] = r; }
}
__global__ void maxdiag_final_kernel(const double* partial, int n, double* out) {   // max is exact: any order gives the same value
  double m = 0;
  for (int k = threadIdx.x; k < n; k += 32) m = fmax(m, partial[k]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_down_sync(0xffffffffu, m, o));
  if (threadIdx.x == 0) out[0] = m;
}
// computeScale: sum_j x_j (lambda x_j + b_j) (levenberg.cpp:177-184)
__global__ void __launch_bounds__(256) scale_partial_kernel(const double* x, const double* b, int64_t n, double lambda, double* partial) {
  __shared__ double sm[8];
  double v = 0;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) v += x[i] * (lambda * x[i] + b[i]);
  const double r = blockSumL<256>(v, sm);
  if (threadIdx.x == 0) partial[blockIdx.x] = r;
}
__global__ void __launch_bounds__(256) final_sum_kernel(const double* partial, int n, double* out) {
  __shared__ double sm[8];
  const double r = sumPartialsAll<256>(partial, n, sm);
  if (threadIdx.x == 0) out[0] = r;
}

// Dogleg step vectors (optimization_algorithm_dogleg.cpp:102,141-157): mode 0: out = a u; mode 1: out = v - u; mode 2: out = u + a (v - u);
// PCG recurrences of the full-system solver (linear_solver_pcg.hpp:138-150): mode 3: out += a u; mode 4: out = u + a out
__global__ void __launch_bounds__(256) lincomb_kernel(double* out, const double* u, const double* v, double a, int mode, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const double ui = u[i];
    double r;
    if (mode == 0) r = a * ui;
    else if (mode == 1) r = v[i] - ui;
    else if (mode == 2) r = ui + a * (v[i] - ui);
    else if (mode == 3) r = out[i] + a * ui;
    else r = ui + a * out[i];
    out[i] = r;
  }
}

// ---- full-system PCG over [Hpp Hpl; Hpl^T Hll] (graphs whose points are not marginalized; the product of the Hpp part is spmv_tma_kernel) ----
// out = (M + lambda I) in for a block-diagonal M of nBlocks D x D column-major blocks: thread per scalar row
__global__ void __launch_bounds__(256) blockdiag_mult_kernel(double* __restrict__ out, const double* __restrict__ M, const double* __restrict__ in, int nBlocks, int D, double lambda) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= (int64_t)nBlocks * D) return;
  const int64_t blk = i / D; const int row = (int)(i - blk * D);
  const double* m = M + blk * D * D; const double* x = in + blk * D;
  double v = lambda * x[row];
  for (int k = 0; k < D; ++k) v += m[row + D * k] * x[k];
  out[i] = v;
}
// Dinv = (Hll + lambda I)^-1 per point block (Eigen's fixed-size inverse(): cofactors), thread per block
template <int L> __global__ void __launch_bounds__(128) point_block_inverse_kernel(double* __restrict__ Dinv, const double* __restrict__ Hll, int n, double lambda) {
  constexpr int LL = L * L;
  const int i = blockIdx.x * 128 + threadIdx.x;
  if (i >= n) return;
  double H[LL], X[LL];
#pragma unroll
  for (int q = 0; q < LL; ++q) H[q] = Hll[(size_t)i * LL + q];
#pragma unroll
  for (int q = 0; q < L; ++q) H[q * (L + 1)] += lambda;
  invSmall<L>(H, X);
#pragma unroll
  for (int q = 0; q < LL; ++q) Dinv[(size_t)i * LL + q] = X[q];
}
// q_p[row] += B d_l[lm], q_l[lm] += B^T d_p[row] for every Hpl block B (P x L, column-major): thread per block
__global__ void __launch_bounds__(128) hpl_mult_kernel(const double* __restrict__ Hpl, const int32_t* __restrict__ hplRow, const int32_t* __restrict__ hplLm, int nBlocks, int P, int L,
                                                        const double* __restrict__ dp, const double* __restrict__ dl, double* qp, double* ql) {
  const int k = blockIdx.x * 128 + threadIdx.x;
  if (k >= nBlocks) return;
  const double* B = Hpl + (size_t)k * P * L;
  const int row = hplRow[k], lm = hplLm[k];
  const double* xp = dp + (size_t)row * P; const double* xl = dl + (size_t)lm * L;
  for (int c = 0; c < L; ++c) {
    double t = 0;
    for (int r = 0; r < P; ++r) t += B[r + P * c] * xp[r];
    atomicAdd(ql + (size_t)lm * L + c, t);
  }
  for (int r = 0; r < P; ++r) {
    double t = 0;
    for (int c = 0; c < L; ++c) t += B[r + P * c] * xl[c];
    atomicAdd(qp + (size_t)row * P + r, t);
  }
}

// ------------------------------------------------------------------------------------------------
// launch wrappers
// ------------------------------------------------------------------------------------------------
void launchLincomb(double* out, const double* u, const double* v, double a, int mode, int64_t n, cudaStream_t st, int64_t* launches) {
  if (n <= 0) return;
  int64_t nb64 = (n + 255) / 256; const int nb = (int)(nb64 < 148 * 8 ? nb64 : 148 * 8);
  lincomb_kernel<<<nb, 256, 0, st>>>(out, u, v, a, mode, n);
  *launches += 1;
}
void launchBlockDiagMult(double* out, const double* M, const double* in, int nBlocks, int D, double lambda, cudaStream_t st, int64_t* launches) {
  if (nBlocks <= 0) return;
  blockdiag_mult_kernel<<<(unsigned)(((int64_t)nBlocks * D + 255) / 256), 256, 0, st>>>(out, M, in, nBlocks, D, lambda);
  *launches += 1;
}
void launchPointBlockInverse(double* Dinv, const double* Hll, int n, int L, double lambda, cudaStream_t st, int64_t* launches) {
  if (n <= 0) return;
  if (L == 2) point_block_inverse_kernel<2><<<(n + 127) / 128, 128, 0, st>>>(Dinv, Hll, n, lambda);
  else point_block_inverse_kernel<3><<<(n + 127) / 128, 128, 0, st>>>(Dinv, Hll, n, lambda);
  *launches += 1;
}
void launchHplMult(const double* Hpl, const int32_t* hplRow, const int32_t* hplLm, int nBlocks, int P, int L, const double* dp, const double* dl, double* qp, double* ql,
                   cudaStream_t st, int64_t* launches) {
  if (nBlocks <= 0) return;
  hpl_mult_kernel<<<(nBlocks + 127) / 128, 128, 0, st>>>(Hpl, hplRow, hplLm, nBlocks, P, L, dp, dl, qp, ql);
  *launches += 1;
}

struct MarkScope {
  const KernelMarks* m;
  MarkScope(const KernelMarks* m_, const char* name) : m(m_) { if (m && m->begin) m->begin(m->ctx, name); }
  ~MarkScope() { if (m && m->end) m->end(m->ctx); }
};
template <int P, int L> static void schurPL(const SchurDev& d, const SystemDev& sys, const int32_t* hplLm, int nBlocks, double lambda, double lambdaDiag, cudaStream_t st, int64_t* launches, const KernelMarks* marks,
                                            const SideStream* side) {
  constexpr int PP = P * P;
  cudaMemsetAsync(d.S, 0, sizeof(double) * (size_t)d.nnzS * PP, st);
  const int64_t tot = max((int64_t)d.nnzHpp * PP, (int64_t)d.numPoses * P);
  schur_init_kernel<P><<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(d, sys.Hpp, sys.b, lambda);
  add_lambda_diag_kernel<P><<<(d.numPoses * P + 255) / 256, 256, 0, st>>>(d.S, d.sDiag, d.numPoses, lambdaDiag);
  *launches += 3;
  if (d.lmEnd > d.lmBegin) { dinv_kernel<P, L><<<(d.lmEnd - d.lmBegin + 255) / 256, 256, 0, st>>>(d, sys.Hll, sys.b, lambda); *launches += 1; }
  const bool mma = schurMmaSupported(P, L);
  // The short-track kernel is bound by the L2 reduction rate, the coefficient pass by HBM and the tile kernel by the FP64 pipe: the first
  // runs on the side stream next to the other two (all of them only add into S / b_schur).  Serial when per-kernel timing is on.
  const bool forked = side && side->stream && d.nPairs > 0 && !(marks && marks->begin);
  const bool kpack = mma && schurKpackEnabled();
  auto pairs = [&](cudaStream_t ps) {
    if (d.nPairSegs <= 0) return;
    MarkScope ms(forked ? nullptr : marks, "schur_pairs");
    const double* W = d.Wshort ? d.Wshort : d.W; const int32_t* wIdx = d.Wshort ? d.pairW : d.pairEdgeI;
    const int nbs = (d.nPairSegs + 7) / 8 < 148 * 8 * 4 ? (d.nPairSegs + 7) / 8 : 148 * 8 * 4;
    schur_pairs_dmma_kernel<P, L><<<nbs, 256, 0, ps>>>(d, sys.Hpl, W, wIdx);
    *launches += 1;
  };
  // coefficient pass first: b_schur, and W of the short tracks for the pair kernel (the older tile path forms all of W in its own pass)
  if (mma && !kpack) launchSchurMma(d, sys, hplLm, nBlocks, st, launches, marks);
  else if (nBlocks > 0) { MarkScope ms(marks, "schur_coeff"); coeff_kernel<P, L><<<(nBlocks + 127) / 128, 128, 0, st>>>(d, sys.Hpl, hplLm, nBlocks); *launches += 1; }
  if (forked) { cudaEventRecord(side->fork, st); cudaStreamWaitEvent(side->stream, side->fork, 0); pairs(side->stream); cudaEventRecord(side->join, side->stream); }
  if (kpack) launchSchurKpack(d, sys, hplLm, nBlocks, st, launches, marks);
  if (!forked) pairs(st);
  if (d.nTileChunks > 0 && !mma) {
    constexpr int kTileSmem = kTileBatch * (kTileCols * ((P * L) | 1) + kTileRows * P * L) * (int)sizeof(double);
    cudaFuncSetAttribute(schur_tile_kernel<P, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileSmem);   // per device, hence on every call
    MarkScope ms(marks, "schur_tiles");
    schur_tile_kernel<P, L><<<d.nTileChunks, kTileThreads, kTileSmem, st>>>(d, sys.Hpl);
    *launches += 1;
  }
  if (forked) cudaStreamWaitEvent(st, side->join, 0);
}
void launchSchur(const SchurDev& d, const SystemDev& sys, const int32_t* hplLm, int nBlocks, double lambda, double lambdaDiag, cudaStream_t st, int64_t* launches, const KernelMarks* marks,
                 const SideStream* side) {
  if (d.P == 9 && d.L == 3) schurPL<9, 3>(d, sys, hplLm, nBlocks, lambda, lambdaDiag, st, launches, marks, side);
  else if (d.P == 6 && d.L == 3) schurPL<6, 3>(d, sys, hplLm, nBlocks, lambda, lambdaDiag, st, launches, marks, side);
  else if (d.P == 3 && d.L == 2) schurPL<3, 2>(d, sys, hplLm, nBlocks, lambda, lambdaDiag, st, launches, marks, side);
}
template <int P, int L> static void backsubPL(const SchurDev& d, const SystemDev& sys, const int32_t* hplLm, int nBlocks, const double* xp, double* xl, cudaStream_t st, int64_t* launches) {
  cudaMemsetAsync(xl, 0, sizeof(double) * (size_t)d.numLandmarks * L, st);
  if (nBlocks > 0) { backsub_accum_kernel<P, L><<<(nBlocks + 127) / 128, 128, 0, st>>>(d, sys.Hpl, hplLm, nBlocks, xp, xl); *launches += 1; }
  if (d.lmEnd > d.lmBegin) backsub_finish_kernel<P, L><<<(d.lmEnd - d.lmBegin + 255) / 256, 256, 0, st>>>(d, sys.b, xl);
  *launches += 2;
}
void launchBacksub(const SchurDev& d, const SystemDev& sys, const int32_t* hplLm, int nBlocks, const double* xp, double* xl, cudaStream_t st, int64_t* launches) {
  if (d.numLandmarks == 0) return;
  if (d.P == 9 && d.L == 3) backsubPL<9, 3>(d, sys, hplLm, nBlocks, xp, xl, st, launches);
  else if (d.P == 6 && d.L == 3) backsubPL<6, 3>(d, sys, hplLm, nBlocks, xp, xl, st, launches);
  else if (d.P == 3 && d.L == 2) backsubPL<3, 2>(d, sys, hplLm, nBlocks, xp, xl, st, launches);
}

#define FOR_P(Pv, CALL) switch (Pv) { case 3: { CALL(3) break; } case 6: { CALL(6) break; } case 9: { CALL(9) break; } default: break; }

void launchBlockInverse(const PcgDev& p, cudaStream_t st, int64_t* launches) {
#define CALL(PV) block_inverse_kernel<PV><<<(p.nb + 63) / 64, 64, 0, st>>>(p);
  FOR_P(p.P, CALL)
#undef CALL
  *launches += 1;
}

void launchSpmv(const PcgDev& p, const double* src, double* dst, cudaStream_t st, int64_t* launches, bool dstIsZero) {
  if (!dstIsZero) cudaMemsetAsync(dst, 0, sizeof(double) * (size_t)p.n, st);
#define CALL(PV) spmv_tma_kernel<PV><<<(p.nItems + 3) / 4, 128, 0, st>>>(p, p.itemRow, p.itemBegin, p.itemEnd, p.nItems, src, dst);
  FOR_P(p.P, CALL)
#undef CALL
  *launches += 1;   // kernels only: the memset is not a kernel of this library
}

void launchPcgInit(const PcgDev& p, const double* b, double tolerance, double residual, int absoluteTolerance, cudaStream_t st, int64_t* launches) {
#define CALL(PV) pcg_init_kernel<PV><<<p.nPartial, 256, 0, st>>>(p, b);
  FOR_P(p.P, CALL)
#undef CALL
  pcg_init_finish_kernel<<<1, 256, 0, st>>>(p, tolerance, residual, absoluteTolerance);
  *launches += 2;
}

bool pcgSingleCtaTail(const PcgDev&) { return true; }   // the tail leaves q zeroed for the next product
static bool pcgFusedTailEnabled() {
  static const bool on = [] { const char* e = getenv("G2OCU_PCG_TAIL"); return !(e && (e[0] == 's' || e[0] == 'S')); }();   // G2OCU_PCG_TAIL=split: the three-kernel path
  return on;
}
// Where the one-launch tail is used: small systems only - its 8 CTAs work on 8 of the 148 SMs, and inside a CUDA graph the three launches of
// the split path cost less than that from about 10^4 unknowns on (C1, 84 unknowns: 1185 vs 1061 LM it/s with / without it; C3, 16 002: 35.3 vs
// 35.6; C2, 60 000: 312 vs 543; C3 in the slab PCG over peer memory, where it also absorbs the wait for the peers: 71.4 vs 72.5 on 2 GPUs, 121.0
// vs 122.8 on 8).
static int pcgFusedTailMax() {
  static const int n = [] { const char* e = getenv("G2OCU_PCG_FUSED_MAX"); const int v = e ? atoi(e) : 0; return v > 0 ? (v < kFusedTailMaxUnknowns ? v : kFusedTailMaxUnknowns) : 8192; }();   // developer switch
  return n;
}
bool pcgFusedTail(const PcgDev& p) { return pcgFusedTailEnabled() && p.n <= pcgFusedTailMax(); }
// slab PCG with the peer-memory exchange and the one-launch tail: push the partial product, the tail does the rest (waits for the peers,
// sums, d.q, recurrences)
void launchP2pPushAndTail(const PcgDev& p, const P2pDev& x, cudaStream_t st, int64_t* launches) {
  p2p_push_kernel<<<p.nPartialDq, 256, 0, st>>>(p, x);
#define CALL(PV) pcg_tail_fused_kernel<PV><<<kFusedCtas, kFusedThreads, 0, st>>>(p, 2, x);
  FOR_P(p.P, CALL)
#undef CALL
  *launches += 2;
}
void launchFullPcgInit(const PcgDev& p, int np, const double* Dinv, int L, const double* b, double tolerance, double residual, int absoluteTolerance, cudaStream_t st, int64_t* launches) {
  pcg_full_init_kernel<<<p.nPartial, 256, 0, st>>>(p, np, Dinv, L, b);
  pcg_init_finish_kernel<<<1, 256, 0, st>>>(p, tolerance, residual, absoluteTolerance);
  *launches += 2;
}
void launchFullPcgTail(const PcgDev& p, int np, const double* Dinv, int L, cudaStream_t st, int64_t* launches) {
  dot_partial_kernel<<<p.nPartialDq, 256, 0, st>>>(p.scal, p.d, p.q, p.n, p.partialDq);
  pcg_full_update1_kernel<<<p.nPartial, 256, 0, st>>>(p, np, Dinv, L, p.partialDq, p.nPartialDq);
  pcg_update2_commit_kernel<<<p.nPartialDq, 256, 0, st>>>(p, p.ticket);
  *launches += 3;
}
void launchPcgTail(const PcgDev& p, cudaStream_t st, int64_t* launches, bool dotDone) {
  if (pcgFusedTail(p)) {
#define CALL(PV) pcg_tail_fused_kernel<PV><<<kFusedCtas, kFusedThreads, 0, st>>>(p, dotDone ? 1 : 0, P2pDev());
    FOR_P(p.P, CALL)
#undef CALL
    *launches += 1;
    return;
  }
  if (!dotDone) dot_partial_kernel<<<p.nPartialDq, 256, 0, st>>>(p.scal, p.d, p.q, p.n, p.partialDq);
#define CALL(PV) pcg_update1_kernel<PV><<<p.nPartial, 256, 0, st>>>(p, p.partialDq, p.nPartialDq);
  FOR_P(p.P, CALL)
#undef CALL
  pcg_update2_commit_kernel<<<p.nPartialDq, 256, 0, st>>>(p, p.ticket);
  *launches += 3;
}

void launchExtractPoseDiag(const SystemDev& sys, double* out, cudaStream_t st, int64_t* launches) {
  extract_pose_diag_kernel<<<(sys.numPoses * sys.P + 255) / 256, 256, 0, st>>>(sys, out);
  *launches += 1;
}
void launchMaxDiag(const SystemDev& sys, const double* poseDiag, int lmBegin, int lmEnd, double* scratch, double* out, cudaStream_t st, int64_t* launches) {
  const int nb = 148;   // partial maxima land in scratch, one per CTA
  maxdiag_kernel<<<nb, 1024, 0, st>>>(sys, poseDiag, lmBegin, lmEnd, scratch);
  maxdiag_final_kernel<<<1, 32, 0, st>>>(scratch, nb, out);
  *launches += 2;
}
void launchScale(const double* x, const double* b, int64_t n, double lambda, double* scratch, double* out, cudaStream_t st, int64_t* launches) {
  int64_t nb64 = (n + 255) / 256; const int nb = (int)(nb64 < 1184 ? nb64 : 1184);
  scale_partial_kernel<<<nb, 256, 0, st>>>(x, b, n, lambda, scratch);
  final_sum_kernel<<<1, 256, 0, st>>>(scratch, nb, out);
  *launches += 2;
}

}  // namespace g2ocu
