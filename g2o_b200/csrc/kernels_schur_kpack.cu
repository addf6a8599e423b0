// Schur complement on the FP64 tensor pipe, second generation: K packed across landmarks, row side compacted per camera.
//   Hschur(i,j) -= (B_i Dinv_l) B_j^T  over all landmarks l seen by cameras i <= j            (block_solver.hpp:357-392)
// Same tiling as kernels_schur_mma.cu (a CTA owns 8 consecutive row cameras x a strip of 32 column cameras and walks the "entries" =
// landmarks seen from both sides, chunks of <= 1024 entries), different use of the DMMA (mma.sync.m8n8k4.f64):
//   * the K dimension of a DMMA holds 4 (landmark, coordinate) pairs taken from the list of landmarks ONE row camera sees: 4 landmarks fill
//     3 DMMAs exactly (the first generation spent one K slot in four on padding), and landmarks the camera does not see cost nothing;
//   * the N dimension runs over the 32 column cameras of the strip STACKED: 32 x 9 scalar rows = 36 tiles of 8 (no ninth-column fringe);
//     absent column cameras are zero rows of a dense operand, and a group of 8 column cameras none of the four landmarks touches is skipped
//     by one warp-uniform branch around its 9 DMMAs;
//   * warp w owns row camera w of the group (rows 0..7 of its blocks: 36 accumulator tiles in registers); the ninth rows of the 8 row
//     cameras form one more M = 8 operand over the union of the batch's landmarks, its 36 tiles split 5/5/5/5/4/4/4/4 over the warps.
// Operands are staged by 8 producer warps (setmaxnreg hands their registers to the 8 consumer warps) that only copy: per batch of 10
// entries, cp.async brings the Hpl blocks of the 32 column cameras and of the 8 row cameras into dense rows (one row per (landmark,
// coordinate), 288 + 4 resp. 72 + 4 doubles; absent cameras are the same copy with source size 0 = zeros) plus Dinv of the landmark.
// Rows k and k' hit different shared-memory banks whenever k != k' mod 4, and lane a of a DMMA fragment only ever reads rows = a mod 4,
// so fragment loads are conflict free by construction.  W = Hpl Dinv is formed by the consumers when they load an A fragment (3 FMAs
// per step) - the W array of the first generation and its 2 GB round trip are gone.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <cstdio>

#include "kernels.hpp"

namespace g2ocu {

#define G2D __device__ __forceinline__

namespace {

constexpr int kKpConsumers = 8, kKpProducers = 8;
constexpr int kKpThreads = (kKpConsumers + kKpProducers) * 32;
constexpr int kKpProducerThreads = kKpProducers * 32;
constexpr int kKpBatch = 10;                 // entries per stage
constexpr int kKpRows = 32;                  // operand rows per stage: row 3 p + c = (entry p of the batch, landmark coordinate c); 30 used
constexpr int kKpStages = 2;
constexpr int kKpZeroRow = kKpStages * kKpRows;   // rows 64..67: zeros, one per bank residue, operand of padded K slots

// Stage header (32-bit words), written by the producers for every batch:
//   [0] entries, [4 + 4 p ..] maskI, maskJ, column groups, first column-side block of entry p, [44 + 2 p ..] first row-side block, landmark,
//   [64 + 2 (4 w + a) ..] queue of row camera w, DMMA lane group a: byte t = t-th operand row = a (mod 4) of an entry the camera is part of,
//   [128 + w] the 4 queue lengths (one byte each), [136 + w] steps = longest of them, [144 + w] column groups touched by step t (nibble t),
//   [152 + w] column cameras the row camera meets in this batch, [160] column cameras of the whole batch
constexpr int kHdrQueue = 64, kHdrLens = 128, kHdrSteps = 136, kHdrGroups = 144, kHdrTouched = 152, kHdrTouchedAll = 160, kHdrWords = 176;

template <int P> struct KpLayout {
  static constexpr int NS = 32 * P;                    // stacked scalar columns of the strip
  static constexpr int NT = NS / 8;                    // DMMA tiles along N: 36 (P = 9) / 24 (P = 6)
  static constexpr int TG = NT / 4;                    // tiles per group of 8 column cameras
  static constexpr int BS = NS + 4;                    // B row stride in doubles (= 4 mod 16)
  static constexpr int AS = 8 * P + 4;                 // W row stride in doubles (= 12 or 4 mod 16)
  static constexpr int kAllRows = kKpStages * kKpRows + 4;
  static constexpr int offB = 0;
  static constexpr int offA = offB + kAllRows * BS * 8;
  static constexpr int offDinv = offA + kAllRows * AS * 8;        // per stage and entry: Dinv of the landmark (9 doubles, padded to 10)
  static constexpr int offHdr = offDinv + kKpStages * kKpBatch * 10 * 8;
  static constexpr int hdrStage = 4 * kHdrWords;
  static constexpr int offDesc = offHdr + kKpStages * hdrStage;     // entry descriptors of the current and the next batch (2 x 64 words), producers only
  static constexpr int offSlots = offDesc + 2 * 64 * 4;   // write-out: Hschur slot of block (row camera w, column camera n), 8 x 32 ints
  static constexpr int offBar = offSlots + 8 * 32 * 4;
  static constexpr int bytes = offBar + 2 * kKpStages * 8;
};

G2D uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
G2D void mbarInit(uint64_t* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count)); }
G2D void mbarArrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smemAddr(bar)) : "memory"); }
G2D void mbarWait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "KP_WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra KP_WAIT_DONE;\n"
      "bra KP_WAIT_LOOP;\n"
      "KP_WAIT_DONE:\n"
      "}\n" ::"r"(smemAddr(bar)), "r"(parity) : "memory");
}
G2D void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
G2D double ldsF64(uint32_t addr) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr)); return v; }
G2D uint32_t ldsU32(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
G2D uint32_t uniformOr(uint32_t v) { return __reduce_or_sync(0xffffffffu, v); }
G2D int uniformMax(int v) { return __reduce_max_sync(0xffffffffu, v); }


template <int P> __global__ void __launch_bounds__(kKpThreads, 1) schur_kpack_kernel(SchurDev d, const double* __restrict__ Hpl) {
  using LY = KpLayout<P>;
  constexpr int PLn = P * 3, PP = P * P;
  constexpr bool FR = P > 8;                        // ninth row of the row cameras: separate M = 8 operand
  constexpr int MR = P < 8 ? P : 8;                 // rows of a block covered by the main DMMA
  constexpr int NT = LY::NT, TG = LY::TG;
  extern __shared__ __align__(128) unsigned char smemRaw[];
  const uint32_t smemBase = smemAddr(smemRaw);
  uint64_t* sFull = reinterpret_cast<uint64_t*>(smemRaw + LY::offBar);
  uint64_t* sEmpty = sFull + kKpStages;

  const int chunk = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int eBegin = d.chunkBegin[chunk], eEnd = d.chunkEnd[chunk];
  {  // all operand rows start as zeros: rows 64..67 stay zero for good (operand of padded K slots), and a stage row no batch has written yet
     // may be multiplied by a zero of the other operand - it must not hold a NaN bit pattern
    double* z = reinterpret_cast<double*>(smemRaw);
    for (int t = threadIdx.x; t < LY::offHdr / 8; t += kKpThreads) z[t] = 0.0;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < kKpStages; ++s) { mbarInit(sFull + s, kKpProducers); mbarInit(sEmpty + s, kKpConsumers); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp >= kKpConsumers) {
    // ------------------------------------------------------------ producers ------------------------------------------------------------
    asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
    const int tp = threadIdx.x - kKpConsumers * 32;
    int stage = 0; uint32_t phase = 0;
    // descriptors of the next batch travel in registers: their global-memory latency hides behind the staging of the current one
    uint32_t nMi = 0, nMj = 0; int nBj = 0, nBi = 0, nLm = 0;
    auto loadDesc = [&](int e0) {
      if (e0 + tp < eEnd && tp < kKpBatch) { nMi = d.entMaskI[e0 + tp]; nMj = d.entMaskJ[e0 + tp]; nBj = d.entBaseJ[e0 + tp]; nBi = d.entBaseI[e0 + tp]; nLm = d.entLm[e0 + tp]; }
    };
    // ... and reach shared memory one batch ahead (two slots, by batch parity), so that nothing but the copies themselves stands between
    // the release of a stage and its refill: slot layout = [4 + 4 p ..] maskI, maskJ, column groups, first column-side block of entry p,
    // [44 + 2 p ..] first row-side block, landmark
    uint32_t* descRing = reinterpret_cast<uint32_t*>(smemRaw + LY::offDesc);
    auto storeDesc = [&](uint32_t* slot) {
      if (tp < kKpBatch) {
        uint32_t g = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) g |= ((nMj >> (8 * q)) & 0xffu) ? (1u << q) : 0u;
        slot[4 + 4 * tp] = nMi; slot[4 + 4 * tp + 1] = nMj; slot[4 + 4 * tp + 2] = g; slot[4 + 4 * tp + 3] = (uint32_t)nBj;
        slot[4 + kKpBatch * 4 + 2 * tp] = (uint32_t)nBi; slot[4 + kKpBatch * 4 + 2 * tp + 1] = (uint32_t)nLm;
      }
    };
    loadDesc(eBegin);
    storeDesc(descRing);
    loadDesc(eBegin + kKpBatch);
    asm volatile("bar.sync 1, %0;" ::"n"(kKpProducerThreads) : "memory");
    int batchIdx = 0;
    for (int e0 = eBegin; e0 < eEnd; e0 += kKpBatch, ++batchIdx) {
      const int nE = min(kKpBatch, eEnd - e0);
      mbarWait(sEmpty + stage, phase ^ 1u);
      uint32_t* hdr = reinterpret_cast<uint32_t*>(smemRaw + LY::offHdr + stage * LY::hdrStage);
      const uint32_t* desc = descRing + (batchIdx & 1) * 64;
      if (tp == 0) hdr[0] = (uint32_t)nE;
      // descriptors of batch b + 1 into the other slot (its readers finished with batch b - 1 before the barrier of the previous iteration);
      // the barrier further down publishes them.  Those of batch b + 2 start their way through the registers.
      storeDesc(descRing + ((batchIdx + 1) & 1) * 64);
      loadDesc(e0 + 2 * kKpBatch);
      const int rows = 3 * nE;
      if (tp >= kKpProducerThreads - 32) {
        // Queues of the 8 row cameras x 4 lane groups, by the last producer warp (lane = (w, a); a quad of lanes shares a row camera).
        // Branch free: lane p < 10 holds the descriptor of entry p in registers, the others read it with shuffles.
        const int ql = tp - (kKpProducerThreads - 32), qw = ql >> 2, qa = ql & 3;
        const uint32_t eMi = ql < nE ? desc[4 + 4 * ql] : 0u, eMj = ql < nE ? desc[4 + 4 * ql + 1] : 0u, eG = ql < nE ? desc[4 + 4 * ql + 2] : 0u;
        uint32_t qlo = 0, qhi = 0, len = 0, gseq = 0, tch = 0;
#pragma unroll
        for (int t = 0; t < kKpRows / 4; ++t) {
          const int k = 4 * t + qa, pe = (k * 11) >> 5;          // rows past the batch belong to entries >= nE, whose masks read as 0
          const uint32_t mi = __shfl_sync(0xffffffffu, eMi, pe & 31), mj = __shfl_sync(0xffffffffu, eMj, pe & 31), g = __shfl_sync(0xffffffffu, eG, pe & 31);
          const uint32_t on = (mi >> qw) & 1u;
          const uint32_t sh = 8u * (len & 3u), kv = on ? (uint32_t)k << sh : 0u;
          qlo |= len < 4 ? kv : 0u; qhi |= len < 4 ? 0u : kv;
          gseq |= on ? g << (4u * len) : 0u;
          tch |= on ? mj : 0u;
          len += on;
        }
        uint32_t steps = len, all = eMj;
#pragma unroll
        for (int o = 1; o < 4; o <<= 1) {
          gseq |= __shfl_xor_sync(0xffffffffu, gseq, o); tch |= __shfl_xor_sync(0xffffffffu, tch, o);
          steps = max(steps, __shfl_xor_sync(0xffffffffu, steps, o));
        }
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) all |= __shfl_xor_sync(0xffffffffu, all, o);   // lanes 0..15 cover the 10 entries
        uint32_t lens = len << (8 * qa);
        lens |= __shfl_xor_sync(0xffffffffu, lens, 1); lens |= __shfl_xor_sync(0xffffffffu, lens, 2);
        hdr[kHdrQueue + 2 * ql] = qlo; hdr[kHdrQueue + 2 * ql + 1] = qhi;
        if (qa == 0) { hdr[kHdrLens + qw] = lens; hdr[kHdrSteps + qw] = steps; hdr[kHdrGroups + qw] = gseq; hdr[kHdrTouched + qw] = tch; }
        if (ql == 0) hdr[kHdrTouchedAll] = all;
      }
      const uint32_t Bd = smemBase + LY::offB + (uint32_t)stage * kKpRows * LY::BS * 8;
      const uint32_t Ad = smemBase + LY::offA + (uint32_t)stage * kKpRows * LY::AS * 8;
      // Copies: a thread owns a scalar column of the stage (one of the 8 x P of the row cameras' A rows or of the 32 x P of the column
      // cameras' B rows) and walks the rows, so that a warp writes 32 consecutive doubles of a row with each cp.async (8 bytes per lane: the
      // Hpl blocks are only 8-byte aligned).  An absent camera is the same copy with source size 0, which writes zeros.  The row side and
      // Dinv go first, in a cp.async group of their own: W is formed while the rest of the column side is still landing.
      if (tp < nE * 9) {   // Dinv of the landmarks
        const int p = tp / 9, q = tp - 9 * p;
        const double* src = d.Dinv + (size_t)(int)desc[4 + kKpBatch * 4 + 2 * p + 1] * 9 + q;
        const uint32_t dst = smemBase + LY::offDinv + (uint32_t)((stage * kKpBatch + p) * 10 + q) * 8u;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
      }
#pragma unroll 1
      for (int pass = 0; pass < 2; ++pass) {
        const int col = tp + pass * kKpProducerThreads;          // columns 0 .. 8 P - 1: row side, then the column side
        if (col < LY::NS + 8 * P) {
          const bool rowSide = col < 8 * P;
          const int s2 = rowSide ? col : col - 8 * P, cam = s2 / P, r = s2 - cam * P;
          const uint32_t dstCol = (rowSide ? Ad : Bd) + (uint32_t)s2 * 8u;
          const uint32_t rowBytes = rowSide ? LY::AS * 8 : LY::BS * 8;
          const uint32_t below = (1u << cam) - 1u;
#pragma unroll 5
          for (int pe = 0; pe < nE; ++pe) {
            const uint32_t mask = rowSide ? desc[4 + 4 * pe] : desc[4 + 4 * pe + 1];
            const int first = (int)(rowSide ? desc[4 + kKpBatch * 4 + 2 * pe] : desc[4 + 4 * pe + 3]);
            const bool on = (mask >> cam) & 1u;
            const double* src = Hpl + (on ? (size_t)(first + __popc(mask & below)) * PLn + r : 0);
            const uint32_t dst = dstCol + (uint32_t)(3 * pe) * rowBytes, sz = on ? 8u : 0u;
#pragma unroll
            for (int c = 0; c < 3; ++c) asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst + (uint32_t)c * rowBytes), "l"(src + P * c), "r"(sz) : "memory");
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
      asm volatile("cp.async.wait_group 1;" ::: "memory");                        // this thread's first pass (row side, Dinv) has landed
      asm volatile("bar.sync 1, %0;" ::"n"(kKpProducerThreads) : "memory");     // ... and everybody else's
      {  // row side in place: W[r, :] = B[r, :] Dinv (block_solver.hpp:366, BDinv = Bi1 * DInvBlock); item = (entry p, scalar column s of the 8 row cameras)
        double* Arows = reinterpret_cast<double*>(smemRaw + LY::offA) + (size_t)stage * kKpRows * LY::AS;
        const double* Dv = reinterpret_cast<const double*>(smemRaw + LY::offDinv) + (size_t)stage * kKpBatch * 10;
        for (int idx = tp; idx < nE * 8 * P; idx += kKpProducerThreads) {
          const int p = idx / (8 * P), s2 = idx - p * (8 * P);
          double* a0 = Arows + (3 * p) * LY::AS + s2;
          const double b0 = a0[0], b1 = a0[LY::AS], b2 = a0[2 * LY::AS];
          const double* di = Dv + 10 * p;
          a0[0] = b0 * di[0] + b1 * di[1] + b2 * di[2];
          a0[LY::AS] = b0 * di[3] + b1 * di[4] + b2 * di[5];
          a0[2 * LY::AS] = b0 * di[6] + b1 * di[7] + b2 * di[8];
        }
        if (FR && nE < kKpBatch)   // last batch of the chunk: the ninth-row pass reads all 32 row slots, the unused ones must be zero
          for (int t = rows * LY::AS + tp; t < kKpRows * LY::AS; t += kKpProducerThreads) Arows[t] = 0.0;
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbarArrive(sFull + stage);
      if (++stage == kKpStages) { stage = 0; phase ^= 1u; }
    }
    return;
  }

  // ------------------------------------------------------------ consumers ------------------------------------------------------------
  asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
  const int w = uniformMax(warp);                      // row camera of this warp inside the group
  const int m = lane >> 2, a = lane & 3;               // fragment coordinates: A[m][k = a], B[k = a][n = m]
  double Cm[NT][2];
#pragma unroll
  for (int u = 0; u < NT; ++u) { Cm[u][0] = 0; Cm[u][1] = 0; }
  double C9[5][2];
#pragma unroll
  for (int u = 0; u < 5; ++u) { C9[u][0] = 0; C9[u][1] = 0; }
  const int tile9 = w < 4 ? 5 * w : 20 + 4 * (w - 4);  // first of this warp's ninth-row tiles (5 for w < 4, else 4)
  uint32_t groups9 = 0;                                // column-camera groups those tiles lie in
  if (FR) for (int t = 0; t < (w < 4 ? 5 : 4); ++t) groups9 |= 1u << ((tile9 + t) / TG);
  uint32_t touched = 0, touched9 = 0;                  // column cameras that received a product (main / ninth row)

  const uint32_t baseB = smemBase + LY::offB + (uint32_t)m * 8u;
  const uint32_t baseA = smemBase + LY::offA + (uint32_t)(P * w + m) * 8u;
  int stage = 0; uint32_t phase = 0;
  for (int e0 = eBegin; e0 < eEnd; e0 += kKpBatch) {
    mbarWait(sFull + stage, phase);
    const uint32_t hdrAddr = smemBase + LY::offHdr + (uint32_t)stage * LY::hdrStage;
    // ---- my queue (prepared by the producers): the operand rows = a (mod 4) of the entries row camera w is part of ----
    const uint32_t qlo = ldsU32(hdrAddr + 4u * (kHdrQueue + 2 * (4 * w + a))), qhi = ldsU32(hdrAddr + 4u * (kHdrQueue + 2 * (4 * w + a) + 1));
    const int len = (int)((ldsU32(hdrAddr + 4u * (kHdrLens + w)) >> (8 * a)) & 0xffu);
    const int steps = (int)uniformOr(ldsU32(hdrAddr + 4u * (kHdrSteps + w)));
    uint32_t gseq = uniformOr(ldsU32(hdrAddr + 4u * (kHdrGroups + w)));
    touched |= ldsU32(hdrAddr + 4u * (kHdrTouched + w));
    const uint32_t rowBase = (uint32_t)stage * kKpRows;
    auto stepOperands = [&](int pos, double& av, uint32_t& bRow) {
      const uint32_t k = ((pos < 4 ? qlo >> (8 * pos) : qhi >> (8 * (pos - 4))) & 0xffu);
      const uint32_t row = pos < len ? rowBase + k : (uint32_t)kKpZeroRow + a;
      av = ldsF64(baseA + row * (LY::AS * 8));         // A fragment: W[m, c] of my (landmark, coordinate) row
      bRow = baseB + row * (LY::BS * 8);
    };
    double avN = 0.0; uint32_t bRowN = 0;
    if (steps > 0) stepOperands(0, avN, bRowN);
#pragma unroll 1
    for (int pos = 0; pos < steps; ++pos) {
      const double av = avN; const uint32_t bRow = bRowN;
      const uint32_t gm = gseq & 0xfu; gseq >>= 4;
      if (pos + 1 < steps) stepOperands(pos + 1, avN, bRowN);   // in flight behind this step's DMMAs
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if ((gm >> q) & 1u) {                           // warp-uniform: a group of 8 column cameras none of the 4 landmarks touches is skipped
#pragma unroll
          for (int u = 0; u < TG; ++u) dmma(Cm[q * TG + u], av, ldsF64(bRow + (uint32_t)(q * TG + u) * 64u));
        }
      }
    }
    if (FR) {
      // ---- ninth rows of the 8 row cameras over all 32 row slots of the stage, 4 per DMMA in row order (absent cameras and unused
      // slots are zero rows of W) ----
      touched9 |= ldsU32(hdrAddr + 4u * kHdrTouchedAll);
      const int T = (3 * (int)uniformOr(ldsU32(hdrAddr)) + 3) >> 2;
      const uint32_t a9Row = smemBase + LY::offA + (uint32_t)(P * m + 8) * 8u + (rowBase + (uint32_t)a) * (LY::AS * 8);   // W of row camera m, element (8, c)
      const uint32_t b9Row = baseB + (rowBase + (uint32_t)a) * (LY::BS * 8) + (uint32_t)tile9 * 64u;
      double av9[kKpRows / 4];
#pragma unroll
      for (int t = 0; t < kKpRows / 4; ++t) av9[t] = ldsF64(a9Row + (uint32_t)t * (4u * LY::AS * 8));
#pragma unroll
      for (int t = 0; t < kKpRows / 4; ++t) {
        if (t < T) {
          const uint32_t bRow = b9Row + (uint32_t)t * (4u * LY::BS * 8);
#pragma unroll
          for (int u = 0; u < 4; ++u) dmma(C9[u], av9[t], ldsF64(bRow + (uint32_t)u * 64u));
          if (w < 4) dmma(C9[4], av9[t], ldsF64(bRow + 4u * 64u));
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbarArrive(sEmpty + stage);
    if (++stage == kKpStages) { stage = 0; phase ^= 1u; }
  }

  // ------------------------------------------------------------ write-out ------------------------------------------------------------
  touched = uniformOr(touched); touched9 = uniformOr(touched9);
  // lane n holds the Hschur slot of block (row camera w, column camera n); the table of all 8 row cameras serves the ninth-row tiles
  const int mySlot = ((touched >> lane) & 1u) ? d.chunkSlots[((size_t)chunk * kMmaTileRows + w) * kTileCols + lane] : -1;
  int* slotTab = reinterpret_cast<int*>(smemRaw + LY::offSlots);
  if (FR) { slotTab[32 * w + lane] = mySlot; asm volatile("bar.sync 2, %0;" ::"n"(kKpConsumers * 32) : "memory"); }
#pragma unroll
  for (int u = 0; u < NT; ++u) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int s2 = 8 * u + 2 * a + h;                // stacked scalar column held in Cm[u][h]
      const int n = s2 / P, r = s2 - n * P;
      const int slot = __shfl_sync(0xffffffffu, mySlot, n);
      if (slot >= 0 && m < MR) atomicAdd(d.S + (size_t)slot * PP + m + P * r, -Cm[u][h]);
    }
  }
  if (FR) {
    const int nT = w < 4 ? 5 : 4;
#pragma unroll
    for (int u = 0; u < 5; ++u) {
      if (u < nT) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int s2 = 8 * (tile9 + u) + 2 * a + h;
          const int n = s2 / P, r = s2 - n * P;
          const double v = C9[u][h];                   // element (8, r) of block (row camera m, column camera n)
          if (v != 0.0) { const int slot = slotTab[32 * m + n]; if (slot >= 0) atomicAdd(d.S + (size_t)slot * PP + 8 + P * r, -v); }
        }
      }
    }
  }
}

}  // namespace

template <int P> static void launchKpackP(const SchurDev& d, const SystemDev& sys, const int32_t* hplLm, int nBlocks, cudaStream_t st, int64_t* launches, const KernelMarks* marks) {
  if (d.nTileChunks > 0) {
    cudaFuncSetAttribute(schur_kpack_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, KpLayout<P>::bytes);   // per device, hence on every call
    if (marks && marks->begin) marks->begin(marks->ctx, "schur_tiles");
    schur_kpack_kernel<P><<<d.nTileChunks, kKpThreads, KpLayout<P>::bytes, st>>>(d, sys.Hpl);
    if (marks && marks->end) marks->end(marks->ctx);
    *launches += 1;
  }
}
bool schurKpackEnabled() {
  static const bool on = [] { const char* e = getenv("G2OCU_SCHUR_KERNEL"); return !(e && (e[0] == 'm' || e[0] == 'M')); }();
  return on;
}
// K-packed tensor-pipe tile pass; the caller has initialised S, computed Dinv and run the coefficient pass (b_schur)
void launchSchurKpack(const SchurDev& d, const SystemDev& sys, const int32_t* hplLm, int nBlocks, cudaStream_t st, int64_t* launches, const KernelMarks* marks) {
  if (d.P == 9 && d.L == 3) launchKpackP<9>(d, sys, hplLm, nBlocks, st, launches, marks);
  else if (d.P == 6 && d.L == 3) launchKpackP<6>(d, sys, hplLm, nBlocks, st, launches, marks);
}

}  // namespace g2ocu
