// Schur complement on the FP64 tensor pipe (sm_100a DMMA, mma.sync.m8n8k4.f64) for pose blocks of dimension 6..9 and 3-d landmarks.
//   Hschur(i,j) -= B_i Dinv_l B_j^T  over all landmarks l seen by cameras i <= j            (block_solver.hpp:357-392)
// W = Hpl Dinv is formed once per solve by coeff_w_kernel (which also does b_schur -= B_i Dinv b_l, block_solver.hpp:366-374);
// the products W_i B_j^T are then accumulated OUTPUT-STATIONARY: a CTA owns the Hschur blocks of 8 consecutive row cameras x a strip of
// 32 consecutive column cameras and walks the list of landmarks seen from both sides ("entries": first Hpl block + presence mask on
// each side; a landmark's blocks are contiguous in Hpl because they are sorted by camera).
//   * staging: per batch of up to 32 entries, one cp.async.bulk (TMA bulk copy) per side and entry brings the raw, contiguous
//     W / Hpl blocks into a 2-stage shared-memory ring of 96 KB stages (measured: fewer, larger batches beat a deeper ring), issued
//     lane-parallel by warp 0 one batch ahead; completion is tracked by
//     mbarriers (no __syncthreads in the main loop).
//   * 8 warps = (row half: 4 cameras) x (column group: 8 cameras); 2 warps per SM sub-partition leave 255 registers per thread.  One DMMA covers rows 0..7 x columns 0..7 of one (i,j)
//     block with K = the 3 landmark coordinates (+1 zero pad): A[m][k] = W_i[m,k], B[k][n] = B_j[n,k]; absent row cameras are skipped
//     with one warp-uniform branch each, absent column cameras inside a non-empty group multiply a zero fragment.  For P = 9 the ninth rows / columns are gathered into
//     fringe tiles: (ninth row of the 8 row cameras) x block columns, block rows x (ninth column of the 8 column cameras), and the
//     (ninth, ninth) corner - one extra DMMA per present row, per present column and per entry.
//   * accumulators stay in registers for the whole chunk; one RED per element per chunk at the end.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <cstdio>
#include <type_traits>
#include <vector>

#include "kernels.hpp"

namespace g2ocu {

#define G2D __device__ __forceinline__

namespace {

constexpr int kMmaConsumers = 8;                   // consumer warps
constexpr int kMmaThreads = kMmaConsumers * 32;
constexpr int kMmaStages = 2;
constexpr int kMmaStageDoubles = 12 * 1024;        // 96 KB per stage
constexpr int kMmaBatch = 32;                      // entries per stage at most (one per producer lane)

struct __align__(16) MmaHdr { uint32_t offI, offJ, maskI, maskJ; };   // offsets in doubles into the stage buffer

constexpr int kMmaZeroBytes = 256;                 // zeroed region: operand of absent cameras / padded lanes
constexpr int kMmaScratchEntry = 48;               // per warp and entry: 8 column offsets, 8 row offsets (u16 each), presence bits
constexpr int kMmaOffZero = kMmaStages * kMmaStageDoubles * 8;
constexpr int kMmaOffHdr = kMmaOffZero + kMmaZeroBytes;
constexpr int kMmaOffCnt = kMmaOffHdr + kMmaStages * kMmaBatch * (int)sizeof(MmaHdr);
constexpr int kMmaOffBar = kMmaOffCnt + 32;
constexpr int kMmaOffScratch = kMmaOffBar + 2 * kMmaStages * 8;
constexpr int kMmaSmemBytes = kMmaOffScratch + kMmaConsumers * kMmaBatch * kMmaScratchEntry;

G2D uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
G2D void mbarInit(uint64_t* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count)); }
G2D void mbarArrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smemAddr(bar)) : "memory"); }
G2D void mbarArriveExpectTx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory"); }
G2D void mbarWait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smemAddr(bar)), "r"(parity) : "memory");
}
G2D void bulkLoad(void* dstSmem, const void* srcGlobal, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smemAddr(dstSmem)), "l"(srcGlobal), "r"(bytes),
               "r"(smemAddr(bar))
               : "memory");
}
G2D void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
// Measured on B200 (tools/dmma_occ.cu, tools/dmma_branch.cu): a DMMA that is predicated off still occupies the FP64 pipe for its full
// 16 cycles, a real branch around a single DMMA costs more than that, and ptxas puts WARPSYNC.ALL + NOP in front of every mma.sync it
// cannot prove convergent.  Hence: (1) every value that steers control flow is made warp-uniform through redux.sync (result lives in a
// uniform register, no convergence barriers are emitted), (2) absent row cameras are skipped by one real branch per row (9 DMMA slots
// each), (3) inside a present row all 8 column slots are issued, absent columns multiply a zero fragment.
G2D double ldsF64(uint32_t addr) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr)); return v; }
G2D uint32_t ldsU16(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
G2D uint32_t ldsU32(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
G2D uint4 ldsV4(uint32_t addr) { uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr)); return v; }
G2D void stsV4(uint32_t addr, uint4 v) { asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
G2D void stsU32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
template <int K> G2D uint32_t half16(const uint4& v) {   // K-th 16-bit field of a packed 8 x u16 vector
  const uint32_t w = (K >> 1) == 0 ? v.x : (K >> 1) == 1 ? v.y : (K >> 1) == 2 ? v.z : v.w;
  return (K & 1) ? (w >> 16) : (w & 0xffffu);
}
G2D uint32_t uniformOr(uint32_t v) { return __reduce_or_sync(0xffffffffu, v); }
G2D int uniformMax(int v) { return __reduce_max_sync(0xffffffffu, v); }

// W_k = B_k Dinv_l and b_schur[c_k] -= B_k (Dinv_l b_l) for every Hpl block k: thread per block, blocks staged through shared memory so that
// both the read of Hpl and the write of W are coalesced.
template <int P, int L> __global__ void __launch_bounds__(128) coeff_w_kernel(SchurDev d, const double* __restrict__ Hpl, const int32_t* __restrict__ hplLm, int nBlocks) {
  constexpr int PLn = P * L, LL = L * L;
  __shared__ double sB[128 * PLn];
  const int tid = threadIdx.x, k0 = d.blockBegin + blockIdx.x * 128;
  const int nb = min(128, d.blockBegin + nBlocks - k0);
  const double* src = Hpl + (size_t)k0 * PLn;
  for (int t = tid; t < nb * PLn; t += 128) sB[t] = src[t];
  __syncthreads();
  if (tid < nb) {
    const int k = k0 + tid, ci = d.hplRowIdx[k], lm = hplLm[k];
    double dbv[L], Di[LL];
#pragma unroll
    for (int a = 0; a < L; ++a) dbv[a] = d.db[(size_t)lm * L + a];
#pragma unroll
    for (int a = 0; a < LL; ++a) Di[a] = d.Dinv[(size_t)lm * LL + a];
    double* blk = sB + tid * PLn;
#pragma unroll
    for (int r = 0; r < P; ++r) {
      double bv[L], v = 0;
#pragma unroll
      for (int a = 0; a < L; ++a) { bv[a] = blk[r + P * a]; v += bv[a] * dbv[a]; }
      atomicAdd(d.bschur + (size_t)ci * P + r, -v);
#pragma unroll
      for (int a = 0; a < L; ++a) { double w = 0;
#pragma unroll
        for (int a2 = 0; a2 < L; ++a2) w += bv[a2] * Di[a2 + L * a];
        blk[r + P * a] = w; }
    }
  }
  __syncthreads();
  double* dst = d.W + (size_t)k0 * PLn;
  for (int t = tid; t < nb * PLn; t += 128) dst[t] = sB[t];
}

template <int P, int L> __global__ void __launch_bounds__(kMmaThreads, 1) schur_mma_kernel(SchurDev d, const double* __restrict__ Hpl, int dbg, int never, long long* prof) {
  constexpr int PLn = P * L, PP = P * P;
  constexpr bool FR = P > 8;                        // ninth row / column handled by fringe tiles
  constexpr int MR = P < 8 ? P : 8;                 // rows / columns of a block covered by the main DMMA
  static_assert(L <= 4 && P <= 9 && P >= 5, "block shape not supported by the DMMA mapping");
  extern __shared__ __align__(128) unsigned char smemRaw[];
  double* sData = reinterpret_cast<double*>(smemRaw);
  MmaHdr* sHdr = reinterpret_cast<MmaHdr*>(smemRaw + kMmaOffHdr);
  int* sCnt = reinterpret_cast<int*>(smemRaw + kMmaOffCnt);
  uint64_t* sFull = reinterpret_cast<uint64_t*>(smemRaw + kMmaOffBar);
  uint64_t* sEmpty = sFull + kMmaStages;
  const uint32_t smemBase = smemAddr(smemRaw);       // 32-bit shared-window address of the dynamic segment

  const int chunk = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tileI = d.chunkI[chunk], tileJ = d.chunkJ[chunk];
  const int eBegin = d.chunkBegin[chunk], eEnd = d.chunkEnd[chunk];
  if (threadIdx.x < kMmaZeroBytes / 8) reinterpret_cast<double*>(smemRaw + kMmaOffZero)[threadIdx.x] = 0.0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kMmaStages; ++s) { mbarInit(sFull + s, 1); mbarInit(sEmpty + s, kMmaConsumers); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // ------------------------------------------------ producer (warp 0, lane-parallel over the entries of a batch) ------------------------------------------------
  // Lane l keeps the descriptor of entry pe + l in registers; the window is refilled right after a batch has been cut so that the global
  // loads of the next descriptors are in flight while this warp consumes a batch (their latency never sits in front of a bulk copy).
  int pe = eBegin, pstage = 0; uint32_t pphase = 0;
  uint32_t wBaseI = 0, wBaseJ = 0, wMaskI = 0, wMaskJ = 0;
  auto loadDesc = [&](int idx) {
    wBaseI = (uint32_t)d.entBaseI[idx]; wBaseJ = (uint32_t)d.entBaseJ[idx]; wMaskI = d.entMaskI[idx]; wMaskJ = d.entMaskJ[idx];
  };
  if (warp == 0 && eBegin + lane < eEnd) loadDesc(eBegin + lane);
  auto produce = [&]() {
    if (pe >= eEnd) return;
    mbarWait(sEmpty + pstage, pphase ^ 1u);
    const bool valid = pe + lane < eEnd;
    const uint32_t baseI = wBaseI, baseJ = wBaseJ, maskI = valid ? wMaskI : 0u, maskJ = valid ? wMaskJ : 0u;
    // 16-byte aligned spans (even double offsets) covering the blocks of each side
    const int64_t startI = (int64_t)baseI * PLn, startJ = (int64_t)baseJ * PLn;
    const int64_t s0I = startI & ~(int64_t)1, s0J = startJ & ~(int64_t)1;
    const int lenI = valid ? (int)(((startI + (int64_t)__popc(maskI) * PLn + 1) & ~(int64_t)1) - s0I) : 0;
    const int lenJ = valid ? (int)(((startJ + (int64_t)__popc(maskJ) * PLn + 1) & ~(int64_t)1) - s0J) : 0;
    const int len = lenI + lenJ;
    int incl = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    const int excl = incl - len;
    const bool fits = valid && incl <= kMmaStageDoubles;
    const unsigned fm = __ballot_sync(0xffffffffu, fits);
    const int cnt = __popc(fm);                        // >= 1: one entry is at most (8 + 32) blocks
    const int total = __shfl_sync(0xffffffffu, incl, cnt - 1);
    double* buf = sData + (size_t)pstage * kMmaStageDoubles;
    if (lane < cnt) { MmaHdr h; h.offI = (uint32_t)(excl + (int)(startI - s0I)); h.offJ = (uint32_t)(excl + lenI + (int)(startJ - s0J)); h.maskI = maskI; h.maskJ = maskJ; sHdr[pstage * kMmaBatch + lane] = h; }
    if (lane == 0) sCnt[pstage] = cnt;
    __syncwarp();
    if (lane == 0) mbarArriveExpectTx(sFull + pstage, dbg == 1 ? 0u : (uint32_t)total * 8u);
    __syncwarp();
    if (lane < cnt && dbg != 1) {
      bulkLoad(buf + excl, d.W + s0I, (uint32_t)lenI * 8u, sFull + pstage);
      bulkLoad(buf + excl + lenI, Hpl + s0J, (uint32_t)lenJ * 8u, sFull + pstage);
    }
    pe += cnt;
    if (++pstage == kMmaStages) { pstage = 0; pphase ^= 1u; }
    // slide the descriptor window by cnt entries
    const int srcLane = lane + cnt;
    const uint32_t tBI = __shfl_sync(0xffffffffu, wBaseI, srcLane & 31), tBJ = __shfl_sync(0xffffffffu, wBaseJ, srcLane & 31);
    const uint32_t tMI = __shfl_sync(0xffffffffu, wMaskI, srcLane & 31), tMJ = __shfl_sync(0xffffffffu, wMaskJ, srcLane & 31);
    if (srcLane < 32) { wBaseI = tBI; wBaseJ = tBJ; wMaskI = tMI; wMaskJ = tMJ; }
    else if (pe + lane < eEnd) loadDesc(pe + lane);
  };
  if (warp == 0) {
#pragma unroll 1
    for (int s = 0; s < kMmaStages - 1; ++s) produce();
  }

  // -------------------------------------------------- consumers --------------------------------------------------
  // Warp (rh, cg) owns row cameras 2 i + rh (i < 4) of the row group and column cameras 4 j + cg (j < 8) of the strip: the interleaving
  // gives every warp the same share of a landmark's window, whichever part of the tile the window covers.
  const int warpU = uniformMax(warp);
  const int rh = warpU >> 2, cg = warpU & 3;
  const int m = lane >> 2, a = lane & 3;              // fragment coordinates: (row / column inside the block, landmark coordinate)
  const bool fragOk = a < L && m < MR;
  const bool aOk = a < L;
  const uint32_t zeroAddr = smemBase + kMmaOffZero;
  const uint32_t strideMain = fragOk ? 8u : 0u, strideFr = aOk ? 8u : 0u;    // padded lanes stay on the zero region
  const uint32_t scratchWarp = smemBase + kMmaOffScratch + (uint32_t)warpU * (kMmaBatch * kMmaScratchEntry);
  double Cm[4][8][2], Ca[4][2], Cb[4][2], Cc[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { Cm[i][j][0] = 0; Cm[i][j][1] = 0; }
    Ca[i][0] = Ca[i][1] = Cb[i][0] = Cb[i][1] = 0;
  }
  Cc[0] = Cc[1] = 0;
  uint32_t touched = 0;                               // bit 8 i + j: block (row 2 i + rh, column 4 j + cg) received a product

  long long tWait = 0, tProd = 0, tSlots = 0; const long long tStart = clock64();
  auto consume = [&](auto rhc) {
    constexpr int RH = decltype(rhc)::value;
    int e = eBegin, stage = 0; uint32_t phase = 0;
    while (e < eEnd) {
      long long t0 = 0, t1 = 0, t2 = 0;
      if (prof) t0 = clock64();
      if (RH == 0 && cg == 0) produce();                // warp 0 refills the stage released one batch ago
      if (prof) t1 = clock64();
      mbarWait(sFull + stage, phase);
      if (prof) { t2 = clock64(); tProd += t1 - t0; tWait += t2 - t1; }
      const int cntAll = uniformMax(sCnt[stage]);
      const uint32_t bufAddr = smemBase + (uint32_t)stage * (kMmaStageDoubles * 8);
      // ---- batch preparation, lane = entry: block offsets of my cameras (absent -> zero region), presence bits, relevance ----
      bool relevant = false;
      if (lane < cntAll && dbg != 2) {
        const uint4 h = ldsV4(smemBase + kMmaOffHdr + (uint32_t)(stage * kMmaBatch + lane) * 16u);   // offI, offJ, maskI, maskJ
        const uint32_t zeroOff = (zeroAddr - bufAddr) >> 3;
        uint32_t co[8], ro[8], cJ = 0, tI = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t c = 4 * j + cg, on = (h.w >> c) & 1u;
          co[j] = on ? h.y + __popc(h.w & ((1u << c) - 1u)) * PLn : zeroOff;
          cJ |= on << j;
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const uint32_t on = (h.z >> r) & 1u;
          ro[r] = on ? h.x + __popc(h.z & ((1u << r) - 1u)) * PLn : zeroOff;
          if ((r & 1) == RH) tI |= on << (r >> 1);
        }
        relevant = cJ != 0 && (tI != 0 || (FR && (RH == 0 || (cJ >> 4) != 0)));
        const uint32_t sa = scratchWarp + (uint32_t)lane * kMmaScratchEntry;
        stsV4(sa, make_uint4(co[0] | (co[1] << 16), co[2] | (co[3] << 16), co[4] | (co[5] << 16), co[6] | (co[7] << 16)));
        stsV4(sa + 16, make_uint4(ro[0] | (ro[1] << 16), ro[2] | (ro[3] << 16), ro[4] | (ro[5] << 16), ro[6] | (ro[7] << 16)));
        stsU32(sa + 32, tI | (cJ << 8));
      }
      uint32_t rel = uniformOr(__ballot_sync(0xffffffffu, relevant));
      __syncwarp();
      const uint32_t baseMain = fragOk ? bufAddr + (uint32_t)(m + P * a) * 8u : zeroAddr;
      const uint32_t baseFr = aOk ? bufAddr + (uint32_t)(8 + P * a) * 8u : zeroAddr;
      // ---- visits: only the entries this warp has work for ----
#pragma unroll 1
      while (rel) {
        const int q = __ffs(rel) - 1; rel &= rel - 1u;
        const uint32_t sa = scratchWarp + (uint32_t)q * kMmaScratchEntry;
        const uint4 cp = ldsV4(sa), rp = ldsV4(sa + 16);
        const uint32_t bits = ldsU32(sa + 32);
        const uint32_t tI = uniformOr(bits & 0xfu), cJ = (bits >> 8) & 0xffu;
        double Bf[8], Af[4], FA = 0, FB = 0;
        Bf[0] = ldsF64(baseMain + half16<0>(cp) * strideMain); Bf[1] = ldsF64(baseMain + half16<1>(cp) * strideMain);
        Bf[2] = ldsF64(baseMain + half16<2>(cp) * strideMain); Bf[3] = ldsF64(baseMain + half16<3>(cp) * strideMain);
        Bf[4] = ldsF64(baseMain + half16<4>(cp) * strideMain); Bf[5] = ldsF64(baseMain + half16<5>(cp) * strideMain);
        Bf[6] = ldsF64(baseMain + half16<6>(cp) * strideMain); Bf[7] = ldsF64(baseMain + half16<7>(cp) * strideMain);
        Af[0] = ldsF64(baseMain + half16<0 + RH>(rp) * strideMain); Af[1] = ldsF64(baseMain + half16<2 + RH>(rp) * strideMain);
        Af[2] = ldsF64(baseMain + half16<4 + RH>(rp) * strideMain); Af[3] = ldsF64(baseMain + half16<6 + RH>(rp) * strideMain);
        if (FR) {
          // ninth column of my 8 column cameras: B[k][n] = B_{j_n}[8, k];  ninth row of the 8 row cameras of the group: A[m][k] = W_{i_m}[8, k]
          FB = ldsF64(baseFr + ldsU16(sa + 2u * m) * strideFr);
          FA = ldsF64(baseFr + ldsU16(sa + 16u + 2u * m) * strideFr);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if ((tI >> i) & 1u) {
            touched |= cJ << (8 * i);
            if (prof) tSlots += FR ? 9 : 8;
            do {                                        // `never` is 0: the loop form keeps ptxas from predicating the 9 slots of an absent row
#pragma unroll
              for (int j = 0; j < 8; ++j) dmma(Cm[i][j], Af[i], Bf[j]);
              if (FR) dmma(Cb[i], Af[i], FB);
            } while (never);
          }
        }
        if (FR) {
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) dmma(Ca[jj], FA, Bf[4 * RH + jj]);
          if (RH == 0) dmma(Cc, FA, FB);
          if (prof) tSlots += RH == 0 ? 5 : 4;
        }
      }
      __syncwarp();
      if (lane == 0) mbarArrive(sEmpty + stage);
      e += cntAll;
      if (++stage == kMmaStages) { stage = 0; phase ^= 1u; }
    }
  };
  if (rh == 0) consume(std::integral_constant<int, 0>{}); else consume(std::integral_constant<int, 1>{});
  const long long tLoopEnd = clock64();

  // ---------------------------------------------------- write-out ----------------------------------------------------
  auto findSlot = [&](int ci, int cj) -> int {
    if (ci >= d.numPoses || cj >= d.numPoses || cj < ci) return -1;
    int lo = d.sRowPtr[ci]; const int end = d.sRowPtr[ci + 1]; int hi = end;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (d.sColIdx[mid] < cj) lo = mid + 1; else hi = mid; }
    return (lo < end && d.sColIdx[lo] == cj) ? lo : -1;
  };
  const int rowCam0 = tileI * kMmaTileRows, colCam0 = tileJ * kTileCols + cg;   // my row cameras: rowCam0 + 2 i + rh, my column cameras: colCam0 + 4 j
  int mySlot = -1;
  if ((touched >> lane) & 1u) mySlot = findSlot(rowCam0 + 2 * (lane >> 3) + rh, colCam0 + 4 * (lane & 7));
  const int n0 = 2 * a, n1 = 2 * a + 1;               // accumulator columns held by this lane
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int slot = __shfl_sync(0xffffffffu, mySlot, i * 8 + j);
      if (slot >= 0 && m < MR) {
        double* Sb = d.S + (size_t)slot * PP + m;
        if (n0 < MR) atomicAdd(Sb + P * n0, -Cm[i][j][0]);
        if (n1 < MR) atomicAdd(Sb + P * n1, -Cm[i][j][1]);
      }
    }
    if (FR) {   // rows 0..7 of column 8 of the blocks (my row i, my column n)
      const int s0 = __shfl_sync(0xffffffffu, mySlot, i * 8 + n0), s1 = __shfl_sync(0xffffffffu, mySlot, i * 8 + n1);
      if (s0 >= 0) atomicAdd(d.S + (size_t)s0 * PP + m + P * 8, -Cb[i][0]);
      if (s1 >= 0) atomicAdd(d.S + (size_t)s1 * PP + m + P * 8, -Cb[i][1]);
    }
  }
  if (FR) {
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {   // row 8, columns 0..7 of the blocks (row camera m of the group, my column camera 4 rh + jj)
      if (Ca[jj][0] != 0.0 || Ca[jj][1] != 0.0) {
        const int slot = findSlot(rowCam0 + m, colCam0 + 4 * (4 * rh + jj));
        if (slot >= 0) { double* Sb = d.S + (size_t)slot * PP + 8; atomicAdd(Sb + P * n0, -Ca[jj][0]); atomicAdd(Sb + P * n1, -Ca[jj][1]); }
      }
    }
    if (rh == 0) {                      // element (8, 8) of the blocks (row camera m of the group, my column camera n)
      if (Cc[0] != 0.0) { const int slot = findSlot(rowCam0 + m, colCam0 + 4 * n0); if (slot >= 0) atomicAdd(d.S + (size_t)slot * PP + 8 + P * 8, -Cc[0]); }
      if (Cc[1] != 0.0) { const int slot = findSlot(rowCam0 + m, colCam0 + 4 * n1); if (slot >= 0) atomicAdd(d.S + (size_t)slot * PP + 8 + P * 8, -Cc[1]); }
    }
  }
  if (prof && lane == 0) {
    long long* o = prof + ((size_t)blockIdx.x * kMmaConsumers + warpU) * 8;
    unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    o[0] = tLoopEnd - tStart; o[1] = tWait; o[2] = tProd; o[3] = tSlots; o[4] = clock64() - tLoopEnd; o[5] = eEnd - eBegin; o[6] = smid; o[7] = tStart;
  }
}

}  // namespace

bool schurMmaSupported(int P, int L) { return L == 3 && (P == 9 || P == 6); }

template <int P, int L> static void launchMmaPL(const SchurDev& d, const SystemDev& sys, const int32_t* hplLm, int nBlocks, cudaStream_t st, int64_t* launches, const KernelMarks* marks) {
  if (nBlocks > 0) { if (marks && marks->begin) marks->begin(marks->ctx, "schur_coeff"); coeff_w_kernel<P, L><<<(nBlocks + 127) / 128, 128, 0, st>>>(d, sys.Hpl, hplLm, nBlocks); *launches += 1; if (marks && marks->end) marks->end(marks->ctx); }
  if (d.nTileChunks > 0) {
    cudaFuncSetAttribute(schur_mma_kernel<P, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMmaSmemBytes);   // per device, hence on every call
    static const int dbg = getenv("G2OCU_SCHUR_DEBUG") ? atoi(getenv("G2OCU_SCHUR_DEBUG")) : 0;   // developer switches: 1 no loads, 2 no products, 3 per-warp cycle dump (tools/schur_prof.py)
    long long* prof = nullptr;
    if (dbg == 3) cudaMalloc(&prof, sizeof(long long) * 8 * kMmaConsumers * (size_t)d.nTileChunks);
    if (marks && marks->begin) marks->begin(marks->ctx, "schur_tiles");
    schur_mma_kernel<P, L><<<d.nTileChunks, kMmaThreads, kMmaSmemBytes, st>>>(d, sys.Hpl, dbg, 0, prof);
    if (marks && marks->end) marks->end(marks->ctx);
    if (prof) {   // developer instrumentation: per-warp cycle breakdown of every chunk
      cudaStreamSynchronize(st);
      std::vector<long long> h((size_t)8 * kMmaConsumers * d.nTileChunks);
      cudaMemcpy(h.data(), prof, h.size() * sizeof(long long), cudaMemcpyDeviceToHost); cudaFree(prof);
      if (FILE* f = fopen("gpurun_out/schur_prof.bin", "wb")) { fwrite(h.data(), sizeof(long long), h.size(), f); fclose(f); }
    }
    *launches += 1;
  }
}
// coefficient pass (writes W) + tensor-pipe tile pass; the caller has initialised S / b_schur and computed Dinv / db
void launchSchurMma(const SchurDev& d, const SystemDev& sys, const int32_t* hplLm, int nBlocks, cudaStream_t st, int64_t* launches, const KernelMarks* marks) {
  if (d.P == 9 && d.L == 3) launchMmaPL<9, 3>(d, sys, hplLm, nBlocks, st, launches, marks);
  else if (d.P == 6 && d.L == 3) launchMmaPL<6, 3>(d, sys, hplLm, nBlocks, st, launches, marks);
}

}  // namespace g2ocu
