// Launch wrappers for the sm_100a kernels (definitions in kernels_build.cu / kernels_linear.cu).
// Plain C++ signatures; everything runs on the stream passed in.  All data FP64, indices int32.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace g2ocu {

// Device view of one homogeneous edge set (host_structure.hpp: EdgeSet), arrays in kernel order.
struct EdgeSetDev {
  int etype = 0, n = 0;
  int poseLandmark = 0;                  // 1: (pose, landmark) edge set, sorted by landmark slot
  const int32_t* slot0 = nullptr;        // class slot of vertices()[0]
  const int32_t* slot1 = nullptr;
  const int32_t* block = nullptr;        // off-diagonal target block (Hpl index / Hpp-CSR index), -1 none
  const uint8_t* transposed = nullptr;   // pose-pose sets: write the block transposed
  const int32_t* pos = nullptr;          // position in the active edge list
  const double* meas = nullptr;          // n x M
  const double* info = nullptr;          // E x E (uniform) or n x E x E
  int infoMode = 0;                      // 0 identity, 1 uniform, 2 per edge
  const int32_t* kernelKind = nullptr;   // per edge (kernelMode 2)
  const double* kernelDelta = nullptr;
  int kernelMode = 0;                    // 0 none, 1 uniform, 2 per edge
  int kKind = 0; double kDelta = 1.0;
  const double* prm = nullptr;           // NP (uniform) or n x NP
  int prmMode = 0;                       // 0 none, 1 uniform, 2 per edge
  // pose-sorted view (pose-landmark sets)
  const int32_t* byPose = nullptr; int nByPose = 0;
  const int32_t* chunkPose = nullptr; const int32_t* chunkBegin = nullptr; const int32_t* chunkEnd = nullptr; int nChunks = 0;
  const int32_t* poseChunkPtr = nullptr;
  double* partial = nullptr;             // nChunks x (P(P+1)/2 + P)
};

struct SystemDev {
  int numPoses = 0, numLandmarks = 0, numPoseSlots = 0, numLmSlots = 0, P = 0, L = 0;
  double* poseEst = nullptr; double* lmEst = nullptr;    // slots x S
  double* Hpp = nullptr;        // nnzHpp x P x P (CSR-upper order), diagonal blocks full symmetric
  const int32_t* hppDiag = nullptr;
  double* Hll = nullptr;        // numLandmarks x L x L
  double* Hpl = nullptr;        // nnzHpl x P x L (column-major blocks, CCS-by-landmark order)
  double* b = nullptr;          // sizePoses + sizeLandmarks
  int hplShared = 0, hppShared = 0;
};

// ---- build / error kernels (kernels_build.cu) ----
// chi2 partial sums: out2[0] += sum rho0 (robust), out2[1] += sum chi2 (plain); errOut (optional) in active-edge order
void launchErrors(const EdgeSetDev& s, const SystemDev& sys, double* scratch, double* out2, double* errOut, const int64_t* errOff, cudaStream_t st, int64_t* launches);
void launchBuild(const EdgeSetDev& s, const SystemDev& sys, cudaStream_t st, int64_t* launches);
void launchJacobianDump(const EdgeSetDev& s, const SystemDev& sys, double* jacOut, const int64_t* jacOff, cudaStream_t st, int64_t* launches);
void launchUpdate(int vtype, double* est, double* backupOrNull, int* counters, const double* x, int nFree, cudaStream_t st, int64_t* launches);
int errorScratchDoubles(int n);

// ---- linear algebra kernels (kernels_linear.cu) ----
struct SchurDev {
  int numPoses = 0, numLandmarks = 0, P = 0, L = 0;
  int lmBegin = 0, lmEnd = 0, blockBegin = 0;     // landmark slots / first Hpl block owned by this rank
  const int32_t* hplColPtr = nullptr; const int32_t* hplRowIdx = nullptr;
  const int32_t* sRowPtr = nullptr; const int32_t* sColIdx = nullptr; const int32_t* sDiag = nullptr;
  const int32_t* hppToS = nullptr; int nnzHpp = 0; int nnzS = 0;
  // short tracks: flat list of (Hpl block i, Hpl block j) per pair i <= j, sorted by target Hschur block and cut into segments of <= kPairSegment
  // pairs of one block: a warp forms a segment's sum as one K-stacked DMMA product, one RED per element
  int64_t nPairs = 0; const int32_t* pairEdgeI = nullptr; const int32_t* pairEdgeJ = nullptr;
  int nPairSegs = 0; const int32_t* pairSegBegin = nullptr; const int32_t* pairSegSlot = nullptr;
  // W = B Dinv of the blocks of short tracks, formed by the coefficient pass: compact index of every Hpl block (-1: long track), the W blocks
  // in that order, and per pair the compact index of its row-side block
  const int32_t* hplShortIdx = nullptr; double* Wshort = nullptr; const int32_t* pairW = nullptr;
  double* S = nullptr; double* Dinv = nullptr; double* db = nullptr; double* bschur = nullptr;
  double* W = nullptr;          // Hpl Dinv, same block order as Hpl (tensor-pipe path only)
  // long tracks (>= kTileMinTrack observations): output-stationary tiles, see schur_tile_kernel
  int nTileChunks = 0;
  const int32_t* chunkI = nullptr; const int32_t* chunkJ = nullptr; const int32_t* chunkBegin = nullptr; const int32_t* chunkEnd = nullptr;
  const int32_t* chunkSlots = nullptr;   // per chunk: Hschur block of (row camera w, column camera n) of the tile, 8 x 32 ints, -1 = none / lower triangle (tensor-pipe tiles only)
  const int32_t* entLm = nullptr; const int32_t* entBaseI = nullptr; const int32_t* entBaseJ = nullptr; const uint32_t* entMaskJ = nullptr; const uint8_t* entMaskI = nullptr;
};
static const int kTileRows = 4, kTileCols = 32;   // cameras per tile row group / column strip
static const int kMmaTileRows = 8;                // row group of the tensor-pipe tile kernel (kernels_schur_mma.cu)
bool schurMmaSupported(int P, int L);             // block shapes routed through the DMMA tile kernel
struct KernelMarks;
void launchSchurMma(const SchurDev& d, const SystemDev& sys, const int32_t* hplLm, int nBlocks, cudaStream_t st, int64_t* launches, const KernelMarks* marks);
// second generation of the tensor-pipe tile pass (kernels_schur_kpack.cu): K packed across landmarks, W formed on the fly (no W array);
// G2OCU_SCHUR_KERNEL=mma selects the first generation (kernels_schur_mma.cu)
void launchSchurKpack(const SchurDev& d, const SystemDev& sys, const int32_t* hplLm, int nBlocks, cudaStream_t st, int64_t* launches, const KernelMarks* marks);
bool schurKpackEnabled();
static const int kPairSegment = 16;               // pairs per segment of the short-track kernel
static const int kTileMinTrack = 16;              // landmarks with at least this many observations go through the tile kernel
// optional per-kernel timing hooks: begin(ctx, name) / end(ctx) bracket one kernel (CUDA events on the launching stream in api.cu)
struct KernelMarks { void* ctx = nullptr; void (*begin)(void*, const char*) = nullptr; void (*end)(void*) = nullptr; };
// a second stream of the solver + two events: independent kernels of one phase are forked onto it and joined before the phase ends
struct SideStream { cudaStream_t stream = nullptr; cudaEvent_t fork = nullptr, join = nullptr; };
void launchSchur(const SchurDev& d, const SystemDev& sys, const int32_t* hplLm, int nBlocks, double lambda, double lambdaDiag, cudaStream_t st, int64_t* launches, const KernelMarks* marks = nullptr,
                 const SideStream* side = nullptr);
void launchBacksub(const SchurDev& d, const SystemDev& sys, const int32_t* hplLm, int nBlocks, const double* xp, double* xl, cudaStream_t st, int64_t* launches);

struct PcgDev {
  int n = 0, nb = 0, P = 0;                // scalar size, block rows, block size
  const int32_t* rowPtr = nullptr; const int32_t* colIdx = nullptr; const int32_t* diag = nullptr; int nnz = 0;
  const double* A = nullptr;               // upper blocks, CSR order
  double lambda = 0.0;                     // added to the diagonal on the fly (non-Schur mode)
  double* Minv = nullptr;                  // nb x P x P
  double *r = nullptr, *d = nullptr, *q = nullptr, *s = nullptr, *x = nullptr;
  double* scal = nullptr;                  // device scalars: [0] dn, [1] d.q, [2] dn_new, [3] alpha, [4] beta, [5] d0, [6] converged flag, [7] iterations
  double* partial = nullptr; int nPartial = 0;          // one per CTA of pcg_init / pcg_update1: ceil(n/256)
  double* partialDq = nullptr; int nPartialDq = 0;      // one per CTA of the dot kernel
  const int32_t* itemRow = nullptr; const int32_t* itemBegin = nullptr; const int32_t* itemEnd = nullptr; int nItems = 0;   // SpMV work items (row, block range)
  unsigned int* ticket = nullptr;          // zero-initialised counter for the last-CTA commit of pcg_update2_commit_kernel
  int ownLo = 0, ownHi = 0;                // blocks of A this rank owns (slab PCG); the whole matrix when not sharded
};
// Peer-memory exchange of the slab-PCG product (one process per GPU, buffers opened through CUDA IPC): every rank owns
//   data[2][world][cap] doubles | flags[world] uint64 | seq uint64 | ticket[2] uint32
// `peer[r]` is that block on rank r as mapped into this process (peer[rank] is the local one).
struct P2pDev {
  double* peer[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int rank = 0, world = 1; int64_t cap = 0;
};
static inline size_t p2pBytes(int world, int64_t cap) { return sizeof(double) * 2 * (size_t)world * (size_t)cap + 8 * (size_t)world + 8 + 8 + 48; }
// q <- sum over ranks of the per-rank partial products (fixed rank order: bit-identical on every rank), partialDq <- partial sums of d.q;
// replaces an NCCL all-reduce + dot_partial_kernel in the PCG iteration
void launchP2pExchangeDot(const PcgDev& p, const P2pDev& x, cudaStream_t st, int64_t* launches);
// x.peer[r] = rank r's partial Hschur buffer: mine[begin, begin + count) <- sum over the ranks, in rank order
void launchSlabReduce(const P2pDev& x, size_t begin, size_t count, cudaStream_t st, int64_t* launches);
void launchBlockInverse(const PcgDev& p, cudaStream_t st, int64_t* launches);
void launchPcgInit(const PcgDev& p, const double* b, double tolerance, double residual, int absoluteTolerance, cudaStream_t st, int64_t* launches);   // x=0, r=b, d=M^-1 r, dn=r.d, d0
void launchPcgTail(const PcgDev& p, cudaStream_t st, int64_t* launches, bool dotDone = false);   // after q = A d: dot, x/r/s update, d update, commit (no-ops once converged)
// whole-system PCG (points not marginalized): the same with a preconditioner of two block sizes (unknowns [0, np): blocks of p.P from p.Minv, the rest blocks of L from Dinv)
void launchFullPcgInit(const PcgDev& p, int np, const double* Dinv, int L, const double* b, double tolerance, double residual, int absoluteTolerance, cudaStream_t st, int64_t* launches);
void launchFullPcgTail(const PcgDev& p, int np, const double* Dinv, int L, cudaStream_t st, int64_t* launches);
void launchSpmv(const PcgDev& p, const double* src, double* dst, cudaStream_t st, int64_t* launches, bool dstIsZero = false);   // dst = (A + lambda I) src, symmetric upper
bool pcgFusedTail(const PcgDev& p);        // launchPcgTail runs the recurrences as one cluster kernel (single GPU: small systems, see kernels_linear.cu)
void launchP2pPushAndTail(const PcgDev& p, const P2pDev& x, cudaStream_t st, int64_t* launches);   // slab PCG: peer-memory exchange fused into that kernel
bool pcgSingleCtaTail(const PcgDev& p);   // launchPcgTail zeroes q itself (small systems): the next launchSpmv may skip its memset

void launchExtractPoseDiag(const SystemDev& sys, double* out, cudaStream_t st, int64_t* launches);
void launchMaxDiag(const SystemDev& sys, const double* poseDiag, int lmBegin, int lmEnd, double* scratch /* >= 148 doubles */, double* out, cudaStream_t st, int64_t* launches);
void launchScale(const double* x, const double* b, int64_t n, double lambda, double* scratch, double* out, cudaStream_t st, int64_t* launches);   // out[0] = sum_j x_j (lambda x_j + b_j); with lambda = 0 a fixed-order dot product
// Dogleg step vectors: mode 0: out = a u; mode 1: out = v - u; mode 2: out = u + a (v - u); PCG recurrences: mode 3: out += a u; mode 4: out = u + a out
void launchLincomb(double* out, const double* u, const double* v, double a, int mode, int64_t n, cudaStream_t st, int64_t* launches);
// full-system PCG (points not marginalized): block-diagonal products / inverses and the Hpl, Hpl^T products
void launchBlockDiagMult(double* out, const double* M, const double* in, int nBlocks, int D, double lambda, cudaStream_t st, int64_t* launches);   // out = (M + lambda I) in
void launchPointBlockInverse(double* Dinv, const double* Hll, int n, int L, double lambda, cudaStream_t st, int64_t* launches);                      // Dinv = (Hll + lambda I)^-1, L = 2 | 3
void launchHplMult(const double* Hpl, const int32_t* hplRow, const int32_t* hplLm, int nBlocks, int P, int L, const double* dp, const double* dl, double* qp, double* ql,
                   cudaStream_t st, int64_t* launches);                                                                                            // qp += Hpl dl, ql += Hpl^T dp

// dense FP64 Cholesky path (kernels_dense.cu)
void launchDenseAssemble(const PcgDev& p, double* H, cudaStream_t st, int64_t* launches);      // upper blocks -> dense lower-filled n x n
int launchDenseCholeskySolve(double* H, int n, const double* b, double* x, int* info, cudaStream_t st, int64_t* launches);
int launchDenseCholeskyFactor(double* H, int n, int* info, cudaStream_t st, int64_t* launches);
// blocks of the inverse (computeMarginals): unit columns, (L L^T)^-1 applied to many columns, P x P blocks picked out of the result
void launchUnitColumns(double* X, size_t ldx, const int32_t* colScalar, int nCols, cudaStream_t st, int64_t* launches);
void launchDenseSolveMany(const double* H, int n, double* X, size_t ldx, int nrhs, int firstNonZero, int firstNeeded, cudaStream_t st, int64_t* launches);
void launchGatherBlocks(const double* X, size_t ldx, const int32_t* rowScalar, const int32_t* rowDim, const int32_t* colStart, const int32_t* colDim, const int64_t* outOff, int nPairs,
                        double* out, cudaStream_t st, int64_t* launches);
void launchDenseAssembleFull(const PcgDev& hpp, const double* Hll, const double* Hpl, const int32_t* hplRow, const int32_t* hplLm, int nHplBlocks, int numLandmarks, int L, double* H, int n,
                             cudaStream_t st, int64_t* launches);

}  // namespace g2ocu
