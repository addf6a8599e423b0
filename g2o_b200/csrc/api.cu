// C ABI of libg2ocu.so (include/g2ocu.h): handle, device memory, phase sequencing and the LM / GN control flow.
// The control flow restates OptimizationAlgorithmLevenberg::solve (optimization_algorithm_levenberg.cpp:58-150),
// OptimizationAlgorithmGaussNewton::solve (optimization_algorithm_gauss_newton.cpp:50-91),
// OptimizationAlgorithmDogleg::solve (optimization_algorithm_dogleg.cpp:56-197) and
// SparseOptimizer::optimize (sparse_optimizer.cpp:374-439); all per-edge / per-block arithmetic runs in the kernels.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <string>
#include <vector>

#include "../../include/g2ocu.h"
#include "host_structure.hpp"
#include "kernels.hpp"

using namespace g2ocu;

namespace {

std::string g_createError;
const int64_t kDenseMaxN = 40000;   // dense FP64 Cholesky: n x n doubles (12.8 GB at the limit)

double wallNow() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// NCCL is resolved at run time from the library the host names (torch ships one); only the handful of entry points used here.
struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, struct NcclId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*ReduceScatter)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
struct NcclId { char internal[128]; };
NcclApi g_nccl;
bool loadNccl(const char* path, std::string& err) {
  if (g_nccl.handle) return true;
  void* h = dlopen(path && *path ? path : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) { err = std::string("dlopen(") + (path ? path : "libnccl.so.2") + "): " + dlerror(); return false; }
  NcclApi a; a.handle = h;
  a.GetUniqueId = (int (*)(void*))dlsym(h, "ncclGetUniqueId");
  a.CommInitRank = (int (*)(void**, int, NcclId, int))dlsym(h, "ncclCommInitRank");
  a.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclAllReduce");
  a.ReduceScatter = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclReduceScatter");
  a.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
  a.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
  if (!a.GetUniqueId || !a.CommInitRank || !a.AllReduce || !a.ReduceScatter || !a.CommDestroy || !a.GetErrorString) { err = "the NCCL library lacks a required symbol"; dlclose(h); return false; }
  g_nccl = a;
  return true;
}

template <class T> struct DVec {
  T* p = nullptr; size_t n = 0;
  DVec() {}
  DVec(const DVec&) = delete; DVec& operator=(const DVec&) = delete;
  ~DVec() { release(); }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
  cudaError_t alloc(size_t count) {
    if (count == n && p) return cudaSuccess;
    release();
    if (count == 0) return cudaSuccess;
    cudaError_t e = cudaMalloc((void**)&p, count * sizeof(T));
    if (e == cudaSuccess) n = count;
    return e;
  }
  cudaError_t upload(const std::vector<T>& h, cudaStream_t st) {
    cudaError_t e = alloc(h.size());
    if (e != cudaSuccess || h.empty()) return e;
    return cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, st);
  }
  cudaError_t zero(cudaStream_t st) { return n ? cudaMemsetAsync(p, 0, n * sizeof(T), st) : cudaSuccess; }
};

struct EdgeSetState {
  EdgeSetDev dev;
  DVec<int32_t> slot0, slot1, block, pos, byPose, chunkPose, chunkBegin, chunkEnd, poseChunkPtr, kernelKind;
  DVec<uint8_t> transposed;
  DVec<double> meas, info, prm, kernelDelta, partial;
  int scratchDoubles = 0;
};

struct PhaseAcc { double seconds = 0; int64_t launches = 0; int64_t calls = 0; };
struct PendingEvent { std::string phase; cudaEvent_t a, b; int64_t launches; };

}  // namespace

struct g2ocu_solver {
  g2ocu_config cfg;
  std::string err;
  HostGraph g; bool hasGraph = false;
  Structure st; bool optInitialized = false, structureBuilt = false, algoInitialized = false;
  cudaStream_t stream = nullptr; bool ownStream = false, cudaReady = false;
  SideStream side;                // second stream for independent kernels inside a phase (forked from / joined into `stream`)
  // LM properties / state (optimization_algorithm_levenberg.cpp:40-52)
  double userLambdaInit = 0.0; int maxTrialsAfterFailure = 10;
  double currentLambda = -1.0, tau = 1e-5, goodStepUpperScale = 2. / 3., goodStepLowerScale = 1. / 3., ni = 2.0;
  int levenbergIterations = 0;
  // Dogleg properties / state (optimization_algorithm_dogleg.cpp:40-52, optimization_algorithm_dogleg.h:77-91)
  double dlUserDeltaInit = 1e4, dlInitialLambda = 1e-7, dlLambdaFactor = 10., dlDelta = 1e4, dlCurrentLambda = 1e-7;
  int dlMaxTrialsAfterFailure = 100, dlLastStep = 0, dlLastNumTries = 0; bool dlWasPD = true;
  // solver state
  double lambda = 0.0;            // damping currently "set" on the diagonals (0 after restoreDiagonal)
  double pcgResidual = -1.0;      // LinearSolverPCG::_residual, persists across solves until init()
  // CUDA graph of kPcgGraphIters CG iterations (product + one-launch tail, + the peer-memory push when sharded): launched instead of the
  // individual kernels between two convergence polls; re-captured when a kernel argument changes (matrix, lambda, transport)
  cudaGraphExec_t pcgGraph = nullptr; double pcgGraphLambda = 0; bool pcgGraphP2p = false; const double* pcgGraphA = nullptr; int pcgGraphN = 0; int64_t pcgGraphKernels = 0;
  int tileMinTrack = kTileMinTrack;                            // tracks with fewer observations go through the pair kernel
  int lastPcgIterations = 0; int64_t totalPcgIterations = 0;   // of the last solve / of all solves since g2ocu_reset_counters
  bool errorsValid = false; double chi2Robust = 0, chi2Plain = 0;
  // estimates of each class form one contiguous run of the packed host array and together cover it: copies go straight between the caller's buffer and the device
  bool fastEstimates = false; int64_t poseHostOff = 0, lmHostOff = 0;
  // sharding
  int rank = 0, world = 1; g2ocu_allreduce_fn allreduce = nullptr; void* allreduceUser = nullptr;
  const volatile unsigned char* forceStop = nullptr;   // SparseOptimizer::_forceStopFlag (sparse_optimizer.h:186-190); read between trials / iterations
  int expectPoseDim = -1, expectLandmarkDim = -1;      // BlockSolver<BlockSolverTraits<p,l>>: block sizes the solver was registered for (-1 = variable)
  bool kernelTiming = false;      // property "kernelTiming": CUDA events around individual kernels, not only around phases
  int64_t slabBlocks = 0;         // blocks of the reduced system per rank (slab PCG), 0 when not sharded
  void* ncclComm = nullptr;       // set by g2ocu_set_shard_nccl: collectives go straight to NCCL on the solver's stream
  P2pDev schurPeers; bool schurP2pReady = false; void* schurOpened[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // peers' partial-Hschur buffers (CUDA IPC)
  P2pDev p2p; bool p2pReady = false; double* p2pLocal = nullptr; void* p2pOpened[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // CUDA-IPC peer buffers for the slab-PCG exchange
  // device state
  DVec<double> poseEst, lmEst; std::vector<DVec<double>*> poseBackup, lmBackup; int stackDepth = 0;
  DVec<int> poseCounters, lmCounters;
  DVec<double> denseH; DVec<int> denseInfo; DVec<unsigned int> pcgTicket; int* hostInfo = nullptr;
  DVec<double> Hpp, Hll, Hpl, W, Wshort, b, x, S, Dinv, dbv, bschur, Minv, vr, vd, vq, vs, scal, partial, partialDq, scratch, out2, dbg;
  DVec<double> fr, fd, fq, fs, fPartial, fPartialDq;                 // full-system PCG vectors r, d, q, s (vectorSize each) and its partial sums
  DVec<double> hsd, hdl, aux;                                        // Dogleg: steepest-descent step, final step, auxiliary vector (vectorSize each)
  DVec<int32_t> mhRow, mhBegin, mhEnd, mhRowPtr, mhColIdx; bool mhReady = false;   // SpMV work items over the Hpp pattern (multiplyHessian in Schur mode)
  DVec<int32_t> hppDiag, hplColPtr, hplRowIdx, sRowPtr, sColIdx, sDiag, hppToS, pairEdgeI, pairEdgeJ, aRowPtr, aColIdx, aDiag, spRow, spBegin, spEnd, hplLm, tEntLm, tEntBI, tEntBJ, tChunkI, tChunkJ, tChunkB, tChunkE, tChunkSlots, pairSegB, pairSegS, hplShortIdx, pairW;
  DVec<uint32_t> tEntMJ; DVec<uint8_t> tEntMI;
  DVec<int64_t> off64;
  std::vector<EdgeSetState*> sets;
  SystemDev sys; SchurDev schur; PcgDev pcg;
  double* hostScal = nullptr;     // pinned
  // counters
  int64_t launches = 0;
  std::map<std::string, PhaseAcc> phases;
  std::vector<PendingEvent> pending; std::vector<cudaEvent_t> eventPool;

  ~g2ocu_solver() {
    for (auto* s : sets) delete s;
    for (auto* b2 : poseBackup) delete b2;
    for (auto* b2 : lmBackup) delete b2;
    for (auto& pe : pending) { cudaEventDestroy(pe.a); cudaEventDestroy(pe.b); }
    for (auto e : eventPool) cudaEventDestroy(e);
    for (void* o : p2pOpened) if (o) cudaIpcCloseMemHandle(o);
    for (void* o : schurOpened) if (o) cudaIpcCloseMemHandle(o);
    if (p2pLocal) cudaFree(p2pLocal);
    if (pcgGraph) cudaGraphExecDestroy(pcgGraph);
    if (ncclComm && g_nccl.CommDestroy) g_nccl.CommDestroy(ncclComm);
    if (hostScal) cudaFreeHost(hostScal);
    if (hostInfo) cudaFreeHost(hostInfo);
    if (side.fork) cudaEventDestroy(side.fork);
    if (side.join) cudaEventDestroy(side.join);
    if (side.stream) cudaStreamDestroy(side.stream);
    if (ownStream && stream) cudaStreamDestroy(stream);
  }
};

namespace {

int fail(g2ocu_solver* s, int code, const std::string& msg) { if (s) s->err = msg; else g_createError = msg; return code; }
bool terminateRequested(const g2ocu_solver* s) { return s->forceStop && *s->forceStop != 0; }   // SparseOptimizer::terminate()
#define CU(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) return fail(s, G2OCU_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e)); } while (0)

int ensureCuda(g2ocu_solver* s) {
  if (s->cudaReady) return G2OCU_OK;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) return fail(s, G2OCU_E_CUDA, std::string("no usable CUDA device (") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") + "); this backend has no CPU fallback");
  if (s->cfg.device >= 0) CU(cudaSetDevice(s->cfg.device));
  if (s->cfg.stream) { s->stream = (cudaStream_t)s->cfg.stream; s->ownStream = false; }
  else { CU(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking)); s->ownStream = true; }
  CU(cudaStreamCreateWithFlags(&s->side.stream, cudaStreamNonBlocking));
  CU(cudaEventCreateWithFlags(&s->side.fork, cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&s->side.join, cudaEventDisableTiming));
  CU(cudaMallocHost((void**)&s->hostScal, 64 * sizeof(double)));
  CU(cudaMallocHost((void**)&s->hostInfo, 16 * sizeof(int)));
  s->cudaReady = true;
  return G2OCU_OK;
}

cudaEvent_t getEvent(g2ocu_solver* s) {
  if (!s->eventPool.empty()) { cudaEvent_t e = s->eventPool.back(); s->eventPool.pop_back(); return e; }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}
struct PhaseTimer {
  g2ocu_solver* s; std::string phase; cudaEvent_t a; int64_t l0;
  PhaseTimer(g2ocu_solver* s_, const char* ph) : s(s_), phase(ph) { a = getEvent(s); cudaEventRecord(a, s->stream); l0 = s->launches; }
  ~PhaseTimer() { cudaEvent_t b = getEvent(s); cudaEventRecord(b, s->stream); s->pending.push_back({phase, a, b, s->launches - l0}); }
};
// per-kernel timing (inside a phase): two event records per kernel are measurable in a PCG iteration of a few tens of microseconds,
// so these only run when the "kernelTiming" property is set (bench.py does a separate breakdown pass with it)
struct KernelTimer {
  PhaseTimer* t = nullptr;
  KernelTimer(g2ocu_solver* s, const char* ph) { if (s->kernelTiming) t = new PhaseTimer(s, ph); }
  ~KernelTimer() { delete t; }
};
// call only after the stream has been synchronised
void resolveEvents(g2ocu_solver* s) {
  for (auto& pe : s->pending) {
    float ms = 0; if (cudaEventElapsedTime(&ms, pe.a, pe.b) == cudaSuccess) { auto& acc = s->phases[pe.phase]; acc.seconds += ms * 1e-3; acc.launches += pe.launches; acc.calls += 1; }
    s->eventPool.push_back(pe.a); s->eventPool.push_back(pe.b);
  }
  s->pending.clear();
}
int syncStream(g2ocu_solver* s) { CU(cudaStreamSynchronize(s->stream)); resolveEvents(s); return G2OCU_OK; }
double phaseSeconds(g2ocu_solver* s, const char* ph) { auto it = s->phases.find(ph); return it == s->phases.end() ? 0.0 : it->second.seconds; }

// The peer-memory exchange buffers are sized for one structure: any change of graph, shard or structure drops the mappings (the host
// exports / imports again after the next g2ocu_build_structure).
void dropP2p(g2ocu_solver* s) {
  if (s->pcgGraph) { cudaGraphExecDestroy(s->pcgGraph); s->pcgGraph = nullptr; }   // its kernel arguments hold the old pointers
  s->p2pReady = false;
  for (void*& o : s->p2pOpened) if (o) { cudaIpcCloseMemHandle(o); o = nullptr; }
  s->p2p = P2pDev();
  s->schurP2pReady = false;
  for (void*& o : s->schurOpened) if (o) { cudaIpcCloseMemHandle(o); o = nullptr; }
  s->schurPeers = P2pDev();
}
int collectiveDev(g2ocu_solver* s, double* buf, int64_t count, int op);
int allreduceDev(g2ocu_solver* s, double* buf, int64_t count, int op) { return collectiveDev(s, buf, count, op); }
int collectiveDev(g2ocu_solver* s, double* buf, int64_t count, int op) {
  if (s->world <= 1) return G2OCU_OK;
  if (s->ncclComm) {
    const int kF64 = 8, kSum = 0, kMax = 2;   // ncclFloat64, ncclSum, ncclMax (nccl.h)
    int rc;
    if (op == G2OCU_OP_REDUCE_SCATTER_SUM) rc = g_nccl.ReduceScatter(buf, buf + (size_t)s->rank * count, (size_t)count, kF64, kSum, s->ncclComm, s->stream);
    else rc = g_nccl.AllReduce(buf, buf, (size_t)count, kF64, op == G2OCU_OP_MAX ? kMax : kSum, s->ncclComm, s->stream);
    if (rc != 0) return fail(s, G2OCU_E_COMM, std::string("NCCL: ") + g_nccl.GetErrorString(rc));
    return G2OCU_OK;
  }
  if (!s->allreduce) return fail(s, G2OCU_E_COMM, "world > 1 but neither an allreduce hook nor an NCCL communicator was set");
  if (s->allreduce(buf, count, op, (void*)s->stream, s->allreduceUser) != 0) return fail(s, G2OCU_E_COMM, "allreduce hook reported an error");
  return G2OCU_OK;
}

// ---- upload estimates of one class from the host graph ----
int uploadEstimates(g2ocu_solver* s) {
  const Structure& st = s->st; const HostGraph& g = s->g;
  auto pack = [&](const std::vector<int32_t>& verts, int vtype, std::vector<double>& out) {
    const int S = vertexEstimateDim(vtype); out.resize((size_t)verts.size() * S);
    for (size_t i = 0; i < verts.size(); ++i) std::memcpy(&out[i * S], &g.vEst[g.vEstOff[verts[i]]], sizeof(double) * S);
  };
  std::vector<double> hp, hl;
  pack(st.poseVerts, st.poseType, hp);
  CU(s->poseEst.upload(hp, s->stream));
  if (st.lmType) { pack(st.lmVerts, st.lmType, hl); CU(s->lmEst.upload(hl, s->stream)); }
  CU(cudaStreamSynchronize(s->stream));   // host staging vectors go out of scope
  s->sys.poseEst = s->poseEst.p; s->sys.lmEst = s->lmEst.p;
  s->errorsValid = false;
  return G2OCU_OK;
}

// device estimates -> host graph copy (the host copy is what a rebuild uploads again)
// sharded runs: zero the landmarks owned by other ranks and sum, afterwards every rank holds all of them (on the device)
int gatherLandmarks(g2ocu_solver* s) {
  const Structure& st = s->st;
  if (s->world <= 1 || st.numLandmarks == 0) return G2OCU_OK;
  const int Sl = vertexEstimateDim(st.lmType);
  if (st.lmBegin > 0) CU(cudaMemsetAsync(s->lmEst.p, 0, sizeof(double) * (size_t)st.lmBegin * Sl, s->stream));
  if (st.lmEnd < st.numLandmarks) CU(cudaMemsetAsync(s->lmEst.p + (size_t)st.lmEnd * Sl, 0, sizeof(double) * (size_t)(st.numLandmarks - st.lmEnd) * Sl, s->stream));
  return allreduceDev(s, s->lmEst.p, (int64_t)st.numLandmarks * Sl, 0);
}
int downloadEstimates(g2ocu_solver* s) {
  const Structure& st = s->st; HostGraph& g = s->g;
  { int rc = gatherLandmarks(s); if (rc) return rc; }
  std::vector<double> hp(s->poseEst.n), hl(s->lmEst.n);
  if (hp.size()) CU(cudaMemcpyAsync(hp.data(), s->poseEst.p, hp.size() * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  if (hl.size()) CU(cudaMemcpyAsync(hl.data(), s->lmEst.p, hl.size() * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  int rc = syncStream(s); if (rc) return rc;
  const int Sp = vertexEstimateDim(st.poseType), Sl = st.lmType ? vertexEstimateDim(st.lmType) : 0;
  for (size_t i = 0; i < st.poseVerts.size() && hp.size(); ++i) std::memcpy(&g.vEst[g.vEstOff[st.poseVerts[i]]], &hp[i * Sp], sizeof(double) * Sp);
  for (size_t i = 0; i < st.lmVerts.size() && hl.size(); ++i) std::memcpy(&g.vEst[g.vEstOff[st.lmVerts[i]]], &hl[i * Sl], sizeof(double) * Sl);
  return G2OCU_OK;
}

// G2OCU_TRACE=1: wall time of the phases of the device-side set-up on stderr (host_structure.cpp prints the ones before it)
struct DeviceTrace {
  static bool on() { static const bool v = [] { const char* e = std::getenv("G2OCU_TRACE"); return e && *e && *e != '0'; }(); return v; }
  std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
  void mark(const char* what) {
    if (!on()) return;
    const auto n = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[g2ocu] buildDevice    %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
    t = n;
  }
};

int buildDevice(g2ocu_solver* s) {
  const Structure& st = s->st; const HostGraph& g = s->g;
  cudaStream_t stream = s->stream;
  DeviceTrace trace;
  for (auto* es : s->sets) delete es;
  s->sets.clear(); s->mhReady = false;
  int rc = uploadEstimates(s); if (rc) return rc;
  if (st.poseType == G2OCU_VERTEX_SE3) { CU(s->poseCounters.alloc(st.numPoseSlots)); CU(s->poseCounters.zero(stream)); }
  const int P = st.P, L = st.L;
  // ---- system matrices ----
  CU(s->hppDiag.upload(st.hppDiag, stream));
  CU(s->Hpp.alloc((size_t)st.hppColIdx.size() * P * P + 2));   // +2: see S
  CU(s->b.alloc((size_t)st.sizePoses + st.sizeLandmarks)); CU(s->x.alloc((size_t)st.sizePoses + st.sizeLandmarks));
  CU(s->b.zero(stream)); CU(s->x.zero(stream)); CU(s->Hpp.zero(stream));
  SystemDev& sys = s->sys;
  sys.numPoses = st.numPoses; sys.numLandmarks = st.numLandmarks; sys.numPoseSlots = st.numPoseSlots; sys.numLmSlots = st.numLmSlots; sys.P = P; sys.L = L;
  sys.Hpp = s->Hpp.p; sys.hppDiag = s->hppDiag.p; sys.b = s->b.p; sys.hplShared = st.hplShared; sys.hppShared = st.hppShared;
  if (st.doSchur) {
    CU(s->Hll.alloc((size_t)st.numLandmarks * L * L)); CU(s->Hpl.alloc((size_t)st.hplRowIdx.size() * P * L + 2));   // +2: 16-byte aligned bulk copies may read one double past the last block
    CU(s->Hll.zero(stream)); CU(s->Hpl.zero(stream));
    sys.Hll = s->Hll.p; sys.Hpl = s->Hpl.p;
  } else { sys.Hll = nullptr; sys.Hpl = nullptr; }

  trace.mark("estimates, matrices");
  // ---- edge sets ----
  size_t maxScratch = 4096;
  for (const EdgeSet& hs : st.sets) {
    if (hs.pos.empty()) continue;                 // sharded run: this rank owns no edge of the type
    auto* es = new EdgeSetState; s->sets.push_back(es);
    const int n = (int)hs.pos.size(); const int t = hs.etype;
    const int E = edgeDim(t), M = edgeMeasDim(t), NP = edgeParamDim(t);
    EdgeSetDev& d = es->dev;
    d.etype = t; d.n = n; d.poseLandmark = hs.poseLandmark;
    CU(es->slot0.upload(hs.slot0, stream)); CU(es->slot1.upload(hs.slot1, stream)); CU(es->block.upload(hs.block, stream));
    CU(es->transposed.upload(hs.transposed, stream)); CU(es->pos.upload(hs.pos, stream));
    d.slot0 = es->slot0.p; d.slot1 = es->slot1.p; d.block = es->block.p; d.transposed = es->transposed.p; d.pos = es->pos.p;
    std::vector<double> meas((size_t)n * M), info((size_t)n * E * E), prm((size_t)n * NP), delta(n);
    std::vector<int32_t> kind(n);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
      const int e = st.activeEdges[hs.pos[i]];
      std::memcpy(&meas[(size_t)i * M], &g.eMeas[g.eMeasOff[e]], sizeof(double) * M);
      std::memcpy(&info[(size_t)i * E * E], &g.eInfo[g.eInfoOff[e]], sizeof(double) * E * E);
      if (NP) std::memcpy(&prm[(size_t)i * NP], &g.ePrm[g.ePrmOff[e]], sizeof(double) * NP);
      kind[i] = g.eKernel[e]; delta[i] = g.eDelta[e];
    }
    CU(es->meas.upload(meas, stream)); d.meas = es->meas.p;
    // information: identity / uniform / per edge
    bool identity = true;
    int notUniform = 0;
#pragma omp parallel for schedule(static) reduction(| : notUniform)
    for (int i = 0; i < n; ++i) notUniform |= std::memcmp(&info[(size_t)i * E * E], &info[0], sizeof(double) * E * E) != 0;
    const bool uniform = !notUniform;
    for (int k = 0; k < E * E && identity; ++k) identity = info[k] == ((k % (E + 1)) == 0 ? 1.0 : 0.0);
    if (uniform && identity) d.infoMode = 0;
    else if (uniform) { info.resize((size_t)E * E); d.infoMode = 1; }
    else d.infoMode = 2;
    if (d.infoMode) { CU(es->info.upload(info, stream)); d.info = es->info.p; }
    int kDiffers = 0;
#pragma omp parallel for schedule(static) reduction(| : kDiffers)
    for (int i = 1; i < n; ++i) kDiffers |= !(kind[i] == kind[0] && delta[i] == delta[0]);
    const bool kUniform = !kDiffers;
    if (kUniform) { d.kernelMode = kind[0] ? 1 : 0; d.kKind = kind[0]; d.kDelta = delta[0]; }
    else { d.kernelMode = 2; CU(es->kernelKind.upload(kind, stream)); CU(es->kernelDelta.upload(delta, stream)); d.kernelKind = es->kernelKind.p; d.kernelDelta = es->kernelDelta.p; }
    if (NP) {
      int pDiffers = 0;
#pragma omp parallel for schedule(static) reduction(| : pDiffers)
      for (int i = 1; i < n; ++i) pDiffers |= std::memcmp(&prm[(size_t)i * NP], &prm[0], sizeof(double) * NP) != 0;
      const bool pUniform = !pDiffers;
      if (pUniform) { prm.resize(NP); d.prmMode = 1; } else d.prmMode = 2;
      CU(es->prm.upload(prm, stream)); d.prm = es->prm.p;
    }
    if (hs.poseLandmark) {
      CU(es->byPose.upload(hs.byPose, stream)); CU(es->chunkPose.upload(hs.chunkPose, stream)); CU(es->chunkBegin.upload(hs.chunkBegin, stream));
      CU(es->chunkEnd.upload(hs.chunkEnd, stream)); CU(es->poseChunkPtr.upload(hs.poseChunkPtr, stream));
      d.byPose = es->byPose.p; d.nByPose = (int)hs.byPose.size(); d.chunkPose = es->chunkPose.p; d.chunkBegin = es->chunkBegin.p; d.chunkEnd = es->chunkEnd.p;
      d.nChunks = (int)hs.chunkPose.size(); d.poseChunkPtr = es->poseChunkPtr.p;
      CU(es->partial.alloc((size_t)d.nChunks * (P * (P + 1) / 2 + P))); d.partial = es->partial.p;
    }
    CU(cudaStreamSynchronize(stream));   // staging vectors die here
    maxScratch = std::max(maxScratch, (size_t)errorScratchDoubles(n));
  }
  trace.mark("edge sets");
  CU(s->scratch.alloc(maxScratch)); CU(s->out2.alloc(16));   // [0,1] chi2, [4] max diagonal, [5,6] computeScale, [8..15] Dogleg dot products
  {
    auto contiguous = [&](const std::vector<int32_t>& verts, int vtype, int64_t& off) {
      if (verts.empty()) { off = 0; return true; }
      const int S = vertexEstimateDim(vtype); off = g.vEstOff[verts[0]];
      for (size_t i = 0; i < verts.size(); ++i) if (g.vEstOff[verts[i]] != off + (int64_t)i * S) return false;
      return true;
    };
    const bool cp = contiguous(st.poseVerts, st.poseType, s->poseHostOff), cl = contiguous(st.lmVerts, st.lmType ? st.lmType : st.poseType, s->lmHostOff);
    s->fastEstimates = cp && cl && (s->poseEst.n + s->lmEst.n == g.vEst.size());
  }

  // ---- Schur structures ----
  SchurDev& sd = s->schur; sd = SchurDev();
  if (st.doSchur) {
    CU(s->hplColPtr.upload(st.hplColPtr, stream)); CU(s->hplRowIdx.upload(st.hplRowIdx, stream));
    { std::vector<int32_t> lmOf(st.hplRowIdx.size()); for (int l = 0; l < st.numLandmarks; ++l) for (int k = st.hplColPtr[l]; k < st.hplColPtr[l + 1]; ++k) lmOf[k] = l;
      CU(s->hplLm.upload(lmOf, stream)); CU(cudaStreamSynchronize(stream)); }
    CU(s->sRowPtr.upload(st.sRowPtr, stream)); CU(s->sColIdx.upload(st.sColIdx, stream)); CU(s->sDiag.upload(st.sDiag, stream)); CU(s->hppToS.upload(st.hppToS, stream));
    // short tracks (< kTileMinTrack observations): flat pair list, landmarks visited in the order of their first camera;
    // long tracks: entries of the output-stationary tile kernel
    const bool useMma = schurMmaSupported(P, L);
    const int tileRows = useMma ? kMmaTileRows : kTileRows;
    static const int tileMinTrackEnv = [] { const char* e = getenv("G2OCU_TILE_MIN_TRACK"); return e ? atoi(e) : 0; }();   // developer switch
    const int tileMinTrack = tileMinTrackEnv > 0 ? tileMinTrackEnv : kTileMinTrack;
    s->tileMinTrack = tileMinTrack;
    struct TileEntry { int64_t key; int32_t lm, baseI, baseJ; uint32_t maskJ; uint8_t maskI; };
    std::vector<TileEntry> entries;
    std::vector<int32_t> shortLm;
    for (int l = st.lmBegin; l < st.lmEnd; ++l) {   // owned landmarks only
      const int cb = st.hplColPtr[l]; const int64_t k = st.hplColPtr[l + 1] - cb;
      if (k == 0) continue;
      if (k < tileMinTrack) { shortLm.push_back(l); continue; }
      struct Run { int32_t tile, base; uint32_t mask; };
      std::vector<Run> rowsR, colsR;
      for (int i = 0; i < k; ++i) {
        const int c = st.hplRowIdx[cb + i];
        const int ti = c / tileRows, tj = c / kTileCols;
        if (rowsR.empty() || rowsR.back().tile != ti) rowsR.push_back({ti, cb + i, 0u});
        rowsR.back().mask |= 1u << (c % tileRows);
        if (colsR.empty() || colsR.back().tile != tj) colsR.push_back({tj, cb + i, 0u});
        colsR.back().mask |= 1u << (c % kTileCols);
      }
      for (const Run& ri : rowsR)
        for (const Run& rj : colsR) {
          if ((rj.tile + 1) * kTileCols - 1 < ri.tile * tileRows) continue;       // strip entirely left of the row tile: lower triangle
          int32_t baseJ = rj.base; uint32_t maskJ = rj.mask;
          if (useMma && rj.tile * kTileCols < ri.tile * tileRows) {
            // diagonal strip: columns left of the row group only form lower-triangle pairs - drop them from the entry
            const uint32_t drop = maskJ & ((1u << (ri.tile * tileRows - rj.tile * kTileCols)) - 1u);
            baseJ += __builtin_popcount(drop); maskJ &= ~drop;
            if (maskJ == 0) continue;
          }
          entries.push_back({((int64_t)ri.tile << 32) | (uint32_t)rj.tile, l, ri.base, baseJ, maskJ, (uint8_t)ri.mask});
        }
    }
    trace.mark("tile entries");
    {
      // the pairs (block i, block j), i <= j, of every short track, landmarks in index order (their order inside a target block only decides the
      // order of a segment's sum); offsets by a prefix sum so that the lists fill in parallel
      std::vector<int64_t> pairOff(shortLm.size() + 1, 0);
      for (size_t q = 0; q < shortLm.size(); ++q) { const int64_t k = st.hplColPtr[shortLm[q] + 1] - st.hplColPtr[shortLm[q]]; pairOff[q + 1] = pairOff[q] + k * (k + 1) / 2; }
      std::vector<int32_t> pI((size_t)pairOff.back()), pJ((size_t)pairOff.back());
#pragma omp parallel for schedule(static, 4096)
      for (int64_t q = 0; q < (int64_t)shortLm.size(); ++q) {
        const int cb = st.hplColPtr[shortLm[q]], ce = st.hplColPtr[shortLm[q] + 1];
        int64_t o = pairOff[q];
        for (int i = cb; i < ce; ++i) for (int j = i; j < ce; ++j) { pI[o] = i; pJ[o] = j; ++o; }
      }
      {  // sort the pairs by target Hschur block (counting sort) and cut them into segments of one block each
        const size_t np = pI.size();
        std::vector<int32_t> slot(np);
#pragma omp parallel for schedule(static)
        for (int64_t k = 0; k < (int64_t)np; ++k) {
          const int ci = st.hplRowIdx[pI[k]], cj = st.hplRowIdx[pJ[k]];
          const int32_t* rb = &st.sColIdx[st.sRowPtr[ci]]; const int32_t* re = &st.sColIdx[st.sRowPtr[ci + 1]];
          slot[k] = (int32_t)(std::lower_bound(rb, re, cj) - st.sColIdx.data());
        }
        std::vector<int64_t> cnt(st.sColIdx.size() + 1, 0);
        for (size_t k = 0; k < np; ++k) cnt[slot[k] + 1]++;
        for (size_t b2 = 0; b2 < st.sColIdx.size(); ++b2) cnt[b2 + 1] += cnt[b2];
        std::vector<int32_t> sI(np), sJ(np), segB, segS;
        { std::vector<int64_t> fill(cnt.begin(), cnt.end() - 1);
          for (size_t k = 0; k < np; ++k) { const int64_t o = fill[slot[k]]++; sI[o] = pI[k]; sJ[o] = pJ[k]; } }
        for (size_t b2 = 0; b2 < st.sColIdx.size(); ++b2)
          for (int64_t o = cnt[b2]; o < cnt[b2 + 1]; o += kPairSegment) { segB.push_back((int32_t)o); segS.push_back((int32_t)b2); }
        segB.push_back((int32_t)np);
        pI.swap(sI); pJ.swap(sJ);
        CU(s->pairSegB.upload(segB, stream)); CU(s->pairSegS.upload(segS, stream));
        CU(cudaStreamSynchronize(stream));
        sd.nPairSegs = (int)segS.size(); sd.pairSegBegin = s->pairSegB.p; sd.pairSegSlot = s->pairSegS.p;
      }
      CU(s->pairEdgeI.upload(pI, stream)); CU(s->pairEdgeJ.upload(pJ, stream));
      CU(cudaStreamSynchronize(stream));
      sd.nPairs = (int64_t)pI.size(); sd.pairEdgeI = s->pairEdgeI.p; sd.pairEdgeJ = s->pairEdgeJ.p;
      // W = B Dinv of the short tracks' blocks is formed by the coefficient pass for the pair kernel: compact index per Hpl block (in block
      // order, so that the pass writes it front to back), and per pair the index of its row-side block.  (The older tile path forms all of W.)
      if (!pI.empty() && !(useMma && !schurKpackEnabled())) {
        std::vector<int32_t> shortIdx(st.hplRowIdx.size(), -1); int32_t nShort = 0;
        for (int l = st.lmBegin; l < st.lmEnd; ++l) { const int cb = st.hplColPtr[l], ce = st.hplColPtr[l + 1]; if (ce - cb > 0 && ce - cb < tileMinTrack) for (int k = cb; k < ce; ++k) shortIdx[k] = nShort++; }
        std::vector<int32_t> pW(pI.size());
        for (size_t k = 0; k < pI.size(); ++k) pW[k] = shortIdx[pI[k]];
        CU(s->hplShortIdx.upload(shortIdx, stream)); CU(s->pairW.upload(pW, stream)); CU(s->Wshort.alloc((size_t)nShort * P * L + 2)); CU(s->Wshort.zero(stream));
        CU(cudaStreamSynchronize(stream));
        sd.hplShortIdx = s->hplShortIdx.p; sd.pairW = s->pairW.p; sd.Wshort = s->Wshort.p;
      }
    }
    trace.mark("pair lists");
    {  // tile entries grouped by (row tile, column strip), split in chunks of at most kTileChunk entries (one CTA each)
      const int kTileChunk = useMma ? 1024 : 256;
      // inside a tile: landmarks that touch the same groups of 8 column cameras next to each other (the K-packed tile kernel skips a group
      // none of the landmarks sharing a DMMA touches)
      auto groupsOf = [](uint32_t m) { return ((m & 0xffu) ? 1u : 0u) | ((m & 0xff00u) ? 2u : 0u) | ((m & 0xff0000u) ? 4u : 0u) | ((m & 0xff000000u) ? 8u : 0u); };
      std::stable_sort(entries.begin(), entries.end(), [&](const TileEntry& a, const TileEntry& b) { return a.key != b.key ? a.key < b.key : groupsOf(a.maskJ) < groupsOf(b.maskJ); });
      std::vector<int32_t> eLm(entries.size()), eBI(entries.size()), eBJ(entries.size()), cI, cJ, cB, cE; std::vector<uint32_t> eMJ(entries.size()); std::vector<uint8_t> eMI(entries.size());
      for (size_t i = 0; i < entries.size(); ++i) { eLm[i] = entries[i].lm; eBI[i] = entries[i].baseI; eBJ[i] = entries[i].baseJ; eMJ[i] = entries[i].maskJ; eMI[i] = entries[i].maskI; }
      for (size_t i = 0; i < entries.size();) {
        size_t j = i; while (j < entries.size() && entries[j].key == entries[i].key) ++j;
        for (size_t c0 = i; c0 < j; c0 += kTileChunk) { cI.push_back((int32_t)(entries[i].key >> 32)); cJ.push_back((int32_t)(entries[i].key & 0xffffffff)); cB.push_back((int32_t)c0); cE.push_back((int32_t)std::min(j, c0 + kTileChunk)); }
        i = j;
      }
      if (useMma) {   // largest chunks first: one CTA per chunk, dispatched in index order - the small ones fill the gaps at the end of the launch
        std::vector<int32_t> order(cI.size());
        for (size_t i = 0; i < order.size(); ++i) order[i] = (int32_t)i;
        std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return cE[a] - cB[a] > cE[b] - cB[b]; });
        std::vector<int32_t> nI(cI.size()), nJ(cI.size()), nB(cI.size()), nE2(cI.size());
        for (size_t i = 0; i < order.size(); ++i) { nI[i] = cI[order[i]]; nJ[i] = cJ[order[i]]; nB[i] = cB[order[i]]; nE2[i] = cE[order[i]]; }
        cI.swap(nI); cJ.swap(nJ); cB.swap(nB); cE.swap(nE2);
      }
      CU(s->tEntLm.upload(eLm, stream)); CU(s->tEntBI.upload(eBI, stream)); CU(s->tEntBJ.upload(eBJ, stream)); CU(s->tEntMJ.upload(eMJ, stream)); CU(s->tEntMI.upload(eMI, stream));
      CU(s->tChunkI.upload(cI, stream)); CU(s->tChunkJ.upload(cJ, stream)); CU(s->tChunkB.upload(cB, stream)); CU(s->tChunkE.upload(cE, stream));
      if (useMma) {   // Hschur slot of every block of every chunk's tile: the tile kernel's write-out needs no search
        std::vector<int32_t> slots(cI.size() * (size_t)kMmaTileRows * kTileCols, -1);
#pragma omp parallel for schedule(static)
        for (int64_t c = 0; c < (int64_t)cI.size(); ++c)
          for (int w = 0; w < kMmaTileRows; ++w) {
            const int ci = cI[c] * kMmaTileRows + w;
            if (ci >= st.numPoses) continue;
            const int32_t* rb = &st.sColIdx[st.sRowPtr[ci]]; const int32_t* re = &st.sColIdx[st.sRowPtr[ci + 1]];
            for (int n = 0; n < kTileCols; ++n) {
              const int cj = cJ[c] * kTileCols + n;
              if (cj < ci || cj >= st.numPoses) continue;
              const int32_t* it = std::lower_bound(rb, re, cj);
              if (it != re && *it == cj) slots[((size_t)c * kMmaTileRows + w) * kTileCols + n] = (int32_t)(it - st.sColIdx.data());
            }
          }
        CU(s->tChunkSlots.upload(slots, stream)); sd.chunkSlots = s->tChunkSlots.p;
      }
      CU(cudaStreamSynchronize(stream));
      sd.nTileChunks = (int)cI.size();
      sd.chunkI = s->tChunkI.p; sd.chunkJ = s->tChunkJ.p; sd.chunkBegin = s->tChunkB.p; sd.chunkEnd = s->tChunkE.p;
      sd.entLm = s->tEntLm.p; sd.entBaseI = s->tEntBI.p; sd.entBaseJ = s->tEntBJ.p; sd.entMaskJ = s->tEntMJ.p; sd.entMaskI = s->tEntMI.p;
    }
    trace.mark("tile chunks, slot tables");
    if (useMma && !schurKpackEnabled()) { CU(s->W.alloc((size_t)st.hplRowIdx.size() * P * L + 2)); CU(s->W.zero(stream)); sd.W = s->W.p; }
    // multi-GPU: the reduced system is reduce-scattered into equal block ranges (the last one padded), rank r solves with blocks [r c, (r+1) c)
    CU(s->S.alloc((s->world > 1 ? (size_t)s->slabBlocks * s->world * P * P : (size_t)st.sColIdx.size() * P * P) + 2)); CU(s->S.zero(stream));   // +2: 16-byte aligned bulk copies may read one double past the last block
    CU(s->Dinv.alloc((size_t)st.numLandmarks * L * L)); CU(s->dbv.alloc((size_t)st.numLandmarks * L)); CU(s->bschur.alloc((size_t)st.sizePoses));
    sd.numPoses = st.numPoses; sd.numLandmarks = st.numLandmarks; sd.P = P; sd.L = L;
    sd.lmBegin = st.lmBegin; sd.lmEnd = st.lmEnd; sd.blockBegin = st.hplColPtr[st.lmBegin];
    sd.hplColPtr = s->hplColPtr.p; sd.hplRowIdx = s->hplRowIdx.p; sd.sRowPtr = s->sRowPtr.p; sd.sColIdx = s->sColIdx.p; sd.sDiag = s->sDiag.p;
    sd.hppToS = s->hppToS.p; sd.nnzHpp = (int)st.hppColIdx.size(); sd.nnzS = (int)st.sColIdx.size();
    sd.S = s->S.p; sd.Dinv = s->Dinv.p; sd.db = s->dbv.p; sd.bschur = s->bschur.p;
    CU(cudaStreamSynchronize(stream));
  }
  trace.mark("Schur buffers");
  // ---- PCG structures over A = Hschur (Schur) or Hpp ----
  PcgDev& pc = s->pcg; pc = PcgDev();
  const std::vector<int32_t>& rowPtr = st.doSchur ? st.sRowPtr : st.hppRowPtr;
  const std::vector<int32_t>& colIdx = st.doSchur ? st.sColIdx : st.hppColIdx;
  const std::vector<int32_t>& diag = st.doSchur ? st.sDiag : st.hppDiag;
  CU(s->aRowPtr.upload(rowPtr, stream)); CU(s->aColIdx.upload(colIdx, stream)); CU(s->aDiag.upload(diag, stream));
  {
    const int G = 32 / P, chunk = G * 16;
    std::vector<int32_t> r, bgn, en;
    const bool slab = s->world > 1 && st.doSchur;
    const int64_t lo = slab ? s->slabBlocks * s->rank : 0, hi = slab ? std::min<int64_t>(s->slabBlocks * (s->rank + 1), (int64_t)colIdx.size()) : (int64_t)colIdx.size();
    for (int i = 0; i < st.numPoses; ++i) {
      const int kb = (int)std::max<int64_t>(rowPtr[i], lo), ke = (int)std::min<int64_t>(rowPtr[i + 1], hi);
      for (int k = kb; k < ke; k += chunk) { r.push_back(i); bgn.push_back(k); en.push_back(std::min(k + chunk, ke)); }
    }
    pc.ownLo = (int)lo; pc.ownHi = (int)hi;
    CU(s->spRow.upload(r, stream)); CU(s->spBegin.upload(bgn, stream)); CU(s->spEnd.upload(en, stream));
    pc.nItems = (int)r.size();
    CU(cudaStreamSynchronize(stream));
  }
  pc.n = st.sizePoses; pc.nb = st.numPoses; pc.P = P; pc.rowPtr = s->aRowPtr.p; pc.colIdx = s->aColIdx.p; pc.diag = s->aDiag.p; pc.nnz = (int)colIdx.size();
  pc.A = st.doSchur ? s->S.p : s->Hpp.p;
  CU(s->Minv.alloc((size_t)st.numPoses * P * P)); CU(s->vr.alloc(pc.n)); CU(s->vd.alloc(pc.n)); CU(s->vq.alloc(pc.n)); CU(s->vs.alloc(pc.n)); CU(s->scal.alloc(16));
  pc.nPartial = (pc.n + 255) / 256;   // one partial per CTA of the thread-per-scalar-row kernels
  pc.nPartialDq = std::min(296, (pc.n + 255) / 256);
  CU(s->partial.alloc(pc.nPartial)); CU(s->partialDq.alloc(pc.nPartialDq));
  pc.Minv = s->Minv.p; pc.r = s->vr.p; pc.d = s->vd.p; pc.q = s->vq.p; pc.s = s->vs.p; pc.x = s->x.p; pc.scal = s->scal.p; pc.partial = s->partial.p; pc.partialDq = s->partialDq.p;
  pc.itemRow = s->spRow.p; pc.itemBegin = s->spBegin.p; pc.itemEnd = s->spEnd.p;
  CU(s->scal.zero(stream));
  CU(s->pcgTicket.alloc(4)); CU(s->pcgTicket.zero(stream)); pc.ticket = s->pcgTicket.p;
  CU(cudaStreamSynchronize(stream));
  CU(cudaGetLastError());
  trace.mark("PCG structures");
  return G2OCU_OK;
}

// SparseOptimizer::computeActiveErrors etc. work right after initializeOptimization in the reference; the device state they need is
// created on demand here (the index map is rebuilt by the next explicit buildStructure anyway)
int requireBuilt(g2ocu_solver* s) {
  if (!s) return G2OCU_E_INVALID;
  if (s->structureBuilt) return G2OCU_OK;
  if (!s->optInitialized) return fail(s, G2OCU_E_INVALID, "initializeOptimization has not been called");
  return g2ocu_build_structure(s);
}

int computeErrors(g2ocu_solver* s, double* errOut, const int64_t* errOff) {
  PhaseTimer pt(s, "errors");
  CU(s->out2.zero(s->stream));
  for (auto* es : s->sets) launchErrors(es->dev, s->sys, s->scratch.p, s->out2.p, errOut, errOff, s->stream, &s->launches);
  int rc = allreduceDev(s, s->out2.p, 2, 0); if (rc) return rc;
  CU(cudaMemcpyAsync(s->hostScal, s->out2.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  return G2OCU_OK;
}
int finishErrors(g2ocu_solver* s) {
  int rc = syncStream(s); if (rc) return rc;
  s->chi2Robust = s->hostScal[0]; s->chi2Plain = s->hostScal[1]; s->errorsValid = true;
  return G2OCU_OK;
}

int buildSystem(g2ocu_solver* s) {
  PhaseTimer pt(s, "build");
  CU(s->Hpp.zero(s->stream)); CU(s->b.zero(s->stream));
  if (s->st.doSchur) { CU(s->Hll.zero(s->stream)); if (s->st.hplShared) CU(s->Hpl.zero(s->stream)); }
  for (auto* es : s->sets) launchBuild(es->dev, s->sys, s->stream, &s->launches);
  CU(cudaGetLastError());
  if (s->world > 1 && !s->st.doSchur) {   // pose graphs are sharded by edge: every rank needs the full Hpp and b for the replicated solve
    int rc = allreduceDev(s, s->Hpp.p, (int64_t)s->st.hppColIdx.size() * s->st.P * s->st.P, 0); if (rc) return rc;
    rc = allreduceDev(s, s->b.p, (int64_t)s->b.n, 0); if (rc) return rc;
  }
  return G2OCU_OK;
}

const int kPcgGraphIters = 4;
// The CG iterations between two convergence polls as one graph launch (the kernels are tens of microseconds long at most - at 8 GPUs
// shorter than their launch gaps).  Single GPU and the peer-memory slab PCG; not with NCCL in the loop, not while per-kernel timing is on.
int ensurePcgGraph(g2ocu_solver* s, bool p2p, int64_t* kernelsPerLaunch) {
  PcgDev& pc = s->pcg;
  *kernelsPerLaunch = s->pcgGraphKernels;
  if (s->pcgGraph && s->pcgGraphLambda == pc.lambda && s->pcgGraphP2p == p2p && s->pcgGraphA == pc.A && s->pcgGraphN == pc.n) return G2OCU_OK;
  if (s->pcgGraph) { cudaGraphExecDestroy(s->pcgGraph); s->pcgGraph = nullptr; }
  int64_t dummy = 0;
  CU(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
  for (int k = 0; k < kPcgGraphIters; ++k) {
    launchSpmv(pc, pc.d, pc.q, s->stream, &dummy, true);
    if (p2p && pcgFusedTail(pc)) launchP2pPushAndTail(pc, s->p2p, s->stream, &dummy);
    else if (p2p) { launchP2pExchangeDot(pc, s->p2p, s->stream, &dummy); launchPcgTail(pc, s->stream, &dummy, true); }   // exchange fused with d.q, then the split tail
    else launchPcgTail(pc, s->stream, &dummy, false);
  }
  cudaGraph_t graph = nullptr;
  CU(cudaStreamEndCapture(s->stream, &graph));
  const cudaError_t e = cudaGraphInstantiate(&s->pcgGraph, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) { s->pcgGraph = nullptr; return fail(s, G2OCU_E_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e)); }
  s->pcgGraphLambda = pc.lambda; s->pcgGraphP2p = p2p; s->pcgGraphA = pc.A; s->pcgGraphN = pc.n;
  s->pcgGraphKernels = dummy; *kernelsPerLaunch = dummy;     // what the launch helpers counted during the capture
  return G2OCU_OK;
}

int solvePcg(g2ocu_solver* s, const double* rhs) {
  PcgDev& pc = s->pcg;
  pc.lambda = s->st.doSchur ? 0.0 : s->lambda;
  const bool slab = s->world > 1 && s->st.doSchur;   // row-range SpMV per rank, q summed over the ranks; all vector recurrences replicated
  { KernelTimer pt(s, "pcg_setup");
    launchBlockInverse(pc, s->stream, &s->launches);
    if (slab) { int rc = allreduceDev(s, pc.Minv, (int64_t)pc.nb * pc.P * pc.P, 0); if (rc) return rc; }   // every rank inverts the diagonal blocks it owns
    launchPcgInit(pc, rhs, s->cfg.pcg_tolerance, s->pcgResidual, s->cfg.pcg_absolute_tolerance, s->stream, &s->launches); }
  // (Measured and dropped: pinning a share of the matrix in L2 with an access-policy window for the duration of the solve - 761 MB against a
  // 126 MB L2 on C3 - changes the product by less than 2 % and costs the Schur phase 4 % through the persisting carve-out.)
  const int maxIter = s->cfg.pcg_max_iterations < 0 ? pc.n : s->cfg.pcg_max_iterations;
  int issued = 0; bool done = false;
  const int kCheckEvery = 4;
  while (!done) {
    // the first poll comes where the previous solve converged (the counts grow slowly from one LM iteration to the next); launches past
    // convergence are no-ops on the device, so overshooting costs microseconds while every poll drains the stream
    const int want = issued == 0 ? std::min(std::max(kCheckEvery, s->lastPcgIterations - 1), 256) : kCheckEvery;
    const int batch = std::min(want, maxIter - issued);
    const bool p2p = slab && s->p2pReady && pc.n <= s->p2p.cap;
    static const bool graphsOn = [] { const char* e = getenv("G2OCU_PCG_GRAPH"); return !(e && e[0] == '0'); }();
    const bool useGraph = graphsOn && !s->kernelTiming && (slab ? p2p : true);
    for (int k = 0; k < batch; ++k) {
      if (useGraph && issued + k > 0 && batch - k >= kPcgGraphIters) {      // (the first product of a solve clears q itself)
        int64_t graphLaunches = 0;
        int rc = ensurePcgGraph(s, p2p, &graphLaunches); if (rc) return rc;
        CU(cudaGraphLaunch(s->pcgGraph, s->stream));
        s->launches += graphLaunches; k += kPcgGraphIters - 1;
        continue;
      }
      { KernelTimer pt(s, "pcg_spmv"); launchSpmv(pc, pc.d, pc.q, s->stream, &s->launches, issued + k > 0 && pcgSingleCtaTail(pc)); }
      if (p2p && pcgFusedTail(pc)) { KernelTimer pt(s, "pcg_vec"); launchP2pPushAndTail(pc, s->p2p, s->stream, &s->launches); continue; }   // push + one kernel: wait for the peers, sum, d.q, recurrences
      if (p2p) { KernelTimer pt(s, "pcg_exchange"); launchP2pExchangeDot(pc, s->p2p, s->stream, &s->launches); }   // peer-memory all-reduce of q fused with d.q
      else if (slab) { KernelTimer pt(s, "pcg_exchange"); int rc = allreduceDev(s, pc.q, pc.n, 0); if (rc) return rc; }
      { KernelTimer pt(s, "pcg_vec"); launchPcgTail(pc, s->stream, &s->launches, p2p); }
    }
    issued += batch;
    CU(cudaMemcpyAsync(s->hostScal + 8, pc.scal, 8 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    int rc = syncStream(s); if (rc) return rc;
    if (s->hostScal[8 + 6] == 3.0) return fail(s, G2OCU_E_COMM, "slab PCG: a peer rank did not publish its partial product within the wait budget (about 2 s) - the rank has probably failed");
    const bool converged = s->hostScal[8 + 6] != 0.0;
    if (converged || issued >= maxIter) done = true;
    if (!std::isfinite(s->hostScal[8 + 2])) done = true;   // NaN/Inf in the recurrence: stop issuing work (the reference would spin to maxIter)
  }
  s->lastPcgIterations = (int)s->hostScal[8 + 7]; s->totalPcgIterations += s->lastPcgIterations;
  s->pcgResidual = 0.5 * s->hostScal[8 + 2];
  return G2OCU_OK;
}

int solveFullPcg(g2ocu_solver* s);

int solveSystem(g2ocu_solver* s, int* solved) {
  const Structure& st = s->st;
  *solved = 1;
  const bool dense = s->cfg.linear_solver == G2OCU_LINEAR_DENSE;
  // Points that are not marginalized: the reference's BlockSolver has no Schur step and hands the whole matrix to its LinearSolver.
  // PCG iterates on that whole system (its iterates differ from those of a reduced system); an exact factorisation gives the same
  // x either way, so the dense solver eliminates the points (Schur + Cholesky of the reduced system + back-substitution below).
  if (st.fullSystem && !dense) { PhaseTimer pt(s, "linear_solver"); return solveFullPcg(s); }
  // LinearSolverDense::solve (linear_solver_dense.h:65-115): dense copy of the solved matrix, Cholesky, x = A^-1 rhs; false when not positive
  auto solveDense = [&](const double* rhs) -> int {
    PcgDev& pc = s->pcg;
    pc.lambda = st.doSchur ? 0.0 : s->lambda;
    const int64_t n = pc.n;
    if (n > kDenseMaxN) return fail(s, G2OCU_E_UNSUPPORTED, "dense Cholesky is limited to systems of dimension <= " + std::to_string(kDenseMaxN) + " (this one has " + std::to_string(n) + "); use the PCG solver");
    CU(s->denseH.alloc((size_t)n * n)); CU(s->denseInfo.alloc(1));
    { PhaseTimer pt(s, "dense_assemble"); launchDenseAssemble(pc, s->denseH.p, s->stream, &s->launches); }
    { PhaseTimer pt(s, "dense_cholesky"); launchDenseCholeskySolve(s->denseH.p, (int)n, rhs, s->x.p, s->denseInfo.p, s->stream, &s->launches); }
    CU(cudaMemcpyAsync(s->hostInfo, s->denseInfo.p, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    int rc = syncStream(s); if (rc) return rc;
    if (*s->hostInfo != 0) *solved = 0;
    s->lastPcgIterations = 0;
    return G2OCU_OK;
  };
  auto broadcastPoseStep = [&]() -> int {   // the pose system is solved redundantly on every rank; rank 0's solution wins so that replicas stay bitwise identical
    if (s->world <= 1) return G2OCU_OK;
    if (s->rank != 0) CU(cudaMemsetAsync(s->x.p, 0, sizeof(double) * (size_t)st.sizePoses, s->stream));
    return allreduceDev(s, s->x.p, st.sizePoses, 0);
  };
  if (!st.doSchur) {
    PhaseTimer pt(s, "linear_solver");
    int rc = dense ? solveDense(s->b.p) : solvePcg(s, s->b.p); if (rc) return rc;
    return broadcastPoseStep();
  }
  { PhaseTimer pt(s, "schur");
    struct MarkCtx { g2ocu_solver* s; PhaseTimer* t; } mc{s, nullptr};
    KernelMarks marks; marks.ctx = &mc;
    marks.begin = [](void* c, const char* name) { auto* m = (MarkCtx*)c; m->t = m->s->kernelTiming ? new PhaseTimer(m->s, name) : nullptr; };
    marks.end = [](void* c) { auto* m = (MarkCtx*)c; delete m->t; m->t = nullptr; };
    launchSchur(s->schur, s->sys, s->hplLm.p, st.hplColPtr[st.lmEnd] - st.hplColPtr[st.lmBegin], s->lambda, s->rank == 0 ? s->lambda : 0.0, s->stream, &s->launches, s->kernelTiming ? &marks : nullptr, &s->side);
    if (s->world > 1) {
      KernelTimer pt2(s, "schur_exchange");
      if (s->schurP2pReady && !dense) {
        // Peer-memory reduction.  The all-reduce of b_schur doubles as the barrier in front of it (it completes on this rank only after
        // every rank has enqueued it behind its own Schur kernels); the all-reduce of the block inverses at the start of solvePcg is the
        // barrier behind it (no rank clears its buffer for the next trial before every rank has read it).
        int rc = allreduceDev(s, s->bschur.p, (int64_t)s->bschur.n, 0); if (rc) return rc;
        const size_t bs = (size_t)st.P * st.P, lo = (size_t)s->slabBlocks * s->rank * bs, hi = std::min((size_t)s->slabBlocks * (s->rank + 1), st.sColIdx.size()) * bs;
        if (hi > lo) launchSlabReduce(s->schurPeers, lo, hi - lo, s->stream, &s->launches);
      } else {
        int rc = collectiveDev(s, s->S.p, s->slabBlocks * st.P * st.P, G2OCU_OP_REDUCE_SCATTER_SUM); if (rc) return rc;   // rank r keeps the sum of its block range
        rc = allreduceDev(s, s->bschur.p, (int64_t)s->bschur.n, 0); if (rc) return rc;
      }
    } }
  { PhaseTimer pt(s, "linear_solver");
    int rc = dense ? solveDense(s->bschur.p) : solvePcg(s, s->bschur.p); if (rc) return rc;
    rc = broadcastPoseStep(); if (rc) return rc; }
  { PhaseTimer pt(s, "backsub");
    launchBacksub(s->schur, s->sys, s->hplLm.p, st.hplColPtr[st.lmEnd] - st.hplColPtr[st.lmBegin], s->x.p, s->x.p + st.sizePoses, s->stream, &s->launches); }
  CU(cudaGetLastError());
  return G2OCU_OK;
}

int applyUpdate(g2ocu_solver* s, const double* step = nullptr) {   // step: device vector of vectorSize doubles, default the solver's x
  PhaseTimer pt(s, "update");
  const Structure& st = s->st;
  if (!step) step = s->x.p;
  launchUpdate(st.poseType, s->poseEst.p, nullptr, s->poseCounters.p, step, st.numPoses, s->stream, &s->launches);
  if (st.lmEnd > st.lmBegin) launchUpdate(st.lmType, s->lmEst.p + (size_t)st.lmBegin * vertexEstimateDim(st.lmType), nullptr, nullptr, step + st.sizePoses + (size_t)st.lmBegin * st.L, st.lmEnd - st.lmBegin, s->stream, &s->launches);
  s->errorsValid = false;
  CU(cudaGetLastError());
  return G2OCU_OK;
}

int pushEstimates(g2ocu_solver* s) {
  if ((int)s->poseBackup.size() <= s->stackDepth) { s->poseBackup.push_back(new DVec<double>); s->lmBackup.push_back(new DVec<double>); }
  DVec<double>& pb = *s->poseBackup[s->stackDepth]; DVec<double>& lb = *s->lmBackup[s->stackDepth];
  CU(pb.alloc(s->poseEst.n)); CU(lb.alloc(s->lmEst.n));
  CU(cudaMemcpyAsync(pb.p, s->poseEst.p, s->poseEst.n * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
  if (s->lmEst.n) CU(cudaMemcpyAsync(lb.p, s->lmEst.p, s->lmEst.n * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
  s->stackDepth++;
  return G2OCU_OK;
}
int popEstimates(g2ocu_solver* s, bool restore) {
  if (s->stackDepth == 0) return fail(s, G2OCU_E_INVALID, "pop/discardTop on an empty backup stack");
  s->stackDepth--;
  if (restore) {
    DVec<double>& pb = *s->poseBackup[s->stackDepth]; DVec<double>& lb = *s->lmBackup[s->stackDepth];
    CU(cudaMemcpyAsync(s->poseEst.p, pb.p, s->poseEst.n * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
    if (s->lmEst.n) CU(cudaMemcpyAsync(s->lmEst.p, lb.p, s->lmEst.n * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
    s->errorsValid = false;
  }
  return G2OCU_OK;
}

int lambdaInit(g2ocu_solver* s, double* out) {
  if (s->userLambdaInit > 0) { *out = s->userLambdaInit; return G2OCU_OK; }
  const double* poseDiag = nullptr;
  if (s->world > 1 && s->st.doSchur) {   // the pose diagonals are partial sums on each rank: reduce them first, then max over ranks
    launchExtractPoseDiag(s->sys, s->vq.p, s->stream, &s->launches);
    int rc = allreduceDev(s, s->vq.p, s->st.sizePoses, 0); if (rc) return rc;
    poseDiag = s->vq.p;
  }
  launchMaxDiag(s->sys, poseDiag, s->st.lmBegin, s->st.lmEnd, s->scratch.p, s->out2.p + 4, s->stream, &s->launches);
  { int rc = allreduceDev(s, s->out2.p + 4, 1, 1); if (rc) return rc; }
  CU(cudaMemcpyAsync(s->hostScal + 4, s->out2.p + 4, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  int rc = syncStream(s); if (rc) return rc;
  *out = s->tau * s->hostScal[4];
  return G2OCU_OK;
}
// sum_j x_j (lambda x_j + b_j) -> hostScal[5] + hostScal[6] after the next stream sync.
// Sharded: b_p is a partial sum (adds up over ranks), lambda x_p^2 is counted on rank 0 only, landmarks are owned.
int enqueueScale(g2ocu_solver* s, double lambda) {
  const Structure& st = s->st;
  CU(cudaMemsetAsync(s->out2.p + 5, 0, 2 * sizeof(double), s->stream));
  if (s->world <= 1 || !st.doSchur) launchScale(s->x.p, s->b.p, (int64_t)s->x.n, lambda, s->scratch.p, s->out2.p + 5, s->stream, &s->launches);
  else {
    launchScale(s->x.p, s->b.p, (int64_t)st.sizePoses, s->rank == 0 ? lambda : 0.0, s->scratch.p, s->out2.p + 5, s->stream, &s->launches);
    const size_t o = (size_t)st.sizePoses + (size_t)st.lmBegin * st.L;
    if (st.lmEnd > st.lmBegin) launchScale(s->x.p + o, s->b.p + o, (int64_t)(st.lmEnd - st.lmBegin) * st.L, lambda, s->scratch.p + 2048, s->out2.p + 6, s->stream, &s->launches);
    int rc = allreduceDev(s, s->out2.p + 5, 2, 0); if (rc) return rc;
  }
  CU(cudaMemcpyAsync(s->hostScal + 5, s->out2.p + 5, 2 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  return G2OCU_OK;
}
int computeScale(g2ocu_solver* s, double lambda, double* out) {
  int rc0 = enqueueScale(s, lambda); if (rc0) return rc0;
  int rc = syncStream(s); if (rc) return rc;
  *out = s->hostScal[5] + s->hostScal[6];
  return G2OCU_OK;
}

// OptimizationAlgorithmLevenberg::solve, optimization_algorithm_levenberg.cpp:58-150
int solveLevenberg(g2ocu_solver* s, int iteration, int* result) {
  int rc;
  if (iteration == 0 && !s->structureBuilt) { rc = g2ocu_build_structure(s); if (rc) { *result = G2OCU_RESULT_FAIL; return rc; } }   // the map stays valid until the next initializeOptimization / set_graph
  // computeActiveErrors (levenberg.cpp:72): when the previous iteration ended with an accepted step, its last trial already evaluated
  // exactly these estimates (the error kernels are deterministic, the result would be the same bits) - skip the pass and the sync
  const bool fresh = s->errorsValid;
  if (!fresh) { rc = computeErrors(s, nullptr, nullptr); if (rc) return rc; }
  rc = buildSystem(s); if (rc) return rc;          // enqueued behind the error kernels; one sync serves both
  if (!fresh) { rc = finishErrors(s); if (rc) return rc; }
  double currentChi = s->chi2Robust, tempChi = currentChi;
  if (iteration == 0) { rc = lambdaInit(s, &s->currentLambda); if (rc) return rc; s->ni = 2; }
  double rho = 0; int& qmax = s->levenbergIterations; qmax = 0;
  do {
    rc = pushEstimates(s); if (rc) return rc;
    s->lambda = s->currentLambda;                      // setLambda(_currentLambda, true)
    int ok2 = 1;
    rc = solveSystem(s, &ok2); if (rc) return rc;
    rc = applyUpdate(s); if (rc) return rc;
    s->lambda = 0.0;                                   // restoreDiagonal
    rc = computeErrors(s, nullptr, nullptr); if (rc) return rc;
    double scale = 0;
    rc = enqueueScale(s, s->currentLambda); if (rc) return rc;
    rc = finishErrors(s); if (rc) return rc;
    scale = s->hostScal[5] + s->hostScal[6];
    tempChi = s->chi2Robust;
    if (!ok2) tempChi = std::numeric_limits<double>::max();
    rho = (currentChi - tempChi);
    scale += 1e-3;
    rho /= scale;
    if (rho > 0 && std::isfinite(tempChi)) {
      double alpha = 1. - std::pow((2 * rho - 1), 3);
      alpha = (std::min)(alpha, s->goodStepUpperScale);
      const double scaleFactor = (std::max)(s->goodStepLowerScale, alpha);
      s->currentLambda *= scaleFactor; s->ni = 2; currentChi = tempChi;
      rc = popEstimates(s, false); if (rc) return rc;
    } else {
      s->currentLambda *= s->ni; s->ni *= 2;
      rc = popEstimates(s, true); if (rc) return rc;
      if (!std::isfinite(s->currentLambda)) break;
    }
    qmax++;
  } while (rho < 0 && qmax < s->maxTrialsAfterFailure && !terminateRequested(s));   // levenberg.cpp:145
  if (qmax == s->maxTrialsAfterFailure || rho == 0 || !std::isfinite(s->currentLambda)) *result = G2OCU_RESULT_TERMINATE;
  else *result = G2OCU_RESULT_OK;
  return G2OCU_OK;
}

// OptimizationAlgorithmGaussNewton::solve, optimization_algorithm_gauss_newton.cpp:50-91
int solveGaussNewton(g2ocu_solver* s, int iteration, int* result) {
  int rc;
  if (iteration == 0 && !s->structureBuilt) { rc = g2ocu_build_structure(s); if (rc) { *result = G2OCU_RESULT_FAIL; return rc; } }   // the map stays valid until the next initializeOptimization / set_graph
  rc = buildSystem(s); if (rc) return rc;
  s->lambda = 0.0;
  int ok = 1;
  rc = solveSystem(s, &ok); if (rc) return rc;
  rc = applyUpdate(s); if (rc) return rc;
  *result = ok ? G2OCU_RESULT_OK : G2OCU_RESULT_FAIL;
  return G2OCU_OK;
}

// PCG view of Hpp (pattern, diagonal blocks, SpMV work items).  Without Schur that is the solver's own view; with the landmark class
// present the solver's view covers Hschur and the Hpp pattern gets its own work items (one per block row), uploaded once per structure.
int hppView(g2ocu_solver* s, PcgDev& pc) {
  const Structure& st = s->st;
  pc = s->pcg; pc.A = s->Hpp.p; pc.lambda = s->lambda; pc.scal = nullptr;
  if (!st.doSchur) return G2OCU_OK;
  if (!s->mhReady) {
    std::vector<int32_t> hr, hb, he;
    for (int i = 0; i < st.numPoses; ++i) { hr.push_back(i); hb.push_back(st.hppRowPtr[i]); he.push_back(st.hppRowPtr[i + 1]); }
    CU(s->mhRow.upload(hr, s->stream)); CU(s->mhBegin.upload(hb, s->stream)); CU(s->mhEnd.upload(he, s->stream));
    CU(s->mhRowPtr.upload(st.hppRowPtr, s->stream)); CU(s->mhColIdx.upload(st.hppColIdx, s->stream));
    CU(cudaStreamSynchronize(s->stream));                            // host staging vectors go out of scope
    s->mhReady = true;
  }
  pc.itemRow = s->mhRow.p; pc.itemBegin = s->mhBegin.p; pc.itemEnd = s->mhEnd.p; pc.nItems = st.numPoses;
  pc.rowPtr = s->mhRowPtr.p; pc.colIdx = s->mhColIdx.p; pc.diag = s->hppDiag.p; pc.nnz = (int)st.hppColIdx.size();
  pc.ownLo = 0; pc.ownHi = pc.nnz;
  return G2OCU_OK;
}

// q = ([Hpp Hpl; Hpl^T Hll] + lambda I) d over the internal [poses | points] layout (full-system mode)
int multFullSystem(g2ocu_solver* s, const double* d, double* q) {
  const Structure& st = s->st; const int64_t np = st.sizePoses;
  PcgDev pc; int rc = hppView(s, pc); if (rc) return rc;
  launchSpmv(pc, d, q, s->stream, &s->launches);                                                                     // q_p = (Hpp + lambda I) d_p
  launchBlockDiagMult(q + np, s->Hll.p, d + np, st.numLandmarks, st.L, s->lambda, s->stream, &s->launches);          // q_l = (Hll + lambda I) d_l
  launchHplMult(s->Hpl.p, s->hplRowIdx.p, s->hplLm.p, (int)st.hplRowIdx.size(), st.P, st.L, d, d + np, q, q + np, s->stream, &s->launches);   // += Hpl d_l, Hpl^T d_p
  CU(cudaGetLastError());
  return G2OCU_OK;
}

// BlockSolver::multiplyHessian (block_solver.h:146) = _Hpp->multiplySymmetricUpperTriangle (sparse_block_matrix.hpp:289-313) on device
// vectors.  With marginalized landmarks the reference's Hpp is the pose block only: dst[0, sizePoses) = (Hpp + lambda I) src[0, sizePoses)
// and the landmark part of its dest is never touched.  In full-system mode the reference's Hpp is the whole matrix.
int multiplyHessianDev(g2ocu_solver* s, const double* src, double* dst) {
  if (s->st.fullSystem) return multFullSystem(s, src, dst);
  PcgDev pc; int rc = hppView(s, pc); if (rc) return rc;
  launchSpmv(pc, src, dst, s->stream, &s->launches);
  CU(cudaGetLastError());
  return G2OCU_OK;
}

// LinearSolverPCG::solve (linear_solver_pcg.hpp:80-156) on the whole system of a graph whose points are not marginalized:
// block-Jacobi preconditioner from the P x P and L x L diagonal blocks, _residual carried from one solve to the next.  As in the Schur-path
// PCG the scalars of the recurrences and the convergence flag live on the device (dot_partial / pcg_full_update1 / pcg_update2_commit, every
// sum in a fixed order); the host polls every 4 iterations, the first time where the previous solve converged.  The product is three
// kernels (Hpp part, point blocks, Hpl blocks with atomics into both halves).
int solveFullPcg(g2ocu_solver* s) {
  const Structure& st = s->st;
  const int64_t np = st.sizePoses, n = np + st.sizeLandmarks;
  CU(s->fr.alloc((size_t)n)); CU(s->fd.alloc((size_t)n)); CU(s->fq.alloc((size_t)n)); CU(s->fs.alloc((size_t)n));
  PcgDev pc; int rc = hppView(s, pc); if (rc) return rc;
  launchBlockInverse(pc, s->stream, &s->launches);                                                                   // J_i = (Hpp_ii + lambda I)^-1 -> Minv
  launchPointBlockInverse(s->Dinv.p, s->Hll.p, st.numLandmarks, st.L, s->lambda, s->stream, &s->launches);           // and the point blocks -> Dinv
  PcgDev fp;                                                                                                        // the whole system as the tail kernels see it
  fp.n = (int)n; fp.nb = st.numPoses + st.numLandmarks; fp.P = st.P; fp.Minv = s->Minv.p;
  fp.r = s->fr.p; fp.d = s->fd.p; fp.q = s->fq.p; fp.s = s->fs.p; fp.x = s->x.p; fp.scal = s->scal.p; fp.ticket = s->pcgTicket.p;
  fp.nPartial = (int)((n + 255) / 256); fp.nPartialDq = std::min(296, fp.nPartial);
  CU(s->fPartial.alloc(fp.nPartial)); CU(s->fPartialDq.alloc(fp.nPartialDq));
  fp.partial = s->fPartial.p; fp.partialDq = s->fPartialDq.p;
  launchFullPcgInit(fp, (int)np, s->Dinv.p, st.L, s->b.p, s->cfg.pcg_tolerance, s->pcgResidual, s->cfg.pcg_absolute_tolerance, s->stream, &s->launches);   // x = 0, r = b, d = M^-1 r, dn, d0
  const int64_t maxIter = s->cfg.pcg_max_iterations < 0 ? n : s->cfg.pcg_max_iterations;
  int64_t issued = 0; bool done = false;
  const int kCheckEvery = 4;
  while (!done) {
    const int64_t want = issued == 0 ? std::min<int64_t>(std::max(kCheckEvery, s->lastPcgIterations - 1), 256) : kCheckEvery;
    const int64_t batch = std::min(want, maxIter - issued);
    for (int64_t k = 0; k < batch; ++k) {                                  // launches past convergence: the tail kernels return at once, the product runs idle
      { KernelTimer pt(s, "pcg_spmv"); rc = multFullSystem(s, fp.d, fp.q); if (rc) return rc; }
      { KernelTimer pt(s, "pcg_vec"); launchFullPcgTail(fp, (int)np, s->Dinv.p, st.L, s->stream, &s->launches); }
    }
    issued += batch;
    CU(cudaMemcpyAsync(s->hostScal + 8, fp.scal, 8 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    rc = syncStream(s); if (rc) return rc;
    if (s->hostScal[8 + 6] != 0.0 || issued >= maxIter) done = true;
    if (!std::isfinite(s->hostScal[8 + 2])) done = true;                   // NaN/Inf in the recurrence: stop issuing work (the reference would spin to maxIter)
  }
  s->lastPcgIterations = (int)s->hostScal[8 + 7]; s->totalPcgIterations += s->lastPcgIterations;
  s->pcgResidual = 0.5 * s->hostScal[8 + 2];
  CU(cudaGetLastError());
  return G2OCU_OK;
}

// OptimizationAlgorithmDogleg::solve, optimization_algorithm_dogleg.cpp:56-197.  Dot products go through the fixed-order reduction of
// computeScale (lambda = 0), results land in out2[8..15] / hostScal[16..23]; the vectors never leave the device.
int solveDogleg(g2ocu_solver* s, int iteration, int* result) {
  int rc;
  *result = G2OCU_RESULT_FAIL;
  if (s->world > 1) return fail(s, G2OCU_E_UNSUPPORTED, "the Dogleg algorithm is not available on a sharded solver");
  if (iteration == 0 && !s->structureBuilt) { rc = g2ocu_build_structure(s); if (rc) return rc; }
  const Structure& st = s->st;
  const int64_t n = (int64_t)st.sizePoses + st.sizeLandmarks;
  const int64_t np = st.fullSystem ? n : st.sizePoses;     // length of the Hessian product: the reference's Hpp (the whole system when nothing is marginalized)
  if (iteration == 0) { s->dlDelta = s->dlUserDeltaInit; s->dlCurrentLambda = s->dlInitialLambda; s->dlWasPD = true; }
  CU(s->hsd.alloc((size_t)n)); CU(s->hdl.alloc((size_t)n)); CU(s->aux.alloc((size_t)n));
  int slot = 0;
  auto dot = [&](const double* u, const double* v, int64_t len) { launchScale(u, v, len, 0.0, s->scratch.p, s->out2.p + 8 + slot, s->stream, &s->launches); return slot++; };
  // results of the dots enqueued since the last fetch -> hostScal[16 + k]; the copy is enqueued right behind them (computeErrors clears out2)
  auto enqueueFetch = [&]() -> int { CU(cudaMemcpyAsync(s->hostScal + 16, s->out2.p + 8, 8 * sizeof(double), cudaMemcpyDeviceToHost, s->stream)); slot = 0; return G2OCU_OK; };
  auto fetch = [&]() -> int { int r = enqueueFetch(); if (r) return r; return syncStream(s); };
  const double* hs = s->hostScal + 16;

  const bool fresh = s->errorsValid;
  if (!fresh) { rc = computeErrors(s, nullptr, nullptr); if (rc) return rc; }
  rc = buildSystem(s); if (rc) return rc;
  // alpha = |b|^2 / (b^T Hpp b) (:97-101)
  s->lambda = 0.0;
  rc = multiplyHessianDev(s, s->b.p, s->aux.p); if (rc) return rc;
  dot(s->b.p, s->b.p, n); dot(s->aux.p, s->b.p, np);
  rc = enqueueFetch(); if (rc) return rc;
  if (!fresh) { rc = finishErrors(s); if (rc) return rc; } else { rc = syncStream(s); if (rc) return rc; }
  const double currentChi = s->chi2Robust;
  const double bNormSquared = hs[0], alpha = bNormSquared / hs[1];
  launchLincomb(s->hsd.p, s->b.p, nullptr, alpha, 0, n, s->stream, &s->launches);          // _hsd = alpha * b
  dot(s->hsd.p, s->hsd.p, n);
  rc = fetch(); if (rc) return rc;
  const double hsdSqrNorm = hs[0], hsdNorm = std::sqrt(hsdSqrNorm);
  double hgnNorm = -1.;
  bool solvedGaussNewton = false, goodStep = false;
  int& numTries = s->dlLastNumTries; numTries = 0;
  do {
    ++numTries;
    if (!solvedGaussNewton) {
      const double minLambda = 1e-12, maxLambda = 1e3;
      solvedGaussNewton = true;
      bool solverOk = false;
      while (!solverOk) {
        // damping only after the system was found not positive definite once (:117-135)
        s->lambda = s->dlWasPD ? 0.0 : s->dlCurrentLambda;
        int ok = 1;
        rc = solveSystem(s, &ok); s->lambda = 0.0; if (rc) return rc;
        solverOk = ok != 0;
        s->dlWasPD = s->dlWasPD && solverOk;
        if (!s->dlWasPD) {
          if (solverOk) s->dlCurrentLambda = std::max(minLambda, s->dlCurrentLambda / (0.5 * s->dlLambdaFactor));
          else { s->dlCurrentLambda *= s->dlLambdaFactor; if (s->dlCurrentLambda > maxLambda) { s->dlCurrentLambda = maxLambda; *result = G2OCU_RESULT_FAIL; return G2OCU_OK; } }
        }
      }
      dot(s->x.p, s->x.p, n);
      rc = fetch(); if (rc) return rc;
      hgnNorm = std::sqrt(hs[0]);
    }
    const double delta = s->dlDelta;
    if (hgnNorm < delta) {
      CU(cudaMemcpyAsync(s->hdl.p, s->x.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
      s->dlLastStep = G2OCU_DOGLEG_STEP_GN;
    } else if (hsdNorm > delta) {
      launchLincomb(s->hdl.p, s->hsd.p, nullptr, delta / hsdNorm, 0, n, s->stream, &s->launches);
      s->dlLastStep = G2OCU_DOGLEG_STEP_SD;
    } else {
      launchLincomb(s->aux.p, s->hsd.p, s->x.p, 0.0, 1, n, s->stream, &s->launches);       // _auxVector = hgn - _hsd
      dot(s->hsd.p, s->aux.p, n); dot(s->aux.p, s->aux.p, n);
      rc = fetch(); if (rc) return rc;
      const double c = hs[0], bmaSquaredNorm = hs[1];
      double beta;
      if (c <= 0.) beta = (-c + std::sqrt(c * c + bmaSquaredNorm * (delta * delta - hsdSqrNorm))) / bmaSquaredNorm;
      else beta = (delta * delta - hsdSqrNorm) / (c + std::sqrt(c * c + bmaSquaredNorm * (delta * delta - hsdSqrNorm)));
      launchLincomb(s->hdl.p, s->hsd.p, s->x.p, beta, 2, n, s->stream, &s->launches);      // _hdl = _hsd + beta * (hgn - _hsd)
      s->dlLastStep = G2OCU_DOGLEG_STEP_DL;
    }
    // linear gain = -(Hpp hdl).hdl + 2 b.hdl (:165-168)
    rc = multiplyHessianDev(s, s->hdl.p, s->aux.p); if (rc) return rc;
    dot(s->aux.p, s->hdl.p, np); dot(s->b.p, s->hdl.p, n); dot(s->hdl.p, s->hdl.p, n);
    rc = enqueueFetch(); if (rc) return rc;
    rc = pushEstimates(s); if (rc) return rc;
    rc = applyUpdate(s, s->hdl.p); if (rc) return rc;
    rc = computeErrors(s, nullptr, nullptr); if (rc) return rc;
    rc = finishErrors(s); if (rc) return rc;                 // one sync serves the dot products and the new chi2
    double linearGain = -1 * hs[0] + 2 * hs[1];
    const double hdlNorm = std::sqrt(hs[2]);
    const double newChi = s->chi2Robust;
    const double nonLinearGain = currentChi - newChi;
    if (std::fabs(linearGain) < 1e-12) linearGain = 1e-12;
    const double rho = nonLinearGain / linearGain;
    if (rho > 0) { rc = popEstimates(s, false); if (rc) return rc; goodStep = true; }
    else { rc = popEstimates(s, true); if (rc) return rc; }
    if (rho > 0.75) s->dlDelta = std::max(s->dlDelta, 3 * hdlNorm);
    else if (rho < 0.25) s->dlDelta *= 0.5;
  } while (!goodStep && numTries < s->dlMaxTrialsAfterFailure);
  *result = (numTries == s->dlMaxTrialsAfterFailure || !goodStep) ? G2OCU_RESULT_TERMINATE : G2OCU_RESULT_OK;
  return G2OCU_OK;
}

}  // namespace

// =================================================================================================
extern "C" {

void g2ocu_default_config(g2ocu_config* cfg) {
  if (!cfg) return;
  cfg->device = -1; cfg->linear_solver = G2OCU_LINEAR_PCG; cfg->pcg_tolerance = 1e-6; cfg->pcg_max_iterations = -1; cfg->pcg_absolute_tolerance = 1; cfg->stream = nullptr;
}
int g2ocu_version(void) { return G2OCU_VERSION; }
const char* g2ocu_last_error(const g2ocu_solver* s) { return s ? s->err.c_str() : g_createError.c_str(); }

int g2ocu_create(const g2ocu_config* cfg, g2ocu_solver** out) {
  if (!out) return fail(nullptr, G2OCU_E_INVALID, "null output pointer");
  g2ocu_solver* s = new g2ocu_solver;
  if (cfg) s->cfg = *cfg; else g2ocu_default_config(&s->cfg);
  if (s->cfg.linear_solver != G2OCU_LINEAR_PCG && s->cfg.linear_solver != G2OCU_LINEAR_DENSE) { delete s; return fail(nullptr, G2OCU_E_INVALID, "unknown linear solver kind"); }
  *out = s;
  return G2OCU_OK;
}
void g2ocu_destroy(g2ocu_solver* s) { delete s; }

int g2ocu_set_graph(g2ocu_solver* s, const g2ocu_graph* g) {
  if (!s) return G2OCU_E_INVALID;
  std::string err;
  s->hasGraph = false; s->optInitialized = false; s->structureBuilt = false; s->algoInitialized = false; dropP2p(s);
  if (!s->g.assign(g, err)) return fail(s, err.find("unsupported") != std::string::npos ? G2OCU_E_UNSUPPORTED : G2OCU_E_INVALID, err);
  s->hasGraph = true;
  return G2OCU_OK;
}
int g2ocu_set_property(g2ocu_solver* s, const char* name, double value) {
  if (!s || !name) return G2OCU_E_INVALID;
  const std::string n(name);
  if (n == "initialLambda") s->userLambdaInit = value;
  else if (n == "maxTrialsAfterFailure") s->maxTrialsAfterFailure = (int)value;
  else if (n == "doglegInitialDelta") s->dlUserDeltaInit = value;                   // OptimizationAlgorithmDogleg "initialDelta" (dogleg.cpp:44)
  else if (n == "doglegMaxTrialsAfterFailure") s->dlMaxTrialsAfterFailure = (int)value;   // its own "maxTrialsAfterFailure", default 100 (:45)
  else if (n == "doglegInitialLambda") s->dlInitialLambda = value;                  // its own "initialLambda", default 1e-7 (:46)
  else if (n == "doglegLambdaFactor") s->dlLambdaFactor = value;                    // "lambdaFactor" (:47)
  else if (n == "pcgTolerance") s->cfg.pcg_tolerance = value;
  else if (n == "pcgMaxIterations") s->cfg.pcg_max_iterations = (int)value;
  else if (n == "pcgAbsoluteTolerance") s->cfg.pcg_absolute_tolerance = (int)value;
  else if (n == "kernelTiming") s->kernelTiming = value != 0.0;
  else if (n == "poseDim") s->expectPoseDim = (int)value;
  else if (n == "landmarkDim") s->expectLandmarkDim = (int)value;
  else if (n == "linearSolver") {
    if ((int)value != G2OCU_LINEAR_PCG && (int)value != G2OCU_LINEAR_DENSE) return fail(s, G2OCU_E_INVALID, "unknown linear solver kind");
    s->cfg.linear_solver = (int)value;
  }
  else return fail(s, G2OCU_E_INVALID, "unknown property " + n);
  return G2OCU_OK;
}
int g2ocu_set_force_stop_flag(g2ocu_solver* s, const unsigned char* flag) { if (!s) return G2OCU_E_INVALID; s->forceStop = flag; return G2OCU_OK; }
int g2ocu_set_shard(g2ocu_solver* s, int32_t rank, int32_t world, g2ocu_allreduce_fn fn, void* user) {
  if (!s || world < 1 || rank < 0 || rank >= world) return fail(s, G2OCU_E_INVALID, "bad rank/world");
  if (world > 1 && !fn) return fail(s, G2OCU_E_INVALID, "world > 1 needs an allreduce hook");
  s->rank = rank; s->world = world; s->allreduce = fn; s->allreduceUser = user;
  s->structureBuilt = false; dropP2p(s);
  return G2OCU_OK;
}

int g2ocu_nccl_unique_id(const char* nccl_library, unsigned char unique_id[128]) {
  std::string err;
  if (!unique_id) return G2OCU_E_INVALID;
  if (!loadNccl(nccl_library, err)) return fail(nullptr, G2OCU_E_COMM, err);
  NcclId id;
  const int rc = g_nccl.GetUniqueId(&id);
  if (rc != 0) return fail(nullptr, G2OCU_E_COMM, std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(rc));
  std::memcpy(unique_id, id.internal, 128);
  return G2OCU_OK;
}
int g2ocu_set_shard_nccl(g2ocu_solver* s, int32_t rank, int32_t world, const char* nccl_library, const unsigned char unique_id[128]) {
  if (!s || world < 1 || rank < 0 || rank >= world || !unique_id) return fail(s, G2OCU_E_INVALID, "bad rank/world/unique id");
  std::string err;
  if (!loadNccl(nccl_library, err)) return fail(s, G2OCU_E_COMM, err);
  int rc = ensureCuda(s); if (rc) return rc;
  if (s->ncclComm) { g_nccl.CommDestroy(s->ncclComm); s->ncclComm = nullptr; }
  NcclId id; std::memcpy(id.internal, unique_id, 128);
  const int nrc = g_nccl.CommInitRank(&s->ncclComm, world, id, rank);
  if (nrc != 0) { s->ncclComm = nullptr; return fail(s, G2OCU_E_COMM, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(nrc)); }
  s->rank = rank; s->world = world; s->allreduce = nullptr; s->allreduceUser = nullptr;
  s->structureBuilt = false; dropP2p(s);
  return G2OCU_OK;
}

int g2ocu_p2p_export(g2ocu_solver* s, unsigned char handle[64]) {
  if (!s || !handle) return G2OCU_E_INVALID;
  if (!s->structureBuilt) return fail(s, G2OCU_E_INVALID, "g2ocu_p2p_export needs the structure (call g2ocu_build_structure first)");
  if (s->world < 2 || s->world > 8) return fail(s, G2OCU_E_INVALID, "the peer-memory exchange supports 2..8 ranks");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  dropP2p(s);                                      // peers' mappings of the old buffer are closed by their own export
  if (s->p2pLocal) { cudaFree(s->p2pLocal); s->p2pLocal = nullptr; }
  const int64_t cap = (s->st.sizePoses + 1) & ~(int64_t)1;
  const size_t bytes = p2pBytes(s->world, cap);
  CU(cudaMalloc((void**)&s->p2pLocal, bytes));
  CU(cudaMemsetAsync(s->p2pLocal, 0, bytes, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  cudaIpcMemHandle_t h;
  CU(cudaIpcGetMemHandle(&h, s->p2pLocal));
  std::memcpy(handle, &h, 64);
  s->p2p = P2pDev(); s->p2p.rank = s->rank; s->p2p.world = s->world; s->p2p.cap = cap;
  return G2OCU_OK;
}
int g2ocu_p2p_import(g2ocu_solver* s, const unsigned char* handles) {
  if (!s || !handles) return G2OCU_E_INVALID;
  if (!s->p2pLocal) return fail(s, G2OCU_E_INVALID, "g2ocu_p2p_export has not been called");
  for (int r = 0; r < s->world; ++r) {
    if (r == s->rank) { s->p2p.peer[r] = s->p2pLocal; continue; }
    cudaIpcMemHandle_t h; std::memcpy(&h, handles + 64 * (size_t)r, 64);
    void* ptr = nullptr;
    CU(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    s->p2pOpened[r] = ptr; s->p2p.peer[r] = (double*)ptr;
  }
  s->p2pReady = true;
  return G2OCU_OK;
}

// The same for the reduction of the reduced camera system: export this rank's partial-Hschur buffer, import the peers'.
int g2ocu_p2p_export_schur(g2ocu_solver* s, unsigned char handle[64]) {
  if (!s || !handle) return G2OCU_E_INVALID;
  if (!s->structureBuilt || !s->st.doSchur) return fail(s, G2OCU_E_INVALID, "g2ocu_p2p_export_schur needs the structure of a graph with marginalized landmarks");
  if (s->world < 2 || s->world > 8) return fail(s, G2OCU_E_INVALID, "the peer-memory exchange supports 2..8 ranks");
  cudaIpcMemHandle_t h;
  CU(cudaIpcGetMemHandle(&h, s->S.p));
  std::memcpy(handle, &h, 64);
  return G2OCU_OK;
}
int g2ocu_p2p_import_schur(g2ocu_solver* s, const unsigned char* handles) {
  if (!s || !handles) return G2OCU_E_INVALID;
  if (!s->structureBuilt || !s->st.doSchur) return fail(s, G2OCU_E_INVALID, "g2ocu_p2p_import_schur needs the structure of a graph with marginalized landmarks");
  s->schurP2pReady = false;
  s->schurPeers = P2pDev(); s->schurPeers.rank = s->rank; s->schurPeers.world = s->world; s->schurPeers.cap = (int64_t)s->S.n;
  for (int r = 0; r < s->world; ++r) {
    if (r == s->rank) { s->schurPeers.peer[r] = s->S.p; continue; }
    cudaIpcMemHandle_t h; std::memcpy(&h, handles + 64 * (size_t)r, 64);
    void* ptr = nullptr;
    CU(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    s->schurOpened[r] = ptr; s->schurPeers.peer[r] = (double*)ptr;
  }
  s->schurP2pReady = true;
  return G2OCU_OK;
}

int g2ocu_initialize_optimization(g2ocu_solver* s, int32_t level) {
  if (!s) return G2OCU_E_INVALID;
  if (!s->hasGraph) return fail(s, G2OCU_E_INVALID, "no graph set");
  std::string err;
  if (s->structureBuilt) { int rc0 = downloadEstimates(s); if (rc0) return rc0; }   // vertices keep their estimates
  s->optInitialized = false; s->structureBuilt = false; s->algoInitialized = false; dropP2p(s);
  if (!initializeOptimization(s->g, level, s->st, err)) return fail(s, G2OCU_E_INVALID, err);
  s->optInitialized = true;
  return G2OCU_OK;
}

// OptimizationAlgorithmWithHessian::init (+ Solver::init, LinearSolver::init): chooses Schur, resets the PCG residual.
int g2ocu_init(g2ocu_solver* s, int32_t online) {
  if (!s) return G2OCU_E_INVALID;
  if (!s->optInitialized) return fail(s, G2OCU_E_INVALID, "0 vertices to optimize, maybe forgot to call initializeOptimization()");
  (void)online;
  s->pcgResidual = -1.0;        // LinearSolverPCG::init, linear_solver_pcg.h:64-70
  s->algoInitialized = true;
  return G2OCU_OK;
}

int g2ocu_build_structure(g2ocu_solver* s) {
  if (!s) return G2OCU_E_INVALID;
  if (!s->optInitialized) return fail(s, G2OCU_E_INVALID, "initializeOptimization has not been called");
  std::string err;
  if (s->structureBuilt) { int rc0 = downloadEstimates(s); if (rc0) return rc0; }   // vertices keep their state across optimize() calls
  s->structureBuilt = false; dropP2p(s);
  if (!buildStructure(s->g, s->st, err, s->rank, s->world)) return fail(s, G2OCU_E_UNSUPPORTED, err);
  const Structure& st = s->st;
  const bool okDims = (st.doSchur && ((st.P == 9 && st.L == 3) || (st.P == 6 && st.L == 3) || (st.P == 3 && st.L == 2))) || (!st.doSchur && (st.P == 3 || st.P == 6 || st.P == 9));
  // BlockSolver<BlockSolverTraits<p,l>> maps fixed-size blocks: a graph with other block sizes does not fit a fixN_M solver (block_solver.hpp:103-256)
  if ((s->expectPoseDim > 0 && st.numPoses > 0 && st.P != s->expectPoseDim) || (s->expectLandmarkDim > 0 && st.numLandmarks > 0 && st.L != s->expectLandmarkDim))
    return fail(s, G2OCU_E_UNSUPPORTED, "block sizes P=" + std::to_string(st.P) + " L=" + std::to_string(st.L) + " do not match the solver's fixed sizes " + std::to_string(s->expectPoseDim) + "_" + std::to_string(s->expectLandmarkDim));
  if (!okDims) return fail(s, G2OCU_E_UNSUPPORTED, "unsupported block sizes P=" + std::to_string(st.P) + " L=" + std::to_string(st.L));
  for (const EdgeSet& es : st.sets) {
    const bool naturalPL = es.etype == G2OCU_EDGE_SE2_POINT_XY || es.etype == G2OCU_EDGE_PROJECT_XYZ2UV || es.etype == G2OCU_EDGE_SE3_PROJECT_XYZ || es.etype == G2OCU_EDGE_BAL;
    const int naturalSide = (es.etype == G2OCU_EDGE_PROJECT_XYZ2UV || es.etype == G2OCU_EDGE_SE3_PROJECT_XYZ) ? 1 : 0;
    if (naturalPL != es.poseLandmark || (naturalPL && naturalSide != es.poseSide))
      return fail(s, G2OCU_E_UNSUPPORTED, "edge type " + std::to_string(es.etype) + ": the landmark-side vertices must be marginalized and the pose-side vertices must not (mixed block sizes are not supported)");
  }
  if (st.fullSystem && s->world > 1) return fail(s, G2OCU_E_UNSUPPORTED, "a graph whose points are not marginalized cannot be sharded (mark the points as marginalized)");
  s->slabBlocks = st.doSchur ? (int64_t)((st.sColIdx.size() + s->world - 1) / s->world) : 0;   // equal block ranges of the reduced system (slab PCG)
  int rc = ensureCuda(s); if (rc) return rc;
  rc = buildDevice(s); if (rc) return rc;
  s->structureBuilt = true; s->lambda = 0.0; s->stackDepth = 0;
  return G2OCU_OK;
}

int g2ocu_compute_active_errors(g2ocu_solver* s) {
  int rc = requireBuilt(s); if (rc) return rc;
  rc = computeErrors(s, nullptr, nullptr); if (rc) return rc;
  return finishErrors(s);
}
int g2ocu_active_robust_chi2(g2ocu_solver* s, double* chi2) {
  int rc = requireBuilt(s); if (rc) return rc;
  if (!s->errorsValid) { rc = g2ocu_compute_active_errors(s); if (rc) return rc; }
  if (chi2) *chi2 = s->chi2Robust;
  return G2OCU_OK;
}
int g2ocu_active_chi2(g2ocu_solver* s, double* chi2) {
  int rc = requireBuilt(s); if (rc) return rc;
  if (!s->errorsValid) { rc = g2ocu_compute_active_errors(s); if (rc) return rc; }
  if (chi2) *chi2 = s->chi2Plain;
  return G2OCU_OK;
}
int g2ocu_build_system(g2ocu_solver* s) { int rc = requireBuilt(s); if (rc) return rc; rc = buildSystem(s); if (rc) return rc; return syncStream(s); }
int g2ocu_set_lambda(g2ocu_solver* s, double lambda, int32_t) { int rc = requireBuilt(s); if (rc) return rc; s->lambda += lambda; return G2OCU_OK; }
int g2ocu_restore_diagonal(g2ocu_solver* s) { int rc = requireBuilt(s); if (rc) return rc; s->lambda = 0.0; return G2OCU_OK; }
int g2ocu_solve(g2ocu_solver* s, int32_t* solved) {
  int rc = requireBuilt(s); if (rc) return rc;
  int ok = 1; rc = solveSystem(s, &ok); if (rc) return rc;
  if (solved) *solved = ok;
  return syncStream(s);
}
int g2ocu_update(g2ocu_solver* s, const double* host) {
  int rc = requireBuilt(s); if (rc) return rc;
  std::vector<double> tmp;
  if (host && s->st.fullSystem) {   // the caller's vector is in the reference's order (all vertices by id)
    tmp.resize(s->x.n);
    for (size_t i = 0; i < tmp.size(); ++i) tmp[s->st.refToInternal[i]] = host[i];
    host = tmp.data();
  }
  if (host) CU(cudaMemcpyAsync(s->x.p, host, s->x.n * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  rc = applyUpdate(s); if (rc) return rc;
  return syncStream(s);
}
int g2ocu_push(g2ocu_solver* s) { int rc = requireBuilt(s); if (rc) return rc; return pushEstimates(s); }
int g2ocu_pop(g2ocu_solver* s) { int rc = requireBuilt(s); if (rc) return rc; return popEstimates(s, true); }
int g2ocu_discard_top(g2ocu_solver* s) { int rc = requireBuilt(s); if (rc) return rc; return popEstimates(s, false); }
int g2ocu_compute_lambda_init(g2ocu_solver* s, double* lambda) { int rc = requireBuilt(s); if (rc) return rc; double v = 0; rc = lambdaInit(s, &v); if (lambda) *lambda = v; return rc; }
int g2ocu_compute_scale(g2ocu_solver* s, double lambda, double* scale) { int rc = requireBuilt(s); if (rc) return rc; double v = 0; rc = computeScale(s, lambda, &v); if (scale) *scale = v; return rc; }

int g2ocu_multiply_hessian(g2ocu_solver* s, double* hostDest, const double* hostSrc) {
  int rc = requireBuilt(s); if (rc) return rc;
  if (!hostDest || !hostSrc) return fail(s, G2OCU_E_INVALID, "null vector");
  if (s->st.fullSystem) {                                            // the reference's Hpp is the whole matrix; vectors in its order
    const size_t n = s->x.n;
    std::vector<double> tmp(n);
    for (size_t i = 0; i < n; ++i) tmp[s->st.refToInternal[i]] = hostSrc[i];
    CU(s->fd.alloc(n)); CU(s->fq.alloc(n));
    CU(cudaMemcpyAsync(s->fd.p, tmp.data(), n * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    rc = multFullSystem(s, s->fd.p, s->fq.p); if (rc) return rc;
    CU(cudaMemcpyAsync(tmp.data(), s->fq.p, n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    rc = syncStream(s); if (rc) return rc;
    for (size_t i = 0; i < n; ++i) hostDest[i] = tmp[s->st.refToInternal[i]];
    return G2OCU_OK;
  }
  const int n = s->st.sizePoses;                                     // BlockSolverBase::multiplyHessian works on Hpp (block_solver.h:87-95)
  CU(cudaMemcpyAsync(s->vd.p, hostSrc, n * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  rc = multiplyHessianDev(s, s->vd.p, s->vq.p); if (rc) return rc;
  CU(cudaMemcpyAsync(hostDest, s->vq.p, n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  return syncStream(s);
}

// SparseOptimizer::computeMarginals (sparse_optimizer.cpp:594-596) -> BlockSolver::computeMarginals (block_solver.hpp:451-459) ->
// LinearSolver::solvePattern(spinv, blockIndices, *_Hpp) (linear_solver.h:89-98; the CSparse / CHOLMOD solvers answer it through
// MarginalCovarianceCholesky, marginal_covariance_cholesky.cpp:153-222): blocks (row, col) of the inverse of Hpp as it stands - the pose
// block of the last buildSystem, plus whatever setLambda has put on its diagonal.  Here: dense copy of Hpp, FP64 Cholesky on the tensor
// pipe (kernels_dense.cu), (L L^T)^-1 applied to the unit columns of the requested block columns (at most 1 GiB of them at a time), the
// requested blocks gathered on the device.  *computed = 0 where the reference's solvePattern returns false: factorisation failed.  With points
// that are not marginalized the reference's Hpp is the whole system over all vertices in id order: blocks of two sizes, indexed as there.
int g2ocu_compute_marginals(g2ocu_solver* s, int32_t nPairs, const int32_t* blockRows, const int32_t* blockCols, double* out, int32_t* computed) {
  int rc = requireBuilt(s); if (rc) return rc;
  if (computed) *computed = 0;
  if (nPairs < 0 || (nPairs > 0 && (!blockRows || !blockCols || !out))) return fail(s, G2OCU_E_INVALID, "g2ocu_compute_marginals: null argument");
  const Structure& st = s->st;
  const int P = st.P; const bool full = st.fullSystem;
  // the matrix the reference calls Hpp: the pose block, or - points not marginalized - the whole system over all vertices in id order
  const int64_t n = full ? (int64_t)st.sizePoses + st.sizeLandmarks : (int64_t)st.sizePoses;
  const int nBlocks = full ? (int)st.refPoseBlockIndices.size() : st.numPoses;
  if (P != 3 && P != 6 && P != 9) return fail(s, G2OCU_E_UNSUPPORTED, "g2ocu_compute_marginals: pose dimension " + std::to_string(P));
  if (full && s->world > 1) return fail(s, G2OCU_E_UNSUPPORTED, "g2ocu_compute_marginals: a graph whose points are not marginalized cannot be sharded");
  if (n > kDenseMaxN) return fail(s, G2OCU_E_UNSUPPORTED, "g2ocu_compute_marginals: the dense factorisation is limited to systems of dimension <= " + std::to_string(kDenseMaxN) + " (this one has " + std::to_string(n) + ")");
  for (int i = 0; i < nPairs; ++i)
    if (blockRows[i] < 0 || blockRows[i] >= nBlocks || blockCols[i] < 0 || blockCols[i] >= nBlocks)
      return fail(s, G2OCU_E_INVALID, "g2ocu_compute_marginals: block (" + std::to_string(blockRows[i]) + ", " + std::to_string(blockCols[i]) + ") is outside Hpp (" + std::to_string(nBlocks) + " block rows)");
  if (nPairs == 0) { if (computed) *computed = 1; return G2OCU_OK; }
  // block index (the reference's) -> first scalar row in the device's layout, dimension
  auto blockAt = [&](int b, int& scalar, int& dim) {
    if (!full) { scalar = b * P; dim = P; return; }
    const int begin = b ? st.refPoseBlockIndices[b - 1] : 0;
    dim = st.refPoseBlockIndices[b] - begin; scalar = st.refToInternal[begin];   // a vertex's scalars are contiguous in both orders
  };
  PhaseTimer pt(s, "marginals");
  PcgDev pc; rc = hppView(s, pc); if (rc) return rc;
  DVec<double> hppSum;
  if (s->world > 1 && st.doSchur) {   // landmark shards hold partial pose blocks: sum a copy
    const size_t cnt = st.hppColIdx.size() * (size_t)P * P;
    CU(hppSum.alloc(cnt)); CU(cudaMemcpyAsync(hppSum.p, s->Hpp.p, cnt * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
    rc = allreduceDev(s, hppSum.p, (int64_t)cnt, 0); if (rc) return rc;
    pc.A = hppSum.p;
  }
  CU(s->denseH.alloc((size_t)n * n)); CU(s->denseInfo.alloc(1));
  if (full) launchDenseAssembleFull(pc, s->Hll.p, s->Hpl.p, s->hplRowIdx.p, s->hplLm.p, (int)st.hplRowIdx.size(), st.numLandmarks, st.L, s->denseH.p, (int)n, s->stream, &s->launches);
  else launchDenseAssemble(pc, s->denseH.p, s->stream, &s->launches);
  launchDenseCholeskyFactor(s->denseH.p, (int)n, s->denseInfo.p, s->stream, &s->launches);
  CU(cudaMemcpyAsync(s->hostInfo, s->denseInfo.p, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
  rc = syncStream(s); if (rc) return rc;
  if (*s->hostInfo != 0) return G2OCU_OK;                            // not positive definite: solvePattern == false
  // pairs by block column; a batch = as many distinct block columns as fit the column budget (1 GiB of right-hand sides)
  std::vector<int32_t> order(nPairs);
  for (int i = 0; i < nPairs; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return blockCols[a] < blockCols[b]; });
  std::vector<int64_t> outOffAll(nPairs + 1, 0);
  for (int i = 0; i < nPairs; ++i) { int sr, dr, sc, dc; blockAt(blockRows[i], sr, dr); blockAt(blockCols[i], sc, dc); outOffAll[i + 1] = outOffAll[i] + (int64_t)dr * dc; }
  const int64_t colsMax = std::max<int64_t>(16, ((int64_t)1 << 27) / n);
  DVec<double> X, outDev; DVec<int32_t> dColScalar, dRowScalar, dRowDim, dColStart, dColDim; DVec<int64_t> dOutOff;
  CU(outDev.alloc((size_t)outOffAll[nPairs]));
  size_t at = 0;
  while (at < order.size()) {
    std::vector<int32_t> colScalar, rowScalar, rowDim, colStart, colDim; std::vector<int64_t> outOff;
    int lastCol = -1, curStart = 0, curDim = 0, minRow = (int)n, minCol = (int)n;
    size_t e = at;
    for (; e < order.size(); ++e) {
      const int32_t c = blockCols[order[e]];
      if (c != lastCol) {
        int sc, dc; blockAt(c, sc, dc);
        if (!colScalar.empty() && (int64_t)colScalar.size() + dc > colsMax) break;
        curStart = (int)colScalar.size(); curDim = dc; lastCol = c; minCol = std::min(minCol, sc);
        for (int q = 0; q < dc; ++q) colScalar.push_back(sc + q);
      }
      int sr, dr; blockAt(blockRows[order[e]], sr, dr);
      rowScalar.push_back(sr); rowDim.push_back(dr); colStart.push_back(curStart); colDim.push_back(curDim); outOff.push_back(outOffAll[order[e]]);
      minRow = std::min(minRow, sr);
    }
    const int nrhs = (int)colScalar.size();
    CU(X.alloc((size_t)n * nrhs));
    CU(dColScalar.upload(colScalar, s->stream)); CU(dRowScalar.upload(rowScalar, s->stream)); CU(dRowDim.upload(rowDim, s->stream));
    CU(dColStart.upload(colStart, s->stream)); CU(dColDim.upload(colDim, s->stream)); CU(dOutOff.upload(outOff, s->stream));
    launchUnitColumns(X.p, (size_t)n, dColScalar.p, nrhs, s->stream, &s->launches);
    launchDenseSolveMany(s->denseH.p, (int)n, X.p, (size_t)n, nrhs, minCol, minRow, s->stream, &s->launches);
    launchGatherBlocks(X.p, (size_t)n, dRowScalar.p, dRowDim.p, dColStart.p, dColDim.p, dOutOff.p, (int)rowScalar.size(), outDev.p, s->stream, &s->launches);
    CU(cudaGetLastError());
    rc = syncStream(s); if (rc) return rc;                           // the host staging vectors go out of scope
    at = e;
  }
  CU(cudaMemcpyAsync(out, outDev.p, (size_t)outOffAll[nPairs] * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  rc = syncStream(s); if (rc) return rc;
  if (s->cfg.linear_solver != G2OCU_LINEAR_DENSE) s->denseH.release();   // n x n doubles: only the dense solver keeps its matrix between calls
  if (computed) *computed = 1;
  return G2OCU_OK;
}

int g2ocu_solver_iteration(g2ocu_solver* s, int32_t algorithm, int32_t iteration, g2ocu_iteration_stats* stats) {
  if (!s) return G2OCU_E_INVALID;
  if (!s->algoInitialized) return fail(s, G2OCU_E_INVALID, "g2ocu_init has not been called");
  if (iteration != 0) { int rc = requireBuilt(s); if (rc) return rc; }
  const double t0 = wallNow();
  const double e0 = phaseSeconds(s, "errors"), b0 = phaseSeconds(s, "build"), sc0 = phaseSeconds(s, "schur"), l0 = phaseSeconds(s, "linear_solver"), u0 = phaseSeconds(s, "update"), bs0 = phaseSeconds(s, "backsub");
  int result = G2OCU_RESULT_FAIL;
  if (algorithm != G2OCU_ALGORITHM_GN && algorithm != G2OCU_ALGORITHM_LM && algorithm != G2OCU_ALGORITHM_DOGLEG) return fail(s, G2OCU_E_INVALID, "unknown algorithm code");
  int rc = algorithm == G2OCU_ALGORITHM_LM ? solveLevenberg(s, iteration, &result) : algorithm == G2OCU_ALGORITHM_DOGLEG ? solveDogleg(s, iteration, &result) : solveGaussNewton(s, iteration, &result);
  if (rc) return rc;
  if (stats) {
    // SparseOptimizer::optimize computes the errors again for the statistics (sparse_optimizer.cpp:411-417)
    double chi2 = 0; rc = g2ocu_active_robust_chi2(s, &chi2); if (rc) return rc;
    std::memset(stats, 0, sizeof(*stats));
    stats->iteration = iteration; stats->result = result; stats->levenberg_iterations = algorithm == G2OCU_ALGORITHM_LM ? s->levenbergIterations : 0;
    stats->iterations_linear_solver = s->lastPcgIterations; stats->chi2 = chi2; stats->lambda = algorithm == G2OCU_ALGORITHM_DOGLEG ? s->dlCurrentLambda : s->currentLambda;
    stats->time_residuals = phaseSeconds(s, "errors") - e0; stats->time_quadratic_form = phaseSeconds(s, "build") - b0;
    stats->time_schur_complement = phaseSeconds(s, "schur") - sc0; stats->time_linear_solver = phaseSeconds(s, "linear_solver") - l0;
    stats->time_linear_solution = stats->time_schur_complement + stats->time_linear_solver + (phaseSeconds(s, "backsub") - bs0);
    stats->time_update = phaseSeconds(s, "update") - u0; stats->time_iteration = wallNow() - t0;
    stats->hessian_pose_dimension = s->st.fullSystem ? s->st.sizePoses + s->st.sizeLandmarks : s->st.sizePoses;   // full-system mode: everything is in the reference's Hpp
    stats->hessian_landmark_dimension = s->st.fullSystem ? 0 : s->st.sizeLandmarks;
  }
  return G2OCU_OK;
}

// SparseOptimizer::optimize, sparse_optimizer.cpp:374-439
int g2ocu_optimize(g2ocu_solver* s, int32_t algorithm, int32_t iterations, g2ocu_iteration_stats* stats, int32_t* performed) {
  if (!s) return G2OCU_E_INVALID;
  if (performed) *performed = -1;
  if (!s->optInitialized || s->st.ivMap.empty()) return fail(s, G2OCU_E_INVALID, "0 vertices to optimize, maybe forgot to call initializeOptimization()");
  int rc = g2ocu_init(s, 0); if (rc) return rc;
  int cj = 0; bool ok = true; int result = G2OCU_RESULT_OK;
  for (int i = 0; i < iterations && !terminateRequested(s) && ok; ++i) {   // sparse_optimizer.cpp:396
    g2ocu_iteration_stats local;
    rc = g2ocu_solver_iteration(s, algorithm, i, stats ? &stats[i] : &local); if (rc) return rc;
    result = stats ? stats[i].result : local.result;
    ok = (result == G2OCU_RESULT_OK);
    ++cj;
  }
  if (performed) *performed = (result == G2OCU_RESULT_FAIL) ? 0 : cj;
  return G2OCU_OK;
}

int64_t g2ocu_vector_size(const g2ocu_solver* s) { return (s && s->structureBuilt) ? (int64_t)s->st.sizePoses + s->st.sizeLandmarks : 0; }

int g2ocu_set_estimates(g2ocu_solver* s, const double* host) {
  if (!s || !host) return G2OCU_E_INVALID;
  if (!s->hasGraph) return fail(s, G2OCU_E_INVALID, "no graph set");
  if (s->structureBuilt && s->fastEstimates) {
    CU(cudaMemcpyAsync(s->poseEst.p, host + s->poseHostOff, s->poseEst.n * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    if (s->lmEst.n) CU(cudaMemcpyAsync(s->lmEst.p, host + s->lmHostOff, s->lmEst.n * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    s->errorsValid = false;
    return syncStream(s);       // the caller may reuse its buffer as soon as this returns
  }
  std::memcpy(s->g.vEst.data(), host, sizeof(double) * s->g.vEst.size());
  if (s->structureBuilt) return uploadEstimates(s);
  return G2OCU_OK;
}
int g2ocu_get_estimates(g2ocu_solver* s, double* host) {
  if (!s || !host) return G2OCU_E_INVALID;
  if (!s->hasGraph) return fail(s, G2OCU_E_INVALID, "no graph set");
  if (s->structureBuilt && s->fastEstimates) {
    { int rc = gatherLandmarks(s); if (rc) return rc; }
    CU(cudaMemcpyAsync(host + s->poseHostOff, s->poseEst.p, s->poseEst.n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    if (s->lmEst.n) CU(cudaMemcpyAsync(host + s->lmHostOff, s->lmEst.p, s->lmEst.n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    return syncStream(s);
  }
  if (s->structureBuilt) { int rc = downloadEstimates(s); if (rc) return rc; }
  std::memcpy(host, s->g.vEst.data(), sizeof(double) * s->g.vEst.size());
  return G2OCU_OK;
}

// Landmark-sharded runs: the same two transfers restricted to what this rank works on - every pose estimate and the estimates of its own
// landmark range [lmBegin, lmEnd).  A rank never reads the other landmarks (its edges are those of its own landmarks), so the job's host
// side moves every estimate once per step instead of once per rank, and the result needs no all-gather.
int g2ocu_set_estimates_owned(g2ocu_solver* s, const double* host) {
  if (!s || !host) return G2OCU_E_INVALID;
  if (!s->hasGraph) return fail(s, G2OCU_E_INVALID, "no graph set");
  if (!(s->structureBuilt && s->fastEstimates) || s->world <= 1) return g2ocu_set_estimates(s, host);
  const Structure& st = s->st;
  CU(cudaMemcpyAsync(s->poseEst.p, host + s->poseHostOff, s->poseEst.n * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  if (st.lmEnd > st.lmBegin) {
    const size_t Sl = vertexEstimateDim(st.lmType), off = (size_t)st.lmBegin * Sl, cnt = (size_t)(st.lmEnd - st.lmBegin) * Sl;
    CU(cudaMemcpyAsync(s->lmEst.p + off, host + s->lmHostOff + off, cnt * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  }
  s->errorsValid = false;
  return syncStream(s);
}
int g2ocu_get_estimates_owned(g2ocu_solver* s, double* host) {
  if (!s || !host) return G2OCU_E_INVALID;
  if (!s->hasGraph) return fail(s, G2OCU_E_INVALID, "no graph set");
  if (!(s->structureBuilt && s->fastEstimates) || s->world <= 1) return g2ocu_get_estimates(s, host);
  const Structure& st = s->st;
  CU(cudaMemcpyAsync(host + s->poseHostOff, s->poseEst.p, s->poseEst.n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  if (st.lmEnd > st.lmBegin) {
    const size_t Sl = vertexEstimateDim(st.lmType), off = (size_t)st.lmBegin * Sl, cnt = (size_t)(st.lmEnd - st.lmBegin) * Sl;
    CU(cudaMemcpyAsync(host + s->lmHostOff + off, s->lmEst.p + off, cnt * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  }
  return syncStream(s);
}

static int64_t copyOutI32(const std::vector<int32_t>& v, int32_t* out, int64_t cap) { if (out) std::memcpy(out, v.data(), sizeof(int32_t) * (size_t)std::min<int64_t>(cap, (int64_t)v.size())); return (int64_t)v.size(); }

int64_t g2ocu_get_i32(g2ocu_solver* s, const char* name, int32_t* out, int64_t cap) {
  if (!s || !name) return G2OCU_E_INVALID;
  if (!s->optInitialized) return fail(s, G2OCU_E_INVALID, "initializeOptimization has not been called");
  const std::string n(name); const Structure& st = s->st;
  if (n == "hessian_index") return copyOutI32(st.hessianIndex, out, cap);
  if (n == "active_vertices") return copyOutI32(st.activeVertices, out, cap);
  if (n == "active_edges") return copyOutI32(st.activeEdges, out, cap);
  if (n == "index_mapping") return copyOutI32(st.ivMap, out, cap);
  if (n == "linear_solver_iterations_total") return copyOutI32({(int32_t)std::min<int64_t>(s->totalPcgIterations, INT32_MAX)}, out, cap);   // all solves since g2ocu_reset_counters
  if (n == "linear_solver_iterations") return copyOutI32({(int32_t)s->lastPcgIterations}, out, cap);   // G2OBatchStatistics::iterationsLinearSolver of the last solve (linear_solver_pcg.hpp:150-153)
  if (st.classOf.empty()) return fail(s, G2OCU_E_INVALID, "buildStructure has not been called");
  if (st.fullSystem) {   // what the reference's buildStructure holds for this graph: one Hpp over all vertices in id order
    if (n == "dims") return copyOutI32(st.refDims, out, cap);
    if (n == "pose_block_indices") return copyOutI32(st.refPoseBlockIndices, out, cap);
    if (n == "landmark_block_indices") return copyOutI32({}, out, cap);
    if (n == "hpp_colptr") return copyOutI32(st.refHppColPtr, out, cap);
    if (n == "hpp_rowidx") return copyOutI32(st.refHppRowIdx, out, cap);
    if (n == "edge_targets") return copyOutI32(st.refEdgeTargets, out, cap);
    if (n == "full_system_permutation") return copyOutI32(st.refToInternal, out, cap);
    if (n == "internal_dims") return copyOutI32({st.numPoses, st.numLandmarks, st.sizePoses, st.sizeLandmarks}, out, cap);
  }
  if (n == "tile_min_track") return copyOutI32({s->tileMinTrack}, out, cap);
  if (n == "dims" || n == "internal_dims") return copyOutI32({st.numPoses, st.numLandmarks, st.sizePoses, st.sizeLandmarks}, out, cap);
  if (n == "pose_block_indices") return copyOutI32(st.poseBlockIndices, out, cap);
  if (n == "landmark_block_indices") return copyOutI32(st.landmarkBlockIndices, out, cap);
  if (n == "hpp_colptr") return copyOutI32(st.hppColPtr, out, cap);
  if (n == "hpp_rowidx") return copyOutI32(st.hppRowIdx, out, cap);
  if (n == "hpl_colptr") return copyOutI32(st.hplColPtr, out, cap);
  if (n == "hpl_rowidx") return copyOutI32(st.hplRowIdx, out, cap);
  if (n == "hschur_colptr") return copyOutI32(st.sColPtr, out, cap);
  if (n == "hschur_rowidx") return copyOutI32(st.sRowIdx, out, cap);
  if (n == "hschur_t_colptr") return copyOutI32(st.sTRefRowPtr, out, cap);
  if (n == "hschur_t_rowidx") return copyOutI32(st.sTRefColIdx, out, cap);
  if (n == "edge_targets") return copyOutI32(st.edgeTargets, out, cap);
  if (n == "shard_landmark_range") return copyOutI32({st.lmBegin, st.lmEnd}, out, cap);
  if (n == "slab_block_range") {   // blocks of the reduced system this rank solves with (CSR order); everything when not sharded
    const int64_t nnz = (int64_t)st.sColIdx.size();
    const int64_t lo = s->world > 1 ? std::min(nnz, s->slabBlocks * s->rank) : 0, hi = s->world > 1 ? std::min(nnz, s->slabBlocks * (s->rank + 1)) : nnz;
    return copyOutI32({(int32_t)lo, (int32_t)hi}, out, cap);
  }
  if (n == "shard_edge_positions") { std::vector<int32_t> v; for (const EdgeSet& es : st.sets) v.insert(v.end(), es.pos.begin(), es.pos.end()); return copyOutI32(v, out, cap); }
  return fail(s, G2OCU_E_INVALID, "unknown int32 array " + n);
}

static int64_t downloadF64(g2ocu_solver* s, const double* dev, size_t count, double* out, int64_t cap) {
  if (out && count) {
    const size_t m = (size_t)std::min<int64_t>(cap, (int64_t)count);
    if (cudaMemcpyAsync(out, dev, m * sizeof(double), cudaMemcpyDeviceToHost, s->stream) != cudaSuccess || cudaStreamSynchronize(s->stream) != cudaSuccess)
      return fail(s, G2OCU_E_CUDA, "device to host copy failed");
    resolveEvents(s);
  }
  return (int64_t)count;
}
// block values in the reference's CCS order (block index permutation ccsToCsr), each block P x Q column-major
static int64_t downloadBlocks(g2ocu_solver* s, const double* dev, const std::vector<int32_t>& ccsToCsr, int bs, double lambdaDiag, const std::vector<int32_t>* diagCsr, double* out, int64_t cap) {
  const size_t count = ccsToCsr.size() * (size_t)bs;
  if (!out) return (int64_t)count;
  std::vector<double> tmp(count);
  if (count && (cudaMemcpyAsync(tmp.data(), dev, count * sizeof(double), cudaMemcpyDeviceToHost, s->stream) != cudaSuccess || cudaStreamSynchronize(s->stream) != cudaSuccess))
    return fail(s, G2OCU_E_CUDA, "device to host copy failed");
  resolveEvents(s);
  if (lambdaDiag != 0.0 && diagCsr) { int P = (int)std::lround(std::sqrt((double)bs)); for (int32_t k : *diagCsr) for (int q = 0; q < P; ++q) tmp[(size_t)k * bs + q * (P + 1)] += lambdaDiag; }
  for (size_t k = 0; k < ccsToCsr.size(); ++k) {
    if ((int64_t)((k + 1) * bs) > cap) break;
    std::memcpy(out + k * bs, &tmp[(size_t)ccsToCsr[k] * bs], sizeof(double) * bs);
  }
  return (int64_t)count;
}

int64_t g2ocu_get_f64(g2ocu_solver* s, const char* name, double* out, int64_t cap) {
  int rc = requireBuilt(s); if (rc) return rc;
  const std::string n(name ? name : ""); const Structure& st = s->st; const int P = st.P, L = st.L;
  if (st.fullSystem) {
    if (n == "x" || n == "b") {   // reference order: all vertices by id
      const size_t cnt = s->x.n;
      if (!out) return (int64_t)cnt;
      std::vector<double> tmp(cnt);
      const int64_t got = downloadF64(s, n == "x" ? s->x.p : s->b.p, cnt, tmp.data(), (int64_t)cnt); if (got < 0) return got;
      for (size_t i = 0; i < cnt && (int64_t)i < cap; ++i) out[i] = tmp[st.refToInternal[i]];
      return (int64_t)cnt;
    }
    if (n == "bschur" || n == "hpp_values" || n == "hschur_values" || n == "hpl_values" || n == "hll_values" || n == "dinv_values")
      return fail(s, G2OCU_E_UNSUPPORTED, "block values are not available in the reference's layout when the points are not marginalized (" + n + ")");
  }
  if (n == "x") return downloadF64(s, s->x.p, s->x.n, out, cap);
  if (n == "b") return downloadF64(s, s->b.p, s->b.n, out, cap);
  if (n == "bschur") return downloadF64(s, s->bschur.p, s->bschur.n, out, cap);
  if (n == "hpp_values") return downloadBlocks(s, s->Hpp.p, st.hppCcsToCsr, P * P, s->lambda, &st.hppDiag, out, cap);
  if (n == "hschur_values") return downloadBlocks(s, s->S.p, st.sCcsToCsr, P * P, 0.0, nullptr, out, cap);
  if (n == "hpl_values") return downloadF64(s, s->Hpl.p, s->st.hplRowIdx.size() * (size_t)s->st.P * s->st.L, out, cap);
  if (n == "hll_values") {
    const int64_t cnt = downloadF64(s, s->Hll.p, s->Hll.n, out, cap);
    if (out && s->lambda != 0.0) for (int64_t i = 0; i < st.numLandmarks && (i + 1) * L * L <= cap; ++i) for (int q = 0; q < L; ++q) out[i * L * L + q * (L + 1)] += s->lambda;
    return cnt;
  }
  if (n == "dinv_values") return downloadF64(s, s->Dinv.p, s->Dinv.n, out, cap);
  if (n == "diagonal_blocks") {   // the Hessian block of every vertex of the index mapping, in its order (what OptimizableGraph::Vertex::hessian maps, block_solver.hpp:150-170)
    int64_t total = 0;
    for (int32_t v : st.ivMap) { const int D = st.classOf[v] == 0 ? P : L; total += (int64_t)D * D; }
    if (!out) return total;
    std::vector<double> hpp(s->Hpp.n), hll(s->Hll.n);
    if (downloadF64(s, s->Hpp.p, hpp.size(), hpp.data(), (int64_t)hpp.size()) < 0) return G2OCU_E_CUDA;
    if (hll.size() && downloadF64(s, s->Hll.p, hll.size(), hll.data(), (int64_t)hll.size()) < 0) return G2OCU_E_CUDA;
    int64_t o = 0;
    for (int32_t v : st.ivMap) {
      const bool pose = st.classOf[v] == 0; const int D = pose ? P : L;
      if (o + D * D > cap) break;
      const double* src = pose ? &hpp[(size_t)st.hppDiag[st.slotOf[v]] * P * P] : &hll[(size_t)st.slotOf[v] * L * L];
      for (int k = 0; k < D * D; ++k) out[o + k] = src[k] + ((k % (D + 1)) == 0 ? s->lambda : 0.0);
      o += D * D;
    }
    return total;
  }
  if (n == "errors" || n == "jacobians") {
    const bool jac = n == "jacobians";
    std::vector<int64_t> off(st.activeEdges.size() + 1, 0);
    for (size_t k = 0; k < st.activeEdges.size(); ++k) {
      const int e = st.activeEdges[k], t = s->g.eType[e], E = edgeDim(t);
      off[k + 1] = off[k] + (jac ? E * (vertexDim(edgeVertexType(t, 0)) + vertexDim(edgeVertexType(t, 1))) : E);
    }
    if (!out) return off.back();
    if (s->off64.upload(off, s->stream) != cudaSuccess || s->dbg.alloc((size_t)off.back()) != cudaSuccess) return fail(s, G2OCU_E_CUDA, "allocation failed");
    if (jac) { for (auto* es : s->sets) launchJacobianDump(es->dev, s->sys, s->dbg.p, s->off64.p, s->stream, &s->launches); }
    else { rc = computeErrors(s, s->dbg.p, s->off64.p); if (rc) return rc; rc = finishErrors(s); if (rc) return rc; }
    return downloadF64(s, s->dbg.p, (size_t)off.back(), out, cap);
  }
  if (n == "estimates") { if (out && cap >= (int64_t)s->g.vEst.size()) { rc = g2ocu_get_estimates(s, out); if (rc) return rc; } return (int64_t)s->g.vEst.size(); }
  if (n == "lambda") { if (out && cap >= 1) out[0] = s->currentLambda; return 1; }
  if (n == "dogleg") {   // trustRegion(), lastStep() (optimization_algorithm_dogleg.h:66-68), tries of the last iteration, damping, positive-definite flag
    const double v[5] = {s->dlDelta, (double)s->dlLastStep, (double)s->dlLastNumTries, s->dlCurrentLambda, s->dlWasPD ? 1.0 : 0.0};
    if (out) std::memcpy(out, v, sizeof(double) * (size_t)std::min<int64_t>(cap, 5));
    return 5;
  }
  return fail(s, G2OCU_E_INVALID, "unknown double array " + n);
}

int64_t g2ocu_launch_count(const g2ocu_solver* s) { return s ? s->launches : 0; }
int g2ocu_phase_time(g2ocu_solver* s, const char* phase, double* seconds, int64_t* launches, int64_t* calls) {
  if (!s || !phase) return G2OCU_E_INVALID;
  auto it = s->phases.find(phase);
  if (seconds) *seconds = it == s->phases.end() ? 0.0 : it->second.seconds;
  if (launches) *launches = it == s->phases.end() ? 0 : it->second.launches;
  if (calls) *calls = it == s->phases.end() ? 0 : it->second.calls;
  return G2OCU_OK;
}
int g2ocu_reset_counters(g2ocu_solver* s) { if (!s) return G2OCU_E_INVALID; s->phases.clear(); s->launches = 0; s->totalPcgIterations = 0; return G2OCU_OK; }

// =================================================================================================
// LinearSolver<MatrixType> level (core/linear_solver.h:42-105): the block-Jacobi PCG of solvers/pcg/linear_solver_pcg.hpp:80-156 as a
// stand-alone solve of a symmetric block matrix handed over in the reference's own layout - for callers that keep g2o's BlockSolver
// (its CPU buildSystem and Schur complement) and only swap the linear solver.  The matrix crosses the bus on every solve.
struct g2ocu_linear_solver {
  g2ocu_solver core;                                  // stream, PCG state (_residual carry-over), counters
  std::vector<int32_t> colptr, rowidx;                // pattern of the last solve (the flattened view is rebuilt only when it changes)
  std::vector<int32_t> csrOfCcs;                      // CCS position -> CSR position of a block
  DVec<double> A, rhs; std::vector<double> staged;
};

int g2ocu_linear_create(const g2ocu_config* cfg, g2ocu_linear_solver** out) {
  if (!out) return fail(nullptr, G2OCU_E_INVALID, "null output pointer");
  g2ocu_linear_solver* L = new g2ocu_linear_solver;
  if (cfg) L->core.cfg = *cfg; else g2ocu_default_config(&L->core.cfg);
  L->core.cfg.linear_solver = G2OCU_LINEAR_PCG;
  *out = L;
  return G2OCU_OK;
}
void g2ocu_linear_destroy(g2ocu_linear_solver* L) { delete L; }
const char* g2ocu_linear_last_error(const g2ocu_linear_solver* L) { return L ? L->core.err.c_str() : g_createError.c_str(); }
int g2ocu_linear_init(g2ocu_linear_solver* L) { if (!L) return G2OCU_E_INVALID; L->core.pcgResidual = -1.0; return G2OCU_OK; }   // LinearSolverPCG::init, linear_solver_pcg.h:64-70
int g2ocu_linear_set_property(g2ocu_linear_solver* L, const char* name, double value) { return L ? g2ocu_set_property(&L->core, name, value) : G2OCU_E_INVALID; }

int g2ocu_linear_solve(g2ocu_linear_solver* L, int32_t nBlockCols, int32_t blockDim, const int32_t* colptr, const int32_t* rowidx, const double* values,
                       const double* b, double* x, int32_t* solved, int32_t* iterations) {
  if (!L || !colptr || !rowidx || !values || !b || !x || nBlockCols <= 0) return fail(L ? &L->core : nullptr, G2OCU_E_INVALID, "g2ocu_linear_solve: bad argument");
  g2ocu_solver* s = &L->core;
  if (blockDim != 3 && blockDim != 6 && blockDim != 9) return fail(s, G2OCU_E_UNSUPPORTED, "g2ocu_linear_solve: block dimension " + std::to_string(blockDim) + " (supported: 3, 6, 9; blocks of one size)");
  int rc = ensureCuda(s); if (rc) return rc;
  const int P = blockDim, PP = P * P; const int64_t nnz = colptr[nBlockCols];
  const bool samePattern = (int)L->colptr.size() == nBlockCols + 1 && std::memcmp(L->colptr.data(), colptr, sizeof(int32_t) * (nBlockCols + 1)) == 0 &&
                           (int64_t)L->rowidx.size() == nnz && std::memcmp(L->rowidx.data(), rowidx, sizeof(int32_t) * nnz) == 0 && s->pcg.P == P;
  if (!samePattern) {
    // upper blocks by column (ascending rows, the diagonal block last)  ->  CSR over the same blocks (row r: columns c >= r ascending, diagonal first)
    L->colptr.assign(colptr, colptr + nBlockCols + 1); L->rowidx.assign(rowidx, rowidx + nnz);
    std::vector<int32_t> rowPtr(nBlockCols + 1, 0), colIdx(nnz), diag(nBlockCols, -1);
    for (int c = 0; c < nBlockCols; ++c) for (int k = colptr[c]; k < colptr[c + 1]; ++k) { if (rowidx[k] > c || rowidx[k] < 0) return fail(s, G2OCU_E_INVALID, "g2ocu_linear_solve: expected the upper blocks (row <= column)"); rowPtr[rowidx[k] + 1]++; }
    for (int r = 0; r < nBlockCols; ++r) rowPtr[r + 1] += rowPtr[r];
    L->csrOfCcs.resize(nnz);
    { std::vector<int32_t> fill(rowPtr.begin(), rowPtr.end() - 1);
      for (int c = 0; c < nBlockCols; ++c) for (int k = colptr[c]; k < colptr[c + 1]; ++k) { const int o = fill[rowidx[k]]++; colIdx[o] = c; L->csrOfCcs[k] = o; } }
    for (int r = 0; r < nBlockCols; ++r) { if (rowPtr[r] == rowPtr[r + 1] || colIdx[rowPtr[r]] != r) return fail(s, G2OCU_E_INVALID, "g2ocu_linear_solve: block row " + std::to_string(r) + " has no diagonal block"); diag[r] = rowPtr[r]; }
    CU(s->aRowPtr.upload(rowPtr, s->stream)); CU(s->aColIdx.upload(colIdx, s->stream)); CU(s->aDiag.upload(diag, s->stream));
    std::vector<int32_t> ir, ib, ie; const int chunk = (32 / P) * 16;
    for (int r = 0; r < nBlockCols; ++r) for (int k = rowPtr[r]; k < rowPtr[r + 1]; k += chunk) { ir.push_back(r); ib.push_back(k); ie.push_back(std::min(k + chunk, rowPtr[r + 1])); }
    CU(s->spRow.upload(ir, s->stream)); CU(s->spBegin.upload(ib, s->stream)); CU(s->spEnd.upload(ie, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    PcgDev& pc = s->pcg; pc = PcgDev();
    pc.n = nBlockCols * P; pc.nb = nBlockCols; pc.P = P; pc.nnz = (int)nnz; pc.nItems = (int)ir.size(); pc.ownLo = 0; pc.ownHi = (int)nnz;
    CU(L->A.alloc((size_t)nnz * PP + 2)); CU(L->rhs.alloc(pc.n)); CU(s->x.alloc(pc.n));
    CU(s->Minv.alloc((size_t)nBlockCols * PP)); CU(s->vr.alloc(pc.n)); CU(s->vd.alloc(pc.n)); CU(s->vq.alloc(pc.n)); CU(s->vs.alloc(pc.n)); CU(s->scal.alloc(16));
    pc.nPartial = (pc.n + 255) / 256; pc.nPartialDq = std::min(296, (pc.n + 255) / 256);
    CU(s->partial.alloc(pc.nPartial)); CU(s->partialDq.alloc(pc.nPartialDq)); CU(s->pcgTicket.alloc(4)); CU(s->pcgTicket.zero(s->stream)); CU(s->scal.zero(s->stream));
    pc.rowPtr = s->aRowPtr.p; pc.colIdx = s->aColIdx.p; pc.diag = s->aDiag.p; pc.A = L->A.p; pc.Minv = s->Minv.p;
    pc.r = s->vr.p; pc.d = s->vd.p; pc.q = s->vq.p; pc.s = s->vs.p; pc.x = s->x.p; pc.scal = s->scal.p; pc.partial = s->partial.p; pc.partialDq = s->partialDq.p;
    pc.itemRow = s->spRow.p; pc.itemBegin = s->spBegin.p; pc.itemEnd = s->spEnd.p; pc.ticket = s->pcgTicket.p;
  }
  L->staged.resize((size_t)nnz * PP);
#pragma omp parallel for schedule(static)
  for (int64_t k = 0; k < nnz; ++k) std::memcpy(&L->staged[(size_t)L->csrOfCcs[k] * PP], values + (size_t)k * PP, sizeof(double) * PP);
  CU(cudaMemcpyAsync(L->A.p, L->staged.data(), sizeof(double) * (size_t)nnz * PP, cudaMemcpyHostToDevice, s->stream));
  CU(cudaMemcpyAsync(L->rhs.p, b, sizeof(double) * (size_t)s->pcg.n, cudaMemcpyHostToDevice, s->stream));
  s->st.doSchur = false; s->lambda = 0.0; s->world = 1;
  rc = solvePcg(s, L->rhs.p); if (rc) return rc;
  CU(cudaMemcpyAsync(x, s->x.p, sizeof(double) * (size_t)s->pcg.n, cudaMemcpyDeviceToHost, s->stream));
  rc = syncStream(s); if (rc) return rc;
  if (solved) *solved = 1;                           // LinearSolverPCG::solve always reports success (linear_solver_pcg.hpp:155)
  if (iterations) *iterations = s->lastPcgIterations;
  return G2OCU_OK;
}

}  // extern "C"
