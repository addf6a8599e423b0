// Host-side structure builder.  See host_structure.hpp.
#include "host_structure.hpp"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace g2ocu {

static const int kVEst[8] = {0, 3, 2, 12, 7, 3, 9, 3};
static const int kVDim[8] = {0, 3, 2, 6, 6, 3, 9, 3};
static const int kEDim[8] = {0, 3, 2, 6, 6, 2, 2, 2};
static const int kEMeas[8] = {0, 3, 2, 12, 7, 2, 2, 2};
static const int kEPrm[8] = {0, 0, 0, 0, 0, 3, 4, 0};
static const int kEVert[8][2] = {{0, 0}, {1, 1}, {1, 2}, {3, 3}, {4, 4}, {5, 4}, {5, 4}, {6, 7}};
int vertexEstimateDim(int t) { return (t >= 1 && t <= 7) ? kVEst[t] : -1; }
int vertexDim(int t) { return (t >= 1 && t <= 7) ? kVDim[t] : -1; }
int edgeDim(int t) { return (t >= 1 && t <= 7) ? kEDim[t] : -1; }
int edgeMeasDim(int t) { return (t >= 1 && t <= 7) ? kEMeas[t] : -1; }
int edgeParamDim(int t) { return (t >= 1 && t <= 7) ? kEPrm[t] : -1; }
int edgeVertexType(int t, int side) { return (t >= 1 && t <= 7) ? kEVert[t][side] : -1; }

bool HostGraph::assign(const g2ocu_graph* g, std::string& err) {
  if (!g || g->n_vertices < 0 || g->n_edges < 0) { err = "null or negative-sized graph"; return false; }
  nV = g->n_vertices; nE = g->n_edges;
  if (nV && (!g->v_id || !g->v_type || !g->v_fixed || !g->v_marginalized || !g->v_estimate)) { err = "null vertex array"; return false; }
  if (nE && (!g->e_type || !g->e_v0 || !g->e_v1 || !g->e_measurement || !g->e_information)) { err = "null edge array"; return false; }
  vId.assign(g->v_id, g->v_id + nV); vType.assign(g->v_type, g->v_type + nV);
  vFixed.assign(g->v_fixed, g->v_fixed + nV); vMarg.assign(g->v_marginalized, g->v_marginalized + nV);
  vEstOff.assign(nV + 1, 0);
  for (int i = 0; i < nV; ++i) {
    int d = vertexEstimateDim(vType[i]);
    if (d < 0) { err = "unsupported vertex type " + std::to_string(vType[i]) + " at vertex index " + std::to_string(i); return false; }
    vEstOff[i + 1] = vEstOff[i] + d;
  }
  vEst.assign(g->v_estimate, g->v_estimate + vEstOff[nV]);
  eType.assign(g->e_type, g->e_type + nE); eV0.assign(g->e_v0, g->e_v0 + nE); eV1.assign(g->e_v1, g->e_v1 + nE);
  if (g->e_level) eLevel.assign(g->e_level, g->e_level + nE); else eLevel.assign(nE, 0);
  if (g->e_kernel) eKernel.assign(g->e_kernel, g->e_kernel + nE); else eKernel.assign(nE, 0);
  if (g->e_kernel_delta) eDelta.assign(g->e_kernel_delta, g->e_kernel_delta + nE); else eDelta.assign(nE, 1.0);
  eMeasOff.assign(nE + 1, 0); eInfoOff.assign(nE + 1, 0); ePrmOff.assign(nE + 1, 0);
  for (int i = 0; i < nE; ++i) {
    int t = eType[i]; int D = edgeDim(t);
    if (D < 0) { err = "unsupported edge type " + std::to_string(t) + " at edge " + std::to_string(i) + " (rejected, no CPU fallback)"; return false; }
    if (eV0[i] < 0 || eV0[i] >= nV || eV1[i] < 0 || eV1[i] >= nV) { err = "edge vertex index out of range at edge " + std::to_string(i); return false; }
    if (vType[eV0[i]] != edgeVertexType(t, 0) || vType[eV1[i]] != edgeVertexType(t, 1)) {
      err = "edge " + std::to_string(i) + " of type " + std::to_string(t) + " connects vertices of the wrong type"; return false; }
    if (eV0[i] == eV1[i]) { err = "edge " + std::to_string(i) + " connects a vertex to itself"; return false; }
    if (eKernel[i] < 0 || eKernel[i] > 9) { err = "unsupported robust kernel code at edge " + std::to_string(i); return false; }
    eMeasOff[i + 1] = eMeasOff[i] + edgeMeasDim(t); eInfoOff[i + 1] = eInfoOff[i] + D * D; ePrmOff[i + 1] = ePrmOff[i] + edgeParamDim(t);
  }
  eMeas.assign(g->e_measurement, g->e_measurement + eMeasOff[nE]);
  eInfo.assign(g->e_information, g->e_information + eInfoOff[nE]);
  if (ePrmOff[nE]) { if (!g->e_param) { err = "edge parameters required but e_param is null"; return false; } ePrm.assign(g->e_param, g->e_param + ePrmOff[nE]); }
  else ePrm.clear();
  // adjacency (counting sort)
  adjPtr.assign(nV + 1, 0);
  for (int i = 0; i < nE; ++i) { adjPtr[eV0[i] + 1]++; adjPtr[eV1[i] + 1]++; }
  for (int i = 0; i < nV; ++i) adjPtr[i + 1] += adjPtr[i];
  adjEdge.assign(adjPtr[nV], 0);
  std::vector<int64_t> fill(adjPtr.begin(), adjPtr.end() - 1);
  for (int i = 0; i < nE; ++i) { adjEdge[fill[eV0[i]]++] = i; adjEdge[fill[eV1[i]]++] = i; }
  return true;
}

// sparse_optimizer.cpp:208-279 (vset = all vertices), sortVectorContainers :504-509, buildIndexMapping :168-193
bool initializeOptimization(const HostGraph& g, int level, Structure& st, std::string& err) {
  st = Structure();
  if (g.nE == 0) { err = "Attempt to initialize an empty graph"; return false; }
  std::vector<uint8_t> vActive(g.nV, 0);
  st.activeEdges.clear();
  for (int e = 0; e < g.nE; ++e) {
    if (!(level < 0 || g.eLevel[e] == level)) continue;
    if (g.vFixed[g.eV0[e]] && g.vFixed[g.eV1[e]]) continue;         // allVerticesFixed
    st.activeEdges.push_back(e);                                    // already in internalId order
    vActive[g.eV0[e]] = 1; vActive[g.eV1[e]] = 1;
  }
  for (int v = 0; v < g.nV; ++v) if (vActive[v]) st.activeVertices.push_back(v);
  std::stable_sort(st.activeVertices.begin(), st.activeVertices.end(), [&](int a, int b) { return g.vId[a] < g.vId[b]; });
  st.hessianIndex.assign(g.nV, -1);
  if (st.activeVertices.empty()) { err = "no active vertices"; return false; }
  int i = 0;
  for (int k = 0; k < 2; ++k)
    for (int v : st.activeVertices)
      if (!g.vFixed[v] && (int)(g.vMarg[v] != 0) == k) { st.hessianIndex[v] = i++; st.ivMap.push_back(v); }
  return true;
}

// G2OCU_TRACE=1: wall time of the phases of buildStructure on stderr
struct TracePhase {
  static bool on() { static const bool v = [] { const char* e = std::getenv("G2OCU_TRACE"); return e && *e && *e != '0'; }(); return v; }
  std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
  void mark(const char* what) {
    if (!on()) return;
    const auto n = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[g2ocu] buildStructure %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
    t = n;
  }
};

// Stable order of [0, n) by (hi, lo), keys in [0, nHi) x [0, nLo): two counting-sort passes (least significant key first) instead of a
// comparison sort - the edge lists are millions long and the keys are vertex slots.
static void orderByTwoKeys(const std::vector<int32_t>& hi, int nHi, const std::vector<int32_t>& lo, int nLo, std::vector<int32_t>& ord) {
  const size_t n = hi.size();
  std::vector<int32_t> tmp(n); ord.resize(n);
  { std::vector<int64_t> cnt((size_t)nLo + 1, 0);
    for (size_t i = 0; i < n; ++i) cnt[lo[i] + 1]++;
    for (int b = 0; b < nLo; ++b) cnt[b + 1] += cnt[b];
    for (size_t i = 0; i < n; ++i) tmp[cnt[lo[i]]++] = (int32_t)i; }
  { std::vector<int64_t> cnt((size_t)nHi + 1, 0);
    for (size_t i = 0; i < n; ++i) cnt[hi[i] + 1]++;
    for (int b = 0; b < nHi; ++b) cnt[b + 1] += cnt[b];
    for (size_t j = 0; j < n; ++j) { const int32_t i = tmp[j]; ord[cnt[hi[i]]++] = i; } }
}

template <class T> static void sortUnique(std::vector<T>& v) { std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end()); }

bool buildStructure(const HostGraph& g, Structure& st, std::string& err, int rank, int world) {
  if (st.ivMap.empty()) { err = "0 vertices to optimize, maybe forgot to call initializeOptimization()"; return false; }
  TracePhase trace;
  // Schur iff any active vertex is marginalized (optimization_algorithm_with_hessian.cpp:48-66)
  st.doSchur = false; st.fullSystem = false;
  for (int v : st.activeVertices) if (g.vMarg[v]) { st.doSchur = true; break; }
  auto isPoint = [](int t) { return t == G2OCU_VERTEX_POINT_XY || t == G2OCU_VERTEX_POINT_XYZ || t == G2OCU_VERTEX_POINT_BAL; };
  if (!st.doSchur) {
    bool anyPoint = false, anyPose = false;
    for (int v : st.ivMap) { if (isPoint(g.vType[v])) anyPoint = true; else anyPose = true; }   // free vertices only
    st.fullSystem = anyPoint && anyPose;
  }
  // class of a vertex (1 = landmark class) and its index in the [poses | landmarks] order.  Full-system mode: what the flags and
  // indices would be had the caller marginalized the points - a permutation of the reference's order, undone at the boundary.
  std::vector<uint8_t> effMarg(g.nV, 0); std::vector<int32_t> effH(st.hessianIndex);
  for (int v = 0; v < g.nV; ++v) effMarg[v] = st.fullSystem ? (uint8_t)isPoint(g.vType[v]) : (uint8_t)(g.vMarg[v] != 0);
  if (st.fullSystem) {
    st.doSchur = true;
    int i = 0;
    for (int k = 0; k < 2; ++k) for (int v : st.ivMap) if ((int)effMarg[v] == k) effH[v] = i++;
  }
  st.classOf.assign(g.nV, -1); st.slotOf.assign(g.nV, -1);
  st.numPoses = st.numLandmarks = 0; st.poseType = st.lmType = 0;
  for (int v : st.ivMap) { if (!effMarg[v]) st.numPoses++; else st.numLandmarks++; }
  st.poseVerts.clear(); st.lmVerts.clear();
  std::vector<int32_t> byEff(st.ivMap.size());
  for (int v : st.ivMap) byEff[effH[v]] = v;
  for (int v : byEff) {
    int c = effMarg[v] ? 1 : 0; st.classOf[v] = c;
    int& ty = c ? st.lmType : st.poseType;
    if (ty == 0) ty = g.vType[v];
    else if (ty != g.vType[v]) { err = std::string("mixed vertex types inside the ") + (c ? "landmark" : "pose") + " block are not supported (uniform block size required)"; return false; }
    st.slotOf[v] = c ? effH[v] - st.numPoses : effH[v];
    (c ? st.lmVerts : st.poseVerts).push_back(v);
  }
  for (int v : st.activeVertices) {
    if (!g.vFixed[v]) continue;
    int c = effMarg[v] ? 1 : 0; st.classOf[v] = c;
    int& ty = c ? st.lmType : st.poseType;
    if (ty == 0) ty = g.vType[v];
    else if (ty != g.vType[v]) { err = "mixed vertex types inside one block class are not supported"; return false; }
    auto& list = c ? st.lmVerts : st.poseVerts;
    st.slotOf[v] = (int)list.size(); list.push_back(v);
  }
  st.numPoseSlots = (int)st.poseVerts.size(); st.numLmSlots = (int)st.lmVerts.size();
  if (st.numPoses == 0) { err = "no free pose vertex (everything is fixed or marginalized)"; return false; }
  st.P = vertexDim(st.poseType); st.L = st.lmType ? vertexDim(st.lmType) : 0;
  st.sizePoses = st.numPoses * st.P; st.sizeLandmarks = st.numLandmarks * st.L;
  st.poseBlockIndices.resize(st.numPoses); st.landmarkBlockIndices.resize(st.numLandmarks);
  for (int i = 0; i < st.numPoses; ++i) st.poseBlockIndices[i] = (i + 1) * st.P;
  for (int i = 0; i < st.numLandmarks; ++i) st.landmarkBlockIndices[i] = (i + 1) * st.L;

  trace.mark("classes");
  // ---- per-edge targets (block_solver.hpp:166-214) ----
  const int nA = (int)st.activeEdges.size();
  st.edgeTargets.assign((size_t)nA * 4, -1);
  std::vector<int64_t> ppPairs;                // encoded (row<<32|col) for Hpp (row<=col)
  std::vector<int32_t> plLm, plPose;           // (landmark, pose) of every Hpl contribution
  bool lmLmEdge = false;
#pragma omp parallel for schedule(static) reduction(|| : lmLmEdge)
  for (int k = 0; k < nA; ++k) {
    int e = st.activeEdges[k]; int v0 = g.eV0[e], v1 = g.eV1[e];
    int h0 = effH[v0], h1 = effH[v1];
    int* t = &st.edgeTargets[(size_t)k * 4]; t[3] = 0;
    if (h0 == -1 || h1 == -1) continue;
    bool m0 = effMarg[v0], m1 = effMarg[v1];
    if (!m0 && !m1) {
      int a = h0, b = h1; bool tr = a > b; if (tr) std::swap(a, b);
      t[0] = 0; t[1] = a; t[2] = b; t[3] = tr;
    } else if (m0 && m1) lmLmEdge = true;
    else if (m0) { t[0] = 2; t[1] = h1; t[2] = h0 - st.numPoses; t[3] = 1; }
    else { t[0] = 2; t[1] = h0; t[2] = h1 - st.numPoses; t[3] = 0; }
  }
  if (lmLmEdge) { err = st.fullSystem ? "edge between two point vertices is not supported" : "edge between two marginalized vertices (landmark-landmark block) is not supported"; return false; }
  for (int k = 0; k < nA; ++k) {
    const int* t = &st.edgeTargets[(size_t)k * 4];
    if (t[0] == 0) ppPairs.push_back(((int64_t)t[1] << 32) | (uint32_t)t[2]);
    else if (t[0] == 2) { plLm.push_back(t[2]); plPose.push_back(t[1]); }
  }
  trace.mark("edge targets");
  // ---- Hpp pattern: diagonal + pose-pose edges, upper ----
  const size_t ppEdges = ppPairs.size();
  for (int i = 0; i < st.numPoses; ++i) ppPairs.push_back(((int64_t)i << 32) | (uint32_t)i);
  sortUnique(ppPairs);
  st.hppShared = false;
  {
    // duplicates among the edge pairs => two edges share one off-diagonal block
    std::vector<int64_t> chk; chk.reserve(ppEdges);
    for (int k = 0; k < nA; ++k) { const int* t = &st.edgeTargets[(size_t)k * 4]; if (t[0] == 0) chk.push_back(((int64_t)t[1] << 32) | (uint32_t)t[2]); }
    std::sort(chk.begin(), chk.end());
    st.hppShared = std::adjacent_find(chk.begin(), chk.end()) != chk.end();
  }
  const int nnzPP = (int)ppPairs.size();
  st.hppRowPtr.assign(st.numPoses + 1, 0); st.hppColIdx.resize(nnzPP); st.hppDiag.assign(st.numPoses, -1);
  for (int k = 0; k < nnzPP; ++k) { int r = (int)(ppPairs[k] >> 32), c = (int)(ppPairs[k] & 0xffffffff); st.hppRowPtr[r + 1]++; st.hppColIdx[k] = c; if (r == c) st.hppDiag[r] = k; }
  for (int i = 0; i < st.numPoses; ++i) st.hppRowPtr[i + 1] += st.hppRowPtr[i];
  auto transposeToCcs = [](int n, const std::vector<int32_t>& rowPtr, const std::vector<int32_t>& colIdx,
                           std::vector<int32_t>& colPtr, std::vector<int32_t>& rowIdx, std::vector<int32_t>& ccsToCsr) {
    const int nnz = (int)colIdx.size();
    colPtr.assign(n + 1, 0); rowIdx.resize(nnz); ccsToCsr.resize(nnz);
    for (int k = 0; k < nnz; ++k) colPtr[colIdx[k] + 1]++;
    for (int i = 0; i < n; ++i) colPtr[i + 1] += colPtr[i];
    std::vector<int32_t> fill(colPtr.begin(), colPtr.end() - 1);
    for (int r = 0; r < n; ++r) for (int k = rowPtr[r]; k < rowPtr[r + 1]; ++k) { int p = fill[colIdx[k]]++; rowIdx[p] = r; ccsToCsr[p] = k; }
  };
  transposeToCcs(st.numPoses, st.hppRowPtr, st.hppColIdx, st.hppColPtr, st.hppRowIdx, st.hppCcsToCsr);

  trace.mark("Hpp pattern");
  // ---- Hpl pattern: CCS by landmark, ascending pose rows ----
  std::vector<int64_t> plPairs(plLm.size());   // (lm<<32|pose), sorted, unique (a comparison sort of packed keys beats two counting passes over 1M buckets here)
  {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)plLm.size(); ++i) plPairs[i] = ((int64_t)plLm[i] << 32) | (uint32_t)plPose[i];
    std::sort(plPairs.begin(), plPairs.end());
    st.hplShared = std::adjacent_find(plPairs.begin(), plPairs.end()) != plPairs.end();
    plPairs.erase(std::unique(plPairs.begin(), plPairs.end()), plPairs.end());
  }
  const int nnzPL = (int)plPairs.size();
  st.hplColPtr.assign(st.numLandmarks + 1, 0); st.hplRowIdx.resize(nnzPL);
  for (int k = 0; k < nnzPL; ++k) { st.hplColPtr[(int)(plPairs[k] >> 32) + 1]++; st.hplRowIdx[k] = (int)(plPairs[k] & 0xffffffff); }
  for (int i = 0; i < st.numLandmarks; ++i) st.hplColPtr[i + 1] += st.hplColPtr[i];
  st.schurPairs = 0;
  for (int l = 0; l < st.numLandmarks; ++l) { int64_t k = st.hplColPtr[l + 1] - st.hplColPtr[l]; st.schurPairs += k * (k + 1) / 2; }

  trace.mark("Hpl pattern");
  // ---- Schur pattern (block_solver.hpp:190-192, 224-251) ----
  if (st.doSchur) {
    // cams(l): pose hessian indices reachable through ANY incident edge of the landmark vertex (the reference walks
    // v->edges(), not the active edge set), other endpoint must be indexed (hessianIndex != -1)
    std::vector<int32_t> lcPtr(st.numLandmarks + 1, 0), lcIdx;
    {
      std::vector<std::vector<int32_t>> tmp;   // built per landmark, small
      lcIdx.reserve(nnzPL);
      std::vector<int32_t> buf;
      for (int l = 0; l < st.numLandmarks; ++l) {
        int v = st.lmVerts[l]; buf.clear();
        for (int64_t a = g.adjPtr[v]; a < g.adjPtr[v + 1]; ++a) {
          int e = g.adjEdge[a]; int o = g.eV0[e] == v ? g.eV1[e] : g.eV0[e];
          int h = effH[o];
          if (h == -1) continue;
          if (h >= st.numPoses) { err = "edge between two marginalized vertices is not supported"; return false; }
          buf.push_back(h);
        }
        sortUnique(buf);
        lcIdx.insert(lcIdx.end(), buf.begin(), buf.end());
        lcPtr[l + 1] = (int32_t)lcIdx.size();
      }
    }
    trace.mark("  cams per landmark");
    // pose -> landmarks (transpose)
    std::vector<int64_t> clPtr(st.numPoses + 1, 0); std::vector<int32_t> clIdx(lcIdx.size());
    for (int32_t c : lcIdx) clPtr[c + 1]++;
    for (int i = 0; i < st.numPoses; ++i) clPtr[i + 1] += clPtr[i];
    { std::vector<int64_t> fill(clPtr.begin(), clPtr.end() - 1);
      for (int l = 0; l < st.numLandmarks; ++l) for (int k = lcPtr[l]; k < lcPtr[l + 1]; ++k) clIdx[fill[lcIdx[k]]++] = l; }
    trace.mark("  landmarks per cam");
    std::vector<std::vector<int32_t>> rows(st.numPoses);
#pragma omp parallel
    {
      std::vector<int32_t> mark(st.numPoses, -1);
#pragma omp for schedule(dynamic, 8)
      for (int i1 = 0; i1 < st.numPoses; ++i1) {
        auto& out = rows[i1];
        for (int k = st.hppRowPtr[i1]; k < st.hppRowPtr[i1 + 1]; ++k) { int c = st.hppColIdx[k]; if (mark[c] != i1) { mark[c] = i1; out.push_back(c); } }
        for (int64_t a = clPtr[i1]; a < clPtr[i1 + 1]; ++a) {
          int l = clIdx[a];
          const int32_t* cb = &lcIdx[lcPtr[l]]; const int32_t* ce = &lcIdx[lcPtr[l + 1]];
          for (const int32_t* p = std::lower_bound(cb, ce, i1); p != ce; ++p) if (mark[*p] != i1) { mark[*p] = i1; out.push_back(*p); }
        }
        std::sort(out.begin(), out.end());
      }
    }
    trace.mark("  marker sweep");
    st.sRowPtr.assign(st.numPoses + 1, 0);
    for (int i = 0; i < st.numPoses; ++i) st.sRowPtr[i + 1] = st.sRowPtr[i] + (int)rows[i].size();
    st.sColIdx.resize(st.sRowPtr[st.numPoses]); st.sDiag.assign(st.numPoses, -1);
    for (int i = 0; i < st.numPoses; ++i) { std::copy(rows[i].begin(), rows[i].end(), st.sColIdx.begin() + st.sRowPtr[i]); st.sDiag[i] = st.sRowPtr[i]; }
    transposeToCcs(st.numPoses, st.sRowPtr, st.sColIdx, st.sColPtr, st.sRowIdx, st.sCcsToCsr);
    // The reference's _HschurTransposedCCS is filled once in buildStructure (block_solver.hpp:253) from the co-observation pairs and the
    // pose-pose edges; the diagonal blocks of poses that observe no landmark join Hschur only in the first solve (_Hpp->add, :333-335)
    // and never reach that mirror.  The read-back arrays reproduce it as the reference holds it.
    st.sTRefRowPtr.assign(1, 0); st.sTRefColIdx.clear();
    for (int i = 0; i < st.numPoses; ++i) {
      for (int k = st.sRowPtr[i]; k < st.sRowPtr[i + 1]; ++k) if (st.sColIdx[k] != i || clPtr[i + 1] > clPtr[i]) st.sTRefColIdx.push_back(st.sColIdx[k]);
      st.sTRefRowPtr.push_back((int32_t)st.sTRefColIdx.size());
    }
    st.hppToS.resize(nnzPP);
    for (int r = 0; r < st.numPoses; ++r)
      for (int k = st.hppRowPtr[r]; k < st.hppRowPtr[r + 1]; ++k) {
        const int32_t* b = &st.sColIdx[st.sRowPtr[r]]; const int32_t* e = &st.sColIdx[st.sRowPtr[r + 1]];
        st.hppToS[k] = (int)(std::lower_bound(b, e, st.hppColIdx[k]) - st.sColIdx.data());
      }
  } else {
    st.sRowPtr.clear(); st.sColIdx.clear(); st.sColPtr.clear(); st.sRowIdx.clear(); st.sCcsToCsr.clear(); st.sDiag.clear(); st.hppToS.clear(); st.sTRefRowPtr.clear(); st.sTRefColIdx.clear();
  }

  trace.mark("Hschur pattern");
  // ---- landmark ownership for sharded runs: contiguous slot ranges balanced by a cost model of the per-landmark work ----
  // k observations cost ~k in the build / coefficient / back-substitution passes and k (k + 1) / 2 block products in the Schur complement;
  // the weights are the measured single-GPU times per unit on C3 (0.40 ns per observation, 0.085 ns per block product).
  st.lmBegin = 0; st.lmEnd = st.numLandmarks;
  if (world > 1 && st.doSchur) {
    std::vector<int64_t> cum((size_t)st.numLandmarks + 1, 0);
    for (int l = 0; l < st.numLandmarks; ++l) {
      const int64_t k = st.hplColPtr[l + 1] - st.hplColPtr[l];
      cum[l + 1] = cum[l] + 400 * k + 85 * (k * (k + 1) / 2);
    }
    const int64_t total = cum[st.numLandmarks];
    auto splitAt = [&](int r) -> int {
      if (r <= 0) return 0;
      if (r >= world) return st.numLandmarks;
      const int64_t target = total / world * r;
      return (int)(std::lower_bound(cum.begin(), cum.end(), target) - cum.begin());
    };
    st.lmBegin = std::min(splitAt(rank), st.numLandmarks); st.lmEnd = std::min(splitAt(rank + 1), st.numLandmarks);
    if (rank == world - 1) st.lmEnd = st.numLandmarks;
  }
  // ---- edge sets in kernel order ----
  st.sets.clear();
  std::vector<int> setOfType(8, -1);
  for (int k = 0; k < nA; ++k) {
    int e = st.activeEdges[k]; int t = g.eType[e];
    if (setOfType[t] < 0) { setOfType[t] = (int)st.sets.size(); st.sets.emplace_back(); st.sets.back().etype = t; }
    st.sets[setOfType[t]].pos.push_back(k);
  }
  for (auto& s : st.sets) {
    int n = (int)s.pos.size();
    int e0 = st.activeEdges[s.pos[0]];
    int c0 = st.classOf[g.eV0[e0]], c1 = st.classOf[g.eV1[e0]];
    s.poseLandmark = (c0 != c1); s.poseSide = (c0 == 0) ? 0 : 1;
    if (c0 == 1 && c1 == 1) { err = "edge between two marginalized vertices is not supported"; return false; }
    if (world > 1) {   // this rank's share: edges of owned landmarks (fixed landmarks: rank 0); pose-pose edges split evenly
      std::vector<int32_t> keep;
      for (int i = 0; i < n; ++i) {
        const int e = st.activeEdges[s.pos[i]];
        bool mine;
        if (s.poseLandmark) { const int vl = s.poseSide == 0 ? g.eV1[e] : g.eV0[e]; const int sl = st.slotOf[vl]; mine = sl < st.numLandmarks ? (sl >= st.lmBegin && sl < st.lmEnd) : rank == 0; }
        else mine = i >= (int64_t)n * rank / world && i < (int64_t)n * (rank + 1) / world;
        if (mine) keep.push_back(s.pos[i]);
      }
      s.pos.swap(keep);
      n = (int)s.pos.size();
    }
    if (s.poseLandmark) {   // kernel order: by (landmark slot, pose slot), ties in internalId order
      std::vector<int32_t> kl(n), kp(n), ord;
#pragma omp parallel for schedule(static)
      for (int i = 0; i < n; ++i) {
        int e = st.activeEdges[s.pos[i]];
        int vp = s.poseSide == 0 ? g.eV0[e] : g.eV1[e], vl = s.poseSide == 0 ? g.eV1[e] : g.eV0[e];
        kl[i] = st.slotOf[vl]; kp[i] = st.slotOf[vp];
      }
      orderByTwoKeys(kl, std::max(st.numLmSlots, 1), kp, std::max(st.numPoseSlots, 1), ord);
      std::vector<int32_t> npos(n); for (int i = 0; i < n; ++i) npos[i] = s.pos[ord[i]];
      s.pos.swap(npos);
    }
    s.slot0.resize(n); s.slot1.resize(n); s.block.assign(n, -1); s.transposed.assign(n, 0);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
      int k = s.pos[i]; int e = st.activeEdges[k];
      s.slot0[i] = st.slotOf[g.eV0[e]]; s.slot1[i] = st.slotOf[g.eV1[e]];
      const int* t = &st.edgeTargets[(size_t)k * 4];
      s.transposed[i] = (uint8_t)t[3];
      if (t[0] == 0) {
        const int32_t* b = &st.hppColIdx[st.hppRowPtr[t[1]]]; const int32_t* en = &st.hppColIdx[st.hppRowPtr[t[1] + 1]];
        s.block[i] = (int)(std::lower_bound(b, en, t[2]) - st.hppColIdx.data());
      } else if (t[0] == 2) {
        const int32_t* b = &st.hplRowIdx[st.hplColPtr[t[2]]]; const int32_t* en = &st.hplRowIdx[st.hplColPtr[t[2] + 1]];
        s.block[i] = (int)(std::lower_bound(b, en, t[1]) - st.hplRowIdx.data());
      }
    }
    if (s.poseLandmark) {
      const std::vector<int32_t>& pslot = s.poseSide == 0 ? s.slot0 : s.slot1;
      std::vector<int32_t> cnt(st.numPoses + 1, 0);
      for (int i = 0; i < n; ++i) if (pslot[i] < st.numPoses) cnt[pslot[i] + 1]++;
      for (int i = 0; i < st.numPoses; ++i) cnt[i + 1] += cnt[i];
      s.byPose.resize(cnt[st.numPoses]);
      { std::vector<int32_t> fill(cnt.begin(), cnt.end() - 1); for (int i = 0; i < n; ++i) if (pslot[i] < st.numPoses) s.byPose[fill[pslot[i]]++] = i; }
      s.poseChunkPtr.assign(st.numPoses + 1, 0);
      for (int p = 0; p < st.numPoses; ++p) {
        for (int b = cnt[p]; b < cnt[p + 1]; b += kChunk) { s.chunkPose.push_back(p); s.chunkBegin.push_back(b); s.chunkEnd.push_back(std::min(b + kChunk, cnt[p + 1])); }
        s.poseChunkPtr[p + 1] = (int)s.chunkPose.size();
      }
    }
  }
  trace.mark("edge sets");
  // ---- full-system mode: the arrays of the reference's own buildStructure (everything in one Hpp, vertices by id) ----
  st.refDims.clear(); st.refPoseBlockIndices.clear(); st.refHppColPtr.clear(); st.refHppRowIdx.clear(); st.refEdgeTargets.clear(); st.refToInternal.clear();
  if (st.fullSystem) {
    const int nAll = (int)st.ivMap.size();
    int off = 0;
    for (int v : st.ivMap) {
      const int d = vertexDim(g.vType[v]);
      const int base = st.classOf[v] ? st.sizePoses + st.slotOf[v] * st.L : st.slotOf[v] * st.P;
      for (int q = 0; q < d; ++q) st.refToInternal.push_back(base + q);
      off += d; st.refPoseBlockIndices.push_back(off);
    }
    st.refDims = {nAll, 0, off, 0};
    st.refEdgeTargets.assign((size_t)nA * 4, -1);
    std::vector<int64_t> pairs;                    // (col << 32 | row), row <= col
    for (int i = 0; i < nAll; ++i) pairs.push_back(((int64_t)i << 32) | (uint32_t)i);
    for (int k = 0; k < nA; ++k) {
      const int e = st.activeEdges[k]; const int h0 = st.hessianIndex[g.eV0[e]], h1 = st.hessianIndex[g.eV1[e]];
      int* t = &st.refEdgeTargets[(size_t)k * 4]; t[3] = 0;
      if (h0 == -1 || h1 == -1) continue;
      const int a = std::min(h0, h1), b = std::max(h0, h1);
      t[0] = 0; t[1] = a; t[2] = b; t[3] = h0 > h1;
      pairs.push_back(((int64_t)b << 32) | (uint32_t)a);
    }
    sortUnique(pairs);
    st.refHppColPtr.assign(nAll + 1, 0); st.refHppRowIdx.resize(pairs.size());
    for (size_t k = 0; k < pairs.size(); ++k) { st.refHppColPtr[(int)(pairs[k] >> 32) + 1]++; st.refHppRowIdx[k] = (int)(pairs[k] & 0xffffffff); }
    for (int i = 0; i < nAll; ++i) st.refHppColPtr[i + 1] += st.refHppColPtr[i];
  }
  return true;
}

}  // namespace g2ocu
