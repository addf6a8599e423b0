// Host-side (integer) half of the backend: active sets + index mapping (SparseOptimizer::initializeOptimization,
// sparse_optimizer.cpp:208-279,168-193) and the block index map (BlockSolver::buildStructure, block_solver.hpp:103-256).
// Pure C++ — no CUDA — so the bit-exact structure tests run without a GPU.  The algorithms are sort/scan based
// (CSR adjacency, per-row marker sweep for the Schur pattern) instead of the reference's std::map / hash-map inserts.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/g2ocu.h"

namespace g2ocu {

int vertexEstimateDim(int vtype);
int vertexDim(int vtype);
int edgeDim(int etype);
int edgeMeasDim(int etype);
int edgeParamDim(int etype);
int edgeVertexType(int etype, int side);

struct HostGraph {
  int nV = 0, nE = 0;
  std::vector<int32_t> vId, vType; std::vector<uint8_t> vFixed, vMarg; std::vector<double> vEst; std::vector<int64_t> vEstOff;
  std::vector<int32_t> eType, eV0, eV1, eLevel, eKernel; std::vector<double> eMeas, eInfo, eDelta, ePrm;
  std::vector<int64_t> eMeasOff, eInfoOff, ePrmOff;
  std::vector<int64_t> adjPtr; std::vector<int32_t> adjEdge;   // vertex -> incident edges (all levels)
  bool assign(const g2ocu_graph* g, std::string& err);
};

// One homogeneous group of active edges (same edge type), in the order the kernels consume it.
struct EdgeSet {
  int etype = 0;
  bool poseLandmark = false;      // true: one vertex in the pose class and one in the landmark class
  int poseSide = 0;               // for poseLandmark sets: which of vertices()[0/1] is the pose
  std::vector<int32_t> pos;       // position in the active-edge list, in kernel order
  std::vector<int32_t> slot0, slot1;   // class slot of vertices()[0], vertices()[1]
  std::vector<int32_t> block;     // target off-diagonal block (internal index into Hpl / Hpp-CSR), -1 = none
  std::vector<uint8_t> transposed;
  // pose-landmark sets only: landmark segments over the kernel order, and the pose-sorted view
  std::vector<int32_t> byPose;    // permutation of [0,n): kernel-order indices sorted by pose slot (non-fixed poses only)
  std::vector<int32_t> chunkPose, chunkBegin, chunkEnd;   // chunks of byPose, one pose per chunk, <= kChunk edges
  std::vector<int32_t> poseChunkPtr;                      // pose slot -> [first chunk, last chunk) (size numPoses+1)
};

struct Structure {
  // ---- initializeOptimization ----
  std::vector<int32_t> hessianIndex, activeVertices, activeEdges, ivMap;
  // ---- buildStructure ----
  bool doSchur = false;
  // Full-system mode: no vertex is marginalized but the graph holds pose-type AND point-type vertices (`lm_var` on a SLAM / BA graph,
  // BlockSolverX with blocks of two sizes).  Internally the points still form the landmark class (the same Hpp / Hpl / Hll storage and
  // build kernels as with Schur, doSchur is set), but the reference solves the whole system: PCG runs over [Hpp Hpl; Hpl^T Hll] and the
  // vectors of the boundary (x, b, update) keep the reference's order (all vertices by id).  ref* hold what the reference's
  // buildStructure produces for this graph (one Hpp over all vertices) for the bit-exact structure check.
  bool fullSystem = false;
  std::vector<int32_t> refDims, refPoseBlockIndices, refHppColPtr, refHppRowIdx, refEdgeTargets;
  std::vector<int32_t> refToInternal;    // scalar index in the reference's x / b  ->  scalar index in the internal [poses | points] layout
  int numPoses = 0, numLandmarks = 0, sizePoses = 0, sizeLandmarks = 0;
  int P = 0, L = 0;                      // uniform block dimensions
  int poseType = 0, lmType = 0;          // vertex type of each class (0 = class empty)
  int numPoseSlots = 0, numLmSlots = 0;  // including fixed vertices (slots >= numPoses / numLandmarks)
  std::vector<int32_t> poseBlockIndices, landmarkBlockIndices;
  std::vector<int32_t> classOf, slotOf;  // per vertex: 0 pose class, 1 landmark class, -1 inactive; slot inside the class
  std::vector<int32_t> poseVerts, lmVerts;   // slot -> vertex index
  // Hpp: reference CCS (column-major, ascending rows, upper) and internal CSR over the same upper blocks
  std::vector<int32_t> hppColPtr, hppRowIdx, hppRowPtr, hppColIdx, hppCcsToCsr, hppDiag;
  // Hpl: CCS by landmark column, ascending pose rows (== internal order)
  std::vector<int32_t> hplColPtr, hplRowIdx;
  // Hschur: reference CCS + internal CSR (== the reference's HschurTransposedCCS)
  std::vector<int32_t> sColPtr, sRowIdx, sRowPtr, sColIdx, sCcsToCsr, sDiag, hppToS;
  std::vector<int32_t> sTRefRowPtr, sTRefColIdx;   // the reference's _HschurTransposedCCS as it holds it (built before the first solve extends Hschur)
  bool hplShared = false;                // some Hpl block receives more than one edge (parallel edges)
  bool hppShared = false;
  // per active edge (internalId order): matrix id (0 Hpp, 1 Hll, 2 Hpl, -1 none), block row, block col, transposed
  std::vector<int32_t> edgeTargets;
  std::vector<EdgeSet> sets;
  int64_t schurPairs = 0;                // sum over landmarks of k(k+1)/2
  int lmBegin = 0, lmEnd = 0;            // landmark slots owned by this rank (all of them when world == 1)
};

static const int kChunk = 1024;          // edges per chunk of the pose-sorted accumulation pass

bool initializeOptimization(const HostGraph& g, int level, Structure& st, std::string& err);
bool buildStructure(const HostGraph& g, Structure& st, std::string& err, int rank = 0, int world = 1);

}  // namespace g2ocu
