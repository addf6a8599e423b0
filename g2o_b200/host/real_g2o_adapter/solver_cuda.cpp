// Adapter for the REAL g2o (B0Bftl/g2o) headers: compile this file inside the reference tree as
//   g2o/solvers/cuda/solver_cuda.cpp  ->  libg2o_solver_cuda.so   (needs Eigen3 + the g2o headers; links libg2ocu.so)
// Here it is compiled against the reference's headers with a stand-in for Eigen3 (`make -C oracle ref_adapter`, oracle/_ref/libg2o_solver_cuda.so)
// and run on the GPU next to the reference core library by tests/test_reference_core.py; a production build uses real Eigen3.
// See INTEGRATION.md for the CMake lines.
//
// Two levels of drop-in, both registered with OptimizationAlgorithmFactory the way solvers/pcg/solver_pcg.cpp:41-98 does:
//   *_cuda          OptimizationAlgorithmCuda : OptimizationAlgorithm (core/optimization_algorithm.h:46-110).  solve(i) runs one whole
//                   LM / GN / Dogleg iteration (all trials) on the device and writes the estimates back into the g2o vertices.
//   *_cuda_solver   the reference's own OptimizationAlgorithmLevenberg / GaussNewton / Dogleg driving CudaBlockSolver<P,L> : BlockSolverBase
//                   (core/solver.h:44-155, core/block_solver.h:87-95): buildStructure / buildSystem / setLambda / solve / multiplyHessian on the
//                   device, computeActiveErrors / update / push / pop stay the reference's host loops (estimates cross the bus per buildSystem).
//   *_cuda_linear   the reference's own algorithm AND its own BlockSolver<BlockSolverTraits<P,L>> (CPU buildSystem and Schur complement) over
//                   LinearSolverCuda<MatrixType> : LinearSolver<MatrixType> (core/linear_solver.h:42-105): only LinearSolverPCG::solve runs on the
//                   device (g2ocu_linear_solve); the reduced system crosses the bus on every solve.
// The first two pack SparseOptimizer::activeVertices() / activeEdges() into the flat g2ocu_graph (include/g2ocu.h) and check that the backend's
// hessianIndex of every vertex equals the one SparseOptimizer::buildIndexMapping assigned (sparse_optimizer.cpp:168-193).
// Types are recognised by dynamic_cast; the BAL types, which the reference defines inside examples/bal/bal_example.cpp (no header),
// by their class name + their BaseVertex / BaseEdge instantiation.  Anything else is rejected at init - there is no CPU fallback.
#include <cstring>
#include <typeinfo>
#include <unordered_map>

#include "g2o/core/base_binary_edge.h"
#include "g2o/core/base_vertex.h"
#include "g2o/core/block_solver.h"
#include "g2o/core/linear_solver.h"
#include "g2o/core/optimization_algorithm.h"
#include "g2o/core/optimization_algorithm_dogleg.h"
#include "g2o/core/optimization_algorithm_factory.h"
#include "g2o/core/optimization_algorithm_gauss_newton.h"
#include "g2o/core/optimization_algorithm_levenberg.h"
#include "g2o/core/robust_kernel_impl.h"
#include "g2o/core/sparse_optimizer.h"
#include "g2o/stuff/macros.h"
#include "g2o/types/sba/types_six_dof_expmap.h"
#include "g2o/types/slam2d/edge_se2.h"
#include "g2o/types/slam2d/edge_se2_pointxy.h"
#include "g2o/types/slam3d/edge_se3.h"
#include "g2ocu.h"

namespace g2o {

namespace {

typedef BaseVertex<9, Eigen::VectorXd> BalCameraBase;     // VertexCameraBAL, bal_example.cpp:65
typedef BaseVertex<3, Eigen::Vector3d> BalPointBase;      // VertexPointBAL, bal_example.cpp:102
typedef BaseEdge<2, Eigen::Vector2d> BalEdgeBase;         // EdgeObservationBAL : BaseBinaryEdge<2, Vector2d, VertexCameraBAL, VertexPointBAL>, bal_example.cpp:148

bool named(const std::type_info& t, const char* cls) { return std::strstr(t.name(), cls) != nullptr; }

int vertexCode(const OptimizableGraph::Vertex* v) {
  if (named(typeid(*v), "VertexCameraBAL") && dynamic_cast<const BalCameraBase*>(v)) return G2OCU_VERTEX_CAM_BAL;
  if (named(typeid(*v), "VertexPointBAL") && dynamic_cast<const BalPointBase*>(v)) return G2OCU_VERTEX_POINT_BAL;
  if (dynamic_cast<const VertexSE2*>(v)) return G2OCU_VERTEX_SE2;
  if (dynamic_cast<const VertexPointXY*>(v)) return G2OCU_VERTEX_POINT_XY;
  if (dynamic_cast<const VertexSE3*>(v)) return G2OCU_VERTEX_SE3;
  if (dynamic_cast<const VertexSE3Expmap*>(v)) return G2OCU_VERTEX_SE3_EXPMAP;
  if (dynamic_cast<const VertexSBAPointXYZ*>(v)) return G2OCU_VERTEX_POINT_XYZ;
  return 0;   // unsupported -> init() fails, no CPU fallback
}
int kernelCode(const RobustKernel* k) {
  if (!k) return G2OCU_KERNEL_NONE;
  if (dynamic_cast<const RobustKernelHuber*>(k)) return G2OCU_KERNEL_HUBER;
  if (dynamic_cast<const RobustKernelPseudoHuber*>(k)) return G2OCU_KERNEL_PSEUDO_HUBER;
  if (dynamic_cast<const RobustKernelCauchy*>(k)) return G2OCU_KERNEL_CAUCHY;
  if (dynamic_cast<const RobustKernelGemanMcClure*>(k)) return G2OCU_KERNEL_GEMAN_MCCLURE;
  if (dynamic_cast<const RobustKernelWelsch*>(k)) return G2OCU_KERNEL_WELSCH;
  if (dynamic_cast<const RobustKernelFair*>(k)) return G2OCU_KERNEL_FAIR;
  if (dynamic_cast<const RobustKernelTukey*>(k)) return G2OCU_KERNEL_TUKEY;
  if (dynamic_cast<const RobustKernelSaturated*>(k)) return G2OCU_KERNEL_SATURATED;
  if (dynamic_cast<const RobustKernelDCS*>(k)) return G2OCU_KERNEL_DCS;
  return -1;
}
void packIsometry(const Isometry3& T, std::vector<double>& out) {
  for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) out.push_back(T.matrix()(r, c));
  for (int r = 0; r < 3; ++r) out.push_back(T.translation()(r));
}
const int kEstimateSize[8] = {0, 3, 2, 12, 7, 3, 9, 3};   // doubles per vertex on the boundary, by G2OCU_VERTEX_* code

// The active part of a SparseOptimizer as the flat g2ocu_graph, plus the way back for the estimates.
class GraphBridge {
 public:
  std::string error;

  // activeVertices() (sorted by id, fixed ones included) and activeEdges() (sorted by internalId) of an initialised optimizer
  bool upload(SparseOptimizer* opt, g2ocu_solver* h, int expectPoseDim, int expectLandmarkDim) {
    _vertices.assign(opt->activeVertices().begin(), opt->activeVertices().end()); _codes.clear();
    std::unordered_map<const HyperGraph::Vertex*, int> index;
    std::vector<int32_t> vId, vType, eType, eV0, eV1, eKernel; std::vector<uint8_t> vFixed, vMarg; std::vector<double> eMeas, eInfo, eDelta, ePrm;
    _estimates.clear();
    for (auto* v : _vertices) {
      const int code = vertexCode(v);
      if (!code) { error = std::string("unsupported vertex type ") + typeid(*v).name(); return false; }
      // BlockSolver<BlockSolverTraits<p,l>> only fits graphs with those block sizes (block_solver.hpp:103-256 maps fixed-size blocks)
      const int want = v->marginalized() ? expectLandmarkDim : expectPoseDim;
      if (want > 0 && v->dimension() != want) { error = "vertex " + std::to_string(v->id()) + " has dimension " + std::to_string(v->dimension()) + ", this solver was registered for " + std::to_string(want); return false; }
      index[v] = (int)_codes.size(); _codes.push_back(code);
      vId.push_back(v->id()); vType.push_back(code); vFixed.push_back(v->fixed()); vMarg.push_back(v->marginalized());
      packEstimate(v, code, _estimates);
    }
    for (auto* e : opt->activeEdges()) {
      int code = 0;
      if (auto* x = dynamic_cast<EdgeSE2*>(e)) { code = G2OCU_EDGE_SE2; const Vector3 m = x->measurement().toVector(); eMeas.insert(eMeas.end(), m.data(), m.data() + 3); }
      else if (auto* x = dynamic_cast<EdgeSE2PointXY*>(e)) { code = G2OCU_EDGE_SE2_POINT_XY; eMeas.push_back(x->measurement()[0]); eMeas.push_back(x->measurement()[1]); }
      else if (auto* x = dynamic_cast<EdgeSE3*>(e)) { code = G2OCU_EDGE_SE3; packIsometry(x->measurement(), eMeas); }
      else if (auto* x = dynamic_cast<EdgeSE3Expmap*>(e)) { code = G2OCU_EDGE_SE3_EXPMAP; const Vector7 m = x->measurement().toVector(); eMeas.insert(eMeas.end(), m.data(), m.data() + 7); }
      else if (auto* x = dynamic_cast<EdgeProjectXYZ2UV*>(e)) {
        code = G2OCU_EDGE_PROJECT_XYZ2UV; eMeas.push_back(x->measurement()[0]); eMeas.push_back(x->measurement()[1]);
        const CameraParameters* cam = static_cast<const CameraParameters*>(x->parameter(0));
        ePrm.push_back(cam->focal_length); ePrm.push_back(cam->principle_point[0]); ePrm.push_back(cam->principle_point[1]);
      } else if (auto* x = dynamic_cast<EdgeSE3ProjectXYZ*>(e)) {
        code = G2OCU_EDGE_SE3_PROJECT_XYZ; eMeas.push_back(x->measurement()[0]); eMeas.push_back(x->measurement()[1]);
        ePrm.push_back(x->fx); ePrm.push_back(x->fy); ePrm.push_back(x->cx); ePrm.push_back(x->cy);
      } else if (named(typeid(*e), "EdgeObservationBAL")) {
        if (auto* x = dynamic_cast<BalEdgeBase*>(e)) { code = G2OCU_EDGE_BAL; eMeas.push_back(x->measurement()[0]); eMeas.push_back(x->measurement()[1]); }
      }
      const int kc = kernelCode(e->robustKernel());
      if (!code || kc < 0) { error = std::string("unsupported edge or robust kernel type ") + typeid(*e).name(); return false; }
      eType.push_back(code); eV0.push_back(index.at(e->vertex(0))); eV1.push_back(index.at(e->vertex(1)));
      const int D = e->dimension(); const number_t* info = e->informationData();      // column-major D x D
      eInfo.insert(eInfo.end(), info, info + D * D);
      eKernel.push_back(kc); eDelta.push_back(e->robustKernel() ? e->robustKernel()->delta() : 1.0);
    }
    g2ocu_graph g;
    g.n_vertices = (int32_t)_vertices.size(); g.v_id = vId.data(); g.v_type = vType.data(); g.v_fixed = vFixed.data(); g.v_marginalized = vMarg.data(); g.v_estimate = _estimates.data();
    g.n_edges = (int32_t)eType.size(); g.e_type = eType.data(); g.e_v0 = eV0.data(); g.e_v1 = eV1.data(); g.e_level = nullptr;   // only the active edges are sent: all on level 0
    g.e_measurement = eMeas.data(); g.e_information = eInfo.data(); g.e_kernel = eKernel.data(); g.e_kernel_delta = eDelta.data(); g.e_param = ePrm.data();
    if (g2ocu_set_graph(h, &g) != G2OCU_OK || g2ocu_initialize_optimization(h, 0) != G2OCU_OK) { error = g2ocu_last_error(h); return false; }
    // the index map the backend derived must be the one the optimizer holds (bit-exact structure)
    std::vector<int32_t> hi(_vertices.size());
    if (g2ocu_get_i32(h, "hessian_index", hi.data(), (int64_t)hi.size()) != (int64_t)hi.size()) { error = g2ocu_last_error(h); return false; }
    for (size_t i = 0; i < _vertices.size(); ++i)
      if (hi[i] != _vertices[i]->hessianIndex()) { error = "index mapping differs from SparseOptimizer's at vertex " + std::to_string(_vertices[i]->id()); return false; }
    return true;
  }

  // host vertices -> packed array (after the reference's own update / pop on the host)
  const std::vector<double>& gather() {
    _estimates.clear();
    for (size_t i = 0; i < _vertices.size(); ++i) packEstimate(_vertices[i], _codes[i], _estimates);
    return _estimates;
  }
  // device estimates -> host vertices
  bool writeBack(g2ocu_solver* h) {
    if (g2ocu_get_estimates(h, _estimates.data()) != G2OCU_OK) { error = g2ocu_last_error(h); return false; }
    const double* est = _estimates.data();
    for (size_t i = 0; i < _vertices.size(); ++i) {
      OptimizableGraph::Vertex* v = _vertices[i];
      switch (_codes[i]) {
        case G2OCU_VERTEX_SE2: static_cast<VertexSE2*>(v)->setEstimate(SE2(est[0], est[1], est[2])); break;
        case G2OCU_VERTEX_POINT_XY: static_cast<VertexPointXY*>(v)->setEstimate(Vector2(est[0], est[1])); break;
        case G2OCU_VERTEX_SE3: {
          Isometry3 T = Isometry3::Identity();
          for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) T.matrix()(r, c) = est[r + 3 * c];
          for (int r = 0; r < 3; ++r) T.translation()(r) = est[9 + r];
          static_cast<VertexSE3*>(v)->setEstimate(T); break; }
        case G2OCU_VERTEX_SE3_EXPMAP: { SE3Quat T; Vector7 t; for (int k = 0; k < 7; ++k) t[k] = est[k]; T.fromVector(t); static_cast<VertexSE3Expmap*>(v)->setEstimate(T); break; }
        case G2OCU_VERTEX_POINT_XYZ: static_cast<VertexSBAPointXYZ*>(v)->setEstimate(Vector3(est[0], est[1], est[2])); break;
        case G2OCU_VERTEX_CAM_BAL: { Eigen::VectorXd c(9); for (int k = 0; k < 9; ++k) c[k] = est[k]; dynamic_cast<BalCameraBase*>(v)->setEstimate(c); break; }
        case G2OCU_VERTEX_POINT_BAL: dynamic_cast<BalPointBase*>(v)->setEstimate(Eigen::Vector3d(est[0], est[1], est[2])); break;
      }
      est += kEstimateSize[_codes[i]];
    }
    return true;
  }
  const std::vector<OptimizableGraph::Vertex*>& vertices() const { return _vertices; }

 private:
  static void packEstimate(const OptimizableGraph::Vertex* v, int code, std::vector<double>& out) {
    switch (code) {
      case G2OCU_VERTEX_SE2: { const SE2& e = static_cast<const VertexSE2*>(v)->estimate(); out.push_back(e[0]); out.push_back(e[1]); out.push_back(e[2]); break; }
      case G2OCU_VERTEX_POINT_XY: { const Vector2& e = static_cast<const VertexPointXY*>(v)->estimate(); out.push_back(e[0]); out.push_back(e[1]); break; }
      case G2OCU_VERTEX_SE3: packIsometry(static_cast<const VertexSE3*>(v)->estimate(), out); break;
      case G2OCU_VERTEX_SE3_EXPMAP: { const Vector7 t = static_cast<const VertexSE3Expmap*>(v)->estimate().toVector(); out.insert(out.end(), t.data(), t.data() + 7); break; }
      case G2OCU_VERTEX_POINT_XYZ: { const Vector3& e = static_cast<const VertexSBAPointXYZ*>(v)->estimate(); out.push_back(e[0]); out.push_back(e[1]); out.push_back(e[2]); break; }
      case G2OCU_VERTEX_CAM_BAL: { const Eigen::VectorXd& e = dynamic_cast<const BalCameraBase*>(v)->estimate(); for (int k = 0; k < 9; ++k) out.push_back(e[k]); break; }
      case G2OCU_VERTEX_POINT_BAL: { const Eigen::Vector3d& e = dynamic_cast<const BalPointBase*>(v)->estimate(); for (int k = 0; k < 3; ++k) out.push_back(e[k]); break; }
    }
  }
  std::vector<OptimizableGraph::Vertex*> _vertices; std::vector<int> _codes; std::vector<double> _estimates;
};

void fillBatchStatistics(const g2ocu_iteration_stats& st) {
  if (G2OBatchStatistics* gs = G2OBatchStatistics::globalStats()) {
    gs->timeResiduals = st.time_residuals; gs->timeQuadraticForm = st.time_quadratic_form; gs->timeSchurComplement = st.time_schur_complement;
    gs->timeLinearSolver = st.time_linear_solver; gs->timeLinearSolution = st.time_linear_solution; gs->timeUpdate = st.time_update;
    gs->levenbergIterations = st.levenberg_iterations; gs->iterationsLinearSolver = st.iterations_linear_solver;
    gs->hessianPoseDimension = st.hessian_pose_dimension; gs->hessianLandmarkDimension = st.hessian_landmark_dimension;
    gs->hessianDimension = st.hessian_pose_dimension + st.hessian_landmark_dimension;
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------------------------------
// Algorithm level: one virtual call per outer iteration, everything else on the device.
// SparseOptimizer::computeMarginals (sparse_optimizer.cpp:594-596): blocks of the inverse of Hpp from the device.  `spinv` gets the layout
// MarginalCovarianceCholesky::computeCovariance gives it (marginal_covariance_cholesky.cpp:153-160: a square block matrix over the block
// rows of Hpp, the requested blocks allocated); false where the reference's solvePattern answers false (no factorisation) and where the
// backend has no answer (points that are not marginalized, pose systems beyond the dense factorisation's limit).
static bool marginalsFromDevice(g2ocu_solver* h, SparseBlockMatrix<MatrixX>& spinv, const std::vector<std::pair<int, int>>& blockIndices) {
  if (!h) return false;
  const int64_t nb = g2ocu_get_i32(h, "pose_block_indices", nullptr, 0);      // cumulative block ends of Hpp (every vertex when no point is marginalized)
  if (nb <= 0) return false;
  std::vector<int32_t> ends((size_t)nb);
  g2ocu_get_i32(h, "pose_block_indices", ends.data(), nb);
  auto dimOf = [&](int b) { return ends[b] - (b ? ends[b - 1] : 0); };
  std::vector<int32_t> rows, cols; size_t total = 0;
  for (const auto& rc : blockIndices) {
    if (rc.first < 0 || rc.first >= nb || rc.second < 0 || rc.second >= nb) return false;
    rows.push_back(rc.first); cols.push_back(rc.second); total += (size_t)dimOf(rc.first) * dimOf(rc.second);
  }
  std::vector<double> out(total ? total : 1);
  int32_t computed = 0;
  if (g2ocu_compute_marginals(h, (int32_t)blockIndices.size(), rows.data(), cols.data(), out.data(), &computed) != G2OCU_OK) { std::cerr << "solver_cuda: " << g2ocu_last_error(h) << std::endl; return false; }
  if (!computed) return false;
  std::vector<int> blockEnds(ends.begin(), ends.end());
  spinv = SparseBlockMatrix<MatrixX>(blockEnds.data(), blockEnds.data(), (int)nb, (int)nb, true);
  size_t off = 0;
  for (size_t i = 0; i < blockIndices.size(); ++i) {
    MatrixX* b = spinv.block(rows[i], cols[i], true);
    const int nr = dimOf(rows[i]), nc = dimOf(cols[i]);
    for (int c = 0; c < nc; ++c) for (int r = 0; r < nr; ++r) (*b)(r, c) = out[off + (size_t)c * nr + r];
    off += (size_t)nr * nc;
  }
  return true;
}

class OptimizationAlgorithmCuda : public OptimizationAlgorithm {
 public:
  OptimizationAlgorithmCuda(int algorithm, int poseDim, int landmarkDim, int linearSolver = G2OCU_LINEAR_PCG) : _algorithm(algorithm), _poseDim(poseDim), _landmarkDim(landmarkDim) {
    g2ocu_config cfg; g2ocu_default_config(&cfg); cfg.linear_solver = linearSolver;   // LinearSolverPCG or LinearSolverDense semantics (solvers/pcg, solvers/dense)
    g2ocu_create(&cfg, &_h);
    _userLambdaInit = _properties.makeProperty<Property<number_t>>("initialLambda", 0.);
    _maxTrialsAfterFailure = _properties.makeProperty<Property<int>>("maxTrialsAfterFailure", algorithm == G2OCU_ALGORITHM_DOGLEG ? 100 : 10);
    if (algorithm == G2OCU_ALGORITHM_DOGLEG) {   // OptimizationAlgorithmDogleg's own defaults (optimization_algorithm_dogleg.cpp:44-47)
      _userLambdaInit->setValue(1e-7);
      _userDeltaInit = _properties.makeProperty<Property<number_t>>("initialDelta", (number_t)1e4);
      _lambdaFactor = _properties.makeProperty<Property<number_t>>("lambdaFactor", 10.);
    }
  }
  ~OptimizationAlgorithmCuda() { g2ocu_destroy(_h); }

  bool init(bool online = false) override {
    if (!_bridge.upload(_optimizer, _h, _poseDim, _landmarkDim)) { std::cerr << "solver_cuda: " << _bridge.error << std::endl; return false; }   // unsupported types are rejected here
    if (_algorithm == G2OCU_ALGORITHM_DOGLEG) {
      g2ocu_set_property(_h, "doglegInitialDelta", _userDeltaInit->value()); g2ocu_set_property(_h, "doglegLambdaFactor", _lambdaFactor->value());
      g2ocu_set_property(_h, "doglegInitialLambda", _userLambdaInit->value()); g2ocu_set_property(_h, "doglegMaxTrialsAfterFailure", _maxTrialsAfterFailure->value());
    } else {
      g2ocu_set_property(_h, "initialLambda", _userLambdaInit->value());
      g2ocu_set_property(_h, "maxTrialsAfterFailure", _maxTrialsAfterFailure->value());
    }
    return ok(g2ocu_init(_h, online));
  }
  SolverResult solve(int iteration, bool /*online*/ = false) override {
    g2ocu_iteration_stats st;
    g2ocu_set_force_stop_flag(_h, reinterpret_cast<const unsigned char*>(_optimizer->forceStopFlag()));            // SparseOptimizer::terminate(), sparse_optimizer.h:186-190 / levenberg.cpp:145
    if (!ok(g2ocu_solver_iteration(_h, _algorithm, iteration, &st))) return Fail;
    _lambda = st.lambda; _levenbergIterations = st.levenberg_iterations;
    if (_algorithm == G2OCU_ALGORITHM_DOGLEG) g2ocu_get_f64(_h, "dogleg", _dogleg, 5);   // trust region, step type, tries, damping, PD flag
    if (!_bridge.writeBack(_h)) { std::cerr << "solver_cuda: " << _bridge.error << std::endl; return Fail; }
    fillBatchStatistics(st);
    return st.result == G2OCU_RESULT_OK ? OK : (st.result == G2OCU_RESULT_TERMINATE ? Terminate : Fail);
  }
  bool computeMarginals(SparseBlockMatrix<MatrixX>& spinv, const std::vector<std::pair<int, int>>& blockIndices) override { return marginalsFromDevice(_h, spinv, blockIndices); }
  bool updateStructure(const std::vector<HyperGraph::Vertex*>&, const HyperGraph::EdgeSet&) override { return false; }    // online mode: out of scope
  void printVerbose(std::ostream& os) const override {
    if (_algorithm == G2OCU_ALGORITHM_DOGLEG) {   // optimization_algorithm_dogleg.cpp:199-217
      static const char* const step[] = {"Undefined", "Descent", "GN", "Dogleg"};
      os << "\t Delta= " << _dogleg[0] << "\t step= " << step[(int)_dogleg[1] & 3] << "\t tries= " << (int)_dogleg[2];
      if (_dogleg[4] == 0.0) os << "\t lambda= " << _dogleg[3];
      return;
    }
    os << "\t lambda= " << FIXED(_lambda) << "\t levenbergIter= " << _levenbergIterations;
  }

 private:
  bool ok(int rc) { if (rc != G2OCU_OK) std::cerr << "solver_cuda: " << g2ocu_last_error(_h) << std::endl; return rc == G2OCU_OK; }

  g2ocu_solver* _h = nullptr; int _algorithm, _poseDim, _landmarkDim;
  GraphBridge _bridge;
  Property<number_t>* _userDeltaInit = nullptr; Property<number_t>* _lambdaFactor = nullptr; double _dogleg[5] = {1e4, 0, 0, 1e-7, 1};
  Property<number_t>* _userLambdaInit; Property<int>* _maxTrialsAfterFailure;
  number_t _lambda = -1; int _levenbergIterations = 0;
};

// ------------------------------------------------------------------------------------------------------------------------------------
// Solver level: BlockSolver<BlockSolverTraits<P,L>> replacement (P, L = Eigen::Dynamic: BlockSolverX) under the reference's own algorithms.
class CudaBlockSolverImpl : public BlockSolverBase {
 public:
  CudaBlockSolverImpl(int poseDim, int landmarkDim, int linearSolver) : _poseDim(poseDim), _landmarkDim(landmarkDim) {
    g2ocu_config cfg; g2ocu_default_config(&cfg); cfg.linear_solver = linearSolver;
    g2ocu_create(&cfg, &_h);
  }
  ~CudaBlockSolverImpl() override { g2ocu_destroy(_h); }

  bool init(SparseOptimizer* optimizer, bool online = false) override {   // BlockSolver::init, block_solver.hpp:577-590
    _optimizer = optimizer; _structureReady = false;
    return true;
  }
  bool buildStructure(bool /*zeroBlocks*/ = false) override {             // block_solver.hpp:103-256
    if (!_bridge.upload(_optimizer, _h, _poseDim, _landmarkDim)) { std::cerr << "solver_cuda: " << _bridge.error << std::endl; return false; }
    if (!ok(g2ocu_init(_h, 0)) || !ok(g2ocu_build_structure(_h))) return false;
    resizeVector((size_t)g2ocu_vector_size(_h));
    // the reference maps every vertex's Hessian block into Hpp / Hll (block_solver.hpp:150-170); OptimizationAlgorithmLevenberg::computeLambdaInit
    // reads their diagonals through the vertices (optimization_algorithm_levenberg.cpp:152-175).  Host copies of the diagonal blocks serve that.
    size_t total = 0;
    for (auto* v : _optimizer->indexMapping()) total += (size_t)v->dimension() * v->dimension();
    _diagBlocks.assign(total, 0.0);
    size_t o = 0;
    for (auto* v : _optimizer->indexMapping()) { v->mapHessianMemory(_diagBlocks.data() + o); o += (size_t)v->dimension() * v->dimension(); }
    _structureReady = true; _diagFresh = false;
    return true;
  }
  bool updateStructure(const std::vector<HyperGraph::Vertex*>&, const HyperGraph::EdgeSet&) override { return false; }   // online mode: out of scope
  bool buildSystem() override {                                            // block_solver.hpp:463-521
    if (!_structureReady) return false;
    if (!ok(g2ocu_set_estimates(_h, _bridge.gather().data()))) return false;   // the host loops of SparseOptimizer moved the estimates
    if (!ok(g2ocu_build_system(_h))) return false;
    if (g2ocu_get_f64(_h, "b", _b, (int64_t)_xSize) != (int64_t)_xSize) return ok(G2OCU_E_INVALID);
    if (!_diagFresh) {   // diagonal blocks for computeLambdaInit (iteration 0 only in the reference's LM)
      if (g2ocu_get_f64(_h, "diagonal_blocks", _diagBlocks.data(), (int64_t)_diagBlocks.size()) != (int64_t)_diagBlocks.size()) return ok(G2OCU_E_INVALID);
      _diagFresh = true;
    }
    return true;
  }
  bool setLambda(number_t lambda, bool backup = false) override { return ok(g2ocu_set_lambda(_h, lambda, backup)); }   // block_solver.hpp:525-553
  void restoreDiagonal() override { ok(g2ocu_restore_diagonal(_h)); }      // block_solver.hpp:555-565
  bool solve() override {                                                  // block_solver.hpp:315-447
    int32_t solved = 0;
    g2ocu_reset_counters(_h);
    if (!ok(g2ocu_solve(_h, &solved))) return false;
    if (g2ocu_get_f64(_h, "x", _x, (int64_t)_xSize) != (int64_t)_xSize) return ok(G2OCU_E_INVALID);
    if (G2OBatchStatistics* gs = G2OBatchStatistics::globalStats()) {
      double sec = 0; int64_t n = 0, c = 0;
      g2ocu_phase_time(_h, "schur", &sec, &n, &c); gs->timeSchurComplement = sec;
      g2ocu_phase_time(_h, "linear_solver", &sec, &n, &c); gs->timeLinearSolver = sec;
      int32_t dims[4] = {0, 0, 0, 0}; g2ocu_get_i32(_h, "internal_dims", dims, 4);
      // block_solver.hpp:323,411-413: without Schur everything sits in Hpp
      gs->hessianPoseDimension = _doSchur ? dims[2] : dims[2] + dims[3]; gs->hessianLandmarkDimension = _doSchur ? dims[3] : 0; gs->hessianDimension = dims[2] + dims[3];
      int32_t its = 0; g2ocu_get_i32(_h, "linear_solver_iterations", &its, 1); gs->iterationsLinearSolver = its;   // what LinearSolverPCG::solve records (linear_solver_pcg.hpp:150-153)
    }
    return solved != 0;
  }
  bool computeMarginals(SparseBlockMatrix<MatrixX>& spinv, const std::vector<std::pair<int, int>>& blockIndices) override { return marginalsFromDevice(_h, spinv, blockIndices); }   // block_solver.hpp:451-459
  bool supportsSchur() override { return true; }
  bool schur() override { return _doSchur; }
  void setSchur(bool s) override { _doSchur = s; }                         // the backend derives it from the marginalized flags exactly like with_hessian.cpp:48-66
  void setWriteDebug(bool b) override { _writeDebug = b; }
  bool writeDebug() const override { return _writeDebug; }
  bool saveHessian(const std::string&) const override { return false; }
  void multiplyHessian(number_t* dest, const number_t* src) const override { g2ocu_multiply_hessian(_h, dest, src); }   // block_solver.h:146

 private:
  bool ok(int rc) const { if (rc != G2OCU_OK) std::cerr << "solver_cuda: " << g2ocu_last_error(_h) << std::endl; return rc == G2OCU_OK; }
  g2ocu_solver* _h = nullptr; int _poseDim, _landmarkDim;
  GraphBridge _bridge; std::vector<double> _diagBlocks;
  bool _doSchur = true, _writeDebug = false, _structureReady = false, _diagFresh = false;
};
template <int P, int L> class CudaBlockSolver : public CudaBlockSolverImpl {
 public:
  static const int PoseDim = P, LandmarkDim = L;
  explicit CudaBlockSolver(int linearSolver = G2OCU_LINEAR_PCG) : CudaBlockSolverImpl(P, L, linearSolver) {}
};

// ------------------------------------------------------------------------------------------------------------------------------------
// LinearSolver level: what LinearSolverPCG<MatrixType> is to the reference's BlockSolver (solvers/pcg/linear_solver_pcg.h), on the device.
template <typename MatrixType> class LinearSolverCuda : public LinearSolver<MatrixType> {
 public:
  LinearSolverCuda() { g2ocu_config cfg; g2ocu_default_config(&cfg); g2ocu_linear_create(&cfg, &_h); }
  ~LinearSolverCuda() override { g2ocu_linear_destroy(_h); }
  bool init() override { return g2ocu_linear_init(_h) == G2OCU_OK; }      // LinearSolverPCG::init: the carried residual starts over
  bool solve(const SparseBlockMatrix<MatrixType>& A, number_t* x, number_t* b) override {
    // the upper blocks by column, exactly the walk of LinearSolverPCG::solve (linear_solver_pcg.hpp:86-110)
    const int nCols = (int)A.blockCols().size();
    _colptr.assign(1, 0); _rowidx.clear(); _values.clear();
    int dim = -1;
    for (int i = 0; i < nCols; ++i) {
      for (const auto& kv : A.blockCols()[i]) {
        if (kv.first > i) break;
        const MatrixType& m = *kv.second;
        if (dim < 0) dim = (int)m.rows();
        if ((int)m.rows() != dim || (int)m.cols() != dim) { std::cerr << "solver_cuda: LinearSolverCuda needs blocks of one size" << std::endl; return false; }
        _rowidx.push_back(kv.first);
        for (int c = 0; c < dim; ++c) for (int r = 0; r < dim; ++r) _values.push_back(m(r, c));
      }
      _colptr.push_back((int32_t)_rowidx.size());
    }
    int32_t solved = 0, iterations = 0;
    if (g2ocu_linear_solve(_h, nCols, dim, _colptr.data(), _rowidx.data(), _values.data(), b, x, &solved, &iterations) != G2OCU_OK) {
      std::cerr << "solver_cuda: " << g2ocu_linear_last_error(_h) << std::endl;
      return false;
    }
    if (G2OBatchStatistics* gs = G2OBatchStatistics::globalStats()) gs->iterationsLinearSolver = iterations;   // linear_solver_pcg.hpp:150-153
    return solved != 0;
  }
  void setTolerance(number_t tolerance) { g2ocu_linear_set_property(_h, "pcgTolerance", tolerance); }
  void setMaxIterations(int maxIter) { g2ocu_linear_set_property(_h, "pcgMaxIterations", maxIter); }
  void setAbsoluteTolerance(bool absoluteTolerance) { g2ocu_linear_set_property(_h, "pcgAbsoluteTolerance", absoluteTolerance); }

 private:
  g2ocu_linear_solver* _h = nullptr;
  std::vector<int32_t> _colptr, _rowidx; std::vector<double> _values;
};
template <int P, int L> std::unique_ptr<Solver> allocateLinearLevelSolver() {
  typedef BlockSolverPL<P, L> BS;
  return std::unique_ptr<Solver>(new BS(std::unique_ptr<typename BS::LinearSolverType>(new LinearSolverCuda<typename BS::PoseMatrixType>())));
}

// ------------------------------------------------------------------------------------------------------------------------------------
class CudaSolverCreator : public AbstractOptimizationAlgorithmCreator {
 public:
  explicit CudaSolverCreator(const OptimizationAlgorithmProperty& p) : AbstractOptimizationAlgorithmCreator(p) {}
  OptimizationAlgorithm* construct() override {
    const std::string& n = property().name;
    const int algorithm = n.substr(0, 2) == "lm" ? G2OCU_ALGORITHM_LM : n.substr(0, 2) == "dl" ? G2OCU_ALGORITHM_DOGLEG : G2OCU_ALGORITHM_GN;
    const int linear = n.find("_dense") != std::string::npos ? G2OCU_LINEAR_DENSE : G2OCU_LINEAR_PCG;
    const int P = property().poseDim, L = property().landmarkDim;
    if (n.size() > 12 && n.compare(n.size() - 12, 12, "_cuda_linear") == 0) {   // g2o's own algorithm and BlockSolver, the CUDA linear solver
      std::unique_ptr<Solver> bs = P == 3 ? allocateLinearLevelSolver<3, 2>() : P == 6 ? allocateLinearLevelSolver<6, 3>() : allocateLinearLevelSolver<9, 3>();
      if (algorithm == G2OCU_ALGORITHM_LM) return new OptimizationAlgorithmLevenberg(std::move(bs));
      return new OptimizationAlgorithmGaussNewton(std::move(bs));
    }
    if (n.size() < 12 || n.compare(n.size() - 12, 12, "_cuda_solver") != 0) return new OptimizationAlgorithmCuda(algorithm, P, L, linear);
    // the reference's own algorithm classes over the CUDA block solver
    std::unique_ptr<BlockSolverBase> bs;
    if (P == 3 && L == 2) bs.reset(new CudaBlockSolver<3, 2>(linear));
    else if (P == 6 && L == 3) bs.reset(new CudaBlockSolver<6, 3>(linear));
    else if (P == 9 && L == 3) bs.reset(new CudaBlockSolver<9, 3>(linear));
    else bs.reset(new CudaBlockSolver<Eigen::Dynamic, Eigen::Dynamic>(linear));
    if (algorithm == G2OCU_ALGORITHM_LM) return new OptimizationAlgorithmLevenberg(std::move(bs));
    if (algorithm == G2OCU_ALGORITHM_DOGLEG) return new OptimizationAlgorithmDogleg(std::move(bs));
    return new OptimizationAlgorithmGaussNewton(std::move(bs));
  }
};

#define G2OCU_REGISTER(name, desc, marg, P, L) \
  G2O_REGISTER_OPTIMIZATION_ALGORITHM(name, new CudaSolverCreator(OptimizationAlgorithmProperty(#name, desc, "CUDA", marg, P, L)))

G2O_REGISTER_OPTIMIZATION_LIBRARY(cuda);
G2OCU_REGISTER(gn_var_cuda, "Gauss-Newton: PCG on the GPU (variable blocksize)", false, Eigen::Dynamic, Eigen::Dynamic);
G2OCU_REGISTER(lm_var_cuda, "Levenberg: PCG on the GPU (variable blocksize)", false, Eigen::Dynamic, Eigen::Dynamic);
G2OCU_REGISTER(dl_var_cuda, "Dogleg: block-Jacobi PCG on the GPU (variable blocksize)", false, Eigen::Dynamic, Eigen::Dynamic);
G2OCU_REGISTER(gn_fix3_2_cuda, "Gauss-Newton: Schur + PCG on the GPU", true, 3, 2);
G2OCU_REGISTER(lm_fix3_2_cuda, "Levenberg: Schur + PCG on the GPU", true, 3, 2);
G2OCU_REGISTER(gn_fix6_3_cuda, "Gauss-Newton: Schur + PCG on the GPU", true, 6, 3);
G2OCU_REGISTER(lm_fix6_3_cuda, "Levenberg: Schur + PCG on the GPU", true, 6, 3);
G2OCU_REGISTER(gn_fix9_3_cuda, "Gauss-Newton: Schur + PCG on the GPU (BAL cameras)", true, 9, 3);
G2OCU_REGISTER(lm_fix9_3_cuda, "Levenberg: Schur + PCG on the GPU (BAL cameras)", true, 9, 3);
G2OCU_REGISTER(gn_dense_cuda, "Gauss-Newton: dense FP64 Cholesky of the (reduced) system on the GPU", false, Eigen::Dynamic, Eigen::Dynamic);
G2OCU_REGISTER(lm_dense_cuda, "Levenberg: dense FP64 Cholesky of the (reduced) system on the GPU", false, Eigen::Dynamic, Eigen::Dynamic);
G2OCU_REGISTER(gn_dense3_2_cuda, "Gauss-Newton: dense FP64 Cholesky of the (reduced) system on the GPU", true, 3, 2);
G2OCU_REGISTER(lm_dense3_2_cuda, "Levenberg: dense FP64 Cholesky of the (reduced) system on the GPU", true, 3, 2);
G2OCU_REGISTER(gn_dense6_3_cuda, "Gauss-Newton: dense FP64 Cholesky of the (reduced) system on the GPU", true, 6, 3);
G2OCU_REGISTER(lm_dense6_3_cuda, "Levenberg: dense FP64 Cholesky of the (reduced) system on the GPU", true, 6, 3);
G2OCU_REGISTER(gn_dense9_3_cuda, "Gauss-Newton: dense FP64 Cholesky of the (reduced) system on the GPU", true, 9, 3);
G2OCU_REGISTER(lm_dense9_3_cuda, "Levenberg: dense FP64 Cholesky of the (reduced) system on the GPU", true, 9, 3);
// the reference's own OptimizationAlgorithmLevenberg / GaussNewton / Dogleg over CudaBlockSolver<P,L>
G2OCU_REGISTER(gn_var_cuda_solver, "Gauss-Newton (g2o's own) over the CUDA block solver: PCG on the GPU (variable blocksize)", false, Eigen::Dynamic, Eigen::Dynamic);
G2OCU_REGISTER(lm_var_cuda_solver, "Levenberg (g2o's own) over the CUDA block solver: PCG on the GPU (variable blocksize)", false, Eigen::Dynamic, Eigen::Dynamic);
G2OCU_REGISTER(dl_var_cuda_solver, "Dogleg (g2o's own) over the CUDA block solver: PCG on the GPU (variable blocksize)", false, Eigen::Dynamic, Eigen::Dynamic);
G2OCU_REGISTER(gn_fix3_2_cuda_solver, "Gauss-Newton (g2o's own) over the CUDA block solver: Schur + PCG on the GPU", true, 3, 2);
G2OCU_REGISTER(lm_fix3_2_cuda_solver, "Levenberg (g2o's own) over the CUDA block solver: Schur + PCG on the GPU", true, 3, 2);
G2OCU_REGISTER(gn_fix6_3_cuda_solver, "Gauss-Newton (g2o's own) over the CUDA block solver: Schur + PCG on the GPU", true, 6, 3);
G2OCU_REGISTER(lm_fix6_3_cuda_solver, "Levenberg (g2o's own) over the CUDA block solver: Schur + PCG on the GPU", true, 6, 3);
G2OCU_REGISTER(gn_fix9_3_cuda_solver, "Gauss-Newton (g2o's own) over the CUDA block solver: Schur + PCG on the GPU (BAL cameras)", true, 9, 3);
G2OCU_REGISTER(lm_fix9_3_cuda_solver, "Levenberg (g2o's own) over the CUDA block solver: Schur + PCG on the GPU (BAL cameras)", true, 9, 3);
// g2o's own algorithm and BlockSolver<BlockSolverTraits<P,L>> over LinearSolverCuda (block-Jacobi PCG on the device)
G2OCU_REGISTER(gn_fix3_2_cuda_linear, "Gauss-Newton, g2o's BlockSolver_3_2 on the CPU, block-Jacobi PCG on the GPU", true, 3, 2);
G2OCU_REGISTER(lm_fix3_2_cuda_linear, "Levenberg, g2o's BlockSolver_3_2 on the CPU, block-Jacobi PCG on the GPU", true, 3, 2);
G2OCU_REGISTER(gn_fix6_3_cuda_linear, "Gauss-Newton, g2o's BlockSolver_6_3 on the CPU, block-Jacobi PCG on the GPU", true, 6, 3);
G2OCU_REGISTER(lm_fix6_3_cuda_linear, "Levenberg, g2o's BlockSolver_6_3 on the CPU, block-Jacobi PCG on the GPU", true, 6, 3);
G2OCU_REGISTER(lm_fix9_3_cuda_linear, "Levenberg, g2o's BlockSolver<9,3> on the CPU, block-Jacobi PCG on the GPU (BAL cameras)", true, 9, 3);

}  // namespace g2o
