// Adapter for the REAL g2o (B0Bftl/g2o) headers: compile this file inside the reference tree as
//   g2o/solvers/cuda/solver_cuda.cpp  ->  libg2o_solver_cuda.so   (needs Eigen3 + the g2o headers; links libg2ocu.so)
// Here it is compiled against the reference's headers with a stand-in for Eigen3 (`make -C oracle ref_adapter`, oracle/_ref/libg2o_solver_cuda.so,
// tests/test_reference_core.py); a production build uses real Eigen3.  The Eigen-free mirror in ../g2o_mirror.* has the same
// structure and is what the tests exercise.  See INTEGRATION.md for the CMake lines.
//
// What it does: packs SparseOptimizer::activeEdges()/indexMapping() into the flat g2ocu_graph, forwards every virtual of
// OptimizationAlgorithm (core/optimization_algorithm.h:46-110) to the C ABI, writes the estimates back into the g2o vertices
// after each solve(i) (SparseOptimizer has no end-of-optimize hook and calls computeActiveErrors itself when verbose / stats
// are on, sparse_optimizer.cpp:411-423), and registers the solver names the way solvers/pcg/solver_pcg.cpp:41-98 does.
#include <typeinfo>

#include "g2o/core/optimization_algorithm.h"
#include "g2o/core/optimization_algorithm_factory.h"
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
int vertexCode(const OptimizableGraph::Vertex* v) {
  if (dynamic_cast<const VertexSE2*>(v)) return G2OCU_VERTEX_SE2;
  if (dynamic_cast<const VertexPointXY*>(v)) return G2OCU_VERTEX_POINT_XY;
  if (dynamic_cast<const VertexSE3*>(v)) return G2OCU_VERTEX_SE3;
  if (dynamic_cast<const VertexSE3Expmap*>(v)) return G2OCU_VERTEX_SE3_EXPMAP;
  if (dynamic_cast<const VertexSBAPointXYZ*>(v)) return G2OCU_VERTEX_POINT_XYZ;
  // VertexCameraBAL / VertexPointBAL live in examples/bal/bal_example.cpp: move them into a header to use them here
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
}  // namespace

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
    if (!packAndUpload()) return false;                       // unsupported types are rejected here
    if (_algorithm == G2OCU_ALGORITHM_DOGLEG) {
      g2ocu_set_property(_h, "doglegInitialDelta", _userDeltaInit->value()); g2ocu_set_property(_h, "doglegLambdaFactor", _lambdaFactor->value());
      g2ocu_set_property(_h, "doglegInitialLambda", _userLambdaInit->value()); g2ocu_set_property(_h, "doglegMaxTrialsAfterFailure", _maxTrialsAfterFailure->value());
    } else {
      g2ocu_set_property(_h, "initialLambda", _userLambdaInit->value());
      g2ocu_set_property(_h, "maxTrialsAfterFailure", _maxTrialsAfterFailure->value());
    }
    return ok(g2ocu_initialize_optimization(_h, 0)) && ok(g2ocu_init(_h, online));
  }
  SolverResult solve(int iteration, bool /*online*/ = false) override {
    g2ocu_iteration_stats st;
    if (!ok(g2ocu_solver_iteration(_h, _algorithm, iteration, &st))) return Fail;
    _lambda = st.lambda; _levenbergIterations = st.levenberg_iterations;
    if (_algorithm == G2OCU_ALGORITHM_DOGLEG) g2ocu_get_f64(_h, "dogleg", _dogleg, 5);   // trust region, step type, tries, damping, PD flag
    writeBack();
    if (G2OBatchStatistics* gs = G2OBatchStatistics::globalStats()) {
      gs->timeResiduals = st.time_residuals; gs->timeQuadraticForm = st.time_quadratic_form; gs->timeSchurComplement = st.time_schur_complement;
      gs->timeLinearSolver = st.time_linear_solver; gs->timeLinearSolution = st.time_linear_solution; gs->timeUpdate = st.time_update;
      gs->levenbergIterations = st.levenberg_iterations; gs->iterationsLinearSolver = st.iterations_linear_solver;
      gs->hessianPoseDimension = st.hessian_pose_dimension; gs->hessianLandmarkDimension = st.hessian_landmark_dimension;
      gs->hessianDimension = st.hessian_pose_dimension + st.hessian_landmark_dimension;
    }
    return st.result == G2OCU_RESULT_OK ? OK : (st.result == G2OCU_RESULT_TERMINATE ? Terminate : Fail);
  }
  bool computeMarginals(SparseBlockMatrix<MatrixX>&, const std::vector<std::pair<int, int>>&) override { return false; }   // out of scope
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
  bool ok(int rc) { if (rc != G2OCU_OK) std::cerr << __PRETTY_FUNCTION__ << ": " << g2ocu_last_error(_h) << std::endl; return rc == G2OCU_OK; }

  bool packAndUpload() {
    // every vertex of the graph (ids are looked up through a dense index), every edge in internalId order
    _vertices.clear(); std::unordered_map<const HyperGraph::Vertex*, int> index;
    for (auto& kv : _optimizer->vertices()) { index[kv.second] = (int)_vertices.size(); _vertices.push_back(static_cast<OptimizableGraph::Vertex*>(kv.second)); }
    std::vector<OptimizableGraph::Edge*> edges;
    for (auto* e : _optimizer->edges()) edges.push_back(static_cast<OptimizableGraph::Edge*>(e));
    std::sort(edges.begin(), edges.end(), OptimizableGraph::EdgeIDCompare());
    std::vector<int32_t> vId, vType, eType, eV0, eV1, eLevel, eKernel; std::vector<uint8_t> vFixed, vMarg; std::vector<double> vEst, eMeas, eInfo, eDelta, ePrm;
    for (auto* v : _vertices) {
      const int code = vertexCode(v);
      if (!code) { std::cerr << "solver_cuda: unsupported vertex type " << typeid(*v).name() << std::endl; return false; }
      vId.push_back(v->id()); vType.push_back(code); vFixed.push_back(v->fixed()); vMarg.push_back(v->marginalized());
      if (code == G2OCU_VERTEX_SE3) packIsometry(static_cast<VertexSE3*>(v)->estimate(), vEst);
      else if (code == G2OCU_VERTEX_SE3_EXPMAP) { const Vector7 t = static_cast<VertexSE3Expmap*>(v)->estimate().toVector(); vEst.insert(vEst.end(), t.data(), t.data() + 7); }
      else { std::vector<double> tmp(v->estimateDimension()); v->getEstimateData(tmp.data()); vEst.insert(vEst.end(), tmp.begin(), tmp.end()); }
    }
    for (auto* e : edges) {
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
      }
      const int kc = kernelCode(e->robustKernel());
      if (!code || kc < 0) { std::cerr << "solver_cuda: unsupported edge or robust kernel type " << typeid(*e).name() << std::endl; return false; }
      eType.push_back(code); eV0.push_back(index[e->vertex(0)]); eV1.push_back(index[e->vertex(1)]); eLevel.push_back(e->level());
      const int D = e->dimension(); const number_t* info = e->informationData();      // column-major D x D
      eInfo.insert(eInfo.end(), info, info + D * D);
      eKernel.push_back(kc); eDelta.push_back(e->robustKernel() ? e->robustKernel()->delta() : 1.0);
    }
    g2ocu_graph g;
    g.n_vertices = (int32_t)_vertices.size(); g.v_id = vId.data(); g.v_type = vType.data(); g.v_fixed = vFixed.data(); g.v_marginalized = vMarg.data(); g.v_estimate = vEst.data();
    g.n_edges = (int32_t)edges.size(); g.e_type = eType.data(); g.e_v0 = eV0.data(); g.e_v1 = eV1.data(); g.e_level = eLevel.data();
    g.e_measurement = eMeas.data(); g.e_information = eInfo.data(); g.e_kernel = eKernel.data(); g.e_kernel_delta = eDelta.data(); g.e_param = ePrm.data();
    _estimateSize = vEst.size();
    return ok(g2ocu_set_graph(_h, &g));
  }

  void writeBack() {
    std::vector<double> est(_estimateSize);
    if (!ok(g2ocu_get_estimates(_h, est.data()))) return;
    size_t o = 0;
    for (auto* v : _vertices) {
      if (auto* x = dynamic_cast<VertexSE3*>(v)) {
        Isometry3 T = Isometry3::Identity();
        for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) T.matrix()(r, c) = est[o + r + 3 * c];
        for (int r = 0; r < 3; ++r) T.translation()(r) = est[o + 9 + r];
        x->setEstimate(T); o += 12;
      } else if (auto* x = dynamic_cast<VertexSE3Expmap*>(v)) {
        SE3Quat T; Vector7 t; for (int i = 0; i < 7; ++i) t[i] = est[o + i]; T.fromVector(t); x->setEstimate(T); o += 7;
      } else { v->setEstimateData(est.data() + o); o += v->estimateDimension(); }
    }
  }

  g2ocu_solver* _h = nullptr; int _algorithm, _poseDim, _landmarkDim;
  Property<number_t>* _userDeltaInit = nullptr; Property<number_t>* _lambdaFactor = nullptr; double _dogleg[5] = {1e4, 0, 0, 1e-7, 1};
  std::vector<OptimizableGraph::Vertex*> _vertices; size_t _estimateSize = 0;
  Property<number_t>* _userLambdaInit; Property<int>* _maxTrialsAfterFailure;
  number_t _lambda = -1; int _levenbergIterations = 0;
};

class CudaSolverCreator : public AbstractOptimizationAlgorithmCreator {
 public:
  explicit CudaSolverCreator(const OptimizationAlgorithmProperty& p) : AbstractOptimizationAlgorithmCreator(p) {}
  OptimizationAlgorithm* construct() override {
    const std::string& n = property().name;
    return new OptimizationAlgorithmCuda(n.substr(0, 2) == "lm" ? G2OCU_ALGORITHM_LM : n.substr(0, 2) == "dl" ? G2OCU_ALGORITHM_DOGLEG : G2OCU_ALGORITHM_GN, property().poseDim, property().landmarkDim,
                                         n.find("_dense") != std::string::npos ? G2OCU_LINEAR_DENSE : G2OCU_LINEAR_PCG);
  }
};

G2O_REGISTER_OPTIMIZATION_LIBRARY(cuda);
G2O_REGISTER_OPTIMIZATION_ALGORITHM(gn_var_cuda, new CudaSolverCreator(OptimizationAlgorithmProperty("gn_var_cuda", "Gauss-Newton: PCG on the GPU (variable blocksize)", "CUDA", false, Eigen::Dynamic, Eigen::Dynamic)));
G2O_REGISTER_OPTIMIZATION_ALGORITHM(lm_var_cuda, new CudaSolverCreator(OptimizationAlgorithmProperty("lm_var_cuda", "Levenberg: PCG on the GPU (variable blocksize)", "CUDA", false, Eigen::Dynamic, Eigen::Dynamic)));
G2O_REGISTER_OPTIMIZATION_ALGORITHM(gn_fix3_2_cuda, new CudaSolverCreator(OptimizationAlgorithmProperty("gn_fix3_2_cuda", "Gauss-Newton: Schur + PCG on the GPU", "CUDA", true, 3, 2)));
G2O_REGISTER_OPTIMIZATION_ALGORITHM(lm_fix3_2_cuda, new CudaSolverCreator(OptimizationAlgorithmProperty("lm_fix3_2_cuda", "Levenberg: Schur + PCG on the GPU", "CUDA", true, 3, 2)));
G2O_REGISTER_OPTIMIZATION_ALGORITHM(gn_fix6_3_cuda, new CudaSolverCreator(OptimizationAlgorithmProperty("gn_fix6_3_cuda", "Gauss-Newton: Schur + PCG on the GPU", "CUDA", true, 6, 3)));
G2O_REGISTER_OPTIMIZATION_ALGORITHM(lm_fix6_3_cuda, new CudaSolverCreator(OptimizationAlgorithmProperty("lm_fix6_3_cuda", "Levenberg: Schur + PCG on the GPU", "CUDA", true, 6, 3)));
G2O_REGISTER_OPTIMIZATION_ALGORITHM(gn_fix7_3_cuda, new CudaSolverCreator(OptimizationAlgorithmProperty("gn_fix7_3_cuda", "Gauss-Newton: Schur + PCG on the GPU (sim3 types are rejected at init)", "CUDA", true, 7, 3)));
G2O_REGISTER_OPTIMIZATION_ALGORITHM(lm_fix7_3_cuda, new CudaSolverCreator(OptimizationAlgorithmProperty("lm_fix7_3_cuda", "Levenberg: Schur + PCG on the GPU (sim3 types are rejected at init)", "CUDA", true, 7, 3)));
G2O_REGISTER_OPTIMIZATION_ALGORITHM(gn_fix9_3_cuda, new CudaSolverCreator(OptimizationAlgorithmProperty("gn_fix9_3_cuda", "Gauss-Newton: Schur + PCG on the GPU (BAL cameras)", "CUDA", true, 9, 3)));
G2O_REGISTER_OPTIMIZATION_ALGORITHM(gn_dense_cuda, new CudaSolverCreator(OptimizationAlgorithmProperty("gn_dense_cuda", "Gauss-Newton: dense FP64 Cholesky of the (reduced) system on the GPU", "CUDA", false, Eigen::Dynamic, Eigen::Dynamic)));
G2O_REGISTER_OPTIMIZATION_ALGORITHM(gn_dense3_2_cuda, new CudaSolverCreator(OptimizationAlgorithmProperty("gn_dense3_2_cuda", "Gauss-Newton: dense FP64 Cholesky of the (reduced) system on the GPU", "CUDA", true, 3, 2)));
G2O_REGISTER_OPTIMIZATION_ALGORITHM(gn_dense6_3_cuda, new CudaSolverCreator(OptimizationAlgorithmProperty("gn_dense6_3_cuda", "Gauss-Newton: dense FP64 Cholesky of the (reduced) system on the GPU", "CUDA", true, 6, 3)));
G2O_REGISTER_OPTIMIZATION_ALGORITHM(gn_dense7_3_cuda, new CudaSolverCreator(OptimizationAlgorithmProperty("gn_dense7_3_cuda", "Gauss-Newton: dense FP64 Cholesky of the (reduced) system on the GPU", "CUDA", true, 7, 3)));
G2O_REGISTER_OPTIMIZATION_ALGORITHM(gn_dense9_3_cuda, new CudaSolverCreator(OptimizationAlgorithmProperty("gn_dense9_3_cuda", "Gauss-Newton: dense FP64 Cholesky of the (reduced) system on the GPU", "CUDA", true, 9, 3)));
G2O_REGISTER_OPTIMIZATION_ALGORITHM(lm_fix9_3_cuda, new CudaSolverCreator(OptimizationAlgorithmProperty("lm_fix9_3_cuda", "Levenberg: Schur + PCG on the GPU (BAL cameras)", "CUDA", true, 9, 3)));
G2O_REGISTER_OPTIMIZATION_ALGORITHM(lm_dense_cuda, new CudaSolverCreator(OptimizationAlgorithmProperty("lm_dense_cuda", "Levenberg: dense FP64 Cholesky of the (reduced) system on the GPU", "CUDA", false, Eigen::Dynamic, Eigen::Dynamic)));
G2O_REGISTER_OPTIMIZATION_ALGORITHM(lm_dense3_2_cuda, new CudaSolverCreator(OptimizationAlgorithmProperty("lm_dense3_2_cuda", "Levenberg: dense FP64 Cholesky of the (reduced) system on the GPU", "CUDA", true, 3, 2)));
G2O_REGISTER_OPTIMIZATION_ALGORITHM(lm_dense6_3_cuda, new CudaSolverCreator(OptimizationAlgorithmProperty("lm_dense6_3_cuda", "Levenberg: dense FP64 Cholesky of the (reduced) system on the GPU", "CUDA", true, 6, 3)));
G2O_REGISTER_OPTIMIZATION_ALGORITHM(lm_dense7_3_cuda, new CudaSolverCreator(OptimizationAlgorithmProperty("lm_dense7_3_cuda", "Levenberg: dense FP64 Cholesky of the (reduced) system on the GPU", "CUDA", true, 7, 3)));
G2O_REGISTER_OPTIMIZATION_ALGORITHM(lm_dense9_3_cuda, new CudaSolverCreator(OptimizationAlgorithmProperty("lm_dense9_3_cuda", "Levenberg: dense FP64 Cholesky of the (reduced) system on the GPU", "CUDA", true, 9, 3)));
G2O_REGISTER_OPTIMIZATION_ALGORITHM(dl_var_cuda, new CudaSolverCreator(OptimizationAlgorithmProperty("dl_var_cuda", "Dogleg: block-Jacobi PCG on the GPU (variable blocksize)", "CUDA", false, Eigen::Dynamic, Eigen::Dynamic)));

}  // namespace g2o
