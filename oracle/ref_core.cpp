// TEST INFRASTRUCTURE.  The REAL reference end to end for 2-D SLAM and bundle-adjustment graphs: g2o/core (SparseOptimizer, OptimizableGraph,
// BlockSolver, OptimizationAlgorithmLevenberg / GaussNewton / Dogleg, robust kernels), g2o/stuff, g2o/solvers/pcg/linear_solver_pcg.h, the slam2d
// types VertexSE2, VertexPointXY, EdgeSE2, EdgeSE2PointXY and the sba types VertexSE3Expmap, VertexSBAPointXYZ, EdgeProjectXYZ2UV (+ CameraParameters),
// EdgeSE3ProjectXYZ, EdgeSE3Expmap, the slam3d types VertexSE3, EdgeSE3 and the BAL types of examples/bal/bal_example.cpp, compiled unmodified from /root/reference against the Eigen stand-in in
// oracle/eigen_shim (NOT Eigen; see its Core header) by `make -C oracle ref_core` into oracle/_ref/libg2o_ref_core.so.
// This file only builds a g2o::SparseOptimizer from the flat graph layout of include/g2ocu.h, runs optimize() and reads the results back.
// tests/test_reference_core.py compares the oracle (and through it the CUDA path) with what comes out of here.
#ifdef _OPENMP
#include <omp.h>
#endif
#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "g2o/core/block_solver.h"
#include "g2o/core/optimization_algorithm_dogleg.h"
#include "g2o/core/optimization_algorithm_factory.h"
#include "g2o/core/optimization_algorithm_gauss_newton.h"
#include "g2o/core/optimization_algorithm_levenberg.h"
#include "g2o/core/robust_kernel_factory.h"
#include "g2o/core/sparse_optimizer.h"
#include "g2o/solvers/csparse/linear_solver_csparse.h"
#include "g2o/solvers/pcg/linear_solver_pcg.h"
#include "g2o/types/sba/types_six_dof_expmap.h"
#include "g2o/types/slam3d/edge_se3.h"
#include "g2o/types/slam3d/vertex_se3.h"
#include "g2o/types/slam2d/edge_se2.h"
#include "g2o/types/slam2d/edge_se2_pointxy.h"
#include "g2o/types/slam2d/vertex_point_xy.h"
#include "g2o/types/slam2d/vertex_se2.h"

// the BAL types come from g2o/examples/bal/bal_example.cpp through oracle/ref_core_bal.cpp
extern "C" {
g2o::OptimizableGraph::Vertex* refbal_new_camera(const double* est9);
g2o::OptimizableGraph::Vertex* refbal_new_point(const double* est3);
g2o::OptimizableGraph::Edge* refbal_new_edge(const double* z2, const double* info4);
void refbal_camera_estimate(const g2o::OptimizableGraph::Vertex* v, double* out9);
void refbal_point_estimate(const g2o::OptimizableGraph::Vertex* v, double* out3);
}

namespace {

// same field order as g2ocu_graph (include/g2ocu.h) / orc_graph (oracle/g2o_oracle.cpp)
struct FlatGraph {
  int32_t n_vertices; const int32_t* v_id; const int32_t* v_type; const uint8_t* v_fixed; const uint8_t* v_marginalized; const double* v_estimate;
  int32_t n_edges; const int32_t* e_type; const int32_t* e_v0; const int32_t* e_v1; const int32_t* e_level;
  const double* e_measurement; const double* e_information; const int32_t* e_kernel; const double* e_kernel_delta; const double* e_param;
};
g2o::Isometry3 isoFrom12(const double* v) {      // R column-major (9) + t (3), the layout of include/g2ocu.h
  g2o::Isometry3 T = g2o::Isometry3::Identity();
  for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) T.linear()(r, c) = v[r + 3 * c];
  for (int r = 0; r < 3; ++r) T.translation()[r] = v[9 + r];
  return T;
}
const char* const kKernelNames[10] = {"", "Huber", "PseudoHuber", "Cauchy", "GemanMcClure", "Welsch", "Fair", "Tukey", "Saturated", "DCS"};

struct Handle {
  g2o::SparseOptimizer optimizer;
  std::vector<g2o::OptimizableGraph::Vertex*> vertices;      // in the caller's order
  std::vector<int> vtype;
  g2o::OptimizationAlgorithmLevenberg* lm = nullptr;
  g2o::OptimizationAlgorithmDogleg* dl = nullptr;
  std::function<void(double, int, bool)> setPcg;             // LinearSolverPCG::setTolerance / setMaxIterations / setAbsoluteTolerance
  g2o::BlockSolverBase* blockSolver = nullptr;               // owned by the algorithm
  std::function<bool(const std::string&, std::vector<int32_t>&)> structureI32;   // block patterns of the BlockSolver's protected matrices
  std::function<bool(const std::string&, std::vector<double>&)> structureF64;    // and their values
  std::vector<std::vector<double> > cameras;                 // distinct (f, cx, cy) of the EdgeProjectXYZ2UV edges -> parameter id
  std::string err;
};

// BlockSolver keeps _Hpp / _Hll / _Hpl / _Hschur protected (block_solver.h:150-170); this subclass adds nothing but read access, so that the
// patterns buildStructure produced (block_solver.hpp:103-256) and the values buildSystem / solve left in them can be compared with the backend's.
template <class M> void patternOf(const M& mat, std::vector<int32_t>& out, bool ptr) {
  out.clear();
  if (ptr) { int n = 0; for (const auto& c : mat.blockCols()) { out.push_back(n); n += (int)c.size(); } out.push_back(n); }
  else for (const auto& c : mat.blockCols()) for (const auto& kv : c) out.push_back(kv.first);
}
template <class M> void valuesOf(const M& mat, std::vector<double>& out) {
  out.clear();
  for (const auto& c : mat.blockCols()) for (const auto& kv : c) { const auto& b = *kv.second; for (int cc = 0; cc < b.cols(); ++cc) for (int r = 0; r < b.rows(); ++r) out.push_back(b(r, cc)); }
}
template <class BlockSolverT> struct ExposedBlockSolver : public BlockSolverT {
  using BlockSolverT::BlockSolverT;
  bool i32(const std::string& n, std::vector<int32_t>& out) const {
    if (n == "pose_block_indices" && this->_Hpp) { out.assign(this->_Hpp->colBlockIndices().begin(), this->_Hpp->colBlockIndices().end()); return true; }
    if (n == "landmark_block_indices" && this->_Hll) { out.assign(this->_Hll->colBlockIndices().begin(), this->_Hll->colBlockIndices().end()); return true; }
    if ((n == "hpp_colptr" || n == "hpp_rowidx") && this->_Hpp) { patternOf(*this->_Hpp, out, n == "hpp_colptr"); return true; }
    if ((n == "hpl_colptr" || n == "hpl_rowidx") && this->_Hpl) { patternOf(*this->_Hpl, out, n == "hpl_colptr"); return true; }
    if ((n == "hll_colptr" || n == "hll_rowidx") && this->_Hll) { patternOf(*this->_Hll, out, n == "hll_colptr"); return true; }
    if ((n == "hschur_colptr" || n == "hschur_rowidx") && this->_Hschur) { patternOf(*this->_Hschur, out, n == "hschur_colptr"); return true; }
    if ((n == "hschur_t_colptr" || n == "hschur_t_rowidx") && this->_HschurTransposedCCS) {   // row-major mirror of the upper pattern (block_solver.hpp:253)
      out.clear(); int c = 0;
      for (const auto& col : this->_HschurTransposedCCS->blockCols()) { if (n == "hschur_t_colptr") { out.push_back(c); c += (int)col.size(); } else for (const auto& rb : col) out.push_back(rb.row); }
      if (n == "hschur_t_colptr") out.push_back(c);
      return true;
    }
    if (n == "dims") { out = {this->_numPoses, this->_numLandmarks, this->_sizePoses, this->_sizeLandmarks}; return true; }
    return false;
  }
  bool f64(const std::string& n, std::vector<double>& out) const {
    if (n == "hpp_values" && this->_Hpp) { valuesOf(*this->_Hpp, out); return true; }
    if (n == "hpl_values" && this->_Hpl) { valuesOf(*this->_Hpl, out); return true; }
    if (n == "hll_values" && this->_Hll) { valuesOf(*this->_Hll, out); return true; }
    if (n == "hschur_values" && this->_Hschur) { valuesOf(*this->_Hschur, out); return true; }
    if (n == "b" || n == "x") { const double* v = n == "b" ? this->b() : this->x(); out.assign(v, v + this->vectorSize()); return true; }
    if (n == "bschur" && this->_bschur) { out.assign(this->_bschur.get(), this->_bschur.get() + this->_sizePoses); return true; }
    return false;
  }
};
template <class BlockSolverT> void expose(Handle& h, ExposedBlockSolver<BlockSolverT>* raw) {
  h.blockSolver = raw;
  h.structureI32 = [raw](const std::string& n, std::vector<int32_t>& out) { return raw->i32(n, out); };
  h.structureF64 = [raw](const std::string& n, std::vector<double>& out) { return raw->f64(n, out); };
}

template <class BlockSolverT> std::unique_ptr<BlockSolverT> makeBlockSolver(Handle& h) {
  typedef g2o::LinearSolverPCG<typename BlockSolverT::PoseMatrixType> Pcg;
  std::unique_ptr<Pcg> linear(new Pcg());
  Pcg* raw = linear.get();                                   // owned by the block solver, which the algorithm owns, which the optimizer owns
  h.setPcg = [raw](double tol, int maxIter, bool absolute) { raw->setTolerance(tol); raw->setMaxIterations(maxIter); raw->setAbsoluteTolerance(absolute); };
  ExposedBlockSolver<BlockSolverT>* bs = new ExposedBlockSolver<BlockSolverT>(std::move(linear));
  expose(h, bs);
  return std::unique_ptr<BlockSolverT>(bs);
}

// BlockSolver + LinearSolverCSparse (solvers/csparse/solver_csparse.cpp:44-52): block ordering for the fixed-size solvers, scalar AMD for `var`
template <class BlockSolverT> std::unique_ptr<BlockSolverT> makeCSparseBlockSolver(bool blockOrdering) {
  typedef g2o::LinearSolverCSparse<typename BlockSolverT::PoseMatrixType> Chol;
  std::unique_ptr<Chol> linear(new Chol());
  linear->setBlockOrdering(blockOrdering);
  return std::unique_ptr<BlockSolverT>(new BlockSolverT(std::move(linear)));
}

}  // namespace

extern "C" {

// algorithm: "factory" with blockSolver = a name known to OptimizationAlgorithmFactory (plugins), or
// algorithm: "gn" | "lm" | "dl"; blockSolver: "3_2" | "6_3" | "9_3" (BlockSolver<BlockSolverTraits<P,L>>) | "var" (BlockSolverX) with LinearSolverPCG, or the same
// names + "_csparse" with LinearSolverCSparse (the reference's gn_/lm_fixP_L and lm_var solvers, solvers/csparse/solver_csparse.cpp)
void* refcore_create(const FlatGraph* g, const char* algorithm, const char* blockSolver) {
  std::unique_ptr<Handle> h(new Handle);
  const std::string alg(algorithm), bs(blockSolver);
  std::unique_ptr<g2o::BlockSolverBase> solver;
  if (bs == "3_2") solver = makeBlockSolver<g2o::BlockSolver<g2o::BlockSolverTraits<3, 2> > >(*h);
  else if (bs == "6_3") solver = makeBlockSolver<g2o::BlockSolver<g2o::BlockSolverTraits<6, 3> > >(*h);
  else if (bs == "9_3") solver = makeBlockSolver<g2o::BlockSolver<g2o::BlockSolverTraits<9, 3> > >(*h);   // bal_example.cpp:301
  else if (bs == "var") solver = makeBlockSolver<g2o::BlockSolverX>(*h);
  else if (bs == "3_2_csparse") solver = makeCSparseBlockSolver<g2o::BlockSolver<g2o::BlockSolverTraits<3, 2> > >(true);
  else if (bs == "6_3_csparse") solver = makeCSparseBlockSolver<g2o::BlockSolver<g2o::BlockSolverTraits<6, 3> > >(true);
  else if (bs == "9_3_csparse") solver = makeCSparseBlockSolver<g2o::BlockSolver<g2o::BlockSolverTraits<9, 3> > >(true);
  else if (bs == "var_csparse") solver = makeCSparseBlockSolver<g2o::BlockSolverX>(false);
  else if (alg != "factory") return nullptr;
  g2o::OptimizationAlgorithm* a = nullptr;
  if (alg == "factory") {      // a solver registered with the reference's OptimizationAlgorithmFactory, e.g. by the CUDA plugin libg2o_solver_cuda.so
    g2o::OptimizationAlgorithmProperty prop;
    a = g2o::OptimizationAlgorithmFactory::instance()->construct(bs, prop);
    if (!a) return nullptr;
  } else
  if (alg == "lm") a = h->lm = new g2o::OptimizationAlgorithmLevenberg(std::move(solver));
  else if (alg == "gn") a = new g2o::OptimizationAlgorithmGaussNewton(std::move(solver));
  else if (alg == "dl") a = h->dl = new g2o::OptimizationAlgorithmDogleg(std::move(solver));
  else if (alg != "factory") return nullptr;
  h->optimizer.setAlgorithm(a);
  size_t eo = 0;
  for (int i = 0; i < g->n_vertices; ++i) {
    g2o::OptimizableGraph::Vertex* v = nullptr;
    if (g->v_type[i] == 1) { g2o::VertexSE2* p = new g2o::VertexSE2; p->setEstimate(g2o::SE2(g->v_estimate[eo], g->v_estimate[eo + 1], g->v_estimate[eo + 2])); eo += 3; v = p; }
    else if (g->v_type[i] == 2) { g2o::VertexPointXY* p = new g2o::VertexPointXY; p->setEstimate(g2o::Vector2(g->v_estimate[eo], g->v_estimate[eo + 1])); eo += 2; v = p; }
    else if (g->v_type[i] == 3) { g2o::VertexSE3* p = new g2o::VertexSE3; p->setEstimate(isoFrom12(g->v_estimate + eo)); eo += 12; v = p; }
    else if (g->v_type[i] == 4) {                           // VertexSE3Expmap, estimate = SE3Quat::toVector (t, qx, qy, qz, qw)
      g2o::VertexSE3Expmap* p = new g2o::VertexSE3Expmap; g2o::Vector7 a; for (int k = 0; k < 7; ++k) a[k] = g->v_estimate[eo + k];
      g2o::SE3Quat T; T.fromVector(a); p->setEstimate(T); eo += 7; v = p;
    } else if (g->v_type[i] == 5) { g2o::VertexSBAPointXYZ* p = new g2o::VertexSBAPointXYZ; p->setEstimate(g2o::Vector3(g->v_estimate[eo], g->v_estimate[eo + 1], g->v_estimate[eo + 2])); eo += 3; v = p; }
    else if (g->v_type[i] == 6) { v = refbal_new_camera(g->v_estimate + eo); eo += 9; }
    else if (g->v_type[i] == 7) { v = refbal_new_point(g->v_estimate + eo); eo += 3; }
    else return nullptr;
    v->setId(g->v_id[i]); v->setFixed(g->v_fixed[i] != 0); v->setMarginalized(g->v_marginalized[i] != 0);
    if (!h->optimizer.addVertex(v)) return nullptr;
    h->vertices.push_back(v); h->vtype.push_back(g->v_type[i]);
  }
  size_t mo = 0, io = 0, po = 0;
  for (int i = 0; i < g->n_edges; ++i) {
    g2o::OptimizableGraph::Edge* e = nullptr;
    if (g->e_type[i] == 1) {
      g2o::EdgeSE2* p = new g2o::EdgeSE2; p->setMeasurement(g2o::SE2(g->e_measurement[mo], g->e_measurement[mo + 1], g->e_measurement[mo + 2]));
      g2o::Matrix3 info; for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) info(r, c) = g->e_information[io + r + 3 * c];
      p->setInformation(info); mo += 3; io += 9; e = p;
    } else if (g->e_type[i] == 2) {
      g2o::EdgeSE2PointXY* p = new g2o::EdgeSE2PointXY; p->setMeasurement(g2o::Vector2(g->e_measurement[mo], g->e_measurement[mo + 1]));
      g2o::Matrix2 info; for (int c = 0; c < 2; ++c) for (int r = 0; r < 2; ++r) info(r, c) = g->e_information[io + r + 2 * c];
      p->setInformation(info); mo += 2; io += 4; e = p;
    } else if (g->e_type[i] == 3) {                         // EdgeSE3, measurement Isometry3 (12)
      g2o::EdgeSE3* p = new g2o::EdgeSE3; p->setMeasurement(isoFrom12(g->e_measurement + mo));
      Eigen::Matrix<number_t, 6, 6> info; for (int c = 0; c < 6; ++c) for (int r = 0; r < 6; ++r) info(r, c) = g->e_information[io + r + 6 * c];
      p->setInformation(info); mo += 12; io += 36; e = p;
    } else if (g->e_type[i] == 4) {                         // EdgeSE3Expmap, measurement SE3Quat (7)
      g2o::EdgeSE3Expmap* p = new g2o::EdgeSE3Expmap; g2o::Vector7 a; for (int k = 0; k < 7; ++k) a[k] = g->e_measurement[mo + k];
      g2o::SE3Quat Z; Z.fromVector(a); p->setMeasurement(Z);
      Eigen::Matrix<number_t, 6, 6> info; for (int c = 0; c < 6; ++c) for (int r = 0; r < 6; ++r) info(r, c) = g->e_information[io + r + 6 * c];
      p->setInformation(info); mo += 7; io += 36; e = p;
    } else if (g->e_type[i] == 5 || g->e_type[i] == 6) {    // (VertexSBAPointXYZ, VertexSE3Expmap) projections
      g2o::Matrix2 info; for (int c = 0; c < 2; ++c) for (int r = 0; r < 2; ++r) info(r, c) = g->e_information[io + r + 2 * c];
      const g2o::Vector2 z(g->e_measurement[mo], g->e_measurement[mo + 1]);
      if (g->e_type[i] == 5) {                              // EdgeProjectXYZ2UV with CameraParameters (f, cx, cy)
        std::vector<double> cam(g->e_param + po, g->e_param + po + 3); po += 3;
        size_t id = 0; while (id < h->cameras.size() && h->cameras[id] != cam) ++id;
        if (id == h->cameras.size()) {
          h->cameras.push_back(cam);
          g2o::CameraParameters* cp = new g2o::CameraParameters(cam[0], g2o::Vector2(cam[1], cam[2]), 0.);
          cp->setId((int)id);
          if (!h->optimizer.addParameter(cp)) return nullptr;
        }
        g2o::EdgeProjectXYZ2UV* p = new g2o::EdgeProjectXYZ2UV; p->setMeasurement(z); p->setInformation(info); p->setParameterId(0, (int)id); e = p;
      } else {                                              // the fork's EdgeSE3ProjectXYZ with fx, fy, cx, cy members
        g2o::EdgeSE3ProjectXYZ* p = new g2o::EdgeSE3ProjectXYZ; p->setMeasurement(z); p->setInformation(info);
        p->fx = g->e_param[po]; p->fy = g->e_param[po + 1]; p->cx = g->e_param[po + 2]; p->cy = g->e_param[po + 3]; po += 4; e = p;
      }
      mo += 2; io += 4;
    } else if (g->e_type[i] == 7) { e = refbal_new_edge(g->e_measurement + mo, g->e_information + io); mo += 2; io += 4; }   // EdgeObservationBAL (camera, point)
    else return nullptr;
    e->setVertex(0, h->vertices[g->e_v0[i]]); e->setVertex(1, h->vertices[g->e_v1[i]]);
    e->setLevel(g->e_level ? g->e_level[i] : 0);
    const int kernel = g->e_kernel ? g->e_kernel[i] : 0;
    if (kernel) {
      g2o::RobustKernel* k = g2o::RobustKernelFactory::instance()->construct(kKernelNames[kernel]);
      if (!k) return nullptr;
      k->setDelta(g->e_kernel_delta ? g->e_kernel_delta[i] : 1.0);
      e->setRobustKernel(k);
    }
    if (!h->optimizer.addEdge(e)) return nullptr;
  }
  return h.release();
}
void refcore_destroy(void* hh) { delete (Handle*)hh; }
// One linearisation through the reference's own virtuals, as OptimizationAlgorithmLevenberg::solve strings them together (levenberg.cpp:58-110):
// init, buildStructure, computeActiveErrors, buildSystem, setLambda(lambda, true), solve, restoreDiagonal.  Afterwards the structure / value
// getters below read the BlockSolver's matrices (Hschur holds the damped reduced system of that solve).
int refcore_linearize(void* hh, double lambda) {
  Handle* h = (Handle*)hh;
  if (!h->blockSolver || !h->optimizer.solver()) return -1;
  if (!h->optimizer.solver()->init(false)) return -2;
  if (!h->blockSolver->buildStructure()) return -3;
  h->optimizer.computeActiveErrors();
  h->blockSolver->buildSystem();                 // returns false unconditionally in the reference (block_solver.hpp:520); its callers ignore it
  h->blockSolver->setLambda(lambda, true);
  const bool ok = h->blockSolver->solve();
  h->blockSolver->restoreDiagonal();
  return ok ? 1 : 0;
}
int64_t refcore_structure_i32(void* hh, const char* name, int32_t* out, int64_t cap) {
  Handle* h = (Handle*)hh; std::vector<int32_t> v;
  if (!h->structureI32 || !h->structureI32(name, v)) return -1;
  if (out) std::memcpy(out, v.data(), sizeof(int32_t) * (size_t)std::min<int64_t>(cap, (int64_t)v.size()));
  return (int64_t)v.size();
}
int64_t refcore_structure_f64(void* hh, const char* name, double* out, int64_t cap) {
  Handle* h = (Handle*)hh; std::vector<double> v;
  if (!h->structureF64 || !h->structureF64(name, v)) return -1;
  if (out) std::memcpy(out, v.data(), sizeof(double) * (size_t)std::min<int64_t>(cap, (int64_t)v.size()));
  return (int64_t)v.size();
}
// LinearSolverPCG properties (linear_solver_pcg.h:53-57 defaults: 1e-6, -1, absolute)
void refcore_set_pcg(void* hh, double tolerance, int maxIterations, int absoluteTolerance) { Handle* h = (Handle*)hh; if (h->setPcg) h->setPcg(tolerance, maxIterations, absoluteTolerance != 0); }
// threads of the reference's OpenMP regions (its summation order, hence its last digits, depends on them); 0 or less: leave as is
void refcore_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

// SparseOptimizer::computeMarginals(spinv, blockIndices) (sparse_optimizer.cpp:594-596 -> block_solver.hpp:451-459 -> the linear solver's
// solvePattern on Hpp, linear_solver_csparse.h:190-216, marginal_covariance_cholesky.cpp:153-222).  Pairs are block (hessian) indices;
// out receives the blocks one after the other, each column-major.  Returns 1 / 0 as the reference does, -1 when `cap` is too small.
int refcore_compute_marginals(void* hh, int nPairs, const int32_t* rows, const int32_t* cols, double* out, int64_t cap) {
  Handle* h = (Handle*)hh;
  std::vector<std::pair<int, int> > idx;
  for (int i = 0; i < nPairs; ++i) idx.push_back(std::make_pair((int)rows[i], (int)cols[i]));
  g2o::SparseBlockMatrix<g2o::MatrixX> spinv;
  if (!h->optimizer.computeMarginals(spinv, idx)) return 0;
  int64_t off = 0;
  for (int i = 0; i < nPairs; ++i) {
    const g2o::MatrixX* b = spinv.block(rows[i], cols[i]);
    if (!b) return 0;
    const int64_t sz = (int64_t)b->rows() * b->cols();
    if (off + sz > cap) return -1;
    for (int c = 0; c < b->cols(); ++c) for (int r = 0; r < b->rows(); ++r) out[off++] = (*b)(r, c);
  }
  return 1;
}

int refcore_initialize_optimization(void* hh, int level) { return ((Handle*)hh)->optimizer.initializeOptimization(level) ? 1 : 0; }

// SparseOptimizer::optimize(iterations); stats: iterations x 13 doubles from G2OBatchStatistics = chi2, levenbergIterations, iterationsLinearSolver,
// hessianPoseDimension, hessianLandmarkDimension, iteration, timeIteration, timeLinearSolution, timeResiduals, timeQuadraticForm,
// timeSchurComplement, timeLinearSolver, timeUpdate (lambda is not in there: refcore_current_lambda)
int refcore_optimize(void* hh, int iterations, double* stats) {
  Handle* h = (Handle*)hh;
  h->optimizer.setComputeBatchStatistics(true);
  const int n = h->optimizer.optimize(iterations);
  const g2o::BatchStatisticsContainer& bs = h->optimizer.batchStatistics();
  for (size_t i = 0; i < bs.size() && (int)i < iterations; ++i) {
    double* s = stats + 13 * i;
    s[0] = bs[i].chi2; s[1] = bs[i].levenbergIterations; s[2] = bs[i].iterationsLinearSolver; s[3] = (double)bs[i].hessianPoseDimension;
    s[4] = (double)bs[i].hessianLandmarkDimension; s[5] = bs[i].iteration; s[6] = bs[i].timeIteration; s[7] = bs[i].timeLinearSolution;
    s[8] = bs[i].timeResiduals; s[9] = bs[i].timeQuadraticForm; s[10] = bs[i].timeSchurComplement; s[11] = bs[i].timeLinearSolver; s[12] = bs[i].timeUpdate;
  }
  return n;
}
// The same, bounded in time with the reference's own stop mechanism: SparseOptimizer::setForceStopFlag (sparse_optimizer.h:186-190) - a watcher
// thread raises the flag after `seconds`, optimize() then starts no further iteration (sparse_optimizer.cpp:396).  Returns what optimize returned.
int refcore_optimize_budget(void* hh, int iterations, double* stats, double seconds) {
  Handle* h = (Handle*)hh;
  bool stop = false; std::atomic<bool> done(false);
  h->optimizer.setForceStopFlag(&stop);
  std::thread watcher([&]() {
    const auto t0 = std::chrono::steady_clock::now();
    while (!done.load()) {
      if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > seconds) { stop = true; break; }
      std::this_thread::sleep_for(std::chrono::milliseconds(50));
    }
  });
  const int n = refcore_optimize(hh, iterations, stats);
  done.store(true); watcher.join();
  h->optimizer.setForceStopFlag(nullptr);
  return n;
}
double refcore_current_lambda(void* hh) { Handle* h = (Handle*)hh; return h->lm ? h->lm->currentLambda() : 0.0; }
// OptimizationAlgorithmDogleg::trustRegion(), lastStep()
void refcore_dogleg_state(void* hh, double* out2) { Handle* h = (Handle*)hh; out2[0] = h->dl ? h->dl->trustRegion() : 0.0; out2[1] = h->dl ? h->dl->lastStep() : 0.0; }
double refcore_active_robust_chi2(void* hh) { Handle* h = (Handle*)hh; h->optimizer.computeActiveErrors(); return h->optimizer.activeRobustChi2(); }
double refcore_active_chi2(void* hh) { Handle* h = (Handle*)hh; h->optimizer.computeActiveErrors(); return h->optimizer.activeChi2(); }
// hessianIndex of every vertex in the caller's order (-1 fixed / inactive)
void refcore_hessian_index(void* hh, int32_t* out) { Handle* h = (Handle*)hh; for (size_t i = 0; i < h->vertices.size(); ++i) out[i] = h->vertices[i]->hessianIndex(); }
// packed estimates in the caller's vertex order (SE2: x y theta; point: x y)
void refcore_estimates(void* hh, double* out) {
  Handle* h = (Handle*)hh; size_t o = 0;
  for (size_t i = 0; i < h->vertices.size(); ++i) {
    if (h->vtype[i] == 1) { const g2o::SE2& e = static_cast<g2o::VertexSE2*>(h->vertices[i])->estimate(); out[o++] = e[0]; out[o++] = e[1]; out[o++] = e[2]; }
    else if (h->vtype[i] == 2) { const g2o::Vector2& e = static_cast<g2o::VertexPointXY*>(h->vertices[i])->estimate(); out[o++] = e[0]; out[o++] = e[1]; }
    else if (h->vtype[i] == 3) { const g2o::Isometry3& e = static_cast<g2o::VertexSE3*>(h->vertices[i])->estimate();
      for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) out[o++] = e.linear()(r, c); for (int r = 0; r < 3; ++r) out[o++] = e.translation()[r]; }
    else if (h->vtype[i] == 4) { const g2o::Vector7 e = static_cast<g2o::VertexSE3Expmap*>(h->vertices[i])->estimate().toVector(); for (int k = 0; k < 7; ++k) out[o++] = e[k]; }
    else if (h->vtype[i] == 6) { refbal_camera_estimate(h->vertices[i], out + o); o += 9; }
    else if (h->vtype[i] == 7) { refbal_point_estimate(h->vertices[i], out + o); o += 3; }
    else { const g2o::Vector3& e = static_cast<g2o::VertexSBAPointXYZ*>(h->vertices[i])->estimate(); out[o++] = e[0]; out[o++] = e[1]; out[o++] = e[2]; }
  }
}

}  // extern "C"
