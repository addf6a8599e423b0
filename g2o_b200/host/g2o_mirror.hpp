// Eigen-free host-side mirror of the part of g2o's plugin interface that the LM hot path uses, implemented over the
// C ABI in include/g2ocu.h.  Same class names, method names, argument meaning and error behaviour as the reference
// (paths relative to the reference root), so that callers written against g2o read the same:
//
//   OptimizationAlgorithmProperty   core/optimization_algorithm_property.h:39-55
//   OptimizationAlgorithmFactory    core/optimization_algorithm_factory.h:72-119 (+ G2O_REGISTER_* macros :159-189)
//   OptimizationAlgorithm           core/optimization_algorithm.h:46-110
//   Solver / BlockSolverBase        core/solver.h:44-155, core/block_solver.h:87-95
//   SparseOptimizer                 core/sparse_optimizer.h (initializeOptimization, optimize, computeActiveErrors, ...)
//   vertex / edge classes           types/slam2d, types/slam3d, types/sba, examples/bal (data holders only: all math is on the GPU)
//
// There is no CPU implementation behind these classes: every numeric call forwards to libg2ocu.so and fails (returns false /
// -1 and prints to std::cerr, like the reference) when no CUDA device is available.
#pragma once
#include <array>
#include <cstdint>
#include <iostream>
#include <limits>
#include <list>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/g2ocu.h"

namespace g2o {

typedef double number_t;

// ---------------------------------------------------------------------------------------------------------------
struct OptimizationAlgorithmProperty {
  std::string name, desc, type;
  bool requiresMarginalize = false;
  int poseDim = -1, landmarkDim = -1;
  OptimizationAlgorithmProperty() {}
  OptimizationAlgorithmProperty(const std::string& name_, const std::string& desc_, const std::string& type_, bool requiresMarginalize_, int poseDim_, int landmarkDim_)
      : name(name_), desc(desc_), type(type_), requiresMarginalize(requiresMarginalize_), poseDim(poseDim_), landmarkDim(landmarkDim_) {}
};

// G2OBatchStatistics, core/batch_stats.h:40-78 (times in seconds, from CUDA events)
struct G2OBatchStatistics {
  int iteration = -1, numVertices = 0, numEdges = 0; number_t chi2 = 0;   // iteration -1: not valid yet (batch_stats.cpp:40-46)
  number_t timeResiduals = 0, timeLinearize = 0, timeQuadraticForm = 0; int levenbergIterations = 0;
  number_t timeSchurComplement = 0, timeSymbolicDecomposition = 0, timeNumericDecomposition = 0, timeLinearSolution = 0, timeLinearSolver = 0;
  number_t timeQrDecomposition = 0;                                       // the fork's JacobiSolver field; always 0 here
  int iterationsLinearSolver = 0; number_t timeUpdate = 0, timeIteration = 0, timeMarginals = 0;
  size_t hessianDimension = 0, hessianPoseDimension = 0, hessianLandmarkDimension = 0, choleskyNNZ = 0;
};
// one line per iteration, "name= value\t " in the reference's field order (batch_stats.cpp:48-83); what `g2o -stats file` writes
std::ostream& operator<<(std::ostream& os, const G2OBatchStatistics& st);
typedef std::vector<G2OBatchStatistics> BatchStatisticsContainer;

// ---------------------------------------------------------------------------------------------------------------
// graph objects: data holders with the reference's accessors
class RobustKernel {
 public:
  explicit RobustKernel(int code, number_t delta = 1.) : _code(code), _delta(delta) {}
  virtual ~RobustKernel() {}
  void setDelta(number_t d) { _delta = d; }
  number_t delta() const { return _delta; }
  int code() const { return _code; }
 protected:
  int _code; number_t _delta;
};
#define G2O_MIRROR_KERNEL(Name, CODE) class RobustKernel##Name : public RobustKernel { public: RobustKernel##Name() : RobustKernel(CODE) {} };
G2O_MIRROR_KERNEL(Huber, G2OCU_KERNEL_HUBER) G2O_MIRROR_KERNEL(PseudoHuber, G2OCU_KERNEL_PSEUDO_HUBER) G2O_MIRROR_KERNEL(Cauchy, G2OCU_KERNEL_CAUCHY)
G2O_MIRROR_KERNEL(GemanMcClure, G2OCU_KERNEL_GEMAN_MCCLURE) G2O_MIRROR_KERNEL(Welsch, G2OCU_KERNEL_WELSCH) G2O_MIRROR_KERNEL(Fair, G2OCU_KERNEL_FAIR)
G2O_MIRROR_KERNEL(Tukey, G2OCU_KERNEL_TUKEY) G2O_MIRROR_KERNEL(Saturated, G2OCU_KERNEL_SATURATED) G2O_MIRROR_KERNEL(DCS, G2OCU_KERNEL_DCS)

class OptimizableGraph {
 public:
  class Vertex {
   public:
    Vertex(int type, int estimateDim, int dim) : _type(type), _dimension(dim), _estimate(estimateDim, 0.0) {}
    virtual ~Vertex() {}
    int id() const { return _id; }
    void setId(int id) { _id = id; }
    bool fixed() const { return _fixed; }
    void setFixed(bool f) { _fixed = f; }
    bool marginalized() const { return _marginalized; }
    void setMarginalized(bool m) { _marginalized = m; }
    int dimension() const { return _dimension; }
    int hessianIndex() const { return _hessianIndex; }
    int estimateDimension() const { return (int)_estimate.size(); }
    bool getEstimateData(number_t* est) const { for (size_t i = 0; i < _estimate.size(); ++i) est[i] = _estimate[i]; return true; }
    bool setEstimateData(const number_t* est) { for (size_t i = 0; i < _estimate.size(); ++i) _estimate[i] = est[i]; return true; }
    const std::vector<number_t>& estimateVector() const { return _estimate; }
    int typeCode() const { return _type; }
   protected:
    friend class SparseOptimizer;
    int _type, _dimension, _id = -1, _hessianIndex = -1, _index = -1;
    bool _fixed = false, _marginalized = false;
    std::vector<number_t> _estimate;   // boundary layout of include/g2ocu.h
  };
  class Edge {
   public:
    Edge(int type, int dim, int measDim, int paramDim) : _type(type), _dimension(dim), _measurement(measDim, 0.0), _information((size_t)dim * dim, 0.0), _param(paramDim, 0.0) {
      for (int i = 0; i < dim; ++i) _information[(size_t)i * dim + i] = 1.0;
    }
    virtual ~Edge() {}
    void setVertex(size_t i, Vertex* v) { _vertices[i] = v; }
    Vertex* vertex(size_t i) const { return _vertices[i]; }
    int dimension() const { return _dimension; }
    int level() const { return _level; }
    void setLevel(int l) { _level = l; }
    void setRobustKernel(RobustKernel* k) { _kernel.reset(k); }   // the edge owns its kernel (optimizable_graph.cpp:182-188)
    RobustKernel* robustKernel() const { return _kernel.get(); }
    void setMeasurementData(const number_t* m) { for (size_t i = 0; i < _measurement.size(); ++i) _measurement[i] = m[i]; }
    const number_t* measurementData() const { return _measurement.data(); }
    number_t* informationData() { return _information.data(); }                   // column-major E x E
    void setInformationDiagonal(const number_t* d) { for (int i = 0; i < _dimension; ++i) _information[(size_t)i * _dimension + i] = d[i]; }
    void setParameterData(const number_t* p) { for (size_t i = 0; i < _param.size(); ++i) _param[i] = p[i]; }
    const number_t* parameterData() const { return _param.data(); }
    int typeCode() const { return _type; }
   protected:
    friend class SparseOptimizer;
    int _type, _dimension, _level = 0;
    Vertex* _vertices[2] = {nullptr, nullptr};
    std::vector<number_t> _measurement, _information, _param;
    std::unique_ptr<RobustKernel> _kernel;
  };
};

// concrete types of the reference that the backend supports (anything else cannot even be constructed here;
// the real-g2o adapter rejects other classes at init)
struct VertexSE2 : OptimizableGraph::Vertex { VertexSE2() : Vertex(G2OCU_VERTEX_SE2, 3, 3) {}
  void setEstimate(number_t x, number_t y, number_t theta) { _estimate = {x, y, theta}; } };
struct VertexPointXY : OptimizableGraph::Vertex { VertexPointXY() : Vertex(G2OCU_VERTEX_POINT_XY, 2, 2) {}
  void setEstimate(number_t x, number_t y) { _estimate = {x, y}; } };
struct VertexSE3 : OptimizableGraph::Vertex { VertexSE3() : Vertex(G2OCU_VERTEX_SE3, 12, 6) { _estimate = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0}; }
  // Isometry3 as rotation (column-major) + translation
  void setEstimate(const number_t R[9], const number_t t[3]) { for (int i = 0; i < 9; ++i) _estimate[i] = R[i]; for (int i = 0; i < 3; ++i) _estimate[9 + i] = t[i]; } };
struct VertexSE3Expmap : OptimizableGraph::Vertex { VertexSE3Expmap() : Vertex(G2OCU_VERTEX_SE3_EXPMAP, 7, 6) { _estimate = {0, 0, 0, 0, 0, 0, 1}; }
  void setEstimate(const number_t t[3], const number_t qxyzw[4]) { for (int i = 0; i < 3; ++i) _estimate[i] = t[i]; for (int i = 0; i < 4; ++i) _estimate[3 + i] = qxyzw[i]; } };
struct VertexSBAPointXYZ : OptimizableGraph::Vertex { VertexSBAPointXYZ() : Vertex(G2OCU_VERTEX_POINT_XYZ, 3, 3) {}
  void setEstimate(number_t x, number_t y, number_t z) { _estimate = {x, y, z}; } };
struct VertexCameraBAL : OptimizableGraph::Vertex { VertexCameraBAL() : Vertex(G2OCU_VERTEX_CAM_BAL, 9, 9) {}
  void setEstimate(const number_t c[9]) { for (int i = 0; i < 9; ++i) _estimate[i] = c[i]; } };
struct VertexPointBAL : OptimizableGraph::Vertex { VertexPointBAL() : Vertex(G2OCU_VERTEX_POINT_BAL, 3, 3) {}
  void setEstimate(number_t x, number_t y, number_t z) { _estimate = {x, y, z}; } };

struct EdgeSE2 : OptimizableGraph::Edge { EdgeSE2() : Edge(G2OCU_EDGE_SE2, 3, 3, 0) {}
  void setMeasurement(number_t x, number_t y, number_t theta) { _measurement = {x, y, theta}; } };
struct EdgeSE2PointXY : OptimizableGraph::Edge { EdgeSE2PointXY() : Edge(G2OCU_EDGE_SE2_POINT_XY, 2, 2, 0) {}
  void setMeasurement(number_t x, number_t y) { _measurement = {x, y}; } };
struct EdgeSE3 : OptimizableGraph::Edge { EdgeSE3() : Edge(G2OCU_EDGE_SE3, 6, 12, 0) { _measurement = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0}; }
  void setMeasurement(const number_t R[9], const number_t t[3]) { for (int i = 0; i < 9; ++i) _measurement[i] = R[i]; for (int i = 0; i < 3; ++i) _measurement[9 + i] = t[i]; } };
struct EdgeSE3Expmap : OptimizableGraph::Edge { EdgeSE3Expmap() : Edge(G2OCU_EDGE_SE3_EXPMAP, 6, 7, 0) { _measurement = {0, 0, 0, 0, 0, 0, 1}; } };
struct EdgeProjectXYZ2UV : OptimizableGraph::Edge { EdgeProjectXYZ2UV() : Edge(G2OCU_EDGE_PROJECT_XYZ2UV, 2, 2, 3) {}
  void setMeasurement(number_t u, number_t v) { _measurement = {u, v}; }
  void setCameraParameters(number_t f, number_t cx, number_t cy) { _param = {f, cx, cy}; } };      // CameraParameters, types_six_dof_expmap.h:45-76
struct EdgeSE3ProjectXYZ : OptimizableGraph::Edge { EdgeSE3ProjectXYZ() : Edge(G2OCU_EDGE_SE3_PROJECT_XYZ, 2, 2, 4) {}
  void setMeasurement(number_t u, number_t v) { _measurement = {u, v}; }
  void setIntrinsics(number_t fx, number_t fy, number_t cx, number_t cy) { _param = {fx, fy, cx, cy}; } };   // fx, fy, cx, cy members, :227
struct EdgeObservationBAL : OptimizableGraph::Edge { EdgeObservationBAL() : Edge(G2OCU_EDGE_BAL, 2, 2, 0) {}
  void setMeasurement(number_t u, number_t v) { _measurement = {u, v}; } };

// ---------------------------------------------------------------------------------------------------------------
class SparseOptimizer;

// core/solver.h:44-155
class Solver {
 public:
  virtual ~Solver() {}
  virtual bool init(SparseOptimizer* optimizer, bool online = false) = 0;
  virtual bool buildStructure(bool zeroBlocks = false) = 0;
  virtual bool buildSystem() = 0;
  virtual bool solve() = 0;
  virtual bool setLambda(number_t lambda, bool backup = false) = 0;
  virtual void restoreDiagonal() = 0;
  virtual bool supportsSchur() { return false; }
  virtual bool schur() = 0;
  virtual void setSchur(bool s) = 0;
  virtual number_t* x() = 0;
  virtual number_t* b() = 0;
  virtual size_t vectorSize() const = 0;
  virtual void multiplyHessian(number_t* dest, const number_t* src) const = 0;   // BlockSolverBase, block_solver.h:87-95
  SparseOptimizer* optimizer() const { return _optimizer; }
 protected:
  SparseOptimizer* _optimizer = nullptr;
};

// core/optimization_algorithm.h:46-110
class OptimizationAlgorithm {
 public:
  enum SolverResult { Terminate = 2, OK = 1, Fail = -1 };
  virtual ~OptimizationAlgorithm() {}
  virtual bool init(bool online = false) = 0;
  virtual SolverResult solve(int iteration, bool online = false) = 0;
  virtual void printVerbose(std::ostream&) const {}
  const SparseOptimizer* optimizer() const { return _optimizer; }
  void setOptimizer(SparseOptimizer* o) { _optimizer = o; }
  // PropertyMap subset: "initialLambda", "maxTrialsAfterFailure" (optimization_algorithm_levenberg.cpp:48-49)
  virtual bool updatePropertiesFromString(const std::string&) { return true; }
 protected:
  SparseOptimizer* _optimizer = nullptr;
};

// SparseOptimizer over the device handle.  The per-edge and per-vertex loops of the reference's SparseOptimizer
// (computeActiveErrors, activeRobustChi2, update, push/pop) run on the GPU; estimates are written back to the host
// vertices after every optimize() (and on demand through pullEstimates()).
class SparseOptimizer : public OptimizableGraph {
 public:
  SparseOptimizer();
  ~SparseOptimizer();
  bool addVertex(Vertex* v);                 // takes ownership, like the reference; false on duplicate id
  bool addEdge(Edge* e);                     // false when a vertex is missing or of the wrong type
  Vertex* vertex(int id) const;
  const std::vector<Vertex*>& vertexList() const { return _vertexList; }
  const std::vector<Edge*>& edgeList() const { return _edgeList; }
  void setAlgorithm(OptimizationAlgorithm* algorithm);   // takes ownership (sparse_optimizer.cpp:57-61)
  OptimizationAlgorithm* algorithm() const { return _algorithm; }
  bool initializeOptimization(int level = 0);
  int optimize(int iterations, bool online = false);
  void computeActiveErrors();
  number_t activeChi2() const;
  number_t activeRobustChi2() const;
  void update(const number_t* update);
  void push();
  void pop();
  void discardTop();
  // computeMarginals(spinv, blockIndices) (sparse_optimizer.h:129, sparse_optimizer.cpp:594-596): block i of `spinv` is block
  // (blockIndices[i].first, .second) of the inverse of Hpp in hessian-index units, column-major; false as the reference's solvePattern
  bool computeMarginals(std::vector<std::vector<number_t> >& spinv, const std::vector<std::pair<int, int> >& blockIndices);
  void clear();
  void setVerbose(bool v) { _verbose = v; }
  bool verbose() const { return _verbose; }
  void setComputeBatchStatistics(bool b) { _computeBatchStatistics = b; }
  const BatchStatisticsContainer& batchStatistics() const { return _batchStatistics; }
  size_t activeEdgeCount() const { return _numActiveEdges; }
  size_t activeVertexCount() const { return _numActiveVertices; }
  size_t indexMappingSize() const { return _ivMapSize; }
  void pullEstimates();                      // device -> host vertices
  g2ocu_solver* handle() const { return _handle; }
  G2OBatchStatistics* currentStats() { return _currentStats; }
 private:
  bool uploadGraph();
  std::vector<Vertex*> _vertexList; std::vector<Edge*> _edgeList; std::unordered_map<int, Vertex*> _vertexById;
  OptimizationAlgorithm* _algorithm = nullptr;
  g2ocu_solver* _handle = nullptr;
  bool _graphDirty = true, _verbose = false, _computeBatchStatistics = false;
  size_t _numActiveEdges = 0, _numActiveVertices = 0, _ivMapSize = 0;
  BatchStatisticsContainer _batchStatistics; G2OBatchStatistics* _currentStats = nullptr;
};

// BlockSolver<BlockSolverTraits<P,L>> replacement; P, L = -1 means variable (BlockSolverX)
template <int P, int L> class CudaBlockSolver : public Solver {
 public:
  static const int PoseDim = P, LandmarkDim = L;
  // which LinearSolver the reference BlockSolver would own (block_solver.h:124): LinearSolverPCG (default) or LinearSolverDense
  explicit CudaBlockSolver(int linearSolverKind = G2OCU_LINEAR_PCG) : _linearKind(linearSolverKind) {}
  bool init(SparseOptimizer* optimizer, bool online = false) override;
  bool buildStructure(bool zeroBlocks = false) override;
  bool buildSystem() override;
  bool solve() override;
  bool setLambda(number_t lambda, bool backup = false) override;
  void restoreDiagonal() override;
  bool supportsSchur() override { return true; }
  bool schur() override { return _doSchur; }
  void setSchur(bool s) override { _doSchur = s; }
  number_t* x() override;
  number_t* b() override;
  size_t vectorSize() const override;
  void multiplyHessian(number_t* dest, const number_t* src) const override;
 private:
  bool _doSchur = true; int _linearKind = G2OCU_LINEAR_PCG;
  std::vector<number_t> _x, _b;
};
typedef CudaBlockSolver<-1, -1> CudaBlockSolverX;
typedef CudaBlockSolver<6, 3> CudaBlockSolver_6_3;
typedef CudaBlockSolver<3, 2> CudaBlockSolver_3_2;
typedef CudaBlockSolver<9, 3> CudaBlockSolver_9_3;

// OptimizationAlgorithmLevenberg / GaussNewton whose solve(iteration) runs the whole iteration (all LM trials) on the device
class OptimizationAlgorithmWithHessianCuda : public OptimizationAlgorithm {
 public:
  explicit OptimizationAlgorithmWithHessianCuda(std::unique_ptr<Solver> solver, int algorithm) : _solver(std::move(solver)), _algorithmCode(algorithm) {}
  bool init(bool online = false) override;
  SolverResult solve(int iteration, bool online = false) override;
  Solver& solver() { return *_solver; }
  bool updatePropertiesFromString(const std::string& s) override;
 protected:
  std::unique_ptr<Solver> _solver; int _algorithmCode;
  number_t _currentLambda = -1; int _levenbergIterations = 0;
};
class OptimizationAlgorithmLevenberg : public OptimizationAlgorithmWithHessianCuda {
 public:
  explicit OptimizationAlgorithmLevenberg(std::unique_ptr<Solver> solver) : OptimizationAlgorithmWithHessianCuda(std::move(solver), G2OCU_ALGORITHM_LM) {}
  number_t currentLambda() const { return _currentLambda; }            // optimization_algorithm_levenberg.h:56
  int levenbergIteration() const { return _levenbergIterations; }      // :70
  void setMaxTrialsAfterFailure(int max_trials);
  void setUserLambdaInit(number_t lambda);
  void printVerbose(std::ostream& os) const override;                  // optimization_algorithm_levenberg.cpp:196-202
 private:
  int _maxTrials = 10; number_t _userLambdaInit = 0;
  friend class OptimizationAlgorithmWithHessianCuda;
};
class OptimizationAlgorithmGaussNewton : public OptimizationAlgorithmWithHessianCuda {
 public:
  explicit OptimizationAlgorithmGaussNewton(std::unique_ptr<Solver> solver) : OptimizationAlgorithmWithHessianCuda(std::move(solver), G2OCU_ALGORITHM_GN) {}
  void printVerbose(std::ostream& os) const override { os << "\t schur= " << _solver->schur(); }
};

// core/optimization_algorithm_dogleg.h:43-97: Powell's dogleg; the trial loop, the step vectors and their dot products stay on the device
class OptimizationAlgorithmDogleg : public OptimizationAlgorithmWithHessianCuda {
 public:
  enum { STEP_UNDEFINED, STEP_SD, STEP_GN, STEP_DL };
  explicit OptimizationAlgorithmDogleg(std::unique_ptr<Solver> solver) : OptimizationAlgorithmWithHessianCuda(std::move(solver), G2OCU_ALGORITHM_DOGLEG) {}
  SolverResult solve(int iteration, bool online = false) override;
  int lastStep() const { return _lastStep; }                           // :64
  number_t trustRegion() const { return _delta; }                      // :66
  static const char* stepType2Str(int stepType);                       // optimization_algorithm_dogleg.cpp:209-217
  void printVerbose(std::ostream& os) const override;                  // :199-207
  // properties "initialDelta", "maxTrialsAfterFailure", "initialLambda", "lambdaFactor" (:44-47)
  bool updatePropertiesFromString(const std::string& s) override;
 private:
  number_t _delta = 1e4; int _lastStep = STEP_UNDEFINED, _lastNumTries = 0; bool _wasPDInAllIterations = true;
};

// ---------------------------------------------------------------------------------------------------------------
// core/optimization_algorithm_factory.h:52-137
class AbstractOptimizationAlgorithmCreator {
 public:
  explicit AbstractOptimizationAlgorithmCreator(const OptimizationAlgorithmProperty& p) : _property(p) {}
  virtual ~AbstractOptimizationAlgorithmCreator() {}
  virtual OptimizationAlgorithm* construct() = 0;
  const OptimizationAlgorithmProperty& property() const { return _property; }
 protected:
  OptimizationAlgorithmProperty _property;
};
class OptimizationAlgorithmFactory {
 public:
  typedef std::list<std::unique_ptr<AbstractOptimizationAlgorithmCreator>> CreatorList;
  static OptimizationAlgorithmFactory* instance();
  void registerSolver(const std::shared_ptr<AbstractOptimizationAlgorithmCreator>& c);
  OptimizationAlgorithm* construct(const std::string& tag, OptimizationAlgorithmProperty& solverProperty) const;   // nullptr if unknown
  void listSolvers(std::ostream& os) const;
  const std::vector<std::shared_ptr<AbstractOptimizationAlgorithmCreator>>& creatorList() const { return _creator; }
 private:
  std::vector<std::shared_ptr<AbstractOptimizationAlgorithmCreator>> _creator;
};
class RegisterOptimizationAlgorithmProxy {
 public:
  explicit RegisterOptimizationAlgorithmProxy(AbstractOptimizationAlgorithmCreator* c) { OptimizationAlgorithmFactory::instance()->registerSolver(std::shared_ptr<AbstractOptimizationAlgorithmCreator>(c)); }
};
#define G2O_REGISTER_OPTIMIZATION_LIBRARY(libraryname) extern "C" void g2o_optimization_library_##libraryname(void) {}
#define G2O_REGISTER_OPTIMIZATION_ALGORITHM(optimizername, instance) \
  extern "C" void g2o_optimization_algorithm_##optimizername(void) {}  \
  static g2o::RegisterOptimizationAlgorithmProxy g_optimization_algorithm_proxy_##optimizername(instance);

}  // namespace g2o
