// Implementation of the host-side mirror (g2o_mirror.hpp) over the C ABI.
#include "g2o_mirror.hpp"

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <iomanip>
#include <sstream>

namespace g2o {

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static bool check(g2ocu_solver* h, int rc, const char* where) {
  if (rc == G2OCU_OK) return true;
  std::cerr << where << ": " << g2ocu_last_error(h) << std::endl;      // the reference reports through std::cerr + bool/enum returns
  return false;
}

// ---------------------------------------------------------------------------------------------------------------
SparseOptimizer::SparseOptimizer() {
  g2ocu_config cfg; g2ocu_default_config(&cfg);
  if (g2ocu_create(&cfg, &_handle) != G2OCU_OK) { std::cerr << "SparseOptimizer: " << g2ocu_last_error(nullptr) << std::endl; _handle = nullptr; }
}
SparseOptimizer::~SparseOptimizer() {
  delete _algorithm;                                  // sparse_optimizer.cpp:57-61
  for (Edge* e : _edgeList) delete e;
  for (Vertex* v : _vertexList) delete v;
  if (_handle) g2ocu_destroy(_handle);
}
bool SparseOptimizer::addVertex(Vertex* v) {
  if (!v || _vertexById.count(v->id())) return false;                  // hyper_graph.cpp: duplicate ids are refused
  v->_index = (int)_vertexList.size();
  _vertexList.push_back(v); _vertexById[v->id()] = v; _graphDirty = true;
  return true;
}
bool SparseOptimizer::addEdge(Edge* e) {
  if (!e || !e->vertex(0) || !e->vertex(1)) return false;
  for (int i = 0; i < 2; ++i) { auto it = _vertexById.find(e->vertex(i)->id()); if (it == _vertexById.end() || it->second != e->vertex(i)) return false; }
  _edgeList.push_back(e); _graphDirty = true;                          // position in _edgeList == internalId (optimizable_graph.cpp:267-292)
  return true;
}
OptimizableGraph::Vertex* SparseOptimizer::vertex(int id) const { auto it = _vertexById.find(id); return it == _vertexById.end() ? nullptr : it->second; }
void SparseOptimizer::setAlgorithm(OptimizationAlgorithm* algorithm) {
  if (_algorithm && _algorithm != algorithm) delete _algorithm;
  _algorithm = algorithm;
  if (_algorithm) _algorithm->setOptimizer(this);
}
void SparseOptimizer::clear() {
  for (Edge* e : _edgeList) delete e;
  for (Vertex* v : _vertexList) delete v;
  _edgeList.clear(); _vertexList.clear(); _vertexById.clear(); _graphDirty = true; _ivMapSize = 0; _numActiveEdges = 0;
}

bool SparseOptimizer::uploadGraph() {
  if (!_handle) return false;
  std::vector<int32_t> vId, vType, eType, eV0, eV1, eLevel, eKernel;
  std::vector<uint8_t> vFixed, vMarg; std::vector<double> vEst, eMeas, eInfo, eDelta, ePrm;
  for (Vertex* v : _vertexList) {
    vId.push_back(v->id()); vType.push_back(v->typeCode()); vFixed.push_back(v->fixed()); vMarg.push_back(v->marginalized());
    vEst.insert(vEst.end(), v->_estimate.begin(), v->_estimate.end());
  }
  for (Edge* e : _edgeList) {
    eType.push_back(e->typeCode()); eV0.push_back(e->vertex(0)->_index); eV1.push_back(e->vertex(1)->_index); eLevel.push_back(e->level());
    eMeas.insert(eMeas.end(), e->_measurement.begin(), e->_measurement.end());
    eInfo.insert(eInfo.end(), e->_information.begin(), e->_information.end());
    ePrm.insert(ePrm.end(), e->_param.begin(), e->_param.end());
    eKernel.push_back(e->robustKernel() ? e->robustKernel()->code() : 0); eDelta.push_back(e->robustKernel() ? e->robustKernel()->delta() : 1.0);
  }
  g2ocu_graph g;
  g.n_vertices = (int32_t)_vertexList.size(); g.v_id = vId.data(); g.v_type = vType.data(); g.v_fixed = vFixed.data(); g.v_marginalized = vMarg.data(); g.v_estimate = vEst.data();
  g.n_edges = (int32_t)_edgeList.size(); g.e_type = eType.data(); g.e_v0 = eV0.data(); g.e_v1 = eV1.data(); g.e_level = eLevel.data();
  g.e_measurement = eMeas.data(); g.e_information = eInfo.data(); g.e_kernel = eKernel.data(); g.e_kernel_delta = eDelta.data(); g.e_param = ePrm.data();
  if (!check(_handle, g2ocu_set_graph(_handle, &g), "SparseOptimizer::uploadGraph")) return false;
  _graphDirty = false;
  return true;
}

#define G2O_PTHING(s) #s << "= " << (st.s) << "\t "
std::ostream& operator<<(std::ostream& os, const G2OBatchStatistics& st) {   // batch_stats.cpp:48-83, same fields in the same order
  os << G2O_PTHING(iteration) << G2O_PTHING(numVertices) << G2O_PTHING(numEdges) << G2O_PTHING(chi2);
  os << G2O_PTHING(timeLinearSolution) << G2O_PTHING(iterationsLinearSolver) << G2O_PTHING(timeQrDecomposition) << G2O_PTHING(timeResiduals) << G2O_PTHING(timeLinearize)
     << G2O_PTHING(timeQuadraticForm);
  os << G2O_PTHING(timeSchurComplement);
  os << G2O_PTHING(timeSymbolicDecomposition) << G2O_PTHING(timeNumericDecomposition);
  os << G2O_PTHING(timeUpdate) << G2O_PTHING(timeIteration);
  os << G2O_PTHING(levenbergIterations) << G2O_PTHING(timeLinearSolver);
  os << G2O_PTHING(hessianDimension) << G2O_PTHING(hessianPoseDimension) << G2O_PTHING(hessianLandmarkDimension) << G2O_PTHING(choleskyNNZ) << G2O_PTHING(timeMarginals);
  return os;
}
#undef G2O_PTHING

bool SparseOptimizer::initializeOptimization(int level) {
  if (_edgeList.empty()) { std::cerr << "SparseOptimizer::initializeOptimization: Attempt to initialize an empty graph" << std::endl; return false; }
  if (!uploadGraph()) return false;                                     // vertex estimates / fixed flags may have changed since the last call
  if (!check(_handle, g2ocu_initialize_optimization(_handle, level), "SparseOptimizer::initializeOptimization")) return false;
  std::vector<int32_t> hidx(_vertexList.size());
  g2ocu_get_i32(_handle, "hessian_index", hidx.data(), (int64_t)hidx.size());
  for (size_t i = 0; i < _vertexList.size(); ++i) _vertexList[i]->_hessianIndex = hidx[i];
  _numActiveEdges = (size_t)g2ocu_get_i32(_handle, "active_edges", nullptr, 0);
  _numActiveVertices = (size_t)g2ocu_get_i32(_handle, "active_vertices", nullptr, 0);
  _ivMapSize = (size_t)g2ocu_get_i32(_handle, "index_mapping", nullptr, 0);
  return _ivMapSize > 0;
}

int SparseOptimizer::optimize(int iterations, bool online) {
  if (_ivMapSize == 0) { std::cerr << "SparseOptimizer::optimize: 0 vertices to optimize, maybe forgot to call initializeOptimization()" << std::endl; return -1; }
  if (!_algorithm) { std::cerr << "SparseOptimizer::optimize: no algorithm set" << std::endl; return -1; }
  int cjIterations = 0; double cumTime = 0;
  bool ok = _algorithm->init(online);
  if (!ok) { std::cerr << "SparseOptimizer::optimize Error while initializing" << std::endl; return -1; }
  _batchStatistics.clear();
  if (_computeBatchStatistics) _batchStatistics.resize(iterations);
  OptimizationAlgorithm::SolverResult result = OptimizationAlgorithm::OK;
  for (int i = 0; i < iterations && ok; i++) {
    G2OBatchStatistics local; _currentStats = _computeBatchStatistics ? &_batchStatistics[i] : &local;
    _currentStats->iteration = i; _currentStats->numEdges = (int)_numActiveEdges; _currentStats->numVertices = (int)_numActiveVertices;   // sparse_optimizer.cpp:399-403
    const double ts = now();
    result = _algorithm->solve(i, online);
    ok = (result == OptimizationAlgorithm::OK);
    if (verbose()) {
      const double dts = now() - ts; cumTime += dts;
      std::cerr << "iteration= " << i << "\t chi2= " << std::fixed << activeRobustChi2() << "\t time= " << dts << "\t cumTime= " << cumTime << "\t edges= " << _numActiveEdges;
      _algorithm->printVerbose(std::cerr);
      std::cerr << std::endl;
    }
    ++cjIterations;
  }
  _currentStats = nullptr;
  pullEstimates();
  if (result == OptimizationAlgorithm::Fail) return 0;
  return cjIterations;
}

void SparseOptimizer::pullEstimates() {
  size_t total = 0; for (Vertex* v : _vertexList) total += v->_estimate.size();
  std::vector<double> est(total);
  if (!check(_handle, g2ocu_get_estimates(_handle, est.data()), "SparseOptimizer::pullEstimates")) return;
  size_t o = 0; for (Vertex* v : _vertexList) { std::copy(est.begin() + o, est.begin() + o + v->_estimate.size(), v->_estimate.begin()); o += v->_estimate.size(); }
}
void SparseOptimizer::computeActiveErrors() { check(_handle, g2ocu_compute_active_errors(_handle), "SparseOptimizer::computeActiveErrors"); }
number_t SparseOptimizer::activeChi2() const { double v = std::numeric_limits<double>::quiet_NaN(); check(_handle, g2ocu_active_chi2(_handle, &v), "SparseOptimizer::activeChi2"); return v; }
number_t SparseOptimizer::activeRobustChi2() const { double v = std::numeric_limits<double>::quiet_NaN(); check(_handle, g2ocu_active_robust_chi2(_handle, &v), "SparseOptimizer::activeRobustChi2"); return v; }
void SparseOptimizer::update(const number_t* u) { check(_handle, g2ocu_update(_handle, u), "SparseOptimizer::update"); }
void SparseOptimizer::push() { check(_handle, g2ocu_push(_handle), "SparseOptimizer::push"); }
void SparseOptimizer::pop() { check(_handle, g2ocu_pop(_handle), "SparseOptimizer::pop"); }
void SparseOptimizer::discardTop() { check(_handle, g2ocu_discard_top(_handle), "SparseOptimizer::discardTop"); }
bool SparseOptimizer::computeMarginals(std::vector<std::vector<number_t> >& spinv, const std::vector<std::pair<int, int> >& blockIndices) {
  spinv.clear();
  if (!_handle) return false;
  const int64_t nb = g2ocu_get_i32(_handle, "pose_block_indices", nullptr, 0);   // cumulative block ends of Hpp (every vertex when no point is marginalized)
  if (nb <= 0) return false;
  std::vector<int32_t> ends((size_t)nb);
  g2ocu_get_i32(_handle, "pose_block_indices", ends.data(), nb);
  std::vector<int32_t> rows, cols; std::vector<size_t> sizes; size_t total = 0;
  for (const auto& rc : blockIndices) {
    if (rc.first < 0 || rc.first >= nb || rc.second < 0 || rc.second >= nb) { std::cerr << "SparseOptimizer::computeMarginals: block index outside Hpp" << std::endl; return false; }
    rows.push_back(rc.first); cols.push_back(rc.second);
    sizes.push_back((size_t)(ends[rc.first] - (rc.first ? ends[rc.first - 1] : 0)) * (size_t)(ends[rc.second] - (rc.second ? ends[rc.second - 1] : 0)));
    total += sizes.back();
  }
  std::vector<number_t> out(total ? total : 1);
  int32_t computed = 0;
  if (!check(_handle, g2ocu_compute_marginals(_handle, (int32_t)blockIndices.size(), rows.data(), cols.data(), out.data(), &computed), "SparseOptimizer::computeMarginals") || !computed) return false;
  size_t off = 0;
  for (size_t i = 0; i < blockIndices.size(); ++i) { spinv.emplace_back(out.begin() + off, out.begin() + off + sizes[i]); off += sizes[i]; }
  return true;
}

// ---------------------------------------------------------------------------------------------------------------
template <int P, int L> bool CudaBlockSolver<P, L>::init(SparseOptimizer* optimizer, bool online) {
  _optimizer = optimizer;
  if (!optimizer || !check(optimizer->handle(), g2ocu_set_property(optimizer->handle(), "linearSolver", _linearKind), "CudaBlockSolver::init")) return false;
  g2ocu_set_property(optimizer->handle(), "poseDim", P); g2ocu_set_property(optimizer->handle(), "landmarkDim", L);   // BlockSolverTraits<P,L>: fixed block sizes
  return check(optimizer->handle(), g2ocu_init(optimizer->handle(), online), "CudaBlockSolver::init");
}
template <int P, int L> bool CudaBlockSolver<P, L>::buildStructure(bool) {
  g2ocu_solver* h = _optimizer->handle();
  if (!check(h, g2ocu_build_structure(h), "CudaBlockSolver::buildStructure")) return false;
  int32_t dims[4] = {0, 0, 0, 0}; g2ocu_get_i32(h, "dims", dims, 4);
  // BlockSolverTraits<P,L> fixes the block sizes; the graph must agree (isSolverSuitable, optimizable_graph.cpp:835-856)
  if (P > 0 && dims[0] > 0 && dims[2] != dims[0] * P) { std::cerr << "CudaBlockSolver<" << P << "," << L << ">::buildStructure: pose dimension of the graph does not match the solver" << std::endl; return false; }
  if (L > 0 && dims[1] > 0 && dims[3] != dims[1] * L) { std::cerr << "CudaBlockSolver<" << P << "," << L << ">::buildStructure: landmark dimension of the graph does not match the solver" << std::endl; return false; }
  _x.assign((size_t)dims[2] + dims[3], 0.0); _b.assign(_x.size(), 0.0);
  return true;
}
template <int P, int L> bool CudaBlockSolver<P, L>::buildSystem() { return check(_optimizer->handle(), g2ocu_build_system(_optimizer->handle()), "CudaBlockSolver::buildSystem"); }
template <int P, int L> bool CudaBlockSolver<P, L>::solve() {
  int32_t ok = 0;
  if (!check(_optimizer->handle(), g2ocu_solve(_optimizer->handle(), &ok), "CudaBlockSolver::solve")) return false;
  return ok != 0;
}
template <int P, int L> bool CudaBlockSolver<P, L>::setLambda(number_t lambda, bool backup) { return check(_optimizer->handle(), g2ocu_set_lambda(_optimizer->handle(), lambda, backup), "CudaBlockSolver::setLambda"); }
template <int P, int L> void CudaBlockSolver<P, L>::restoreDiagonal() { check(_optimizer->handle(), g2ocu_restore_diagonal(_optimizer->handle()), "CudaBlockSolver::restoreDiagonal"); }
template <int P, int L> number_t* CudaBlockSolver<P, L>::x() { g2ocu_get_f64(_optimizer->handle(), "x", _x.data(), (int64_t)_x.size()); return _x.data(); }
template <int P, int L> number_t* CudaBlockSolver<P, L>::b() { g2ocu_get_f64(_optimizer->handle(), "b", _b.data(), (int64_t)_b.size()); return _b.data(); }
template <int P, int L> size_t CudaBlockSolver<P, L>::vectorSize() const { return (size_t)g2ocu_vector_size(_optimizer->handle()); }
template <int P, int L> void CudaBlockSolver<P, L>::multiplyHessian(number_t* dest, const number_t* src) const { check(_optimizer->handle(), g2ocu_multiply_hessian(_optimizer->handle(), dest, src), "CudaBlockSolver::multiplyHessian"); }
template class CudaBlockSolver<-1, -1>;
template class CudaBlockSolver<6, 3>;
template class CudaBlockSolver<3, 2>;
template class CudaBlockSolver<9, 3>;

// ---------------------------------------------------------------------------------------------------------------
bool OptimizationAlgorithmWithHessianCuda::init(bool online) {
  if (!_optimizer) return false;
  return _solver->init(_optimizer, online);            // Schur is switched on iff an active vertex is marginalized (with_hessian.cpp:48-66), inside the backend
}
OptimizationAlgorithm::SolverResult OptimizationAlgorithmWithHessianCuda::solve(int iteration, bool) {
  g2ocu_solver* h = _optimizer->handle();
  if (iteration == 0) {                                 // "built up the CCS structure, here due to easy time measure" (levenberg.cpp:63-69)
    if (!_solver->buildStructure()) { std::cerr << "OptimizationAlgorithm::solve: Failure while building CCS structure" << std::endl; return Fail; }
  }
  g2ocu_iteration_stats st;
  if (!check(h, g2ocu_solver_iteration(h, _algorithmCode, iteration, &st), "OptimizationAlgorithm::solve")) return Fail;
  _currentLambda = st.lambda; _levenbergIterations = st.levenberg_iterations;
  if (G2OBatchStatistics* gs = _optimizer->currentStats()) {
    gs->chi2 = st.chi2; gs->timeResiduals = st.time_residuals; gs->timeQuadraticForm = st.time_quadratic_form; gs->levenbergIterations = st.levenberg_iterations;
    gs->timeSchurComplement = st.time_schur_complement; gs->timeLinearSolver = st.time_linear_solver; gs->timeLinearSolution = st.time_linear_solution;
    gs->iterationsLinearSolver = st.iterations_linear_solver; gs->timeUpdate = st.time_update; gs->timeIteration = st.time_iteration;
    gs->hessianPoseDimension = (size_t)st.hessian_pose_dimension; gs->hessianLandmarkDimension = (size_t)st.hessian_landmark_dimension;
    gs->hessianDimension = gs->hessianPoseDimension + gs->hessianLandmarkDimension;
  }
  return st.result == G2OCU_RESULT_OK ? OK : (st.result == G2OCU_RESULT_TERMINATE ? Terminate : Fail);
}
bool OptimizationAlgorithmWithHessianCuda::updatePropertiesFromString(const std::string& s) {   // "k=v,k=v" as g2o -solverProperties (g2o.cpp:229-237)
  std::stringstream ss(s); std::string kv; bool ok = true;
  while (std::getline(ss, kv, ',')) {
    const size_t eq = kv.find('='); if (eq == std::string::npos) { ok = false; continue; }
    ok = check(_optimizer->handle(), g2ocu_set_property(_optimizer->handle(), kv.substr(0, eq).c_str(), std::atof(kv.substr(eq + 1).c_str())), "updatePropertiesFromString") && ok;
  }
  return ok;
}
void OptimizationAlgorithmLevenberg::setMaxTrialsAfterFailure(int max_trials) { _maxTrials = max_trials; if (_optimizer) g2ocu_set_property(_optimizer->handle(), "maxTrialsAfterFailure", max_trials); }
void OptimizationAlgorithmLevenberg::setUserLambdaInit(number_t lambda) { _userLambdaInit = lambda; if (_optimizer) g2ocu_set_property(_optimizer->handle(), "initialLambda", lambda); }
void OptimizationAlgorithmLevenberg::printVerbose(std::ostream& os) const {
  os << "\t schur= " << _solver->schur() << "\t lambda= " << std::fixed << _currentLambda << "\t levenbergIter= " << _levenbergIterations;
}

OptimizationAlgorithm::SolverResult OptimizationAlgorithmDogleg::solve(int iteration, bool online) {
  const SolverResult r = OptimizationAlgorithmWithHessianCuda::solve(iteration, online);
  double d[5] = {_delta, (double)_lastStep, (double)_lastNumTries, _currentLambda, 1.0};
  if (g2ocu_get_f64(_optimizer->handle(), "dogleg", d, 5) == 5) { _delta = d[0]; _lastStep = (int)d[1]; _lastNumTries = (int)d[2]; _currentLambda = d[3]; _wasPDInAllIterations = d[4] != 0.0; }
  return r;
}
const char* OptimizationAlgorithmDogleg::stepType2Str(int stepType) {
  switch (stepType) { case STEP_SD: return "Descent"; case STEP_GN: return "GN"; case STEP_DL: return "Dogleg"; default: return "Undefined"; }
}
void OptimizationAlgorithmDogleg::printVerbose(std::ostream& os) const {
  os << "\t Delta= " << _delta << "\t step= " << stepType2Str(_lastStep) << "\t tries= " << _lastNumTries;
  if (!_wasPDInAllIterations) os << "\t lambda= " << _currentLambda;
}
bool OptimizationAlgorithmDogleg::updatePropertiesFromString(const std::string& s) {   // the Dogleg property names map onto the backend's dogleg* properties
  static const std::map<std::string, std::string> names{{"initialDelta", "doglegInitialDelta"}, {"maxTrialsAfterFailure", "doglegMaxTrialsAfterFailure"},
                                                        {"initialLambda", "doglegInitialLambda"}, {"lambdaFactor", "doglegLambdaFactor"}};
  std::stringstream ss(s), out; std::string kv; bool first = true;
  while (std::getline(ss, kv, ',')) {
    const size_t eq = kv.find('=');
    if (eq != std::string::npos) { auto it = names.find(kv.substr(0, eq)); if (it != names.end()) kv = it->second + kv.substr(eq); }
    out << (first ? "" : ",") << kv; first = false;
  }
  return OptimizationAlgorithmWithHessianCuda::updatePropertiesFromString(out.str());
}

// ---------------------------------------------------------------------------------------------------------------
OptimizationAlgorithmFactory* OptimizationAlgorithmFactory::instance() { static OptimizationAlgorithmFactory f; return &f; }
void OptimizationAlgorithmFactory::registerSolver(const std::shared_ptr<AbstractOptimizationAlgorithmCreator>& c) {
  const std::string& name = c->property().name;
  for (auto& e : _creator) if (e->property().name == name) { e = c; std::cerr << "SOLVER FACTORY WARNING: Overwriting Solver creator " << name << std::endl; return; }   // factory.cpp:61-73
  _creator.push_back(c);
}
OptimizationAlgorithm* OptimizationAlgorithmFactory::construct(const std::string& name, OptimizationAlgorithmProperty& solverProperty) const {
  for (auto& c : _creator) if (c->property().name == name) { solverProperty = c->property(); return c->construct(); }
  std::cerr << "SOLVER FACTORY WARNING: Unable to create solver " << name << std::endl;                         // factory.cpp:85-94
  return nullptr;
}
void OptimizationAlgorithmFactory::listSolvers(std::ostream& os) const {
  size_t w = 0; for (auto& c : _creator) w = std::max(w, c->property().name.size());
  for (auto& c : _creator) os << c->property().name << std::string(w - c->property().name.size() + 4, ' ') << c->property().desc << std::endl;
}

}  // namespace g2o
