// g2o_cuda: the command line of the reference's `g2o` application (g2o/apps/g2o_cli/g2o.cpp:101-693) for the graphs the CUDA backend
// supports, on top of the host mirror.  Same flags where they make sense:
//   g2o_cuda [-i N] [-solver lm_var_cuda] [-robustKernel Huber] [-robustKernelWidth w] [-solverProperties k=v,...] [-gaugeId id]
//            [-marginalize] [-computeMarginals] [-stats file] [-o out.g2o] [-v] [-listSolvers] [-summary] [-bal] input
// -summary only loads and reports (no GPU needed).  Gauge: as g2o.cpp:283-316, a graph without a fixed vertex gets one fixed; the
// reference picks "the first maximum-dimension vertex" in unordered_map order (sparse_optimizer.cpp:118-137, not reproducible), here it
// is the one with the lowest id.  Landmarks are marginalized when the solver requires it (g2o.cpp:318-331).
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <set>
#include <sstream>

#include "g2o_io.hpp"

using namespace g2o;

static RobustKernel* makeKernel(const std::string& n) {   // RobustKernelFactory names (robust_kernel_impl.cpp:172-181)
  if (n == "Huber") return new RobustKernelHuber;
  if (n == "PseudoHuber") return new RobustKernelPseudoHuber;
  if (n == "Cauchy") return new RobustKernelCauchy;
  if (n == "GemanMcClure") return new RobustKernelGemanMcClure;
  if (n == "Welsch") return new RobustKernelWelsch;
  if (n == "Fair") return new RobustKernelFair;
  if (n == "Tukey") return new RobustKernelTukey;
  if (n == "Saturated") return new RobustKernelSaturated;
  if (n == "DCS") return new RobustKernelDCS;
  return nullptr;
}

int main(int argc, char** argv) {
  int maxIterations = 5, gaugeId = -1; bool verbose = false, listSolvers = false, summary = false, bal = false, marginalize = false, computeMarginals = false;
  std::string solver = "lm_var_cuda", robustKernel, outputFile, statsFile, solverProperties, input; double kernelWidth = -1.0;
  for (int a = 1; a < argc; ++a) {
    const std::string f = argv[a];
    auto next = [&]() -> const char* { if (a + 1 >= argc) { std::cerr << "missing value after " << f << std::endl; std::exit(1); } return argv[++a]; };
    if (f == "-i") maxIterations = std::atoi(next());
    else if (f == "-solver") solver = next();
    else if (f == "-robustKernel") robustKernel = next();
    else if (f == "-robustKernelWidth") kernelWidth = std::atof(next());
    else if (f == "-solverProperties") solverProperties = next();
    else if (f == "-gaugeId") gaugeId = std::atoi(next());
    else if (f == "-o") outputFile = next();
    else if (f == "-stats") statsFile = next();
    else if (f == "-v") verbose = true;
    else if (f == "-marginalize") marginalize = true;
    else if (f == "-computeMarginals") computeMarginals = true;
    else if (f == "-listSolvers") listSolvers = true;
    else if (f == "-summary") summary = true;
    else if (f == "-bal") bal = true;
    else if (f[0] == '-') { std::cerr << "unknown option " << f << std::endl; return 1; }
    else input = f;
  }
  if (listSolvers) { OptimizationAlgorithmFactory::instance()->listSolvers(std::cout); return 0; }
  if (input.empty()) { std::cerr << "usage: g2o_cuda [options] graph.g2o   (-listSolvers, -summary, -bal problem.txt)" << std::endl; return 1; }

  SparseOptimizer optimizer;
  optimizer.setVerbose(verbose);
  std::ifstream ifs(input);
  if (!ifs) { std::cerr << "Failed to open file " << input << std::endl; return 1; }
  LoadReport rep;
  if (!(bal ? loadBal(ifs, optimizer, &rep) : loadG2o(ifs, optimizer, &rep))) { std::cerr << "Error loading graph" << std::endl; return 2; }
  std::cerr << "Loaded " << rep.vertices << " vertices" << std::endl << "Loaded " << rep.edges << " edges" << std::endl;
  if (rep.vertices == 0) { std::cerr << "Graph contains no vertices" << std::endl; return 1; }

  std::set<int> dims; bool anyFixed = false; int maxDim = 0;
  for (auto* v : optimizer.vertexList()) { dims.insert(v->dimension()); anyFixed |= v->fixed(); maxDim = std::max(maxDim, v->dimension()); }
  if (summary) {
    std::cout << "vertices " << rep.vertices << " edges " << rep.edges << " fixed " << rep.fixed << " dimensions";
    for (int d : dims) std::cout << " " << d;
    std::cout << " skipped_lines " << rep.skippedLines << std::endl;
    if (!outputFile.empty()) { std::ofstream ofs(outputFile); if (!saveG2o(ofs, optimizer)) return 4; }
    return 0;
  }

  OptimizationAlgorithmProperty solverProperty;
  OptimizationAlgorithm* algorithm = OptimizationAlgorithmFactory::instance()->construct(solver, solverProperty);
  if (!algorithm) { std::cerr << "Error allocating solver. Allocating \"" << solver << "\" failed!" << std::endl; return 1; }
  optimizer.setAlgorithm(algorithm);
  if (!solverProperties.empty() && !algorithm->updatePropertiesFromString(solverProperties)) { std::cerr << "could not apply -solverProperties " << solverProperties << std::endl; return 1; }

  // gauge (g2o.cpp:283-316)
  if (gaugeId >= 0) { auto* v = optimizer.vertex(gaugeId); if (!v) { std::cerr << "fatal, not found the vertex of id " << gaugeId << std::endl; return -1; } v->setFixed(true); anyFixed = true; }
  if (!anyFixed && !bal) {   // bal_example.cpp leaves the gauge free (the LM damping regularises it)
    OptimizableGraph::Vertex* gauge = nullptr;
    for (auto* v : optimizer.vertexList()) if (v->dimension() == maxDim && (!gauge || v->id() < gauge->id())) gauge = v;
    std::cerr << "# graph is fixed by node " << gauge->id() << std::endl;
    gauge->setFixed(true);
  } else std::cerr << "# graph is fixed by priors or already fixed vertex" << std::endl;
  // marginalization of the landmarks (g2o.cpp:318-331)
  if ((marginalize || solverProperty.requiresMarginalize) && dims.size() > 1) {
    std::cerr << "# Preparing Marginalization of the Landmarks ... ";
    for (auto* v : optimizer.vertexList()) if (v->dimension() != maxDim) v->setMarginalized(true);
    std::cerr << "done." << std::endl;
  }
  if (!robustKernel.empty()) {   // g2o.cpp:333-358
    std::cerr << "# Preparing robust error function ... ";
    for (auto* e : optimizer.edgeList()) {
      RobustKernel* k = makeKernel(robustKernel);
      if (!k) { std::cerr << "Unknown Robust Kernel: " << robustKernel << std::endl; return 1; }
      if (kernelWidth > 0) k->setDelta(kernelWidth);
      e->setRobustKernel(k);
    }
    std::cerr << "done." << std::endl;
  }
  if (!statsFile.empty()) optimizer.setComputeBatchStatistics(true);
  if (!optimizer.initializeOptimization()) { std::cerr << "initializeOptimization failed" << std::endl; return 3; }
  optimizer.computeActiveErrors();
  std::cerr << "Initial chi2 = " << std::fixed << optimizer.activeChi2() << std::endl;
  const int result = optimizer.optimize(maxIterations);
  if (maxIterations > 0 && result <= 0) std::cerr << "optimize() returned " << result << ": the solver failed, result might be invalid" << std::endl;
  if (computeMarginals && !(maxIterations > 0 && result <= 0)) {   // g2o.cpp:581-608: per active vertex the blocks (h, h) and (h - 1, h) of the inverse
    std::vector<std::pair<int, int> > blockIndices; std::vector<int> ids; std::map<int, int> dimOf;
    for (const auto* v : optimizer.vertexList()) if (v->hessianIndex() >= 0) dimOf[v->hessianIndex()] = v->dimension();
    for (const auto* v : optimizer.vertexList()) {
      if (v->hessianIndex() >= 0) { blockIndices.push_back(std::make_pair(v->hessianIndex(), v->hessianIndex())); ids.push_back(v->id()); }
      if (v->hessianIndex() > 0) { blockIndices.push_back(std::make_pair(v->hessianIndex() - 1, v->hessianIndex())); ids.push_back(v->id()); }
    }
    std::vector<std::vector<number_t> > spinv;
    if (optimizer.computeMarginals(spinv, blockIndices)) {
      for (size_t i = 0; i < blockIndices.size(); ++i) {
        if (blockIndices[i].first == blockIndices[i].second) std::cerr << "Vertex id:" << ids[i] << std::endl;
        std::cerr << "inv block :" << blockIndices[i].first << ", " << blockIndices[i].second << std::endl;
        const int nr = dimOf[blockIndices[i].first], nc = nr ? (int)(spinv[i].size() / (size_t)nr) : 0;
        for (int r = 0; r < nr; ++r) { for (int c = 0; c < nc; ++c) std::cerr << (c ? " " : "") << std::setprecision(6) << std::defaultfloat << spinv[i][r + (size_t)nr * c]; std::cerr << std::endl; }
      }
    }
  }
  optimizer.computeActiveErrors();
  std::cout << "iterations " << result << " chi2 " << std::setprecision(17) << optimizer.activeChi2() << " robust_chi2 " << optimizer.activeRobustChi2() << std::endl;
  if (!statsFile.empty()) {
    std::cerr << "writing stats to file \"" << statsFile << "\" ... ";      // g2o.cpp:655-663: one G2OBatchStatistics line per iteration
    std::ofstream os(statsFile);
    for (const G2OBatchStatistics& s : optimizer.batchStatistics()) os << s << std::endl;
    std::cerr << "done." << std::endl;
  }
  if (!outputFile.empty() && !bal) { std::ofstream ofs(outputFile); if (!saveG2o(ofs, optimizer)) { std::cerr << "could not write " << outputFile << std::endl; return 4; } std::cerr << "saved " << outputFile << std::endl; }
  return result > 0 || maxIterations == 0 ? 0 : 5;
}
