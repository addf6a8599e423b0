// Registration of the CUDA solvers with OptimizationAlgorithmFactory, written the way the reference registers its own
// (solvers/pcg/solver_pcg.cpp:41-98, solvers/csparse/solver_csparse.cpp:54-117): a creator class switching on the
// algorithm prefix and the block-solver suffix of the name, and one G2O_REGISTER_OPTIMIZATION_ALGORITHM per name.
// Built as libg2o_solver_cuda.so so that g2o's `*_solver_*` library glob finds it (apps/g2o_cli/g2o_common.cpp:82).
#include "g2o_mirror.hpp"

namespace g2o {

namespace {
template <typename BlockSolverT> std::unique_ptr<Solver> AllocateSolver() { return std::unique_ptr<Solver>(new BlockSolverT()); }
// solvers/dense/solver_dense.cpp:44-50: the same BlockSolver owning a LinearSolverDense (here: the device DMMA Cholesky)
template <typename BlockSolverT> std::unique_ptr<Solver> AllocateDenseSolver() { return std::unique_ptr<Solver>(new BlockSolverT(G2OCU_LINEAR_DENSE)); }

OptimizationAlgorithm* createSolver(const std::string& fullSolverName) {
  static const std::map<std::string, std::unique_ptr<Solver> (*)()> solver_factories{
      {"var_cuda", &AllocateSolver<CudaBlockSolverX>},     {"fix3_2_cuda", &AllocateSolver<CudaBlockSolver_3_2>},
      {"fix6_3_cuda", &AllocateSolver<CudaBlockSolver_6_3>},
      {"fix9_3_cuda", &AllocateSolver<CudaBlockSolver_9_3>},
      {"dense_cuda", &AllocateDenseSolver<CudaBlockSolverX>},     {"dense3_2_cuda", &AllocateDenseSolver<CudaBlockSolver_3_2>},
      {"dense6_3_cuda", &AllocateDenseSolver<CudaBlockSolver_6_3>},
      {"dense9_3_cuda", &AllocateDenseSolver<CudaBlockSolver_9_3>},
  };
  const std::string solverName = fullSolverName.substr(3);
  auto it = solver_factories.find(solverName);
  if (it == solver_factories.end()) return nullptr;
  const std::string methodName = fullSolverName.substr(0, 2);
  if (methodName == "gn") return new OptimizationAlgorithmGaussNewton(it->second());
  if (methodName == "lm") return new OptimizationAlgorithmLevenberg(it->second());
  if (methodName == "dl") return new OptimizationAlgorithmDogleg(it->second());
  return nullptr;
}

class CudaSolverCreator : public AbstractOptimizationAlgorithmCreator {
 public:
  explicit CudaSolverCreator(const OptimizationAlgorithmProperty& p) : AbstractOptimizationAlgorithmCreator(p) {}
  OptimizationAlgorithm* construct() override { return createSolver(property().name); }
};
}  // namespace

G2O_REGISTER_OPTIMIZATION_LIBRARY(cuda)
// (the reference's fix7_3 names are not registered: the backend has no sim3 types, and a name that cannot solve anything is worse than an unknown one)

G2O_REGISTER_OPTIMIZATION_ALGORITHM(gn_var_cuda, new CudaSolverCreator(OptimizationAlgorithmProperty("gn_var_cuda", "Gauss-Newton: block-Jacobi PCG on the GPU (variable blocksize)", "CUDA", false, -1, -1)))
G2O_REGISTER_OPTIMIZATION_ALGORITHM(gn_fix3_2_cuda, new CudaSolverCreator(OptimizationAlgorithmProperty("gn_fix3_2_cuda", "Gauss-Newton: Schur + PCG on the GPU (fixed blocksize)", "CUDA", true, 3, 2)))
G2O_REGISTER_OPTIMIZATION_ALGORITHM(gn_fix6_3_cuda, new CudaSolverCreator(OptimizationAlgorithmProperty("gn_fix6_3_cuda", "Gauss-Newton: Schur + PCG on the GPU (fixed blocksize)", "CUDA", true, 6, 3)))
G2O_REGISTER_OPTIMIZATION_ALGORITHM(gn_fix9_3_cuda, new CudaSolverCreator(OptimizationAlgorithmProperty("gn_fix9_3_cuda", "Gauss-Newton: Schur + PCG on the GPU (BAL cameras)", "CUDA", true, 9, 3)))
G2O_REGISTER_OPTIMIZATION_ALGORITHM(lm_var_cuda, new CudaSolverCreator(OptimizationAlgorithmProperty("lm_var_cuda", "Levenberg: block-Jacobi PCG on the GPU (variable blocksize)", "CUDA", false, -1, -1)))
G2O_REGISTER_OPTIMIZATION_ALGORITHM(lm_fix3_2_cuda, new CudaSolverCreator(OptimizationAlgorithmProperty("lm_fix3_2_cuda", "Levenberg: Schur + PCG on the GPU (fixed blocksize)", "CUDA", true, 3, 2)))
G2O_REGISTER_OPTIMIZATION_ALGORITHM(lm_fix6_3_cuda, new CudaSolverCreator(OptimizationAlgorithmProperty("lm_fix6_3_cuda", "Levenberg: Schur + PCG on the GPU (fixed blocksize)", "CUDA", true, 6, 3)))
G2O_REGISTER_OPTIMIZATION_ALGORITHM(lm_fix9_3_cuda, new CudaSolverCreator(OptimizationAlgorithmProperty("lm_fix9_3_cuda", "Levenberg: Schur + PCG on the GPU (BAL cameras)", "CUDA", true, 9, 3)))
G2O_REGISTER_OPTIMIZATION_ALGORITHM(gn_dense_cuda, new CudaSolverCreator(OptimizationAlgorithmProperty("gn_dense_cuda", "Gauss-Newton: dense FP64 Cholesky of the (reduced) system on the GPU (variable blocksize)", "CUDA", false, -1, -1)))
G2O_REGISTER_OPTIMIZATION_ALGORITHM(gn_dense3_2_cuda, new CudaSolverCreator(OptimizationAlgorithmProperty("gn_dense3_2_cuda", "Gauss-Newton: dense FP64 Cholesky of the (reduced) system on the GPU (fixed blocksize)", "CUDA", true, 3, 2)))
G2O_REGISTER_OPTIMIZATION_ALGORITHM(gn_dense6_3_cuda, new CudaSolverCreator(OptimizationAlgorithmProperty("gn_dense6_3_cuda", "Gauss-Newton: dense FP64 Cholesky of the (reduced) system on the GPU (fixed blocksize)", "CUDA", true, 6, 3)))
G2O_REGISTER_OPTIMIZATION_ALGORITHM(gn_dense9_3_cuda, new CudaSolverCreator(OptimizationAlgorithmProperty("gn_dense9_3_cuda", "Gauss-Newton: dense FP64 Cholesky of the (reduced) system on the GPU (BAL cameras)", "CUDA", true, 9, 3)))
G2O_REGISTER_OPTIMIZATION_ALGORITHM(lm_dense_cuda, new CudaSolverCreator(OptimizationAlgorithmProperty("lm_dense_cuda", "Levenberg: dense FP64 Cholesky of the (reduced) system on the GPU (variable blocksize)", "CUDA", false, -1, -1)))
G2O_REGISTER_OPTIMIZATION_ALGORITHM(lm_dense3_2_cuda, new CudaSolverCreator(OptimizationAlgorithmProperty("lm_dense3_2_cuda", "Levenberg: dense FP64 Cholesky of the (reduced) system on the GPU (fixed blocksize)", "CUDA", true, 3, 2)))
G2O_REGISTER_OPTIMIZATION_ALGORITHM(lm_dense6_3_cuda, new CudaSolverCreator(OptimizationAlgorithmProperty("lm_dense6_3_cuda", "Levenberg: dense FP64 Cholesky of the (reduced) system on the GPU (fixed blocksize)", "CUDA", true, 6, 3)))
G2O_REGISTER_OPTIMIZATION_ALGORITHM(lm_dense9_3_cuda, new CudaSolverCreator(OptimizationAlgorithmProperty("lm_dense9_3_cuda", "Levenberg: dense FP64 Cholesky of the (reduced) system on the GPU (BAL cameras)", "CUDA", true, 9, 3)))
G2O_REGISTER_OPTIMIZATION_ALGORITHM(dl_var_cuda, new CudaSolverCreator(OptimizationAlgorithmProperty("dl_var_cuda", "Dogleg: block-Jacobi PCG on the GPU (variable blocksize)", "CUDA", false, -1, -1)))

}  // namespace g2o
