// TEST INFRASTRUCTURE.  C entry points over leaf functions of the REAL reference, compiled from /root/reference by `make -C oracle ref`
// into oracle/_ref/libg2o_ref_leaves.so (together with this file; see the Makefile for the list of reference sources):
//   robust kernels  g2o/core/robust_kernel_impl.cpp:50-181, constructed by name through the reference's own RobustKernelFactory
//   dq/dR           g2o/types/slam3d/dquat2mat.cpp:35-85 + dquat2mat_maxima_generated.cpp
//   normalize_theta g2o/stuff/misc.h:114-127
//   SE2             g2o/types/slam2d/se2.h:39-131 (header only): composition, inverse and the places where the angle is normalised; the edge /
//                   vertex bodies that call it are restated below from edge_se2.h:45-52, edge_se2_pointxy.h:45-50, vertex_se2.h:51-58
//   SE3Quat         g2o/types/slam3d/se3quat.h:37-290 + se3_ops.hpp:27-85 (header only): exp, log, adj, product, inverse, map; the bodies that call
//                   them are restated below from sba/types_six_dof_expmap.h:98-101,117-124, .cpp:74-80,278-293
//   LinearSolverPCG g2o/solvers/pcg/linear_solver_pcg.h:41-113 + .hpp:27-197 over SparseBlockMatrix (g2o/core/sparse_block_matrix.h/.hpp): the solver
//                   object lives across solves (its _residual and its cached block pointers are part of the behaviour), blocks 3x3, 6x6, 9x9 or MatrixX
//   sampleGaussian  g2o/stuff/sampler.cpp:31-45 (one static std::normal_distribution shared by every engine - the noise source of create_sphere)
// tests/test_reference_leaves.py checks the oracle's restatements (and, on the GPU, the device functions through them) against these.
#include <cstring>

#include "g2o/core/robust_kernel.h"
#include "g2o/core/robust_kernel_factory.h"
#include "g2o/core/batch_stats.h"
#include "g2o/core/sparse_block_matrix.h"
#include "g2o/solvers/pcg/linear_solver_pcg.h"
#include "g2o/stuff/misc.h"
#include "g2o/stuff/sampler.h"
#include "g2o/types/slam2d/se2.h"
#include "g2o/types/slam3d/dquat2mat.h"
#include "g2o/types/slam3d/se3quat.h"

extern "C" {

// rho[3] = (rho, rho', rho'') of the kernel `name` ("Huber", "Cauchy", ... as registered by G2O_REGISTER_ROBUST_KERNEL) at squared error e2;
// returns 0, or -1 when the reference's factory does not know the name
int ref_robustify(const char* name, double delta, double e2, double* rho) {
  g2o::AbstractRobustKernelCreator* creator = g2o::RobustKernelFactory::instance()->creator(name);
  if (!creator) return -1;
  g2o::RobustKernel* k = creator->construct();
  k->setDelta(delta);
  g2o::Vector3 r;
  k->robustify(e2, r);
  rho[0] = r[0]; rho[1] = r[1]; rho[2] = r[2];
  delete k;
  return 0;
}

// R9: rotation matrix, column-major; out27: the 3 x 9 matrix dq_dR, column-major
void ref_dq_dR(const double* R9, double* out27) {
  Eigen::Matrix<number_t, 3, 9, Eigen::ColMajor> D;
  g2o::internal::compute_dq_dR(D, R9[0], R9[1], R9[2], R9[3], R9[4], R9[5], R9[6], R9[7], R9[8]);
  std::memcpy(out27, D.data(), sizeof(double) * 27);
}

double ref_normalize_theta(double theta) { return g2o::normalize_theta(theta); }

// EdgeSE2::computeError (edge_se2.h:45-52) with setMeasurementData (:61-65): x0, x1, z = (x, y, theta)
void ref_edge_se2_error(const double* x0, const double* x1, const double* z, double* e) {
  const g2o::SE2 v1(x0[0], x0[1], x0[2]), v2(x1[0], x1[1], x1[2]), measurement(z[0], z[1], z[2]);
  const g2o::SE2 inverseMeasurement = measurement.inverse();
  g2o::SE2 delta = inverseMeasurement * (v1.inverse() * v2);
  const g2o::Vector3 error = delta.toVector();
  e[0] = error[0]; e[1] = error[1]; e[2] = error[2];
}
// EdgeSE2PointXY::computeError (edge_se2_pointxy.h:45-50)
void ref_edge_se2_pointxy_error(const double* x0, const double* l, const double* z, double* e) {
  const g2o::SE2 v1(x0[0], x0[1], x0[2]);
  const g2o::Vector2 error = (v1.inverse() * g2o::Vector2(l[0], l[1])) - g2o::Vector2(z[0], z[1]);
  e[0] = error[0]; e[1] = error[1];
}
// VertexSE2::oplusImpl (vertex_se2.h:51-58)
void ref_vertex_se2_oplus(double* est, const double* update) {
  g2o::SE2 estimate(est[0], est[1], est[2]);
  g2o::Vector2 t = estimate.translation();
  t += g2o::Vector2(update[0], update[1]);
  const number_t angle = g2o::normalize_theta(estimate.rotation().angle() + update[2]);
  estimate.setTranslation(t);
  estimate.setRotation(g2o::Rotation2D(angle));
  est[0] = estimate[0]; est[1] = estimate[1]; est[2] = estimate[2];
}

static g2o::SE3Quat se3FromVector7(const double* v) { g2o::Vector7 a; for (int i = 0; i < 7; ++i) a[i] = v[i]; g2o::SE3Quat T; T.fromVector(a); return T; }
static void se3ToVector7(const g2o::SE3Quat& T, double* v) { const g2o::Vector7 a = T.toVector(); for (int i = 0; i < 7; ++i) v[i] = a[i]; }
// SE3Quat::exp (se3quat.h:218-257) and log (:173-209); 7-vectors are (t, qx, qy, qz, qw)
void ref_se3quat_exp(const double* u6, double* v7) { g2o::Vector6 u; for (int i = 0; i < 6; ++i) u[i] = u6[i]; se3ToVector7(g2o::SE3Quat::exp(u), v7); }
void ref_se3quat_log(const double* v7, double* u6) { const g2o::Vector6 u = se3FromVector7(v7).log(); for (int i = 0; i < 6; ++i) u6[i] = u[i]; }
// VertexSE3Expmap::oplusImpl (types_six_dof_expmap.h:98-101)
void ref_vertex_se3expmap_oplus(double* est7, const double* update6) {
  g2o::Vector6 update; for (int i = 0; i < 6; ++i) update[i] = update6[i];
  se3ToVector7(g2o::SE3Quat::exp(update) * se3FromVector7(est7), est7);
}
// EdgeSE3Expmap::computeError (types_six_dof_expmap.h:117-124) and linearizeOplus (.cpp:278-293); J0, J1 6 x 6 column-major
void ref_edge_se3expmap(const double* x0, const double* x1, const double* z, double* e6, double* J0, double* J1) {
  const g2o::SE3Quat v1 = se3FromVector7(x0), v2 = se3FromVector7(x1), C = se3FromVector7(z);
  g2o::SE3Quat error_ = v2.inverse() * C * v1;
  const g2o::Vector6 e = error_.log();
  for (int i = 0; i < 6; ++i) e6[i] = e[i];
  const g2o::SE3Quat invTij = C.inverse();
  const g2o::SE3Quat invTj_Tij = v2.inverse() * C, infTi_invTij = v1.inverse() * invTij;
  const Eigen::Matrix<number_t, 6, 6, Eigen::ColMajor> A = invTj_Tij.adj(), B = infTi_invTij.adj();
  for (int i = 0; i < 36; ++i) { J0[i] = A.data()[i]; J1[i] = -B.data()[i]; }
}
// EdgeProjectXYZ2UV::computeError (types_six_dof_expmap.h:140-147) with CameraParameters::cam_map (.cpp:74-80); prm = (f, cx, cy)
void ref_edge_project_xyz2uv_error(const double* X, const double* T7, const double* obs, const double* prm, double* e2) {
  const g2o::Vector3 trans_xyz = se3FromVector7(T7).map(g2o::Vector3(X[0], X[1], X[2]));
  const g2o::Vector2 proj = g2o::project(trans_xyz);
  e2[0] = obs[0] - (proj[0] * prm[0] + prm[1]);
  e2[1] = obs[1] - (proj[1] * prm[0] + prm[2]);
}

}  // extern "C"

namespace {
struct PcgHandleBase {
  virtual ~PcgHandleBase() {}
  virtual int solve(int nBlocks, const int* blockIndices, const int* colptr, const int* rowidx, const double* values, const double* b, double* x,
                    double tolerance, int maxIterations, int absoluteTolerance, int* iterations) = 0;
  virtual void init() = 0;
  virtual void multiplySymmetricUpperTriangle(double* dest, const double* src) = 0;
};
template <class MatrixType> struct PcgHandle : PcgHandleBase {
  g2o::LinearSolverPCG<MatrixType> solver;
  std::unique_ptr<g2o::SparseBlockMatrix<MatrixType> > A;
  PcgHandle() { solver.init(); }
  void init() override { solver.init(); }
  // SparseBlockMatrix::multiplySymmetricUpperTriangle (sparse_block_matrix.hpp:289-313) on the matrix of the last solve: dest += A src
  void multiplySymmetricUpperTriangle(double* dest, const double* src) override { number_t* d = dest; A->multiplySymmetricUpperTriangle(d, src); }
  // upper-triangular block CCS (block column c: rows rowidx[colptr[c] .. colptr[c+1]), ascending), values = the blocks in that order, column-major.
  // The matrix object is created on the first call and only refilled afterwards: BlockSolver keeps its SparseBlockMatrix for the whole
  // optimisation and LinearSolverPCG keeps pointers to its off-diagonal blocks (linear_solver_pcg.hpp:96-99).
  int solve(int nBlocks, const int* blockIndices, const int* colptr, const int* rowidx, const double* values, const double* b, double* x,
            double tolerance, int maxIterations, int absoluteTolerance, int* iterations) override {
    if (!A) A.reset(new g2o::SparseBlockMatrix<MatrixType>(blockIndices, blockIndices, nBlocks, nBlocks));
    size_t off = 0;
    for (int c = 0; c < nBlocks; ++c)
      for (int k = colptr[c]; k < colptr[c + 1]; ++k) {
        MatrixType* blk = A->block(rowidx[k], c, true);
        const int rows = A->rowsOfBlock(rowidx[k]), cols = A->colsOfBlock(c);
        for (int j = 0; j < cols; ++j) for (int i = 0; i < rows; ++i) (*blk)(i, j) = values[off + i + (size_t)rows * j];
        off += (size_t)rows * cols;
      }
    solver.setTolerance(tolerance); solver.setMaxIterations(maxIterations); solver.setAbsoluteTolerance(absoluteTolerance != 0);
    g2o::G2OBatchStatistics stats; g2o::G2OBatchStatistics::setGlobalStats(&stats);
    std::vector<double> rhs(b, b + A->rows());
    const bool ok = solver.solve(*A, x, rhs.data());
    g2o::G2OBatchStatistics::setGlobalStats(0);
    if (iterations) *iterations = stats.iterationsLinearSolver;
    return ok ? 1 : 0;
  }
};
}  // namespace

extern "C" {

// blockSize 3, 6, 9: LinearSolverPCG<Matrix<number_t, P, P>> (BlockSolver_3_2 / _6_3 / <9,3> pose blocks); anything else: LinearSolverPCG<MatrixX> (BlockSolverX)
void* ref_pcg_create(int blockSize) {
  if (blockSize == 3) return new PcgHandle<Eigen::Matrix<number_t, 3, 3, Eigen::ColMajor> >;
  if (blockSize == 6) return new PcgHandle<Eigen::Matrix<number_t, 6, 6, Eigen::ColMajor> >;
  if (blockSize == 9) return new PcgHandle<Eigen::Matrix<number_t, 9, 9, Eigen::ColMajor> >;
  return new PcgHandle<g2o::MatrixX>;
}
void ref_pcg_destroy(void* h) { delete (PcgHandleBase*)h; }
void ref_pcg_init(void* h) { ((PcgHandleBase*)h)->init(); }      // LinearSolverPCG::init(): forgets _residual and the cached blocks
int ref_pcg_solve(void* h, int nBlocks, const int* blockIndices, const int* colptr, const int* rowidx, const double* values, const double* b, double* x,
                  double tolerance, int maxIterations, int absoluteTolerance, int* iterations) {
  return ((PcgHandleBase*)h)->solve(nBlocks, blockIndices, colptr, rowidx, values, b, x, tolerance, maxIterations, absoluteTolerance, iterations);
}

// dest (caller-zeroed, A.rows() doubles) += A src with the matrix handed to the last ref_pcg_solve of this handle: what
// BlockSolver::multiplyHessian does with _Hpp (block_solver.h:146)
void ref_multiply_symmetric_upper(void* h, double* dest, const double* src) { ((PcgHandleBase*)h)->multiplySymmetricUpperTriangle(dest, src); }

// out[i] = sampleGaussian(&engine[which[i]]) for two default-seeded std::mt19937 engines, as the two GaussianSampler objects of
// create_sphere.cpp:117-132 hold them (GaussianSampler() : _generator(new std::mt19937), stuff/sampler.h:47-56).  The static
// distribution inside sampler.cpp keeps its saved value across calls and across engines; call once per process for a clean sequence.
void ref_sample_gaussian_two_engines(const int* which, int n, double* out) {
  std::mt19937 engine[2];
  for (int i = 0; i < n; ++i) out[i] = g2o::sampleGaussian(&engine[which[i] & 1]);
}

}  // extern "C"
