/*
 * g2ocu.h — C ABI of the B200-native g2o solver backend (libg2ocu.so).
 *
 * This is the drop-in boundary for the Levenberg-Marquardt / Gauss-Newton hot path of g2o
 * (reference: B0Bftl/g2o, paths below are relative to the reference root).  Every entry point replaces one
 * virtual of the reference's plugin interface; the g2o-side adapter (g2o_b200/host/g2o_solver_cuda.cpp and the
 * binding sketch in INTEGRATION.md) forwards the virtual to the function named here.
 *
 *   reference interface                                                  entry point
 *   ------------------------------------------------------------------   -------------------------------
 *   SparseOptimizer::initializeOptimization   sparse_optimizer.cpp:208    g2ocu_initialize_optimization
 *   OptimizationAlgorithmWithHessian::init    ..._with_hessian.cpp:48     g2ocu_init
 *   Solver::buildStructure                    core/solver.h:64            g2ocu_build_structure
 *   SparseOptimizer::computeActiveErrors      sparse_optimizer.cpp:63     g2ocu_compute_active_errors
 *   SparseOptimizer::activeRobustChi2/Chi2    sparse_optimizer.cpp:92,102 g2ocu_active_robust_chi2 / g2ocu_active_chi2
 *   Solver::buildSystem                       core/solver.h:81            g2ocu_build_system
 *   Solver::setLambda / restoreDiagonal       core/solver.h:108,113       g2ocu_set_lambda / g2ocu_restore_diagonal
 *   Solver::solve                             core/solver.h:86            g2ocu_solve
 *   Solver::x / b / vectorSize                core/solver.h:94-103        g2ocu_get_f64("x"|"b") / g2ocu_vector_size
 *   SparseOptimizer::update                   sparse_optimizer.cpp:441    g2ocu_update
 *   SparseOptimizer::push/pop/discardTop      sparse_optimizer.cpp:624    g2ocu_push / g2ocu_pop / g2ocu_discard_top
 *   OptimizationAlgorithm::solve(iteration)   optimization_algorithm.h:70 g2ocu_solver_iteration (G2OCU_ALGORITHM_GN / _LM / _DOGLEG:
 *                                             optimization_algorithm_gauss_newton.cpp:50, _levenberg.cpp:58, _dogleg.cpp:56)
 *   SparseOptimizer::optimize                 sparse_optimizer.cpp:374    g2ocu_optimize
 *   BlockSolverBase::multiplyHessian          core/block_solver.h:87-95   g2ocu_multiply_hessian
 *   SparseOptimizer::computeMarginals         sparse_optimizer.cpp:594    g2ocu_compute_marginals (Solver::computeMarginals, solver.h:78;
 *                                             LinearSolver::solvePattern, linear_solver.h:89-98)
 *   LinearSolver<M>::solve                    core/linear_solver.h:59     (inside g2ocu_solve; kind = g2ocu_config.linear_solver) / g2ocu_linear_solve (stand-alone)
 *
 * Conventions: all functions return 0 on success and a negative G2OCU_E_* code on failure; the message is
 * available from g2ocu_last_error().  No function aborts or throws across the boundary.  One handle = one
 * optimizer = one CUDA stream; calls on one handle must be serialised by the caller (the reference is not
 * thread-safe across calls either); different handles are independent.  All floating point is IEEE double
 * (number_t = double, config.h.in:33-41).  There is NO CPU fallback: compute entry points fail with
 * G2OCU_E_CUDA when no CUDA device is usable, and unsupported vertex/edge/kernel types are rejected at
 * g2ocu_set_graph / g2ocu_init with G2OCU_E_UNSUPPORTED.
 */
#ifndef G2OCU_H
#define G2OCU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define G2OCU_VERSION 1

/* status codes */
#define G2OCU_OK 0
#define G2OCU_E_INVALID (-1)      /* bad argument / call order                                   */
#define G2OCU_E_UNSUPPORTED (-2)  /* type / configuration outside the supported set (rejected)     */
#define G2OCU_E_CUDA (-3)         /* CUDA runtime error or no device                               */
#define G2OCU_E_NUMERIC (-4)      /* linear solve failed (non-SPD system)                          */
#define G2OCU_E_COMM (-5)         /* the user-supplied collective reported an error                */

/* vertex types: estimate layout on the boundary */
#define G2OCU_VERTEX_SE2 1         /* slam2d/vertex_se2.h        x y theta                         */
#define G2OCU_VERTEX_POINT_XY 2    /* slam2d/vertex_point_xy.h   x y                               */
#define G2OCU_VERTEX_SE3 3         /* slam3d/vertex_se3.h        Isometry3: R column-major(9), t(3) */
#define G2OCU_VERTEX_SE3_EXPMAP 4  /* sba/types_six_dof_expmap.h:84  SE3Quat::toVector tx ty tz qx qy qz qw */
#define G2OCU_VERTEX_POINT_XYZ 5   /* sba/types_sba.h:137        x y z                             */
#define G2OCU_VERTEX_CAM_BAL 6     /* examples/bal/bal_example.cpp:65   rx ry rz tx ty tz f k1 k2  */
#define G2OCU_VERTEX_POINT_BAL 7   /* examples/bal/bal_example.cpp:102  x y z                      */

/* edge types (vertex order as in the reference class) */
#define G2OCU_EDGE_SE2 1              /* slam2d/edge_se2.h            (SE2, SE2)          meas x y theta     */
#define G2OCU_EDGE_SE2_POINT_XY 2     /* slam2d/edge_se2_pointxy.h    (SE2, PointXY)      meas x y           */
#define G2OCU_EDGE_SE3 3              /* slam3d/edge_se3.h            (SE3, SE3)          meas Isometry3(12) */
#define G2OCU_EDGE_SE3_EXPMAP 4       /* sba/types_six_dof_expmap.h:108 (Expmap, Expmap)  meas SE3Quat(7)    */
#define G2OCU_EDGE_PROJECT_XYZ2UV 5   /* sba/types_six_dof_expmap.h:130 (PointXYZ, Expmap) meas u v; param f cx cy      */
#define G2OCU_EDGE_SE3_PROJECT_XYZ 6  /* sba/types_six_dof_expmap.h:201 (PointXYZ, Expmap) meas u v; param fx fy cx cy */
#define G2OCU_EDGE_BAL 7              /* examples/bal/bal_example.cpp:148 (CamBAL, PointBAL) meas u v      */

/* robust kernels, core/robust_kernel_impl.cpp */
#define G2OCU_KERNEL_NONE 0
#define G2OCU_KERNEL_HUBER 1
#define G2OCU_KERNEL_PSEUDO_HUBER 2
#define G2OCU_KERNEL_CAUCHY 3
#define G2OCU_KERNEL_GEMAN_MCCLURE 4
#define G2OCU_KERNEL_WELSCH 5
#define G2OCU_KERNEL_FAIR 6
#define G2OCU_KERNEL_TUKEY 7
#define G2OCU_KERNEL_SATURATED 8
#define G2OCU_KERNEL_DCS 9

/* algorithms (OptimizationAlgorithmLevenberg / GaussNewton) and linear solvers */
#define G2OCU_ALGORITHM_GN 0
#define G2OCU_ALGORITHM_LM 1
#define G2OCU_ALGORITHM_DOGLEG 2   /* core/optimization_algorithm_dogleg.cpp:56-197 (Powell's dogleg); single-GPU handles only */
/* OptimizationAlgorithmDogleg::lastStep() values (optimization_algorithm_dogleg.h:47-50), element 1 of g2ocu_get_f64("dogleg") */
#define G2OCU_DOGLEG_STEP_UNDEFINED 0
#define G2OCU_DOGLEG_STEP_SD 1
#define G2OCU_DOGLEG_STEP_GN 2
#define G2OCU_DOGLEG_STEP_DL 3
#define G2OCU_LINEAR_PCG 0    /* solvers/pcg/linear_solver_pcg.hpp: block-Jacobi preconditioned CG           */
#define G2OCU_LINEAR_DENSE 1  /* solvers/dense/linear_solver_dense.h semantics: dense FP64 Cholesky (small RCS) */

/* SolverResult, core/optimization_algorithm.h:46 */
#define G2OCU_RESULT_OK 1
#define G2OCU_RESULT_TERMINATE 2
#define G2OCU_RESULT_FAIL (-1)

typedef struct g2ocu_solver g2ocu_solver;

/* Graph as flat arrays.  Vertices in insertion order, edges in internalId order (optimizable_graph.cpp:267-292).
 * Packed arrays are concatenated in vertex / edge order with the per-type strides documented above;
 * information matrices are E x E column-major.  All arrays are copied by g2ocu_set_graph. */
typedef struct g2ocu_graph {
  int32_t n_vertices;
  const int32_t* v_id;            /* g2o vertex id (unique)                                   */
  const int32_t* v_type;          /* G2OCU_VERTEX_*                                           */
  const uint8_t* v_fixed;         /* OptimizableGraph::Vertex::fixed()                        */
  const uint8_t* v_marginalized;  /* OptimizableGraph::Vertex::marginalized()                 */
  const double* v_estimate;       /* packed estimates                                         */
  int32_t n_edges;
  const int32_t* e_type;          /* G2OCU_EDGE_*                                             */
  const int32_t* e_v0;            /* index (not id) of vertices()[0] in the vertex arrays     */
  const int32_t* e_v1;            /* index of vertices()[1]                                   */
  const int32_t* e_level;         /* Edge::level(); NULL = all 0                              */
  const double* e_measurement;    /* packed measurements                                      */
  const double* e_information;    /* packed information matrices                              */
  const int32_t* e_kernel;        /* G2OCU_KERNEL_*; NULL = none                              */
  const double* e_kernel_delta;   /* RobustKernel::delta(); NULL = 1.0                        */
  const double* e_param;          /* packed per-edge parameters (camera intrinsics)           */
} g2ocu_graph;

typedef struct g2ocu_config {
  int32_t device;                 /* CUDA device ordinal, -1 = current device                                   */
  int32_t linear_solver;          /* G2OCU_LINEAR_*                                                             */
  double pcg_tolerance;           /* LinearSolverPCG::_tolerance, default 1e-6 (linear_solver_pcg.h:53)         */
  int32_t pcg_max_iterations;     /* LinearSolverPCG::_maxIter, <0 = matrix rows (linear_solver_pcg.hpp:128)    */
  int32_t pcg_absolute_tolerance; /* LinearSolverPCG::_absoluteTolerance, default 1                             */
  void* stream;                   /* optional cudaStream_t to run on; NULL = the library creates its own        */
} g2ocu_config;

/* Per-iteration record: the G2OBatchStatistics fields (core/batch_stats.h:40-78) filled from CUDA events,
 * plus the LM state the reference prints in its verbose line (optimization_algorithm_levenberg.cpp:196-202). */
typedef struct g2ocu_iteration_stats {
  int32_t iteration;
  int32_t result;                 /* G2OCU_RESULT_*                                           */
  int32_t levenberg_iterations;   /* trials in this iteration                                 */
  int32_t iterations_linear_solver; /* PCG iterations of the last solve                       */
  double chi2;                    /* activeRobustChi2 after the iteration                     */
  double lambda;                  /* currentLambda after the iteration                        */
  double time_residuals;          /* seconds, device time                                     */
  double time_quadratic_form;
  double time_schur_complement;
  double time_linear_solver;
  double time_linear_solution;
  double time_update;
  double time_iteration;          /* host wall clock around the whole iteration                */
  int64_t hessian_pose_dimension;
  int64_t hessian_landmark_dimension;
} g2ocu_iteration_stats;

/* Collective hook for landmark-sharded multi-GPU runs, ordered after all work already enqueued on `stream`; returns 0 on
 * success.  Supplied by the host (torch.distributed / NCCL); never called when world == 1.  `buf` is a DEVICE pointer.
 *   op G2OCU_OP_SUM / G2OCU_OP_MAX        in-place all-reduce of `count` doubles
 *   op G2OCU_OP_REDUCE_SCATTER_SUM        in-place reduce-scatter: `buf` holds world * `count` doubles, afterwards rank r holds the
 *                                         sum over all ranks of buf[r * count, (r + 1) * count) in that same range (the NCCL
 *                                         in-place convention); the other ranges are unspecified                              */
#define G2OCU_OP_SUM 0
#define G2OCU_OP_MAX 1
#define G2OCU_OP_REDUCE_SCATTER_SUM 2
typedef int (*g2ocu_allreduce_fn)(void* buf, int64_t count, int32_t op, void* stream, void* user);

void g2ocu_default_config(g2ocu_config* cfg);
int g2ocu_version(void);
const char* g2ocu_last_error(const g2ocu_solver* s);   /* s may be NULL: error of the last failed g2ocu_create */

int g2ocu_create(const g2ocu_config* cfg, g2ocu_solver** out);
void g2ocu_destroy(g2ocu_solver* s);

int g2ocu_set_graph(g2ocu_solver* s, const g2ocu_graph* g);
int g2ocu_set_property(g2ocu_solver* s, const char* name, double value);  /* "initialLambda", "maxTrialsAfterFailure" (levenberg.cpp:48-49); "pcgTolerance", "pcgMaxIterations", "pcgAbsoluteTolerance" (linear_solver_pcg.h:53-57); "linearSolver" = G2OCU_LINEAR_* (which LinearSolver the BlockSolver owns, block_solver.h:124); "doglegInitialDelta", "doglegMaxTrialsAfterFailure", "doglegInitialLambda", "doglegLambdaFactor" = OptimizationAlgorithmDogleg's "initialDelta" (1e4), "maxTrialsAfterFailure" (100), "initialLambda" (1e-7), "lambdaFactor" (10) (optimization_algorithm_dogleg.cpp:44-47); "poseDim", "landmarkDim" = the fixed block sizes of BlockSolver<BlockSolverTraits<p,l>> (-1 = variable, the default): g2ocu_build_structure rejects a graph whose block sizes differ; "kernelTiming" != 0: g2ocu_phase_time also reports per-kernel phases (schur_tiles, pcg_spmv, ...) at the price of two event records per kernel */
/* SparseOptimizer::setForceStopFlag / terminate() (core/sparse_optimizer.h:186-190): `flag` points at a byte owned by the caller (g2o's bool); while it
 * is non-zero g2ocu_optimize starts no further iteration (sparse_optimizer.cpp:396) and the Levenberg trial loop stops after the current
 * trial (optimization_algorithm_levenberg.cpp:145).  NULL (default) = never. */
int g2ocu_set_force_stop_flag(g2ocu_solver* s, const unsigned char* flag);
int g2ocu_set_shard(g2ocu_solver* s, int32_t rank, int32_t world, g2ocu_allreduce_fn fn, void* user);
/* The same sharding with the collectives issued straight from the library through NCCL (no host callback per collective):
 * `nccl_library` is the path of libnccl.so.2 (dlopen'ed; the build has no link-time NCCL dependency), `unique_id` the 128 bytes of an
 * ncclUniqueId created by rank 0 with g2ocu_nccl_unique_id and distributed by the host (torch.distributed, MPI, a file ...).
 * Collective: every rank of the job must call it; it creates one communicator on the solver's device. */
int g2ocu_nccl_unique_id(const char* nccl_library, unsigned char unique_id[128]);
int g2ocu_set_shard_nccl(g2ocu_solver* s, int32_t rank, int32_t world, const char* nccl_library, const unsigned char unique_id[128]);
/* Optional, single node, 2..8 ranks, after g2ocu_build_structure: the per-iteration exchange of the slab PCG (the 8 Nc P bytes of
 * q = A d) goes through NVLink peer memory instead of an NCCL all-reduce.  export: allocates this rank's exchange buffer and returns
 * its 64-byte cudaIpcMemHandle; the host gathers the handles of all ranks (rank order) and passes them to import, which maps the peers'
 * buffers.  Both are collective in the sense that every rank must do both before the next solve. */
int g2ocu_p2p_export(g2ocu_solver* s, unsigned char handle[64]);
int g2ocu_p2p_import(g2ocu_solver* s, const unsigned char* handles /* world x 64 bytes */);
/* The same for the reduction of the reduced camera system (once per LM trial): every rank exports its partial-Hschur buffer, the peers map it,
 * and each rank sums the slab it solves with by reading the peers' buffers over NVLink - instead of an NCCL reduce-scatter of the whole buffer.
 * PCG solver only; without these calls (or with the dense solver) the reduce-scatter is used. */
int g2ocu_p2p_export_schur(g2ocu_solver* s, unsigned char handle[64]);
int g2ocu_p2p_import_schur(g2ocu_solver* s, const unsigned char* handles /* world x 64 bytes */);

int g2ocu_initialize_optimization(g2ocu_solver* s, int32_t level);
int g2ocu_init(g2ocu_solver* s, int32_t online);
int g2ocu_build_structure(g2ocu_solver* s);
int g2ocu_compute_active_errors(g2ocu_solver* s);
int g2ocu_active_robust_chi2(g2ocu_solver* s, double* chi2);
int g2ocu_active_chi2(g2ocu_solver* s, double* chi2);
int g2ocu_build_system(g2ocu_solver* s);
int g2ocu_set_lambda(g2ocu_solver* s, double lambda, int32_t backup);
int g2ocu_restore_diagonal(g2ocu_solver* s);
int g2ocu_solve(g2ocu_solver* s, int32_t* solved);
int g2ocu_update(g2ocu_solver* s, const double* host_update_or_null);
int g2ocu_push(g2ocu_solver* s);
int g2ocu_pop(g2ocu_solver* s);
int g2ocu_discard_top(g2ocu_solver* s);
int g2ocu_compute_lambda_init(g2ocu_solver* s, double* lambda);
int g2ocu_compute_scale(g2ocu_solver* s, double lambda, double* scale);
int g2ocu_multiply_hessian(g2ocu_solver* s, double* host_dest, const double* host_src);
/* Blocks of the inverse of Hpp (marginal covariances): pair i = (block_rows[i], block_cols[i]) in hessian-index units, as the
 * std::pair<int,int> list of SparseOptimizer::computeMarginals; out receives the n_pairs blocks one after the other, each rows x cols
 * doubles column-major (the MatrixX blocks of `spinv`; the block sizes follow from "pose_block_indices").  Hpp is what the reference calls
 * so: the pose block - or, when no point is marginalized, the whole system over all vertices in id order (blocks of two sizes).
 * *computed = 1 / 0 is the bool the reference returns (0: Hpp is not positive definite).  Needs a built system (g2ocu_build_system or an
 * iteration).  G2OCU_E_UNSUPPORTED for systems beyond the dense factorisation's limit (40 000 scalar rows). */
int g2ocu_compute_marginals(g2ocu_solver* s, int32_t n_pairs, const int32_t* block_rows, const int32_t* block_cols, double* out, int32_t* computed);

int g2ocu_solver_iteration(g2ocu_solver* s, int32_t algorithm, int32_t iteration, g2ocu_iteration_stats* stats);
int g2ocu_optimize(g2ocu_solver* s, int32_t algorithm, int32_t iterations, g2ocu_iteration_stats* stats, int32_t* performed);

int64_t g2ocu_vector_size(const g2ocu_solver* s);
int g2ocu_set_estimates(g2ocu_solver* s, const double* host_packed);
int g2ocu_get_estimates(g2ocu_solver* s, double* host_packed);
/* Landmark-sharded runs (g2ocu_set_shard*): the same two transfers restricted to what this rank works on - every pose estimate and the
 * estimates of its own landmark range ("shard_landmark_range"); the other entries of host_packed are neither read nor written, and the
 * read-back needs no all-gather.  The job's host side then moves every estimate once per step instead of once per rank.  Without a
 * shard they are g2ocu_set_estimates / g2ocu_get_estimates. */
int g2ocu_set_estimates_owned(g2ocu_solver* s, const double* host_packed);
int g2ocu_get_estimates_owned(g2ocu_solver* s, double* host_packed);

/* Named array read-back (structure arrays for the bit-exact check of SURVEY.md Appendix B, block values, vectors).
 * Returns the number of elements the array has (and fills at most `capacity` of them), or a negative status.
 * int32: "hessian_index" "active_vertices" "active_edges" "index_mapping" "dims" "pose_block_indices"
 *        "landmark_block_indices" "hpp_colptr" "hpp_rowidx" "hpl_colptr" "hpl_rowidx" "hschur_colptr" "hschur_rowidx"
 *        "hschur_t_colptr" "hschur_t_rowidx" "edge_targets" "shard_landmark_range" "shard_edge_positions"
 *        "linear_solver_iterations" (1 element: PCG iterations of the last g2ocu_solve, G2OBatchStatistics::iterationsLinearSolver)
 *        "tile_min_track" (1 element: landmarks with fewer observations go through the pair kernel of the Schur product, the others through
 *        the tile kernel - an implementation split the benchmark needs for its per-kernel flop counts)
 *        (graphs with poses and points of which none is marginalized - BlockSolverX with two block sizes: "dims",
 *        "pose_block_indices", "hpp_colptr", "hpp_rowidx", "edge_targets" describe the reference's single Hpp over all
 *        vertices in id order; "full_system_permutation" maps a scalar index of its x / b to the internal [poses | points]
 *        layout; "internal_dims" = poses, points, their scalar sizes, available for every graph)
 * double: "x" "b" "bschur" "hpp_values" "hpl_values" "hll_values" "hschur_values" "errors" "jacobians" "estimates" "lambda"
 *         "diagonal_blocks" = the D x D Hessian block of every vertex of the index mapping, in its order (what the reference maps into
 *                      OptimizableGraph::Vertex::hessian, block_solver.hpp:150-170; computeLambdaInit reads its diagonal, levenberg.cpp:152-175)
 *         "dogleg" = { trustRegion(), lastStep() (G2OCU_DOGLEG_STEP_*), tries of the last iteration, damping factor,
 *                      1 if the system was positive definite in all iterations } (optimization_algorithm_dogleg.h:64-68,84-88) */
int64_t g2ocu_get_i32(g2ocu_solver* s, const char* name, int32_t* out, int64_t capacity);
int64_t g2ocu_get_f64(g2ocu_solver* s, const char* name, double* out, int64_t capacity);

/* Device-side counters for measurement: number of kernel launches issued by this handle so far, and per phase name
 * ("errors" "build" "schur" "pcg_setup" "pcg_spmv" "pcg_vec" "linear_solver" "backsub" "update") the accumulated
 * device time in seconds (CUDA events on the handle's stream), the launches inside it and how many times it ran. */
int64_t g2ocu_launch_count(const g2ocu_solver* s);
int g2ocu_phase_time(g2ocu_solver* s, const char* phase, double* seconds, int64_t* launches, int64_t* calls);
int g2ocu_reset_counters(g2ocu_solver* s);

/* ---- LinearSolver<MatrixType> level (core/linear_solver.h:42-105) -------------------------------------------------------------------
 * For callers that keep g2o's own BlockSolver and only swap the linear solver: LinearSolverPCG::solve (solvers/pcg/linear_solver_pcg.hpp:80-156,
 * block-Jacobi preconditioned CG incl. the _residual carried from solve to solve) on a symmetric block matrix given by its UPPER blocks
 * in the reference's block-column layout (SparseBlockMatrix::blockCols(): per block column the blocks with row <= column, ascending rows,
 * each block column-major, one block size for the whole matrix: 3, 6 or 9).  `b`, `x` are host vectors of n_block_cols * block_dim doubles.
 * g2ocu_linear_init = LinearSolver::init (resets the carried residual); properties: "pcgTolerance", "pcgMaxIterations", "pcgAbsoluteTolerance". */
typedef struct g2ocu_linear_solver g2ocu_linear_solver;
int g2ocu_linear_create(const g2ocu_config* cfg, g2ocu_linear_solver** out);
void g2ocu_linear_destroy(g2ocu_linear_solver* s);
const char* g2ocu_linear_last_error(const g2ocu_linear_solver* s);
int g2ocu_linear_init(g2ocu_linear_solver* s);
int g2ocu_linear_set_property(g2ocu_linear_solver* s, const char* name, double value);
int g2ocu_linear_solve(g2ocu_linear_solver* s, int32_t n_block_cols, int32_t block_dim, const int32_t* colptr, const int32_t* rowidx, const double* values,
                       const double* b, double* x, int32_t* solved, int32_t* iterations);

#ifdef __cplusplus
}
#endif
#endif /* G2OCU_H */
