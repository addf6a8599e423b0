// TEST INFRASTRUCTURE — CPU oracle for the g2o LM/BlockSolver hot path (B0Bftl/g2o).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this library.  The product path (g2o_b200/) never links, imports or calls it.
//
// What it is: an Eigen-free C++17 restatement of the reference's CPU algorithm for the path
//   SparseOptimizer -> OptimizationAlgorithm{Levenberg,GaussNewton} -> BlockSolver -> LinearSolver{PCG,Dense,CSparse}
// following the cited reference lines one for one (paths relative to /root/reference).
// The reference does not build the way it ships (every hot-path header needs Eigen3, which is absent: g2o/core/eigen_types.h:30-31); it is
// compiled against a stand-in for Eigen into oracle/_ref/ (oracle/Makefile, oracle/ref_core.cpp) and this restatement is checked against it
// end to end (tests/test_reference_core.py).  The vendored CSparse is linked from oracle/_ref/ for the Cholesky solver.
//
// Parity pin status: pinned against the reference itself.  The reference compiled into oracle/_ref/libg2o_ref_core.so (g2o/core + BlockSolver +
// LM / GN / Dogleg + LinearSolverPCG / CSparse + the slam2d, slam3d, sba and BAL types) and this file produce the same index map, chi2 per
// iteration, LM trial and PCG iteration counts, lambda and estimates on 28 graphs incl. BASELINE configs[0] (tests/test_reference_core.py); leaf
// functions are compared piecewise in tests/test_reference_leaves.py; the reference's own unit-test properties are re-run in
// tests/test_oracle_types.py.  Pinned by restatement only: LinearSolverDense's pivoted LDL^T (numpy cross-checks in tests/test_oracle_solvers.py).
#include "orc_types.hpp"
#include <vector>
#include <map>
#include <unordered_map>
#include <set>
#include <string>
#include <memory>
#include <chrono>
#include <cstdio>
#include <cstdint>
#include <cassert>
#include <array>
#include <dlfcn.h>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace orc {

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// ---------------------------------------------------------------------------------------------
// graph (hyper_graph.h / optimizable_graph.h, reduced to what the hot path reads)
// ---------------------------------------------------------------------------------------------
struct Vertex {
  int id = 0, type = 0;
  bool fixed = false, marginalized = false;
  int hessianIndex = -1, colInHessian = -1;
  int dim = 0;
  std::vector<double> est;
  std::vector<std::vector<double>> backup;      // base_vertex.h:93-96
  int numOplusCalls = 0;                         // vertex_se3.h:110 (not part of the backup)
  std::vector<int> edges;                        // all incident edges, any level
  std::vector<double> b;                         // base_vertex.h _b
  double* hessian = nullptr;                     // mapHessianMemory target (diag block)
#ifdef _OPENMP
  omp_lock_t lock;                               // base_edge.h:44-56 QuadraticFormLock
#endif
};

struct Edge {
  int internalId = 0, type = 0, v[2] = {0, 0}, level = 0, dim = 0;
  std::vector<double> meas, info, prm;
  int kernel = 0; double delta = 1.0;
  double err[6];
  double J0[54], J1[54];
  double* hessian = nullptr;  // off-diagonal block target
  bool hessianRowMajor = false;
  int hdim0 = 0, hdim1 = 0;
  double chi2() const {                          // base_edge.h:79-82
    double s = 0;
    for (int j = 0; j < dim; ++j) { double t = 0; for (int i = 0; i < dim; ++i) t += info[j + dim*i] * err[i]; s += err[j] * t; }
    return s;
  }
};

struct BatchStats {   // core/batch_stats.h:40-78 (doubles only so that ctypes can read it as an array)
  double iteration, numVertices, numEdges, chi2, timeResiduals, timeLinearize, timeQuadraticForm, levenbergIterations,
      timeSchurComplement, timeSymbolicDecomposition, timeNumericDecomposition, timeLinearSolution, timeLinearSolver,
      iterationsLinearSolver, timeUpdate, timeIteration, hessianDimension, hessianPoseDimension, hessianLandmarkDimension,
      choleskyNNZ, lambda, result;
};
static BatchStats* g_stats = nullptr;

// ---------------------------------------------------------------------------------------------
// SparseBlockMatrix (sparse_block_matrix.h/.hpp) : per block-column std::map<row, block>
// ---------------------------------------------------------------------------------------------
struct SparseBlockMatrix {
  std::vector<int> rowBlockIndices, colBlockIndices;   // cumulative ends
  std::vector<std::map<int, int>> blockCols;            // row -> block id
  std::vector<std::vector<double>> blocks;              // heap block per entry (sparse_block_matrix.h:219-224)
  int rowsOfBlock(int r) const { return r ? rowBlockIndices[r] - rowBlockIndices[r-1] : rowBlockIndices[0]; }
  int colsOfBlock(int c) const { return c ? colBlockIndices[c] - colBlockIndices[c-1] : colBlockIndices[0]; }
  int rowBaseOfBlock(int r) const { return r ? rowBlockIndices[r-1] : 0; }
  int colBaseOfBlock(int c) const { return c ? colBlockIndices[c-1] : 0; }
  int rows() const { return rowBlockIndices.empty() ? 0 : rowBlockIndices.back(); }
  int cols() const { return colBlockIndices.empty() ? 0 : colBlockIndices.back(); }
  void init(const int* rbi, const int* cbi, int rb, int cb) {
    rowBlockIndices.assign(rbi, rbi + rb); colBlockIndices.assign(cbi, cbi + cb);
    blockCols.assign(cb, {}); blocks.clear();
  }
  // sparse_block_matrix.hpp:88-110
  int block(int r, int c, bool alloc) {
    auto it = blockCols[c].find(r);
    if (it != blockCols[c].end()) return it->second;
    if (!alloc) return -1;
    int id = (int)blocks.size();
    blocks.emplace_back((size_t)rowsOfBlock(r) * colsOfBlock(c), 0.0);
    blockCols[c].insert({r, id});
    return id;
  }
  int blockConst(int r, int c) const { auto it = blockCols[c].find(r); return it == blockCols[c].end() ? -1 : it->second; }
  double* data(int id) { return blocks[id].data(); }
  const double* data(int id) const { return blocks[id].data(); }
  void clear() {   // sparse_block_matrix.hpp:67-84 (dealloc=false): zero all blocks
    for (auto& b : blocks) std::fill(b.begin(), b.end(), 0.0);
  }
  size_t nonZeroBlocks() const { size_t n = 0; for (auto& c : blockCols) n += c.size(); return n; }
};

// ---------------------------------------------------------------------------------------------
// linear solvers
// ---------------------------------------------------------------------------------------------
struct LinearSolver {
  virtual ~LinearSolver() {}
  virtual bool init() = 0;
  virtual bool solve(const SparseBlockMatrix& A, double* x, double* b) = 0;
};

// g2o/solvers/pcg/linear_solver_pcg.h:53-70, .hpp:80-197
struct LinearSolverPCG : LinearSolver {
  double tolerance = 1e-6, residual = -1.0; bool absoluteTolerance = true; int maxIter = -1; int lastIterations = 0;
  std::vector<const double*> diag; std::vector<std::vector<double>> J;
  std::vector<std::pair<int,int>> indices; std::vector<const double*> sparseMat; std::vector<std::pair<int,int>> sparseDims;
  bool init() override { residual = -1.0; indices.clear(); sparseMat.clear(); sparseDims.clear(); return true; }
  static void multDiagP(const std::vector<int>& cbi, const std::vector<const double*>& A, const double* src, double* dest) {
    int row = 0;
    for (size_t i = 0; i < A.size(); ++i) { int n = cbi[i] - row; for (int k = 0; k < n; ++k) dest[row+k] = 0; mv_add(A[i], src + row, dest + row, n, n); row = cbi[i]; }
  }
  static void multDiagV(const std::vector<int>& cbi, const std::vector<std::vector<double>>& A, const double* src, double* dest) {
    int row = 0;
    for (size_t i = 0; i < A.size(); ++i) { int n = cbi[i] - row; for (int k = 0; k < n; ++k) dest[row+k] = 0; mv_add(A[i].data(), src + row, dest + row, n, n); row = cbi[i]; }
  }
  void mult(const std::vector<int>& cbi, const double* src, double* dest) {       // .hpp:179-197
    multDiagP(cbi, diag, src, dest);
    for (size_t i = 0; i < sparseMat.size(); ++i) {
      const int srcOffset = indices[i].second, destOffsetT = srcOffset, destOffset = indices[i].first, srcOffsetT = destOffset;
      const double* a = sparseMat[i]; const int r = sparseDims[i].first, c = sparseDims[i].second;
      mv_add(a, src + srcOffset, dest + destOffset, r, c);
      mtv_add(a, src + srcOffsetT, dest + destOffsetT, r, c);
    }
  }
  bool solve(const SparseBlockMatrix& A, double* x, double* b) override {
    const bool indexRequired = indices.size() == 0;
    diag.clear(); J.clear();
    int colIdx = 0;
    for (size_t i = 0; i < A.blockCols.size(); ++i) {
      const auto& col = A.blockCols[i];
      if (col.size() > 0) {
        for (auto it = col.begin(); it != col.end(); ++it) {
          if (it->first == (int)i) {
            diag.push_back(A.data(it->second));
            int n = A.colsOfBlock((int)i); std::vector<double> inv((size_t)n*n);
            inverseN(A.data(it->second), inv.data(), n);
            J.push_back(std::move(inv));
            break;
          }
          if (indexRequired) {
            indices.push_back({it->first > 0 ? A.rowBlockIndices[it->first-1] : 0, colIdx});
            sparseMat.push_back(A.data(it->second));
            sparseDims.push_back({A.rowsOfBlock(it->first), A.colsOfBlock((int)i)});
          }
        }
      }
      colIdx = A.colBlockIndices[i];
    }
    int n = A.rows();
    std::vector<double> r(b, b + n), d(n, 0.0), q(n, 0.0), s(n, 0.0);
    for (int i = 0; i < A.cols(); ++i) x[i] = 0;
    multDiagV(A.colBlockIndices, J, r.data(), d.data());
    double dn = 0; for (int i = 0; i < n; ++i) dn += r[i]*d[i];
    double d0 = tolerance * dn;
    if (absoluteTolerance) { if (residual > 0.0 && residual > d0) d0 = residual; }
    int mIter = maxIter < 0 ? A.rows() : maxIter;
    int iteration;
    for (iteration = 0; iteration < mIter; ++iteration) {
      if (dn <= d0) break;
      mult(A.colBlockIndices, d.data(), q.data());
      double dq = 0; for (int i = 0; i < n; ++i) dq += d[i]*q[i];
      double a = dn / dq;
      for (int i = 0; i < n; ++i) x[i] += a*d[i];
      for (int i = 0; i < n; ++i) r[i] -= a*q[i];
      multDiagV(A.colBlockIndices, J, r.data(), s.data());
      double dold = dn;
      dn = 0; for (int i = 0; i < n; ++i) dn += r[i]*s[i];
      double ba = dn / dold;
      for (int i = 0; i < n; ++i) d[i] = s[i] + ba*d[i];
    }
    residual = 0.5 * dn; lastIterations = iteration;
    if (g_stats) g_stats->iterationsLinearSolver = iteration;
    return true;
  }
};

// g2o/solvers/dense/linear_solver_dense.h:65-115 — Eigen::LDLT restated as a diagonally pivoted LDL^T
struct LinearSolverDense : LinearSolver {
  bool init() override { return true; }
  bool solve(const SparseBlockMatrix& A, double* x, double* b) override {
    const int n = A.cols();
    std::vector<double> H((size_t)n*n, 0.0);
    for (size_t i = 0; i < A.blockCols.size(); ++i) {
      int c_idx = A.colBaseOfBlock((int)i), c_size = A.colsOfBlock((int)i);
      for (auto& kv : A.blockCols[i]) {
        if (kv.first <= (int)i) {
          int r_idx = A.rowBaseOfBlock(kv.first), r_size = A.rowsOfBlock(kv.first);
          const double* B = A.data(kv.second);
          for (int c = 0; c < c_size; ++c) for (int r = 0; r < r_size; ++r) {
            H[(size_t)(r_idx+r) + (size_t)(c_idx+c)*n] = B[r + c*r_size];
            if (r_idx != c_idx) H[(size_t)(c_idx+c) + (size_t)(r_idx+r)*n] = B[r + c*r_size];
          }
        }
      }
    }
    // LDL^T with symmetric (diagonal) pivoting on the lower triangle
    std::vector<int> perm(n); for (int i = 0; i < n; ++i) perm[i] = i;
    bool positive = true;
    for (int k = 0; k < n; ++k) {
      int p = k; double best = std::fabs(H[(size_t)k + (size_t)k*n]);
      for (int i = k+1; i < n; ++i) { double v = std::fabs(H[(size_t)i + (size_t)i*n]); if (v > best) { best = v; p = i; } }
      if (p != k) {
        for (int j = 0; j < n; ++j) std::swap(H[(size_t)k + (size_t)j*n], H[(size_t)p + (size_t)j*n]);
        for (int j = 0; j < n; ++j) std::swap(H[(size_t)j + (size_t)k*n], H[(size_t)j + (size_t)p*n]);
        std::swap(perm[k], perm[p]);
      }
      double dk = H[(size_t)k + (size_t)k*n];
      if (!(dk > 0)) { positive = false; if (dk == 0) break; }
      for (int i = k+1; i < n; ++i) H[(size_t)i + (size_t)k*n] /= dk;
      for (int j = k+1; j < n; ++j) {   // full (both triangles) trailing update keeps the symmetric row/column swaps valid
        double ljk = H[(size_t)j + (size_t)k*n] * dk;
        if (ljk == 0) continue;
        for (int i = k+1; i < n; ++i) H[(size_t)i + (size_t)j*n] -= H[(size_t)i + (size_t)k*n] * ljk;
      }
    }
    if (!positive) return false;
    std::vector<double> y(n);
    for (int i = 0; i < n; ++i) y[i] = b[perm[i]];
    for (int i = 0; i < n; ++i) for (int j = 0; j < i; ++j) y[i] -= H[(size_t)i + (size_t)j*n]*y[j];
    for (int i = 0; i < n; ++i) y[i] /= H[(size_t)i + (size_t)i*n];
    for (int i = n-1; i >= 0; --i) for (int j = i+1; j < n; ++j) y[i] -= H[(size_t)j + (size_t)i*n]*y[j];
    for (int i = 0; i < n; ++i) x[perm[i]] = y[i];
    return true;
  }
};

// g2o/solvers/csparse/linear_solver_csparse.h:106-139,246-344 + csparse_extension.cpp:31-119, calling the
// reference's own CSparse (EXTERNAL/csparse/*.c) compiled unmodified into oracle/_ref/libcsparse_ref.so.
struct cs_ref { int nzmax, m, n; int* p; int* i; double* x; int nz; };                 // EXTERNAL/csparse/cs.h struct cs_sparse
struct css_ref { int* pinv; int* q; int* parent; int* cp; int* leftmost; int m2; double lnz, unz; };
struct CSparseApi {
  void* h = nullptr;
  css_ref* (*cs_schol)(int, const cs_ref*) = nullptr;
  int* (*cs_amd)(int, const cs_ref*) = nullptr;
  int* (*cs_pinv)(const int*, int) = nullptr;
  cs_ref* (*cs_symperm)(const cs_ref*, const int*, int) = nullptr;
  int* (*cs_etree)(const cs_ref*, int) = nullptr;
  int* (*cs_post)(const int*, int) = nullptr;
  int* (*cs_counts)(const cs_ref*, const int*, const int*, int) = nullptr;
  double (*cs_cumsum)(int*, int*, int) = nullptr;
  void* (*cs_calloc)(int, size_t) = nullptr;
  void* (*cs_malloc)(int, size_t) = nullptr;
  void* (*cs_free)(void*) = nullptr;
  cs_ref* (*cs_spfree)(cs_ref*) = nullptr;
  css_ref* (*cs_sfree)(css_ref*) = nullptr;
  int (*cs_ereach)(const cs_ref*, int, const int*, int*, int*) = nullptr;
  cs_ref* (*cs_spalloc)(int, int, int, int, int) = nullptr;
  int (*cs_ipvec)(const int*, const double*, double*, int) = nullptr;
  int (*cs_pvec)(const int*, const double*, double*, int) = nullptr;
  int (*cs_lsolve)(const cs_ref*, double*) = nullptr;
  int (*cs_ltsolve)(const cs_ref*, double*) = nullptr;
  bool load(const char* path) {
    h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h) return false;
#define L(n) *(void**)(&n) = dlsym(h, #n); if (!n) return false;
    L(cs_schol) L(cs_amd) L(cs_pinv) L(cs_symperm) L(cs_etree) L(cs_post) L(cs_counts) L(cs_cumsum) L(cs_calloc) L(cs_malloc)
    L(cs_free) L(cs_spfree) L(cs_sfree) L(cs_ereach) L(cs_spalloc) L(cs_ipvec) L(cs_pvec) L(cs_lsolve) L(cs_ltsolve)
#undef L
    return true;
  }
};
static CSparseApi g_cs;

struct LinearSolverCSparse : LinearSolver {
  bool blockOrdering = false;
  css_ref* symbolic = nullptr;
  std::vector<int> Ap, Ai; std::vector<double> Ax;
  ~LinearSolverCSparse() { if (symbolic && g_cs.h) g_cs.cs_sfree(symbolic); }
  bool init() override { if (symbolic) { g_cs.cs_sfree(symbolic); symbolic = nullptr; } return true; }
  // sparse_block_matrix_ccs.h:139-166 fillCCS(upperTriangle=true)
  void fillCCS(const SparseBlockMatrix& A) {
    Ap.clear(); Ai.clear(); Ax.clear();
    for (size_t i = 0; i < A.blockCols.size(); ++i) {
      int csize = A.colsOfBlock((int)i);
      for (int c = 0; c < csize; ++c) {
        Ap.push_back((int)Ai.size());
        for (auto& kv : A.blockCols[i]) {
          if (kv.first > (int)i) break;   // only upper block rows
          int rstart = A.rowBaseOfBlock(kv.first), rsize = A.rowsOfBlock(kv.first);
          int elemsToCopy = rsize;
          if (kv.first == (int)i) elemsToCopy = c + 1;
          const double* B = A.data(kv.second);
          for (int r = 0; r < elemsToCopy; ++r) { Ax.push_back(B[r + c*rsize]); Ai.push_back(rstart + r); }
        }
      }
    }
    Ap.push_back((int)Ai.size());
  }
  bool solve(const SparseBlockMatrix& A, double* x, double* b) override {
    if (!g_cs.h) return false;
    fillCCS(A);
    cs_ref ccsA; ccsA.m = A.rows(); ccsA.n = A.cols(); ccsA.nzmax = (int)Ai.size(); ccsA.p = Ap.data(); ccsA.i = Ai.data(); ccsA.x = Ax.data(); ccsA.nz = -1;
    const int n = ccsA.n;
    double t = now();
    if (!symbolic) {
      if (!blockOrdering) symbolic = g_cs.cs_schol(1, &ccsA);
      else {
        // fillBlockStructure (sparse_block_matrix.hpp:551-575): upper block pattern, CCS
        std::vector<int> bp, bi;
        for (size_t c = 0; c < A.blockCols.size(); ++c) { bp.push_back((int)bi.size()); for (auto& kv : A.blockCols[c]) if (kv.first <= (int)c) bi.push_back(kv.first); }
        bp.push_back((int)bi.size());
        cs_ref aux; aux.nzmax = (int)bi.size(); aux.m = aux.n = (int)A.blockCols.size(); aux.p = bp.data(); aux.i = bi.data(); aux.x = nullptr; aux.nz = -1;
        int* P = g_cs.cs_amd(1, &aux);
        std::vector<int> scalarPermutation(n); size_t scalarIdx = 0;
        for (int i = 0; i < aux.n; ++i) { int p = P[i]; int base = A.colBaseOfBlock(p), nCols = A.colsOfBlock(p); for (int j = 0; j < nCols; ++j) scalarPermutation[scalarIdx++] = base++; }
        g_cs.cs_free(P);
        symbolic = (css_ref*)g_cs.cs_calloc(1, sizeof(css_ref));
        symbolic->pinv = g_cs.cs_pinv(scalarPermutation.data(), n);
        cs_ref* C = g_cs.cs_symperm(&ccsA, symbolic->pinv, 0);
        symbolic->parent = g_cs.cs_etree(C, 0);
        int* post = g_cs.cs_post(symbolic->parent, n);
        int* c = g_cs.cs_counts(C, symbolic->parent, post, 0);
        g_cs.cs_free(post); g_cs.cs_spfree(C);
        symbolic->cp = (int*)g_cs.cs_malloc(n+1, sizeof(int));
        symbolic->unz = symbolic->lnz = g_cs.cs_cumsum(symbolic->cp, c, n);
        g_cs.cs_free(c);
        if (symbolic->lnz < 0) { g_cs.cs_sfree(symbolic); symbolic = nullptr; }
      }
      if (g_stats) g_stats->timeSymbolicDecomposition = now() - t;
      if (!symbolic) return false;
    }
    t = now();
    if (x != b) std::memcpy(x, b, sizeof(double)*n);
    // csparse_extension.cpp:31-50 cs_cholsolsymb + :52-119 cs_chol_workspace (up-looking numeric Cholesky)
    std::vector<double> work(2*(size_t)n); std::vector<int> iwork(2*(size_t)n);
    bool ok = cholsolsymb(&ccsA, x, symbolic, work.data(), iwork.data());
    if (g_stats) { g_stats->timeNumericDecomposition = now() - t; g_stats->choleskyNNZ = symbolic->lnz; }
    return ok;
  }
  static bool cholsolsymb(const cs_ref* A, double* b, const css_ref* S, double* x, int* work) {
    int n = A->n;
    cs_ref* L = chol_workspace(A, S, work, x);
    if (!L) return false;
    g_cs.cs_ipvec(S->pinv, b, x, n);
    g_cs.cs_lsolve(L, x);
    g_cs.cs_ltsolve(L, x);
    g_cs.cs_pvec(S->pinv, x, b, n);
    g_cs.cs_spfree(L);
    return true;
  }
  static cs_ref* chol_workspace(const cs_ref* A, const css_ref* S, int* cin, double* xin) {
    if (!A || !S || !S->cp || !S->parent) return nullptr;
    int n = A->n;
    int* c = cin; int* s = cin + n; double* x = xin;
    int* cp = S->cp; int* pinv = S->pinv; int* parent = S->parent;
    cs_ref* C = pinv ? g_cs.cs_symperm(A, pinv, 1) : (cs_ref*)A;
    cs_ref* E = pinv ? C : nullptr;
    if (!C) return nullptr;
    int* Cp = C->p; int* Ci = C->i; double* Cx = C->x;
    cs_ref* L = g_cs.cs_spalloc(n, n, cp[n], 1, 0);
    if (!L) { if (E) g_cs.cs_spfree(E); return nullptr; }
    int* Lp = L->p; int* Li = L->i; double* Lx = L->x;
    for (int k = 0; k < n; k++) Lp[k] = c[k] = cp[k];
    for (int k = 0; k < n; k++) {
      int top = g_cs.cs_ereach(C, k, parent, s, c);
      x[k] = 0;
      for (int p = Cp[k]; p < Cp[k+1]; p++) if (Ci[p] <= k) x[Ci[p]] = Cx[p];
      double d = x[k]; x[k] = 0;
      for (; top < n; top++) {
        int i = s[top];
        double lki = x[i] / Lx[Lp[i]];
        x[i] = 0;
        for (int p = Lp[i] + 1; p < c[i]; p++) x[Li[p]] -= Lx[p] * lki;
        d -= lki * lki;
        int p = c[i]++;
        Li[p] = k; Lx[p] = lki;
      }
      if (d <= 0) { if (E) g_cs.cs_spfree(E); g_cs.cs_spfree(L); return nullptr; }
      int p = c[k]++;
      Li[p] = k; Lx[p] = std::sqrt(d);
    }
    Lp[n] = cp[n];
    if (E) g_cs.cs_spfree(E);
    return L;
  }
};

// ---------------------------------------------------------------------------------------------
// SparseOptimizer (sparse_optimizer.cpp) + BlockSolver (block_solver.hpp) + algorithms
// ---------------------------------------------------------------------------------------------
struct Optimizer {
  std::vector<Vertex> vertices; std::vector<Edge> edges;
  std::vector<int> activeVertices, activeEdges, ivMap;      // indices into vertices/edges
  int nthreads = 1;

  // ---- BlockSolver state (block_solver.h:100-185) ----
  bool doSchur = false;
  SparseBlockMatrix Hpp, Hll, Hpl, Hschur;
  struct RowBlock { int row; int block; };
  std::vector<std::vector<RowBlock>> HplCCS, HschurTransposedCCS;
  std::vector<std::vector<double>> DInvSchur;
  int numPoses = 0, numLandmarks = 0, sizePoses = 0, sizeLandmarks = 0;
  std::vector<double> x, b, coefficients, bschur;
  std::vector<std::vector<double>> diagonalBackupPose, diagonalBackupLandmark;
  std::unique_ptr<LinearSolver> linearSolver;
#ifdef _OPENMP
  std::vector<omp_lock_t> coefficientsMutex;
#endif
  // ---- algorithm state ----
  bool levenberg = true, dogleg = false;
  // Dogleg properties / state (optimization_algorithm_dogleg.cpp:40-52, .h:77-91)
  double dlUserDeltaInit = 1e4, dlInitialLambda = 1e-7, dlLambdaFactor = 10., dlDelta = 1e4, dlCurrentLambda = 1e-7;
  int dlMaxTrialsAfterFailure = 100, dlLastStep = 0, dlLastNumTries = 0; bool dlWasPDInAllIterations = true;
  std::vector<double> hsd, hdl, auxVector;
  double currentLambda = -1., tau = 1e-5, goodStepUpperScale = 2./3., goodStepLowerScale = 1./3., userLambdaInit = 0., ni = 2.;
  int maxTrialsAfterFailure = 10, levenbergIterations = 0;
  bool computeBatchStatistics = false;

  ~Optimizer() {
#ifdef _OPENMP
    for (auto& v : vertices) omp_destroy_lock(&v.lock);
    for (auto& m : coefficientsMutex) omp_destroy_lock(&m);
#endif
  }

  // sparse_optimizer.cpp:208-279 (vset = all vertices), :504-509, :168-193
  bool initializeOptimization(int level) {
    if (edges.empty()) return false;
    for (int vi : ivMap) vertices[vi].hessianIndex = -1;          // clearIndexMapping :195-200
    ivMap.clear(); activeVertices.clear(); activeEdges.clear();
    std::set<int> auxEdgeSet;
    for (size_t vi = 0; vi < vertices.size(); ++vi) {
      int levelEdges = 0;
      for (int ei : vertices[vi].edges) {
        const Edge& e = edges[ei];
        if (level < 0 || e.level == level) {
          bool allFixed = vertices[e.v[0]].fixed && vertices[e.v[1]].fixed;
          if (!allFixed) { auxEdgeSet.insert(ei); levelEdges++; }
        }
      }
      if (levelEdges) activeVertices.push_back((int)vi);
    }
    activeEdges.assign(auxEdgeSet.begin(), auxEdgeSet.end());
    std::sort(activeVertices.begin(), activeVertices.end(), [&](int a, int b) { return vertices[a].id < vertices[b].id; });
    std::sort(activeEdges.begin(), activeEdges.end(), [&](int a, int b) { return edges[a].internalId < edges[b].internalId; });
    // buildIndexMapping
    if (activeVertices.empty()) { ivMap.clear(); return false; }
    ivMap.resize(activeVertices.size());
    size_t i = 0;
    for (int k = 0; k < 2; k++)
      for (int vi : activeVertices) {
        Vertex& v = vertices[vi];
        if (!v.fixed) { if ((int)v.marginalized == k) { v.hessianIndex = (int)i; ivMap[i] = vi; i++; } }
        else v.hessianIndex = -1;
      }
    ivMap.resize(i);
    return true;
  }

  // sparse_optimizer.cpp:63-88
  void computeActiveErrors() {
#pragma omp parallel for default(shared) num_threads(nthreads) if (activeEdges.size() > 50)
    for (int k = 0; k < (int)activeEdges.size(); ++k) {
      Edge& e = edges[activeEdges[k]];
      computeError(e.type, vertices[e.v[0]].est.data(), vertices[e.v[1]].est.data(), e.meas.data(), e.prm.data(), e.err);
    }
  }
  // sparse_optimizer.cpp:102-116 (serial)
  double activeRobustChi2() const {
    double chi = 0.0;
    for (int ei : activeEdges) {
      const Edge& e = edges[ei];
      if (e.kernel) { double rho[3]; robustify(e.kernel, e.delta, e.chi2(), rho); chi += rho[0]; }
      else chi += e.chi2();
    }
    return chi;
  }
  double activeChi2() const { double chi = 0; for (int ei : activeEdges) chi += edges[ei].chi2(); return chi; }
  // sparse_optimizer.cpp:441-455
  void update(const double* upd) {
    for (int vi : ivMap) { Vertex& v = vertices[vi]; oplus(v.type, v.est.data(), upd, &v.numOplusCalls); upd += v.dim; }
  }
  void push() { for (int vi : activeVertices) vertices[vi].backup.push_back(vertices[vi].est); }          // :624-627
  void pop() { for (int vi : activeVertices) { Vertex& v = vertices[vi]; v.est = v.backup.back(); v.backup.pop_back(); } }
  void discardTop() { for (int vi : activeVertices) vertices[vi].backup.pop_back(); }

  // optimization_algorithm_with_hessian.cpp:48-72 + block_solver.hpp:568-581
  bool algorithmInit() {
    bool useSchur = false;
    for (int vi : activeVertices) if (vertices[vi].marginalized) { useSchur = true; break; }
    doSchur = useSchur;
    return linearSolver->init();
  }

  // block_solver.hpp:54-81 resize + :103-256 buildStructure
  bool buildStructure() {
    numPoses = numLandmarks = sizePoses = sizeLandmarks = 0;
    std::vector<int> blockPoseIndices, blockLandmarkIndices; size_t sparseDim = 0;
    for (int vi : ivMap) {
      Vertex& v = vertices[vi];
      if (!v.marginalized) { v.colInHessian = sizePoses; sizePoses += v.dim; blockPoseIndices.push_back(sizePoses); ++numPoses; }
      else { v.colInHessian = sizeLandmarks; sizeLandmarks += v.dim; blockLandmarkIndices.push_back(sizeLandmarks); ++numLandmarks; }
      sparseDim += v.dim;
    }
    x.assign(sparseDim, 0.0); b.assign(sparseDim, 0.0);
    Hpp.init(blockPoseIndices.data(), blockPoseIndices.data(), numPoses, numPoses);
    if (doSchur) {
      coefficients.assign(sparseDim, 0.0); bschur.assign(std::max(sizePoses, sizeLandmarks), 0.0);
      Hschur.init(blockPoseIndices.data(), blockPoseIndices.data(), numPoses, numPoses);
      Hll.init(blockLandmarkIndices.data(), blockLandmarkIndices.data(), numLandmarks, numLandmarks);
      Hpl.init(blockPoseIndices.data(), blockLandmarkIndices.data(), numPoses, numLandmarks);
#ifdef _OPENMP
      for (auto& m : coefficientsMutex) omp_destroy_lock(&m);
      coefficientsMutex.resize(numPoses); for (auto& m : coefficientsMutex) omp_init_lock(&m);
#endif
    }
    // diagonal blocks (:138-154); heap blocks are stable (vector<vector>) only if we never reallocate inner vectors
    struct Pending { int vi; int mat; int id; };
    std::vector<Pending> vtargets;
    int poseIdx = 0, landmarkIdx = 0;
    for (int vi : ivMap) {
      Vertex& v = vertices[vi];
      if (!v.marginalized) { vtargets.push_back({vi, 0, Hpp.block(poseIdx, poseIdx, true)}); ++poseIdx; }
      else { vtargets.push_back({vi, 1, Hll.block(landmarkIdx, landmarkIdx, true)}); ++landmarkIdx; }
    }
    std::set<std::pair<int,int>> schurMatrixLookup;   // SparseBlockMatrixHashMap + takePatternFromHash => sorted (col,row) set
    struct ETarget { int ei; int mat; int id; bool transposed; int d0, d1; };
    std::vector<ETarget> etargets;
    for (int ei : activeEdges) {
      Edge& e = edges[ei];
      e.hessian = nullptr;
      for (int viIdx = 0; viIdx < 2; ++viIdx) {
        Vertex& v1 = vertices[e.v[viIdx]];
        int ind1 = v1.hessianIndex;
        if (ind1 == -1) continue;
        int indexV1Bak = ind1;
        for (int vjIdx = viIdx + 1; vjIdx < 2; ++vjIdx) {
          Vertex& v2 = vertices[e.v[vjIdx]];
          int ind2 = v2.hessianIndex;
          if (ind2 == -1) continue;
          ind1 = indexV1Bak;
          bool transposedBlock = ind1 > ind2;
          if (transposedBlock) std::swap(ind1, ind2);
          if (!v1.marginalized && !v2.marginalized) {
            int id = Hpp.block(ind1, ind2, true);
            etargets.push_back({ei, 0, id, transposedBlock, 0, 0});
            if (doSchur) schurMatrixLookup.insert({ind2, ind1});
          } else if (v1.marginalized && v2.marginalized) {
            int id = Hll.block(ind1 - numPoses, ind2 - numPoses, true);
            etargets.push_back({ei, 1, id, false, 0, 0});
          } else {
            if (v1.marginalized) { int id = Hpl.block(v2.hessianIndex, v1.hessianIndex - numPoses, true); etargets.push_back({ei, 2, id, true, 0, 0}); }
            else { int id = Hpl.block(v1.hessianIndex, v2.hessianIndex - numPoses, true); etargets.push_back({ei, 2, id, false, 0, 0}); }
          }
        }
      }
    }
    // resolve pointers now that all blocks exist (the reference hands out raw pointers as it goes)
    for (auto& t : vtargets) vertices[t.vi].hessian = (t.mat == 0 ? Hpp : Hll).data(t.id);
    for (auto& t : etargets) {
      Edge& e = edges[t.ei];
      SparseBlockMatrix& M = t.mat == 0 ? Hpp : (t.mat == 1 ? Hll : Hpl);
      e.hessian = M.data(t.id); e.hessianRowMajor = t.transposed;
    }
    if (!doSchur) return true;

    DInvSchur.assign(landmarkIdx, {});
    // fillSparseBlockMatrixCCS (sparse_block_matrix.hpp:624-640)
    HplCCS.assign(Hpl.blockCols.size(), {});
    for (size_t c = 0; c < Hpl.blockCols.size(); ++c) for (auto& kv : Hpl.blockCols[c]) HplCCS[c].push_back({kv.first, kv.second});

    for (int vi : ivMap) {
      Vertex& v = vertices[vi];
      if (!v.marginalized) continue;
      for (int e1 : v.edges) for (int i = 0; i < 2; ++i) {
        int a = edges[e1].v[i]; const Vertex& v1 = vertices[a];
        if (v1.hessianIndex == -1 || a == vi) continue;
        for (int e2 : v.edges) for (int j = 0; j < 2; ++j) {
          int c = edges[e2].v[j]; const Vertex& v2 = vertices[c];
          if (v2.hessianIndex == -1 || c == vi) continue;
          int i1 = v1.hessianIndex, i2 = v2.hessianIndex;
          if (i1 <= i2) schurMatrixLookup.insert({i2, i1});
        }
      }
    }
    // takePatternFromHash (sparse_block_matrix.hpp:659-687): blocks allocated per column in ascending row order
    for (auto& cr : schurMatrixLookup) Hschur.block(cr.second, cr.first, true);
    // fillSparseBlockMatrixCCSTransposed (:642-657)
    rebuildSchurTransposed();
    return true;
  }
  void rebuildSchurTransposed() {
    HschurTransposedCCS.assign(Hschur.blockCols.size(), {});
    for (size_t c = 0; c < Hschur.blockCols.size(); ++c) for (auto& kv : Hschur.blockCols[c]) HschurTransposedCCS[kv.first].push_back({(int)c, kv.second});
  }

  // base_binary_edge.hpp:62-137
  void constructQuadraticForm(Edge& e) {
    Vertex& from = vertices[e.v[0]]; Vertex& to = vertices[e.v[1]];
    const int D = e.dim, Di = from.dim, Dj = to.dim;
    const double* A = e.J0; const double* B = e.J1;
    bool fromNotFixed = !from.fixed, toNotFixed = !to.fixed;
    if (!(fromNotFixed || toNotFixed)) return;
    double omega[36]; std::memcpy(omega, e.info.data(), sizeof(double)*D*D);
    double omega_r[6];
    for (int i = 0; i < D; ++i) { double s = 0; for (int j = 0; j < D; ++j) s += omega[i + D*j]*e.err[j]; omega_r[i] = -s; }
    if (e.kernel) {
      double rho[3]; robustify(e.kernel, e.delta, e.chi2(), rho);
      for (int i = 0; i < D*D; ++i) omega[i] *= rho[1];             // robustInformation, base_edge.h:117-123
      for (int i = 0; i < D; ++i) omega_r[i] *= rho[1];
    }
    double OA[54], OB[54];
    mm(omega, A, OA, D, D, Di); mm(omega, B, OB, D, D, Dj);
    if (fromNotFixed) {
#ifdef _OPENMP
      omp_set_lock(&from.lock);
#endif
      mtv_add(A, omega_r, from.b.data(), D, Di);
      mtm_add(A, OA, from.hessian, Di, D, Di);
#ifdef _OPENMP
      omp_unset_lock(&from.lock);
#endif
      if (toNotFixed && e.hessian) {
        if (e.hessianRowMajor) mtm_add(B, OA, e.hessian, Dj, D, Di);     // _hessianTransposed (Dj x Di) += B^T Ω A
        else mtm_add(A, OB, e.hessian, Di, D, Dj);                        // _hessian (Di x Dj) += A^T Ω B
      }
    }
    if (toNotFixed) {
#ifdef _OPENMP
      omp_set_lock(&to.lock);
#endif
      mtv_add(B, omega_r, to.b.data(), D, Dj);
      mtm_add(B, OB, to.hessian, Dj, D, Dj);
#ifdef _OPENMP
      omp_unset_lock(&to.lock);
#endif
    }
  }

  // block_solver.hpp:463-521
  void buildSystem() {
    for (int vi : ivMap) std::fill(vertices[vi].b.begin(), vertices[vi].b.end(), 0.0);
    Hpp.clear();
    if (doSchur) { Hll.clear(); Hpl.clear(); }
#pragma omp parallel for default(shared) num_threads(nthreads) if (activeEdges.size() > 100)
    for (int k = 0; k < (int)activeEdges.size(); ++k) {
      Edge& e = edges[activeEdges[k]];
      linearizeOplus(e.type, vertices[e.v[0]].est.data(), vertices[e.v[1]].est.data(), e.meas.data(), e.prm.data(), e.J0, e.J1);
      constructQuadraticForm(e);
    }
    for (int vi : ivMap) {
      Vertex& v = vertices[vi];
      int iBase = v.colInHessian; if (v.marginalized) iBase += sizePoses;
      std::memcpy(b.data() + iBase, v.b.data(), sizeof(double)*v.dim);
    }
  }

  // block_solver.hpp:525-565
  void setLambda(double lambda, bool backup) {
    if (backup) { diagonalBackupPose.resize(numPoses); diagonalBackupLandmark.resize(numLandmarks); }
    for (int i = 0; i < numPoses; ++i) {
      double* blk = Hpp.data(Hpp.blockConst(i, i)); int n = Hpp.colsOfBlock(i);
      if (backup) { diagonalBackupPose[i].resize(n); for (int k = 0; k < n; ++k) diagonalBackupPose[i][k] = blk[k + k*n]; }
      for (int k = 0; k < n; ++k) blk[k + k*n] += lambda;
    }
    for (int i = 0; i < numLandmarks; ++i) {
      double* blk = Hll.data(Hll.blockConst(i, i)); int n = Hll.colsOfBlock(i);
      if (backup) { diagonalBackupLandmark[i].resize(n); for (int k = 0; k < n; ++k) diagonalBackupLandmark[i][k] = blk[k + k*n]; }
      for (int k = 0; k < n; ++k) blk[k + k*n] += lambda;
    }
  }
  void restoreDiagonal() {
    for (int i = 0; i < numPoses; ++i) { double* blk = Hpp.data(Hpp.blockConst(i, i)); int n = Hpp.colsOfBlock(i); for (int k = 0; k < n; ++k) blk[k + k*n] = diagonalBackupPose[i][k]; }
    for (int i = 0; i < numLandmarks; ++i) { double* blk = Hll.data(Hll.blockConst(i, i)); int n = Hll.colsOfBlock(i); for (int k = 0; k < n; ++k) blk[k + k*n] = diagonalBackupLandmark[i][k]; }
  }

  // block_solver.hpp:315-447
  bool solve() {
    if (!doSchur) {
      double t = now();
      bool ok = linearSolver->solve(Hpp, x.data(), b.data());
      if (g_stats) { g_stats->timeLinearSolver = now() - t; g_stats->hessianDimension = g_stats->hessianPoseDimension = Hpp.cols(); }
      return ok;
    }
    double t = now();
    Hschur.clear();
    {  // _Hpp->add(*_Hschur): sparse_block_matrix.hpp:175-187 (allocates missing blocks in dest)
      for (size_t i = 0; i < Hpp.blockCols.size(); ++i) for (auto& kv : Hpp.blockCols[i]) {
        int d = Hschur.block(kv.first, (int)i, true);
        const auto& s = Hpp.blocks[kv.second]; auto& dst = Hschur.blocks[d];
        for (size_t k = 0; k < s.size(); ++k) dst[k] += s[k];
      }
      // _HschurTransposedCCS is NOT rebuilt: the reference fills it once in buildStructure (block_solver.hpp:253), so the diagonal blocks of poses
      // that observe no landmark - added to Hschur here - never appear in it (the landmark loop below only looks up co-observation pairs)
    }
    std::fill(coefficients.begin(), coefficients.begin() + sizePoses, 0.0);
#pragma omp parallel for default(shared) schedule(dynamic, 10) num_threads(nthreads)
    for (int landmarkIndex = 0; landmarkIndex < (int)Hll.blockCols.size(); ++landmarkIndex) {
      const auto& marginalizeColumn = Hll.blockCols[landmarkIndex];
      const double* D = Hll.data(marginalizeColumn.begin()->second);
      const int L = Hll.colsOfBlock(landmarkIndex);
      std::vector<double>& Dinv = DInvSchur[landmarkIndex]; Dinv.resize((size_t)L*L);
      inverseN(D, Dinv.data(), L);
      double db0[3], db[3] = {0, 0, 0};
      for (int j = 0; j < L; ++j) db0[j] = b[Hll.rowBaseOfBlock(landmarkIndex) + sizePoses + j];
      mv_add(Dinv.data(), db0, db, L, L);
      const auto& landmarkColumn = HplCCS[landmarkIndex];
      for (size_t o = 0; o < landmarkColumn.size(); ++o) {
        int i1 = landmarkColumn[o].row;
        const double* Bi = Hpl.data(landmarkColumn[o].block);
        const int P = Hpl.rowsOfBlock(i1);
        double BDinv[27]; mm(Bi, Dinv.data(), BDinv, P, L, L);
#ifdef _OPENMP
        omp_set_lock(&coefficientsMutex[i1]);
#endif
        mv_add(Bi, db, &coefficients[Hpl.rowBaseOfBlock(i1)], P, L);
        auto targetColumnIt = HschurTransposedCCS[i1].begin();
        for (size_t in = o; in < landmarkColumn.size(); ++in) {   // lower_bound(row >= i1) == position o (rows ascending, unique)
          int i2 = landmarkColumn[in].row;
          const double* Bj = Hpl.data(landmarkColumn[in].block);
          const int P2 = Hpl.rowsOfBlock(i2);
          while (targetColumnIt->row < i2) ++targetColumnIt;
          double* Hi1i2 = Hschur.data(targetColumnIt->block);
          for (int c = 0; c < P2; ++c) for (int r = 0; r < P; ++r) {
            double s = 0; for (int k = 0; k < L; ++k) s += BDinv[r + P*k] * Bj[c + P2*k];
            Hi1i2[r + P*c] -= s;
          }
        }
#ifdef _OPENMP
        omp_unset_lock(&coefficientsMutex[i1]);
#endif
      }
    }
    std::memcpy(bschur.data(), b.data(), sizeof(double)*sizePoses);
    for (int i = 0; i < sizePoses; ++i) bschur[i] -= coefficients[i];
    if (g_stats) g_stats->timeSchurComplement = now() - t;
    t = now();
    bool solvedPoses = linearSolver->solve(Hschur, x.data(), bschur.data());
    if (g_stats) { g_stats->timeLinearSolver = now() - t; g_stats->hessianPoseDimension = Hpp.cols(); g_stats->hessianLandmarkDimension = Hll.cols(); g_stats->hessianDimension = Hpp.cols() + Hll.cols(); }
    if (!solvedPoses) return false;
    double* xp = x.data(); double* cp = coefficients.data();
    double* xl = x.data() + sizePoses; double* cl = coefficients.data() + sizePoses; double* bl = b.data() + sizePoses;
    for (int i = 0; i < sizePoses; ++i) cp[i] = -xp[i];
    std::memcpy(cl, bl, sizeof(double)*sizeLandmarks);
    // _HplCCS->rightMultiply(cl, cp): sparse_block_matrix_ccs.h:100-125  dest(col) += B^T src(row)
#pragma omp parallel for default(shared) schedule(dynamic, 10) num_threads(nthreads)
    for (int i = 0; i < (int)HplCCS.size(); ++i) {
      int destOffset = Hpl.colBaseOfBlock(i); const int L = Hpl.colsOfBlock(i);
      for (auto& rb : HplCCS[i]) { int srcOffset = Hpl.rowBaseOfBlock(rb.row); mtv_add(Hpl.data(rb.block), cp + srcOffset, cl + destOffset, Hpl.rowsOfBlock(rb.row), L); }
    }
    std::fill(xl, xl + sizeLandmarks, 0.0);
    // _DInvSchur->multiply(xl, cl): sparse_block_matrix_diagonal.h:77-101
#pragma omp parallel for default(shared) schedule(dynamic, 10) num_threads(nthreads)
    for (int i = 0; i < (int)DInvSchur.size(); ++i) { int off = Hll.colBaseOfBlock(i); const int L = Hll.colsOfBlock(i); mv_add(DInvSchur[i].data(), cl + off, xl + off, L, L); }
    return true;
  }

  // optimization_algorithm_levenberg.cpp:152-175
  double computeLambdaInit() const {
    if (userLambdaInit > 0) return userLambdaInit;
    double maxDiagonal = 0;
    for (int vi : ivMap) { const Vertex& v = vertices[vi]; for (int j = 0; j < v.dim; ++j) maxDiagonal = std::max(std::fabs(v.hessian[j + j*v.dim]), maxDiagonal); }
    return tau * maxDiagonal;
  }
  // :177-184
  double computeScale() const { double scale = 0; for (size_t j = 0; j < x.size(); j++) scale += x[j]*(currentLambda*x[j] + b[j]); return scale; }

  enum { OK = 1, Terminate = 2, Fail = -1 };
  // optimization_algorithm_levenberg.cpp:58-150
  int solveLevenberg(int iteration) {
    if (iteration == 0) { if (!buildStructure()) return Fail; }
    double t = now();
    computeActiveErrors();
    if (g_stats) { g_stats->timeResiduals = now() - t; t = now(); }
    double currentChi = activeRobustChi2(), tempChi = currentChi;
    buildSystem();
    if (g_stats) g_stats->timeQuadraticForm = now() - t;
    if (iteration == 0) { currentLambda = computeLambdaInit(); ni = 2; }
    double rho = 0; int& qmax = levenbergIterations; qmax = 0;
    do {
      push();
      if (g_stats) { g_stats->levenbergIterations++; t = now(); }
      setLambda(currentLambda, true);
      bool ok2 = solve();
      if (g_stats) { g_stats->timeLinearSolution += now() - t; t = now(); }
      update(x.data());
      if (g_stats) g_stats->timeUpdate = now() - t;
      restoreDiagonal();
      computeActiveErrors();
      tempChi = activeRobustChi2();
      if (!ok2) tempChi = std::numeric_limits<double>::max();
      rho = (currentChi - tempChi);
      double scale = computeScale(); scale += 1e-3;
      rho /= scale;
      if (rho > 0 && std::isfinite(tempChi)) {
        double alpha = 1. - std::pow((2*rho - 1), 3);
        alpha = (std::min)(alpha, goodStepUpperScale);
        double scaleFactor = (std::max)(goodStepLowerScale, alpha);
        currentLambda *= scaleFactor; ni = 2; currentChi = tempChi; discardTop();
      } else {
        currentLambda *= ni; ni *= 2; pop();
        if (!std::isfinite(currentLambda)) break;
      }
      qmax++;
    } while (rho < 0 && qmax < maxTrialsAfterFailure);
    if (qmax == maxTrialsAfterFailure || rho == 0 || !std::isfinite(currentLambda)) return Terminate;
    return OK;
  }
  // BlockSolver::multiplyHessian (block_solver.h:146) = _Hpp->multiplySymmetricUpperTriangle (sparse_block_matrix.hpp:289-313):
  // dest(pose part) += Hpp src with the stored upper blocks used on both sides; the landmark part of dest is never touched
  void multiplyHessian(double* dest, const double* src) const {
    for (size_t i = 0; i < Hpp.blockCols.size(); ++i) {
      const int srcOffset = Hpp.colBaseOfBlock((int)i);
      for (const auto& kv : Hpp.blockCols[i]) {
        const double* a = Hpp.data(kv.second);
        const int destOffset = Hpp.rowBaseOfBlock(kv.first);
        if (destOffset > srcOffset) break;
        const int r = Hpp.rowsOfBlock(kv.first), c = Hpp.colsOfBlock((int)i);
        mv_add(a, src + srcOffset, dest + destOffset, r, c);
        if (destOffset < srcOffset) mtv_add(a, src + destOffset, dest + srcOffset, r, c);
      }
    }
  }
  enum { STEP_UNDEFINED = 0, STEP_SD = 1, STEP_GN = 2, STEP_DL = 3 };
  // optimization_algorithm_dogleg.cpp:56-197
  int solveDogleg(int iteration) {
    if (iteration == 0) {
      if (!buildStructure()) return Fail;
      hsd.assign(x.size(), 0.0); hdl.assign(x.size(), 0.0); auxVector.assign(x.size(), 0.0);
      dlDelta = dlUserDeltaInit; dlCurrentLambda = dlInitialLambda; dlWasPDInAllIterations = true;
    }
    const size_t n = x.size();
    auto dot = [&](const std::vector<double>& u, const std::vector<double>& v) { double s = 0; for (size_t i = 0; i < n; ++i) s += u[i]*v[i]; return s; };
    double t = now();
    computeActiveErrors();
    if (g_stats) { g_stats->timeResiduals = now() - t; t = now(); }
    const double currentChi = activeRobustChi2();
    buildSystem();
    if (g_stats) g_stats->timeQuadraticForm = now() - t;
    // alpha (:97-101)
    std::fill(auxVector.begin(), auxVector.end(), 0.0);
    multiplyHessian(auxVector.data(), b.data());
    const double bNormSquared = dot(b, b);
    const double alpha = bNormSquared / dot(auxVector, b);
    for (size_t i = 0; i < n; ++i) hsd[i] = alpha * b[i];
    const double hsdNorm = std::sqrt(dot(hsd, hsd));
    double hgnNorm = -1.;
    bool solvedGaussNewton = false, goodStep = false;
    int& numTries = dlLastNumTries; numTries = 0;
    do {
      ++numTries;
      if (!solvedGaussNewton) {
        const double minLambda = 1e-12, maxLambda = 1e3;
        solvedGaussNewton = true;
        bool solverOk = false;
        while (!solverOk) {
          if (!dlWasPDInAllIterations) setLambda(dlCurrentLambda, true);
          solverOk = solve();
          if (!dlWasPDInAllIterations) restoreDiagonal();
          dlWasPDInAllIterations = dlWasPDInAllIterations && solverOk;
          if (!dlWasPDInAllIterations) {
            if (solverOk) dlCurrentLambda = std::max(minLambda, dlCurrentLambda / (0.5 * dlLambdaFactor));
            else { dlCurrentLambda *= dlLambdaFactor; if (dlCurrentLambda > maxLambda) { dlCurrentLambda = maxLambda; return Fail; } }
          }
        }
        hgnNorm = std::sqrt(dot(x, x));
      }
      const std::vector<double>& hgn = x;
      if (hgnNorm < dlDelta) { hdl = hgn; dlLastStep = STEP_GN; }
      else if (hsdNorm > dlDelta) { const double f = dlDelta / hsdNorm; for (size_t i = 0; i < n; ++i) hdl[i] = f * hsd[i]; dlLastStep = STEP_SD; }
      else {
        for (size_t i = 0; i < n; ++i) auxVector[i] = hgn[i] - hsd[i];
        const double c = dot(hsd, auxVector), bmaSquaredNorm = dot(auxVector, auxVector);
        double beta;
        if (c <= 0.) beta = (-c + std::sqrt(c*c + bmaSquaredNorm * (dlDelta*dlDelta - dot(hsd, hsd)))) / bmaSquaredNorm;
        else { const double hsdSqrNorm = dot(hsd, hsd); beta = (dlDelta*dlDelta - hsdSqrNorm) / (c + std::sqrt(c*c + bmaSquaredNorm * (dlDelta*dlDelta - hsdSqrNorm))); }
        for (size_t i = 0; i < n; ++i) hdl[i] = hsd[i] + beta * (hgn[i] - hsd[i]);
        dlLastStep = STEP_DL;
      }
      // linear gain (:165-168)
      std::fill(auxVector.begin(), auxVector.end(), 0.0);
      multiplyHessian(auxVector.data(), hdl.data());
      double linearGain = -1 * dot(auxVector, hdl) + 2 * dot(b, hdl);
      push();
      update(hdl.data());
      computeActiveErrors();
      const double newChi = activeRobustChi2();
      const double nonLinearGain = currentChi - newChi;
      if (std::fabs(linearGain) < 1e-12) linearGain = 1e-12;
      const double rho = nonLinearGain / linearGain;
      if (rho > 0) { discardTop(); goodStep = true; } else pop();
      if (rho > 0.75) dlDelta = std::max(dlDelta, 3 * std::sqrt(dot(hdl, hdl)));
      else if (rho < 0.25) dlDelta *= 0.5;
    } while (!goodStep && numTries < dlMaxTrialsAfterFailure);
    if (numTries == dlMaxTrialsAfterFailure || !goodStep) return Terminate;
    return OK;
  }
  // optimization_algorithm_gauss_newton.cpp:50-91
  int solveGaussNewton(int iteration) {
    double t = now();
    computeActiveErrors();
    if (g_stats) g_stats->timeResiduals = now() - t;
    if (iteration == 0) { if (!buildStructure()) return Fail; }
    t = now(); buildSystem();
    if (g_stats) { g_stats->timeQuadraticForm = now() - t; t = now(); }
    bool ok = solve();
    if (g_stats) { g_stats->timeLinearSolution = now() - t; t = now(); }
    update(x.data());
    if (g_stats) g_stats->timeUpdate = now() - t;
    return ok ? OK : Fail;
  }
  // sparse_optimizer.cpp:374-439
  int optimize(int iterations, BatchStats* stats) {
    if (ivMap.empty()) return -1;
    int cjIterations = 0; bool ok = algorithmInit();
    if (!ok) return -1;
    int result = OK;
    for (int i = 0; i < iterations && ok; i++) {
      BatchStats local; std::memset(&local, 0, sizeof(local));
      BatchStats* cstat = stats ? &stats[i] : &local;
      std::memset(cstat, 0, sizeof(BatchStats));
      g_stats = cstat; cstat->iteration = i; cstat->numEdges = activeEdges.size(); cstat->numVertices = activeVertices.size();
      double ts = now();
      result = dogleg ? solveDogleg(i) : levenberg ? solveLevenberg(i) : solveGaussNewton(i);
      ok = (result == OK);
      computeActiveErrors();
      cstat->chi2 = activeRobustChi2();
      cstat->timeIteration = now() - ts;
      cstat->lambda = dogleg ? dlCurrentLambda : currentLambda; cstat->result = result;
      if (levenberg) cstat->levenbergIterations = levenbergIterations;
      ++cjIterations;
    }
    g_stats = nullptr;
    if (result == Fail) return 0;
    return cjIterations;
  }
};

}  // namespace orc

// =============================================================================================
// C API (ctypes)
// =============================================================================================
using namespace orc;

struct orc_graph {
  int32_t n_vertices; const int32_t* v_id; const int32_t* v_type; const uint8_t* v_fixed; const uint8_t* v_marginalized; const double* v_estimate;
  int32_t n_edges; const int32_t* e_type; const int32_t* e_v0; const int32_t* e_v1; const int32_t* e_level;
  const double* e_measurement; const double* e_information; const int32_t* e_kernel; const double* e_kernel_delta; const double* e_param;
};

struct Handle { Optimizer opt; std::vector<int32_t> ibuf; std::vector<double> dbuf; std::string err; };

extern "C" {

int orc_load_csparse(const char* path) { if (g_cs.h) return 1; return g_cs.load(path) ? 1 : 0; }

void* orc_create(const orc_graph* g) {
  Handle* h = new Handle; Optimizer& o = h->opt;
  o.vertices.resize(g->n_vertices);
  size_t eo = 0;
  for (int i = 0; i < g->n_vertices; ++i) {
    Vertex& v = o.vertices[i];
    v.id = g->v_id[i]; v.type = g->v_type[i]; v.fixed = g->v_fixed[i]; v.marginalized = g->v_marginalized[i];
    int S = vertexEstimateDim(v.type); if (S < 0) { delete h; return nullptr; }
    v.dim = vertexDim(v.type); v.est.assign(g->v_estimate + eo, g->v_estimate + eo + S); eo += S;
    v.b.assign(v.dim, 0.0);
#ifdef _OPENMP
    omp_init_lock(&v.lock);
#endif
  }
  o.edges.resize(g->n_edges);
  size_t mo = 0, io = 0, po = 0;
  for (int i = 0; i < g->n_edges; ++i) {
    Edge& e = o.edges[i];
    e.internalId = i; e.type = g->e_type[i]; e.v[0] = g->e_v0[i]; e.v[1] = g->e_v1[i]; e.level = g->e_level ? g->e_level[i] : 0;
    int D = edgeDim(e.type); if (D < 0) { delete h; return nullptr; }
    e.dim = D; int M = edgeMeasDim(e.type), Pn = edgeParamDim(e.type);
    e.meas.assign(g->e_measurement + mo, g->e_measurement + mo + M); mo += M;
    e.info.assign(g->e_information + io, g->e_information + io + D*D); io += D*D;
    if (Pn) { e.prm.assign(g->e_param + po, g->e_param + po + Pn); po += Pn; }
    e.kernel = g->e_kernel ? g->e_kernel[i] : 0; e.delta = g->e_kernel_delta ? g->e_kernel_delta[i] : 1.0;
    if (o.vertices[e.v[0]].type != edgeVertexType(e.type, 0) || o.vertices[e.v[1]].type != edgeVertexType(e.type, 1)) { delete h; return nullptr; }
    o.vertices[e.v[0]].edges.push_back(i);
    if (e.v[1] != e.v[0]) o.vertices[e.v[1]].edges.push_back(i);
  }
  o.linearSolver.reset(new LinearSolverPCG);
  return h;
}
void orc_destroy(void* hh) { delete (Handle*)hh; }
void orc_set_num_threads(void* hh, int n) { ((Handle*)hh)->opt.nthreads = n < 1 ? 1 : n; }
int orc_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// name: "<gn|lm|dl>_<anything>" ; linear: "pcg" | "dense" | "csparse" | "csparse_block"
int orc_set_solver(void* hh, const char* algorithm, const char* linear) {
  Optimizer& o = ((Handle*)hh)->opt;
  std::string a(algorithm), l(linear);
  o.dogleg = false;
  if (a.substr(0, 2) == "gn") o.levenberg = false; else if (a.substr(0, 2) == "lm") o.levenberg = true; else if (a.substr(0, 2) == "dl") { o.levenberg = false; o.dogleg = true; } else return -1;
  if (l == "pcg") o.linearSolver.reset(new LinearSolverPCG);
  else if (l == "dense") o.linearSolver.reset(new LinearSolverDense);
  else if (l == "csparse" || l == "csparse_block") { if (!g_cs.h) return -2; auto* s = new LinearSolverCSparse; s->blockOrdering = (l == "csparse_block"); o.linearSolver.reset(s); }
  else return -1;
  return 0;
}
void orc_set_lm_params(void* hh, double userLambdaInit, int maxTrialsAfterFailure) { Optimizer& o = ((Handle*)hh)->opt; o.userLambdaInit = userLambdaInit; o.maxTrialsAfterFailure = maxTrialsAfterFailure; }
void orc_set_dogleg_params(void* hh, double initialDelta, int maxTrialsAfterFailure, double initialLambda, double lambdaFactor) {
  Optimizer& o = ((Handle*)hh)->opt; o.dlUserDeltaInit = initialDelta; o.dlMaxTrialsAfterFailure = maxTrialsAfterFailure; o.dlInitialLambda = initialLambda; o.dlLambdaFactor = lambdaFactor;
}
// dest (zeroed here, vectorSize doubles) = Hpp src on the pose part
void orc_multiply_hessian(void* hh, double* dest, const double* src) { Optimizer& o = ((Handle*)hh)->opt; std::fill(dest, dest + o.x.size(), 0.0); o.multiplyHessian(dest, src); }
void orc_set_pcg_params(void* hh, double tol, int maxIter, int absoluteTolerance) {
  auto* p = dynamic_cast<LinearSolverPCG*>(((Handle*)hh)->opt.linearSolver.get());
  if (p) { p->tolerance = tol; p->maxIter = maxIter; p->absoluteTolerance = absoluteTolerance != 0; }
}
int orc_initialize_optimization(void* hh, int level) { return ((Handle*)hh)->opt.initializeOptimization(level) ? 1 : 0; }
int orc_optimize(void* hh, int iterations, double* stats /* iterations x 22 */) { return ((Handle*)hh)->opt.optimize(iterations, (BatchStats*)stats); }
int orc_stats_stride() { return (int)(sizeof(BatchStats) / sizeof(double)); }

// fine-grained Solver/SparseOptimizer interface
int orc_algorithm_init(void* hh) { return ((Handle*)hh)->opt.algorithmInit() ? 1 : 0; }
int orc_build_structure(void* hh) { return ((Handle*)hh)->opt.buildStructure() ? 1 : 0; }
void orc_compute_active_errors(void* hh) { ((Handle*)hh)->opt.computeActiveErrors(); }
double orc_active_robust_chi2(void* hh) { return ((Handle*)hh)->opt.activeRobustChi2(); }
double orc_active_chi2(void* hh) { return ((Handle*)hh)->opt.activeChi2(); }
void orc_build_system(void* hh) { ((Handle*)hh)->opt.buildSystem(); }
void orc_set_lambda(void* hh, double lambda, int backup) { Optimizer& o = ((Handle*)hh)->opt; o.currentLambda = lambda; o.setLambda(lambda, backup != 0); }
void orc_restore_diagonal(void* hh) { ((Handle*)hh)->opt.restoreDiagonal(); }
int orc_solve(void* hh) { return ((Handle*)hh)->opt.solve() ? 1 : 0; }
void orc_update(void* hh, const double* upd) { Optimizer& o = ((Handle*)hh)->opt; o.update(upd ? upd : o.x.data()); }
void orc_push(void* hh) { ((Handle*)hh)->opt.push(); }
void orc_pop(void* hh) { ((Handle*)hh)->opt.pop(); }
void orc_discard_top(void* hh) { ((Handle*)hh)->opt.discardTop(); }
double orc_compute_lambda_init(void* hh) { return ((Handle*)hh)->opt.computeLambdaInit(); }
double orc_compute_scale(void* hh) { return ((Handle*)hh)->opt.computeScale(); }
int orc_do_schur(void* hh) { return ((Handle*)hh)->opt.doSchur ? 1 : 0; }

// named array getters; pointers stay valid until the next getter call on the same handle
static void ccsOf(const SparseBlockMatrix& M, std::vector<int32_t>& out, bool ptr) {
  out.clear();
  if (ptr) { int n = 0; for (auto& c : M.blockCols) { out.push_back(n); n += (int)c.size(); } out.push_back(n); }
  else for (auto& c : M.blockCols) for (auto& kv : c) out.push_back(kv.first);
}
static void valuesOf(const SparseBlockMatrix& M, std::vector<double>& out) {
  out.clear();
  for (auto& c : M.blockCols) for (auto& kv : c) out.insert(out.end(), M.blocks[kv.second].begin(), M.blocks[kv.second].end());
}
const int32_t* orc_get_i32(void* hh, const char* name, int64_t* n) {
  Handle* h = (Handle*)hh; Optimizer& o = h->opt; std::string s(name); auto& out = h->ibuf; out.clear();
  if (s == "hessian_index") for (auto& v : o.vertices) out.push_back(v.hessianIndex);
  else if (s == "active_vertices") out.assign(o.activeVertices.begin(), o.activeVertices.end());
  else if (s == "active_edges") out.assign(o.activeEdges.begin(), o.activeEdges.end());
  else if (s == "index_mapping") out.assign(o.ivMap.begin(), o.ivMap.end());
  else if (s == "dims") { out = {o.numPoses, o.numLandmarks, o.sizePoses, o.sizeLandmarks}; }
  else if (s == "pose_block_indices") out.assign(o.Hpp.colBlockIndices.begin(), o.Hpp.colBlockIndices.end());
  else if (s == "landmark_block_indices") out.assign(o.Hll.colBlockIndices.begin(), o.Hll.colBlockIndices.end());
  else if (s == "hpp_colptr") ccsOf(o.Hpp, out, true); else if (s == "hpp_rowidx") ccsOf(o.Hpp, out, false);
  else if (s == "hpl_colptr") ccsOf(o.Hpl, out, true); else if (s == "hpl_rowidx") ccsOf(o.Hpl, out, false);
  else if (s == "hll_colptr") ccsOf(o.Hll, out, true); else if (s == "hll_rowidx") ccsOf(o.Hll, out, false);
  else if (s == "hschur_colptr") ccsOf(o.Hschur, out, true); else if (s == "hschur_rowidx") ccsOf(o.Hschur, out, false);
  else if (s == "hschur_t_colptr") { int n2 = 0; for (auto& c : o.HschurTransposedCCS) { out.push_back(n2); n2 += (int)c.size(); } out.push_back(n2); }
  else if (s == "hschur_t_rowidx") { for (auto& c : o.HschurTransposedCCS) for (auto& rb : c) out.push_back(rb.row); }
  else if (s == "edge_targets") {
    // per active edge: matrix id (0 Hpp, 1 Hll, 2 Hpl, -1 none), block row, block col, transposed
    auto find = [&](const SparseBlockMatrix& M, const double* p, int& r, int& c) { for (size_t cc = 0; cc < M.blockCols.size(); ++cc) for (auto& kv : M.blockCols[cc]) if (M.blocks[kv.second].data() == p) { r = kv.first; c = (int)cc; return true; } return false; };
    // build pointer maps once
    std::unordered_map<const double*, std::array<int,3>> where;
    const SparseBlockMatrix* Ms[3] = {&o.Hpp, &o.Hll, &o.Hpl};
    for (int m = 0; m < 3; ++m) for (size_t cc = 0; cc < Ms[m]->blockCols.size(); ++cc) for (auto& kv : Ms[m]->blockCols[cc]) where[Ms[m]->blocks[kv.second].data()] = {m, kv.first, (int)cc};
    (void)find;
    for (int ei : o.activeEdges) { const Edge& e = o.edges[ei]; if (!e.hessian) { out.insert(out.end(), {-1, -1, -1, 0}); continue; } auto w = where[e.hessian]; out.insert(out.end(), {w[0], w[1], w[2], e.hessianRowMajor ? 1 : 0}); }
  }
  else { *n = -1; return nullptr; }
  *n = (int64_t)out.size(); return out.data();
}
const double* orc_get_f64(void* hh, const char* name, int64_t* n) {
  Handle* h = (Handle*)hh; Optimizer& o = h->opt; std::string s(name); auto& out = h->dbuf; out.clear();
  if (s == "x") out = o.x; else if (s == "b") out = o.b;
  else if (s == "bschur") out.assign(o.bschur.begin(), o.bschur.begin() + o.sizePoses);
  else if (s == "hpp_values") valuesOf(o.Hpp, out); else if (s == "hpl_values") valuesOf(o.Hpl, out);
  else if (s == "hll_values") valuesOf(o.Hll, out); else if (s == "hschur_values") valuesOf(o.Hschur, out);
  else if (s == "estimates") for (auto& v : o.vertices) out.insert(out.end(), v.est.begin(), v.est.end());
  else if (s == "errors") for (int ei : o.activeEdges) { const Edge& e = o.edges[ei]; out.insert(out.end(), e.err, e.err + e.dim); }
  else if (s == "jacobians") for (int ei : o.activeEdges) { const Edge& e = o.edges[ei]; out.insert(out.end(), e.J0, e.J0 + e.dim*o.vertices[e.v[0]].dim); out.insert(out.end(), e.J1, e.J1 + e.dim*o.vertices[e.v[1]].dim); }
  else if (s == "lambda") out = {o.currentLambda};
  else if (s == "pcg_state") { auto* p = dynamic_cast<LinearSolverPCG*>(o.linearSolver.get()); if (p) out = {p->residual, (double)p->lastIterations}; else { *n = -1; return nullptr; } }   // _residual, iterations of the last solve
  else if (s == "dogleg") out = {o.dlDelta, (double)o.dlLastStep, (double)o.dlLastNumTries, o.dlCurrentLambda, o.dlWasPDInAllIterations ? 1.0 : 0.0};   // trustRegion(), lastStep(), tries, damping, PD flag
  else { *n = -1; return nullptr; }
  *n = (int64_t)out.size(); return out.data();
}
void orc_set_estimates(void* hh, const double* est) {
  Optimizer& o = ((Handle*)hh)->opt; size_t off = 0;
  for (auto& v : o.vertices) { std::memcpy(v.est.data(), est + off, sizeof(double)*v.est.size()); off += v.est.size(); }
}

// stateless per-edge entry points used by the Jacobian property tests
void orc_edge_error(int etype, const double* x0, const double* x1, const double* z, const double* prm, double* e) { computeError(etype, x0, x1, z, prm, e); }
void orc_edge_jacobian(int etype, const double* x0, const double* x1, const double* z, const double* prm, double* J0, double* J1) { linearizeOplus(etype, x0, x1, z, prm, J0, J1); }
void orc_edge_jacobian_numeric(int etype, const double* x0, const double* x1, const double* z, const double* prm, double* J0, double* J1) { linearizeOplusNumeric(etype, x0, x1, z, prm, J0, J1); }
void orc_vertex_oplus(int vtype, double* est, const double* upd, int* counter) { oplus(vtype, est, upd, counter); }
void orc_dq_dR(const double* R9, double* out27) { M3 R; std::memcpy(R.m, R9, 72); compute_dq_dR(out27, R); }
void orc_quat_from_R(const double* R9, double* q4) { M3 R; std::memcpy(R.m, R9, 72); Quat q = fromRotationMatrix(R); q4[0]=q.x; q4[1]=q.y; q4[2]=q.z; q4[3]=q.w; }
void orc_R_from_quat(const double* q4, double* R9) { M3 R = toRotationMatrix(Quat{q4[0],q4[1],q4[2],q4[3]}); std::memcpy(R9, R.m, 72); }
void orc_robustify(int kind, double delta, double e2, double* rho) { robustify(kind, delta, e2, rho); }
void orc_se3_exp(const double* u6, double* v7) { SE3Quat::exp(u6).toVector(v7); }
void orc_se3_log(const double* v7, double* u6) { SE3Quat::fromVectorRaw(v7).log(u6); }

}  // extern "C"
