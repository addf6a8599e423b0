"""TEST INFRASTRUCTURE — ctypes binding of the CPU oracle (oracle/g2o_oracle.cpp).

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this module.  Nothing under ``g2o_b200/`` does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

STATS_FIELDS = ["iteration", "numVertices", "numEdges", "chi2", "timeResiduals", "timeLinearize", "timeQuadraticForm",
                "levenbergIterations", "timeSchurComplement", "timeSymbolicDecomposition", "timeNumericDecomposition",
                "timeLinearSolution", "timeLinearSolver", "iterationsLinearSolver", "timeUpdate", "timeIteration",
                "hessianDimension", "hessianPoseDimension", "hessianLandmarkDimension", "choleskyNNZ", "lambda", "result"]


def build(force: bool = False) -> None:
    """Compile liboracle.so (and oracle/_ref/libcsparse_ref.so when /root/reference is present)."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("g2o_oracle.cpp", "orc_types.hpp", "orc_math.hpp")]
    stale = force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", _HERE, os.path.join(_HERE, "liboracle.so")] + (["-B"] if force else []), check=True, capture_output=True)
    # the Eigen stand-in's own algorithms as a probe library (checked against numpy / scipy in tests/test_reference_leaves.py); no dependency
    subprocess.run(["make", "-C", _HERE, os.path.join(_HERE, "libshim_probe.so")], check=True, capture_output=True)
    # pieces of the reference itself, compiled from /root/reference when that tree is present (the GPU box only has the prebuilt files).  A failure
    # here must not take the oracle down with it: the tests that need these libraries skip when they are missing.
    if not os.path.isdir("/root/reference/g2o/core"):
        return
    ref = os.path.join(_HERE, "_ref")
    shim = [os.path.join(_HERE, f) for f in ("ref_leaves.cpp", "ref_core.cpp", "ref_core_bal.cpp", "Makefile")] + \
           [os.path.join(_HERE, "eigen_shim", "Eigen", f) for f in os.listdir(os.path.join(_HERE, "eigen_shim", "Eigen"))]

    def outdated(name):
        so = os.path.join(ref, name)
        return force or not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in shim)

    shim.append(os.path.join(os.path.dirname(_HERE), "g2o_b200", "host", "real_g2o_adapter", "solver_cuda.cpp"))
    # ref also builds libcsparse_ref.so; ref_adapter = the CUDA plugin for the real g2o against the reference's headers (needs libg2ocu.so)
    for target, name in (("ref", "libg2o_ref_leaves.so"), ("ref_core", "libg2o_ref_core.so"), ("ref_adapter", "libg2o_solver_cuda.so"), ("ref_tools", "create_sphere")):
        if target == "ref_adapter" and not os.path.exists(os.path.join(os.path.dirname(_HERE), "g2o_b200", "lib", "libg2ocu.so")):
            continue
        if outdated(name) or not os.path.exists(os.path.join(ref, "libcsparse_ref.so")):
            r = subprocess.run(["make", "-C", _HERE, target], capture_output=True, text=True)
            if r.returncode != 0 or not os.path.exists(os.path.join(ref, name)):
                import sys
                print(f"oracle.build: `make {target}` failed, continuing without oracle/_ref/{name}:\n{r.stdout[-1500:]}{r.stderr[-1500:]}", file=sys.stderr)


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = ctypes.CDLL(so)
        L.orc_create.restype = ctypes.c_void_p
        L.orc_create.argtypes = [ctypes.c_void_p]
        for f in ("orc_destroy", "orc_compute_active_errors", "orc_build_system", "orc_restore_diagonal", "orc_push", "orc_pop", "orc_discard_top"):
            getattr(L, f).argtypes = [ctypes.c_void_p]
            getattr(L, f).restype = None
        for f in ("orc_active_robust_chi2", "orc_active_chi2", "orc_compute_lambda_init", "orc_compute_scale"):
            getattr(L, f).argtypes = [ctypes.c_void_p]
            getattr(L, f).restype = ctypes.c_double
        for f in ("orc_algorithm_init", "orc_build_structure", "orc_solve", "orc_do_schur"):
            getattr(L, f).argtypes = [ctypes.c_void_p]
            getattr(L, f).restype = ctypes.c_int
        L.orc_set_num_threads.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.orc_set_solver.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p]
        L.orc_set_lm_params.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_int]
        L.orc_set_dogleg_params.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_int, ctypes.c_double, ctypes.c_double]
        L.orc_multiply_hessian.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.orc_multiply_hessian.restype = None
        L.orc_set_pcg_params.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_int, ctypes.c_int]
        L.orc_initialize_optimization.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.orc_optimize.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.orc_set_lambda.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_int]
        L.orc_set_lambda.restype = None
        L.orc_update.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.orc_update.restype = None
        L.orc_set_estimates.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.orc_set_estimates.restype = None
        L.orc_get_i32.restype = ctypes.POINTER(ctypes.c_int32)
        L.orc_get_i32.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int64)]
        L.orc_get_f64.restype = ctypes.POINTER(ctypes.c_double)
        L.orc_get_f64.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int64)]
        L.orc_load_csparse.argtypes = [ctypes.c_char_p]
        ref = os.path.join(_HERE, "_ref", "libcsparse_ref.so")
        L._has_csparse = bool(os.path.exists(ref) and L.orc_load_csparse(ref.encode()))
        _LIB = L
    return _LIB


_LEAVES = None


def shim_probe() -> ctypes.CDLL:
    """oracle/libshim_probe.so: the numerical routines of oracle/eigen_shim (the stand-in the reference is compiled against), one export each."""
    build()
    L = ctypes.CDLL(os.path.join(_HERE, "libshim_probe.so"))
    L.shim_determinant.restype = ctypes.c_double
    L.shim_determinant.argtypes = [ctypes.c_int, ctypes.c_void_p]
    L.shim_angle_axis_R.argtypes = [ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p]
    L.shim_rotation2d.argtypes = [ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    return L


def reference_leaves():
    """ctypes handle of oracle/_ref/libg2o_ref_leaves.so - leaf functions of the REAL reference (robust kernels, dq/dR, normalize_theta)
    compiled from /root/reference by oracle/Makefile - or None when it was not built (no reference tree at build time)."""
    global _LEAVES
    if _LEAVES is None:
        so = os.path.join(_HERE, "_ref", "libg2o_ref_leaves.so")
        if not os.path.exists(so):
            return None
        L = ctypes.CDLL(so)
        L.ref_robustify.argtypes = [ctypes.c_char_p, ctypes.c_double, ctypes.c_double, ctypes.c_void_p]
        L.ref_robustify.restype = ctypes.c_int
        L.ref_dq_dR.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.ref_dq_dR.restype = None
        L.ref_normalize_theta.argtypes = [ctypes.c_double]
        L.ref_normalize_theta.restype = ctypes.c_double
        for f in ("ref_edge_se2_error", "ref_edge_se2_pointxy_error"):
            getattr(L, f).argtypes = [ctypes.c_void_p] * 4
            getattr(L, f).restype = None
        L.ref_vertex_se2_oplus.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.ref_vertex_se2_oplus.restype = None
        L.ref_pcg_create.argtypes = [ctypes.c_int]; L.ref_pcg_create.restype = ctypes.c_void_p
        L.ref_pcg_destroy.argtypes = [ctypes.c_void_p]; L.ref_pcg_destroy.restype = None
        L.ref_pcg_init.argtypes = [ctypes.c_void_p]; L.ref_pcg_init.restype = None
        L.ref_pcg_solve.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 6 + [ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        L.ref_pcg_solve.restype = ctypes.c_int
        L.ref_multiply_symmetric_upper.argtypes = [ctypes.c_void_p] * 3; L.ref_multiply_symmetric_upper.restype = None
        for f, n in (("ref_se3quat_exp", 2), ("ref_se3quat_log", 2), ("ref_vertex_se3expmap_oplus", 2), ("ref_edge_se3expmap", 6), ("ref_edge_project_xyz2uv_error", 5)):
            getattr(L, f).argtypes = [ctypes.c_void_p] * n
            getattr(L, f).restype = None
        L.ref_sample_gaussian_two_engines.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.ref_sample_gaussian_two_engines.restype = None
        _LEAVES = L
    return _LEAVES


_CORE = None


def reference_core():
    """ctypes handle of oracle/_ref/libg2o_ref_core.so - the REAL reference (g2o/core, BlockSolver, LM / GN / Dogleg, LinearSolverPCG, slam2d
    types) compiled from /root/reference by `make -C oracle ref_core` - or None when it was not built."""
    global _CORE
    if _CORE is None:
        so = os.path.join(_HERE, "_ref", "libg2o_ref_core.so")
        if not os.path.exists(so):
            return None
        L = ctypes.CDLL(so)
        L.refcore_create.restype = ctypes.c_void_p
        L.refcore_create.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p]
        L.refcore_destroy.argtypes = [ctypes.c_void_p]; L.refcore_destroy.restype = None
        L.refcore_set_num_threads.argtypes = [ctypes.c_int]; L.refcore_set_num_threads.restype = None
        L.refcore_set_pcg.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_int, ctypes.c_int]; L.refcore_set_pcg.restype = None
        L.refcore_initialize_optimization.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.refcore_optimize.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        for f in ("refcore_current_lambda", "refcore_active_robust_chi2", "refcore_active_chi2"):
            getattr(L, f).argtypes = [ctypes.c_void_p]; getattr(L, f).restype = ctypes.c_double
        for f in ("refcore_dogleg_state", "refcore_hessian_index", "refcore_estimates"):
            getattr(L, f).argtypes = [ctypes.c_void_p, ctypes.c_void_p]; getattr(L, f).restype = None
        L.refcore_optimize_budget.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_double]
        L.refcore_linearize.argtypes = [ctypes.c_void_p, ctypes.c_double]; L.refcore_linearize.restype = ctypes.c_int
        for f in ("refcore_structure_i32", "refcore_structure_f64"):
            getattr(L, f).argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int64]; getattr(L, f).restype = ctypes.c_int64
        _CORE = L
    return _CORE


class ReferenceG2o:
    """The real reference on a 2-D SLAM graph: ``g2o::SparseOptimizer`` + ``BlockSolver`` (``"3_2"`` or ``"var"``) + ``LinearSolverPCG`` +
    Levenberg / Gauss-Newton / Dogleg, through oracle/ref_core.cpp."""

    def __init__(self, graph, algorithm: str = "lm", block_solver: str = "3_2", threads: int = 0):
        """``threads`` > 0 fixes the number of OpenMP threads of the reference (1 = deterministic summation order); 0 leaves the default (all)."""
        self._L = reference_core()
        if self._L is None:
            raise RuntimeError("oracle/_ref/libg2o_ref_core.so has not been built")
        self._L.refcore_set_num_threads(int(threads))
        self.graph = graph
        cg = graph.as_c()
        self._h = self._L.refcore_create(ctypes.byref(cg), algorithm.encode(), block_solver.encode())
        if not self._h:
            raise ValueError("the reference library was built with the slam2d types only / unknown solver")

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.refcore_destroy(self._h)
            self._h = None

    def initialize_optimization(self, level: int = 0) -> bool:
        return bool(self._L.refcore_initialize_optimization(self._h, level))

    def optimize(self, iterations: int, budget_seconds: float | None = None):
        """``SparseOptimizer::optimize(iterations)``; with ``budget_seconds`` the reference's own forceStopFlag is raised after that time, so
        that no further iteration starts (the returned count / stats then cover the iterations that ran)."""
        buf = np.zeros((max(iterations, 1), 13))
        if budget_seconds is None:
            n = self._L.refcore_optimize(self._h, iterations, _dp(buf))
        else:
            n = self._L.refcore_optimize_budget(self._h, iterations, _dp(buf), float(budget_seconds))
        keys = ("chi2", "levenbergIterations", "iterationsLinearSolver", "hessianPoseDimension", "hessianLandmarkDimension", "iteration", "timeIteration", "timeLinearSolution",
                "timeResiduals", "timeQuadraticForm", "timeSchurComplement", "timeLinearSolver", "timeUpdate")
        return n, [dict(zip(keys, row)) for row in buf[:max(n, 0)]]

    def set_pcg_params(self, tol=1e-6, max_iter=-1, absolute=True):
        self._L.refcore_set_pcg(self._h, tol, max_iter, int(absolute))

    def current_lambda(self) -> float: return self._L.refcore_current_lambda(self._h)
    def active_robust_chi2(self) -> float: return self._L.refcore_active_robust_chi2(self._h)
    def active_chi2(self) -> float: return self._L.refcore_active_chi2(self._h)

    def dogleg_state(self) -> dict:
        d = np.zeros(2); self._L.refcore_dogleg_state(self._h, _dp(d)); return {"delta": d[0], "last_step": int(d[1])}

    def hessian_index(self) -> np.ndarray:
        out = np.zeros(self.graph.n_vertices, dtype=np.int32); self._L.refcore_hessian_index(self._h, _dp(out)); return out

    def estimates(self) -> np.ndarray:
        out = np.zeros_like(self.graph.v_estimate); self._L.refcore_estimates(self._h, _dp(out)); return out

    def linearize(self, lam: float) -> bool:
        """init, buildStructure, computeActiveErrors, buildSystem, setLambda(lam, true), solve, restoreDiagonal through the reference's own
        Solver virtuals; the BlockSolver's matrices can be read afterwards with structure_i32 / structure_f64."""
        return self._L.refcore_linearize(self._h, float(lam)) == 1

    def compute_marginals(self, pairs):
        """SparseOptimizer::computeMarginals (sparse_optimizer.cpp:594-596) of the compiled reference; needs a `*_csparse` block solver (the PCG
        linear solver has no solvePattern, linear_solver.h:89-98).  Returns the blocks (column-major MatrixX -> 2-d arrays) or None."""
        rows = np.array([p[0] for p in pairs], dtype=np.int32); cols = np.array([p[1] for p in pairs], dtype=np.int32)
        cap = 81 * max(len(pairs), 1); out = np.zeros(cap)
        self._L.refcore_compute_marginals.restype = ctypes.c_int
        self._L.refcore_compute_marginals.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
        rc = self._L.refcore_compute_marginals(self._h, len(pairs), _dp(rows), _dp(cols), _dp(out), cap)
        if rc != 1:
            return None
        VERTEX_DIM = (0, 3, 2, 6, 6, 3, 9, 3)   # minimal dimension by vertex type code (include/g2ocu.h), as the reference's BaseVertex<D, T>
        hi = self.hessian_index(); dim_of = {int(hi[v]): int(VERTEX_DIM[int(self.graph.v_type[v])]) for v in range(len(hi)) if hi[v] >= 0}
        blocks, off = [], 0
        for r, c in pairs:
            dr, dc = dim_of[int(r)], dim_of[int(c)]
            blocks.append(out[off:off + dr * dc].reshape(dr, dc, order="F").copy()); off += dr * dc
        return blocks

    def structure_i32(self, name: str) -> np.ndarray:
        """Block pattern arrays of the reference's BlockSolver (its protected _Hpp / _Hll / _Hpl / _Hschur / _HschurTransposedCCS), in the format of
        g2ocu_get_i32: pose_block_indices, landmark_block_indices, hpp_/hpl_/hll_/hschur_/hschur_t_ colptr + rowidx, dims."""
        n = self._L.refcore_structure_i32(self._h, name.encode(), None, 0)
        if n < 0:
            raise KeyError(name)
        out = np.zeros(n, dtype=np.int32); self._L.refcore_structure_i32(self._h, name.encode(), _dp(out), n); return out

    def structure_f64(self, name: str) -> np.ndarray:
        """hpp_values, hpl_values, hll_values, hschur_values (blocks in CCS order, column-major), b, x, bschur of the reference's BlockSolver."""
        n = self._L.refcore_structure_f64(self._h, name.encode(), None, 0)
        if n < 0:
            raise KeyError(name)
        out = np.zeros(n); self._L.refcore_structure_f64(self._h, name.encode(), _dp(out), n); return out


def has_csparse() -> bool:
    return lib()._has_csparse


def max_threads() -> int:
    return int(lib().orc_max_threads())


def _dp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class Oracle:
    """One reference ``SparseOptimizer`` + algorithm + ``BlockSolver`` + linear solver, CPU."""

    def __init__(self, graph, algorithm: str = "lm", linear: str = "pcg", threads: int = 1):
        self._L = lib()
        self.graph = graph
        cg = graph.as_c()
        self._h = self._L.orc_create(ctypes.byref(cg))
        if not self._h:
            raise ValueError("oracle rejected the graph (unknown type or vertex/edge type mismatch)")
        rc = self._L.orc_set_solver(self._h, algorithm.encode(), linear.encode())
        if rc != 0:
            raise ValueError(f"oracle: cannot configure solver {algorithm}/{linear} (rc={rc})")
        self._L.orc_set_num_threads(self._h, threads)

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orc_destroy(self._h)
            self._h = None

    # SparseOptimizer
    def initialize_optimization(self, level: int = 0) -> bool:
        return bool(self._L.orc_initialize_optimization(self._h, level))

    def optimize(self, iterations: int):
        stride = self._L.orc_stats_stride()
        buf = np.zeros((max(iterations, 1), stride))
        n = self._L.orc_optimize(self._h, iterations, _dp(buf))
        stats = [dict(zip(STATS_FIELDS, row)) for row in buf[:max(n, 0)]]
        return n, stats

    def set_lm_params(self, initial_lambda=0.0, max_trials=10):
        self._L.orc_set_lm_params(self._h, initial_lambda, max_trials)

    def set_dogleg_params(self, initial_delta=1e4, max_trials=100, initial_lambda=1e-7, lambda_factor=10.0):
        self._L.orc_set_dogleg_params(self._h, initial_delta, max_trials, initial_lambda, lambda_factor)

    def dogleg_state(self) -> dict:
        d = self.get_f64("dogleg")
        return {"delta": d[0], "last_step": int(d[1]), "tries": int(d[2]), "lambda": d[3], "was_pd": bool(d[4])}

    def multiply_hessian(self, src) -> np.ndarray:
        src = np.ascontiguousarray(src, dtype=np.float64); dst = np.zeros_like(src)
        self._L.orc_multiply_hessian(self._h, _dp(dst), _dp(src))
        return dst

    def set_pcg_params(self, tol=1e-6, max_iter=-1, absolute=True):
        self._L.orc_set_pcg_params(self._h, tol, max_iter, int(absolute))

    def compute_active_errors(self): self._L.orc_compute_active_errors(self._h)
    def active_robust_chi2(self) -> float: return self._L.orc_active_robust_chi2(self._h)
    def active_chi2(self) -> float: return self._L.orc_active_chi2(self._h)
    def update(self, x=None):
        self._L.orc_update(self._h, None if x is None else _dp(np.ascontiguousarray(x, dtype=np.float64)))
    def push(self): self._L.orc_push(self._h)
    def pop(self): self._L.orc_pop(self._h)
    def discard_top(self): self._L.orc_discard_top(self._h)

    # OptimizationAlgorithm / Solver
    def algorithm_init(self) -> bool: return bool(self._L.orc_algorithm_init(self._h))
    def build_structure(self) -> bool: return bool(self._L.orc_build_structure(self._h))
    def build_system(self): self._L.orc_build_system(self._h)
    def set_lambda(self, lam: float, backup: bool = True): self._L.orc_set_lambda(self._h, lam, int(backup))
    def restore_diagonal(self): self._L.orc_restore_diagonal(self._h)
    def solve(self) -> bool: return bool(self._L.orc_solve(self._h))
    def compute_lambda_init(self) -> float: return self._L.orc_compute_lambda_init(self._h)
    def compute_scale(self) -> float: return self._L.orc_compute_scale(self._h)
    def do_schur(self) -> bool: return bool(self._L.orc_do_schur(self._h))

    def get_i32(self, name: str) -> np.ndarray:
        n = ctypes.c_int64()
        p = self._L.orc_get_i32(self._h, name.encode(), ctypes.byref(n))
        if n.value < 0:
            raise KeyError(name)
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy() if n.value else np.zeros(0, dtype=np.int32)

    def get_f64(self, name: str) -> np.ndarray:
        n = ctypes.c_int64()
        p = self._L.orc_get_f64(self._h, name.encode(), ctypes.byref(n))
        if n.value < 0:
            raise KeyError(name)
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy() if n.value else np.zeros(0)

    def dense_hpp(self) -> np.ndarray:
        """The reference's Hpp (the pose block; the whole system when no point is marginalized) as a dense symmetric matrix, from its upper blocks
        in CCS order (sparse_block_matrix_ccs.h:48-199), block sizes from the cumulative block ends."""
        ends = self.get_i32("pose_block_indices").astype(np.int64); starts = np.concatenate([[0], ends[:-1]])
        colptr, rowidx, vals = self.get_i32("hpp_colptr"), self.get_i32("hpp_rowidx"), self.get_f64("hpp_values")
        n = int(ends[-1]); H = np.zeros((n, n)); off = 0
        for c in range(len(ends)):
            dc = int(ends[c] - starts[c])
            for k in range(int(colptr[c]), int(colptr[c + 1])):
                r = int(rowidx[k]); dr = int(ends[r] - starts[r])
                blk = vals[off:off + dr * dc].reshape(dr, dc, order="F"); off += dr * dc
                H[starts[r]:ends[r], starts[c]:ends[c]] = blk
                H[starts[c]:ends[c], starts[r]:ends[r]] = blk.T
        return H

    def compute_marginals(self, pairs):
        """SparseOptimizer::computeMarginals (sparse_optimizer.cpp:594-596 -> block_solver.hpp:451-459 -> LinearSolver::solvePattern on Hpp): the
        blocks (row, col) of the inverse of Hpp as it stands.  The reference's CSparse / CHOLMOD back-ends get the requested entries from the
        Cholesky factor by the Takahashi recursion (marginal_covariance_cholesky.cpp:52-100, 153-222), which yields entries of the exact inverse;
        restated here as that definition: Cholesky (numpy), inverse, blocks.  None when Hpp is not positive definite (solvePattern == false)."""
        H = self.dense_hpp()
        try:
            Lc = np.linalg.cholesky(H)
        except np.linalg.LinAlgError:
            return None
        Li = np.linalg.inv(Lc); inv = Li.T @ Li
        ends = self.get_i32("pose_block_indices").astype(np.int64); starts = np.concatenate([[0], ends[:-1]])
        return [inv[starts[r]:ends[r], starts[c]:ends[c]].copy() for r, c in pairs]

    def set_estimates(self, est: np.ndarray):
        self._L.orc_set_estimates(self._h, _dp(np.ascontiguousarray(est, dtype=np.float64)))

    def estimates(self) -> np.ndarray:
        return self.get_f64("estimates")


# ---- stateless per-edge helpers (Jacobian / mapping property tests) ----
def _out(n):
    return np.zeros(n)


def edge_error(etype, x0, x1, z, prm=None):
    from g2o_b200.graph import EDGE_DIM
    L = lib(); e = _out(int(EDGE_DIM[etype]))
    prm = np.zeros(4) if prm is None else np.ascontiguousarray(prm, dtype=np.float64)
    L.orc_edge_error(ctypes.c_int(etype), _dp(np.ascontiguousarray(x0)), _dp(np.ascontiguousarray(x1)), _dp(np.ascontiguousarray(z)), _dp(prm), _dp(e))
    return e


def edge_jacobian(etype, x0, x1, z, prm=None, numeric=False):
    from g2o_b200.graph import EDGE_DIM, EDGE_VERTEX_TYPES, VERTEX_DIM
    L = lib(); E = int(EDGE_DIM[etype]); t0, t1 = EDGE_VERTEX_TYPES[etype]
    J0 = _out(E * int(VERTEX_DIM[t0])); J1 = _out(E * int(VERTEX_DIM[t1]))
    prm = np.zeros(4) if prm is None else np.ascontiguousarray(prm, dtype=np.float64)
    f = L.orc_edge_jacobian_numeric if numeric else L.orc_edge_jacobian
    f(ctypes.c_int(etype), _dp(np.ascontiguousarray(x0)), _dp(np.ascontiguousarray(x1)), _dp(np.ascontiguousarray(z)), _dp(prm), _dp(J0), _dp(J1))
    return J0.reshape(-1, E).T.copy(), J1.reshape(-1, E).T.copy()   # (E x D) matrices


def vertex_oplus(vtype, est, upd, counter=0):
    L = lib(); est = np.array(est, dtype=np.float64); c = ctypes.c_int(counter)
    L.orc_vertex_oplus(ctypes.c_int(vtype), _dp(est), _dp(np.ascontiguousarray(upd, dtype=np.float64)), ctypes.byref(c))
    return est, c.value


def dq_dR(R):
    out = _out(27); lib().orc_dq_dR(_dp(np.asfortranarray(R).ravel(order="F").copy()), _dp(out))
    return out.reshape(9, 3).T.copy()   # 3 x 9


def quat_from_R(R):
    q = _out(4); lib().orc_quat_from_R(_dp(np.asfortranarray(R).ravel(order="F").copy()), _dp(q)); return q


def R_from_quat(q):
    R = _out(9); lib().orc_R_from_quat(_dp(np.ascontiguousarray(q, dtype=np.float64)), _dp(R)); return R.reshape(3, 3).T.copy()


def robustify(kind, delta, e2):
    rho = _out(3); lib().orc_robustify(ctypes.c_int(kind), ctypes.c_double(delta), ctypes.c_double(e2), _dp(rho)); return rho


def se3_exp(u):
    v = _out(7); lib().orc_se3_exp(_dp(np.ascontiguousarray(u, dtype=np.float64)), _dp(v)); return v


def se3_log(v):
    u = _out(6); lib().orc_se3_log(_dp(np.ascontiguousarray(v, dtype=np.float64)), _dp(u)); return u
