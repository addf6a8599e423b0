"""Thin Python binding over the C ABI (include/g2ocu.h) — the same calls the C++ adapter classes in
``g2o_b200/host`` make.  Used by the tests and by ``bench.py``; it adds no arithmetic of its own.

Naming follows the reference's plugin interface: ``SparseOptimizer`` (``initialize_optimization``, ``optimize``,
``compute_active_errors``, ``active_robust_chi2``, ``update``, ``push``/``pop``/``discard_top``), ``Solver``
(``build_structure``, ``build_system``, ``set_lambda``, ``restore_diagonal``, ``solve``, ``x``, ``b``) and
``OptimizationAlgorithmFactory`` solver names (``lm_var_cuda``, ``lm_fix6_3_cuda``, ...).
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .graph import Graph

# name -> (algorithm, poseDim, landmarkDim, requiresMarginalize); mirrors solvers/pcg/solver_pcg.cpp:41-98 and
# solvers/csparse/solver_csparse.cpp:54-117 (the reference has no named 9_3 solver: bal_example.cpp:301 instantiates it).  The reference's
# fix7_3 names (sim3) are not registered: the backend has no 7-dof types, and a name that cannot solve anything is worse than an unknown one.
SOLVER_NAMES = {
    "gn_var_cuda": ("gn", -1, -1, False), "lm_var_cuda": ("lm", -1, -1, False),
    "gn_fix3_2_cuda": ("gn", 3, 2, True), "lm_fix3_2_cuda": ("lm", 3, 2, True),
    "gn_fix6_3_cuda": ("gn", 6, 3, True), "lm_fix6_3_cuda": ("lm", 6, 3, True),
    "gn_fix9_3_cuda": ("gn", 9, 3, True), "lm_fix9_3_cuda": ("lm", 9, 3, True),
    # solvers/dense/solver_dense.cpp:91-98 (+ 9_3): BlockSolver + LinearSolverDense -> device DMMA Cholesky
    "gn_dense_cuda": ("gn", -1, -1, False), "lm_dense_cuda": ("lm", -1, -1, False),
    "gn_dense3_2_cuda": ("gn", 3, 2, True), "lm_dense3_2_cuda": ("lm", 3, 2, True),
    "gn_dense6_3_cuda": ("gn", 6, 3, True), "lm_dense6_3_cuda": ("lm", 6, 3, True),
    "gn_dense9_3_cuda": ("gn", 9, 3, True), "lm_dense9_3_cuda": ("lm", 9, 3, True),
    # Powell's dogleg is registered for the variable-size block solver only (solvers/csparse/solver_csparse.cpp:117, dl_var)
    "dl_var_cuda": ("dl", -1, -1, False),
}
_ALGORITHM_CODES = {"gn": _lib.ALGORITHM_GN, "lm": _lib.ALGORITHM_LM, "dl": _lib.ALGORITHM_DOGLEG}
DOGLEG_STEPS = {0: "Undefined", 1: "Descent", 2: "GN", 3: "Dogleg"}   # OptimizationAlgorithmDogleg::stepType2Str, dogleg.cpp:209-217


class G2oCudaError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"g2ocu error {code}: {message}")
        self.code = code


class CudaSolver:
    """One ``g2ocu_solver`` handle = one optimizer on one CUDA stream."""

    def __init__(self, graph: Graph | None = None, solver_name: str = "lm_var_cuda", linear: str = "pcg", device: int = -1,
                 pcg_tolerance: float = 1e-6, pcg_max_iterations: int = -1, pcg_absolute_tolerance: bool = True, stream: int = 0):
        if solver_name not in SOLVER_NAMES:
            raise KeyError(f"unknown solver {solver_name!r}; registered: {sorted(SOLVER_NAMES)}")
        self.solver_name = solver_name
        self.algorithm = _ALGORITHM_CODES[SOLVER_NAMES[solver_name][0]]
        self._L = _lib.lib()
        cfg = _lib.Config()
        self._L.g2ocu_default_config(ctypes.byref(cfg))
        cfg.device = device
        if "_dense" in solver_name:
            linear = "dense"
        cfg.linear_solver = {"pcg": _lib.LINEAR_PCG, "dense": _lib.LINEAR_DENSE}[linear]
        cfg.pcg_tolerance = pcg_tolerance
        cfg.pcg_max_iterations = pcg_max_iterations
        cfg.pcg_absolute_tolerance = int(pcg_absolute_tolerance)
        cfg.stream = stream or None
        h = ctypes.c_void_p()
        rc = self._L.g2ocu_create(ctypes.byref(cfg), ctypes.byref(h))
        if rc != 0:
            raise G2oCudaError(rc, self._L.g2ocu_last_error(None).decode())
        self._h = h
        self.graph = None
        self._hook = None
        if graph is not None:
            self.set_graph(graph)

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.g2ocu_destroy(self._h)
            self._h = None

    def _ck(self, rc):
        if rc < 0:
            raise G2oCudaError(int(rc), self._L.g2ocu_last_error(self._h).decode())
        return rc

    # ---- graph ----
    def set_graph(self, graph: Graph):
        # BlockSolver<BlockSolverTraits<p,l>> holds blocks of exactly these sizes: build_structure rejects a graph with others
        pd, ld = SOLVER_NAMES[self.solver_name][1:3]
        self.set_property("poseDim", pd); self.set_property("landmarkDim", ld)
        self.graph = graph
        cg = graph.as_c()
        self._ck(self._L.g2ocu_set_graph(self._h, ctypes.byref(cg)))

    def set_force_stop_flag(self, flag):
        """``SparseOptimizer::setForceStopFlag`` (sparse_optimizer.h:186-190): ``flag`` is a ``ctypes.c_ubyte`` owned by the caller (kept alive
        here); while it is non-zero no further iteration / LM trial starts.  ``None`` removes it."""
        self._stop = flag
        self._ck(self._L.g2ocu_set_force_stop_flag(self._h, ctypes.addressof(flag) if flag is not None else None))

    def set_property(self, name: str, value: float):
        self._ck(self._L.g2ocu_set_property(self._h, name.encode(), float(value)))

    def set_shard(self, rank: int, world: int, allreduce=None):
        """``allreduce(ptr, count, op, stream)`` -> 0 on success; kept alive by this object."""
        if allreduce is not None:
            def _cb(buf, count, op, stream, user):
                try:
                    return int(allreduce(buf, count, op, stream) or 0)
                except Exception:   # never let an exception cross the C boundary
                    import traceback
                    traceback.print_exc()
                    return 1
            self._hook = _lib.ALLREDUCE_FN(_cb)
        else:
            self._hook = _lib.ALLREDUCE_FN(0)
        self._ck(self._L.g2ocu_set_shard(self._h, rank, world, self._hook, None))

    def p2p_export(self) -> bytes:
        buf = ctypes.create_string_buffer(64)
        self._ck(self._L.g2ocu_p2p_export(self._h, buf))
        return buf.raw

    def p2p_import(self, handles: bytes):
        self._ck(self._L.g2ocu_p2p_import(self._h, handles))

    def p2p_export_schur(self) -> bytes:
        buf = ctypes.create_string_buffer(64)
        self._ck(self._L.g2ocu_p2p_export_schur(self._h, buf))
        return buf.raw

    def p2p_import_schur(self, handles: bytes):
        self._ck(self._L.g2ocu_p2p_import_schur(self._h, handles))

    def set_shard_nccl(self, rank: int, world: int, nccl_library: str, unique_id: bytes):
        """Collectives issued by the library itself through NCCL (g2ocu_set_shard_nccl); every rank must call it."""
        assert len(unique_id) == 128
        self._ck(self._L.g2ocu_set_shard_nccl(self._h, rank, world, nccl_library.encode(), unique_id))

    # ---- SparseOptimizer ----
    def initialize_optimization(self, level: int = 0) -> bool:
        self._ck(self._L.g2ocu_initialize_optimization(self._h, level))
        return True

    def optimize(self, iterations: int):
        """Returns (number of iterations performed as ``SparseOptimizer::optimize`` would, per-iteration stats)."""
        stats = (_lib.IterationStats * max(iterations, 1))()
        performed = ctypes.c_int32(-1)
        self._ck(self._L.g2ocu_optimize(self._h, self.algorithm, iterations, stats, ctypes.byref(performed)))
        n = performed.value
        done = max(n, 0) if n != 0 else 0
        out = [stats[i].as_dict() for i in range(iterations) if i < done or (n == 0 and i == 0)]
        return n, out

    def compute_active_errors(self): self._ck(self._L.g2ocu_compute_active_errors(self._h))

    def active_robust_chi2(self) -> float:
        v = ctypes.c_double(); self._ck(self._L.g2ocu_active_robust_chi2(self._h, ctypes.byref(v))); return v.value

    def active_chi2(self) -> float:
        v = ctypes.c_double(); self._ck(self._L.g2ocu_active_chi2(self._h, ctypes.byref(v))); return v.value

    def update(self, x=None):
        if x is None:
            self._ck(self._L.g2ocu_update(self._h, None))
        else:
            x = np.ascontiguousarray(x, dtype=np.float64)
            if x.shape[0] != self.vector_size():
                raise ValueError("update vector has the wrong length")
            self._ck(self._L.g2ocu_update(self._h, x.ctypes.data_as(ctypes.c_void_p)))

    def push(self): self._ck(self._L.g2ocu_push(self._h))
    def pop(self): self._ck(self._L.g2ocu_pop(self._h))
    def discard_top(self): self._ck(self._L.g2ocu_discard_top(self._h))

    # ---- OptimizationAlgorithm / Solver ----
    def init(self, online: bool = False): self._ck(self._L.g2ocu_init(self._h, int(online)))
    def build_structure(self): self._ck(self._L.g2ocu_build_structure(self._h))
    def build_system(self): self._ck(self._L.g2ocu_build_system(self._h))
    def set_lambda(self, lam: float, backup: bool = True): self._ck(self._L.g2ocu_set_lambda(self._h, lam, int(backup)))
    def restore_diagonal(self): self._ck(self._L.g2ocu_restore_diagonal(self._h))

    def solve(self) -> bool:
        ok = ctypes.c_int32(); self._ck(self._L.g2ocu_solve(self._h, ctypes.byref(ok))); return bool(ok.value)

    def solver_iteration(self, iteration: int) -> dict:
        st = _lib.IterationStats()
        self._ck(self._L.g2ocu_solver_iteration(self._h, self.algorithm, iteration, ctypes.byref(st)))
        return st.as_dict()

    def compute_lambda_init(self) -> float:
        v = ctypes.c_double(); self._ck(self._L.g2ocu_compute_lambda_init(self._h, ctypes.byref(v))); return v.value

    def compute_scale(self, lam: float) -> float:
        v = ctypes.c_double(); self._ck(self._L.g2ocu_compute_scale(self._h, lam, ctypes.byref(v))); return v.value

    def dogleg_state(self) -> dict:
        """``trustRegion()``, ``lastStep()``, tries of the last iteration, damping factor, PD flag (optimization_algorithm_dogleg.h:64-68)."""
        d = self.get_f64("dogleg")
        return {"delta": float(d[0]), "last_step": int(d[1]), "tries": int(d[2]), "lambda": float(d[3]), "was_pd": bool(d[4])}

    def multiply_hessian(self, src) -> np.ndarray:
        src = np.ascontiguousarray(src, dtype=np.float64); dst = np.zeros_like(src)
        self._ck(self._L.g2ocu_multiply_hessian(self._h, dst.ctypes.data_as(ctypes.c_void_p), src.ctypes.data_as(ctypes.c_void_p)))
        return dst

    def compute_marginals(self, pairs):
        """``SparseOptimizer::computeMarginals(spinv, blockIndices)`` (sparse_optimizer.cpp:594-596): ``pairs`` = (row, col) hessian indices of pose
        vertices (of any vertex when no point is marginalized: the reference's Hpp is then the whole system); returns the list of blocks of the inverse
        of Hpp, or None where the reference returns false."""
        pairs = [(int(r), int(c)) for r, c in pairs]
        rows = np.array([p[0] for p in pairs], dtype=np.int32); cols = np.array([p[1] for p in pairs], dtype=np.int32)
        ends = self.get_i32("pose_block_indices").astype(np.int64)          # cumulative block ends of the reference's Hpp (all vertices when no point is marginalized)
        dim = np.diff(np.concatenate([[0], ends]))
        if len(pairs) and (rows.min() < 0 or cols.min() < 0 or rows.max() >= len(dim) or cols.max() >= len(dim)):
            raise G2oCudaError(_lib.E_INVALID, f"block index outside Hpp ({len(dim)} block rows)")
        sizes = [int(dim[r] * dim[c]) for r, c in pairs]
        out = np.zeros(max(sum(sizes), 1)); ok = ctypes.c_int32(0)
        self._ck(self._L.g2ocu_compute_marginals(self._h, len(pairs), rows.ctypes.data_as(ctypes.c_void_p), cols.ctypes.data_as(ctypes.c_void_p),
                                                 out.ctypes.data_as(ctypes.c_void_p), ctypes.byref(ok)))
        if not ok.value:
            return None
        blocks, off = [], 0
        for (r, c), sz in zip(pairs, sizes):
            blocks.append(out[off:off + sz].reshape(int(dim[r]), int(dim[c]), order="F").copy()); off += sz
        return blocks

    def vector_size(self) -> int: return int(self._L.g2ocu_vector_size(self._h))
    def x(self) -> np.ndarray: return self.get_f64("x")
    def b(self) -> np.ndarray: return self.get_f64("b")

    # ---- data ----
    def set_estimates(self, est):
        est = np.ascontiguousarray(est, dtype=np.float64)
        if est.shape[0] != self.graph.v_estimate.shape[0]:
            raise ValueError("estimate vector has the wrong length")
        self._ck(self._L.g2ocu_set_estimates(self._h, est.ctypes.data_as(ctypes.c_void_p)))

    def get_estimates(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty_like(self.graph.v_estimate)
        self._ck(self._L.g2ocu_get_estimates(self._h, out.ctypes.data_as(ctypes.c_void_p)))
        return out

    def set_estimates_owned(self, est):
        """Sharded runs: upload every pose and this rank's own landmark range only (``g2ocu_set_estimates_owned``)."""
        est = np.ascontiguousarray(est, dtype=np.float64)
        if est.shape[0] != self.graph.v_estimate.shape[0]:
            raise ValueError("estimate vector has the wrong length")
        self._ck(self._L.g2ocu_set_estimates_owned(self._h, est.ctypes.data_as(ctypes.c_void_p)))

    def get_estimates_owned(self, out: np.ndarray) -> np.ndarray:
        """Sharded runs: read back every pose and this rank's own landmark range into ``out``; the other entries are left as they are."""
        self._ck(self._L.g2ocu_get_estimates_owned(self._h, out.ctypes.data_as(ctypes.c_void_p)))
        return out

    def get_i32(self, name: str) -> np.ndarray:
        n = self._ck(self._L.g2ocu_get_i32(self._h, name.encode(), None, 0))
        out = np.zeros(n, dtype=np.int32)
        if n:
            self._ck(self._L.g2ocu_get_i32(self._h, name.encode(), out.ctypes.data_as(ctypes.c_void_p), n))
        return out

    def get_f64(self, name: str) -> np.ndarray:
        n = self._ck(self._L.g2ocu_get_f64(self._h, name.encode(), None, 0))
        out = np.zeros(n, dtype=np.float64)
        if n:
            self._ck(self._L.g2ocu_get_f64(self._h, name.encode(), out.ctypes.data_as(ctypes.c_void_p), n))
        return out

    # ---- measurement ----
    def launch_count(self) -> int: return int(self._L.g2ocu_launch_count(self._h))

    def phase_time(self, phase: str):
        """(device seconds, kernel launches, number of timed intervals) accumulated for ``phase``."""
        s = ctypes.c_double(); n = ctypes.c_int64(); c = ctypes.c_int64()
        self._ck(self._L.g2ocu_phase_time(self._h, phase.encode(), ctypes.byref(s), ctypes.byref(n), ctypes.byref(c)))
        return s.value, n.value, c.value

    def reset_counters(self): self._ck(self._L.g2ocu_reset_counters(self._h))
