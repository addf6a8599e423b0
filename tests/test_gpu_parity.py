"""Parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.
Tolerances follow BASELINE.json's north_star: structure bit-exact (tests/test_structure.py), chi2 relative 1e-8 at
iteration 1 and 1e-6 at the end, lambda trajectory / accept-reject decisions identical, estimates 1e-6."""
import numpy as np
import pytest

from g2o_b200 import graph as G
from g2o_b200 import workloads as W
from g2o_b200.binding import CudaSolver
from oracle.oracle import Oracle

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(1e-300, float(np.max(np.abs(b))))) if a.size else 0.0


GRAPHS = {
    "ba_demo": (lambda: W.ba_demo(), "lm_fix6_3_cuda"),
    "ba_demo_xyz2uv_huber": (lambda: W.ba_demo(edge_type=G.EDGE_PROJECT_XYZ2UV, robust_kernel=True, outlier_ratio=0.05), "lm_fix6_3_cuda"),
    "bal_small": (lambda: W.bal_small(), "lm_fix9_3_cuda"),
    "bal_medium": (lambda: W.bal_synthetic(n_cameras=60, n_points=6000, n_obs=30000, seed=5, k_max=40, min_window=4), "lm_fix9_3_cuda"),
    # long tracks over several tile row groups / column strips, wrapping around the camera ring
    "bal_ring": (lambda: W.bal_synthetic(n_cameras=150, n_points=8000, n_obs=60000, seed=9, k_max=120, min_window=6), "lm_fix9_3_cuda"),
    "sphere": (lambda: W.sphere(nodes_per_level=16, laps=8), "lm_var_cuda"),
    "slam2d": (lambda: W.slam2d(n_poses=800, n_landmarks=200, world_size=30.0), "lm_fix3_2_cuda"),
    # VertexSE3Expmap / EdgeSE3Expmap (types_six_dof_expmap.h:108-127, .cpp:278-293): as a pose graph, and next to projection edges in one BA
    "sphere_expmap": (lambda: W.sphere_expmap(nodes_per_level=16, laps=8), "lm_var_cuda"),
    "ba_pose_constraints": (lambda: W.ba_demo_with_pose_constraints(), "lm_fix6_3_cuda"),
}


def make(name):
    fn, solver = GRAPHS[name]
    g = fn()
    s = CudaSolver(g, solver, device=0)
    s.initialize_optimization()
    o = Oracle(g, "lm", "pcg")
    assert o.initialize_optimization()
    return g, s, o


@pytest.mark.parametrize("name", list(GRAPHS))
def test_linear_system_blocks(name):
    g, s, o = make(name)
    s.init(); s.build_structure()
    assert o.algorithm_init() and o.build_structure()
    # errors and chi2
    s.compute_active_errors(); o.compute_active_errors()
    assert rel(s.get_f64("errors"), o.get_f64("errors")) < 1e-11
    assert abs(s.active_robust_chi2() - o.active_robust_chi2()) <= 1e-11 * abs(o.active_robust_chi2())
    assert abs(s.active_chi2() - o.active_chi2()) <= 1e-11 * abs(o.active_chi2())
    # Jacobians (the device BAL Jacobian is hand-derived, the oracle's is forward-mode AD like the reference)
    s.build_system(); o.build_system()
    assert rel(s.get_f64("jacobians"), o.get_f64("jacobians")) < 1e-10
    for n in ["b", "hpp_values"] + (["hll_values", "hpl_values"] if o.do_schur() else []):
        assert rel(s.get_f64(n), o.get_f64(n)) < 1e-11, n
    lam_s, lam_o = s.compute_lambda_init(), o.compute_lambda_init()
    assert abs(lam_s - lam_o) <= 1e-11 * lam_o
    # damped solve
    s.set_lambda(lam_o); o.set_lambda(lam_o)
    assert rel(s.get_f64("hpp_values"), o.get_f64("hpp_values")) < 1e-11
    assert s.solve() and o.solve()
    if o.do_schur():
        assert rel(s.get_f64("hschur_values"), o.get_f64("hschur_values")) < 1e-10
        assert rel(s.get_f64("bschur"), o.get_f64("bschur")) < 1e-10
    xs, xo = s.x(), o.get_f64("x")
    assert rel(xs, xo) < 1e-6            # PCG stops at a relative residual of 1e-6: solutions agree to that order
    assert abs(s.compute_scale(lam_o) - o.compute_scale()) <= 1e-6 * abs(o.compute_scale())
    s.restore_diagonal(); o.restore_diagonal()
    # apply the oracle's step on both sides: the oplus operators must agree
    s.push(); o.push()
    s.update(xo); o.update(xo)
    assert rel(s.get_estimates(), o.estimates()) < 1e-12
    s.compute_active_errors(); o.compute_active_errors()
    assert abs(s.active_robust_chi2() - o.active_robust_chi2()) <= 1e-10 * abs(o.active_robust_chi2())
    s.pop(); o.pop()
    assert rel(s.get_estimates(), o.estimates()) == 0.0


@pytest.mark.parametrize("name", list(GRAPHS))
def test_lm_trajectory(name):
    g, s, o = make(name)
    iters = 8
    n, st = s.optimize(iters)
    no, sto = o.optimize(iters)
    assert n == no
    assert len(st) == len(sto)
    for i, (a, b) in enumerate(zip(st, sto)):
        tol = 1e-8 if i == 0 else 1e-6
        assert abs(a["chi2"] - b["chi2"]) <= tol * abs(b["chi2"]), (i, a["chi2"], b["chi2"])
        assert a["levenberg_iterations"] == int(b["levenbergIterations"]), i
        assert abs(a["lambda"] - b["lambda"]) <= 1e-6 * abs(b["lambda"]), i
        assert a["result"] == int(b["result"])
    eo = o.estimates()
    assert np.max(np.abs(s.get_estimates() - eo) / (1.0 + np.abs(eo))) < 1e-6


KERNELS = {"Huber": G.KERNEL_HUBER, "PseudoHuber": G.KERNEL_PSEUDO_HUBER, "Cauchy": G.KERNEL_CAUCHY, "GemanMcClure": G.KERNEL_GEMAN_MCCLURE,
           "Welsch": G.KERNEL_WELSCH, "Fair": G.KERNEL_FAIR, "Tukey": G.KERNEL_TUKEY, "Saturated": G.KERNEL_SATURATED, "DCS": G.KERNEL_DCS}


@pytest.mark.parametrize("kernel", list(KERNELS))
@pytest.mark.parametrize("shape", ["bal", "slam2d"])
def test_every_robust_kernel_on_the_device(kernel, shape):
    """The nine kernels of core/robust_kernel_impl.cpp:50-170 in the device error / build kernels (edge_math.cuh): robustified chi2, the
    weighted right-hand side and Hessian blocks, and three LM iterations against the oracle (whose kernels are pinned to the reference's
    RobustKernelFactory classes by tests/test_reference_leaves.py).  The width sits inside the residual distribution so that both branches
    of the piecewise kernels (Huber, Tukey, Saturated) are taken."""
    if shape == "bal":
        g = W.bal_synthetic(n_cameras=24, n_points=1200, n_obs=6000, seed=4, k_max=16, min_window=4, outlier_fraction=0.05); solver = "lm_fix9_3_cuda"
    else:
        g = W.slam2d(n_poses=300, n_landmarks=80, world_size=20.0); solver = "lm_fix3_2_cuda"
    g.e_kernel = np.full(g.n_edges, KERNELS[kernel], dtype=np.int32); g.e_kernel_delta = np.full(g.n_edges, 1.5)
    s = CudaSolver(g, solver, device=0); s.initialize_optimization(); s.init(); s.build_structure()
    o = Oracle(g, "lm", "pcg"); assert o.initialize_optimization() and o.algorithm_init() and o.build_structure()
    s.compute_active_errors(); o.compute_active_errors()
    assert abs(s.active_robust_chi2() - o.active_robust_chi2()) <= 1e-11 * abs(o.active_robust_chi2())
    assert abs(s.active_chi2() - o.active_chi2()) <= 1e-11 * abs(o.active_chi2())
    assert abs(o.active_robust_chi2() - o.active_chi2()) > 1e-3 * o.active_chi2()      # the kernel does change the cost here
    s.build_system(); o.build_system()
    for n in ["b", "hpp_values", "hll_values", "hpl_values"]:
        assert rel(s.get_f64(n), o.get_f64(n)) < 1e-11, n
    s = CudaSolver(g, solver, device=0); s.initialize_optimization()
    o = Oracle(g, "lm", "pcg"); assert o.initialize_optimization()
    n, st = s.optimize(3); no, sto = o.optimize(3)
    assert n == no
    for i, (a, b) in enumerate(zip(st, sto)):
        assert abs(a["chi2"] - b["chi2"]) <= (1e-8 if i == 0 else 1e-6) * abs(b["chi2"]), (kernel, i, a["chi2"], b["chi2"])
        assert a["levenberg_iterations"] == int(b["levenbergIterations"]) and abs(a["lambda"] - b["lambda"]) <= 1e-6 * abs(b["lambda"])


def test_force_stop_flag_ends_the_optimisation():
    """SparseOptimizer::terminate() (sparse_optimizer.h:186-190): with the flag raised optimize() starts no iteration (sparse_optimizer.cpp:396)."""
    import ctypes
    g = W.bal_small()
    s = CudaSolver(g, "lm_fix9_3_cuda", device=0); s.initialize_optimization()
    flag = ctypes.c_ubyte(0); s.set_force_stop_flag(flag)
    n, st = s.optimize(2); assert n == 2
    flag.value = 1
    n, st = s.optimize(5); assert n == 0
    flag.value = 0
    n, st = s.optimize(1); assert n == 1


def test_fused_pcg_tail_matches_the_three_kernel_path():
    """The CG recurrences of one iteration can run as ONE cluster kernel (distributed shared memory for the two reductions) for systems of up to
    65 536 unknowns - it is used for small systems (below 8192 unknowns); G2OCU_PCG_FUSED_MAX raises the
    limit, G2OCU_PCG_TAIL=split selects the three-kernel path.  Every sum of the tail is formed in the same order on both; the
    products before it add with atomics, so two runs agree to rounding, not to the bit: same LM trials, same PCG iteration counts, chi2 to 1e-9."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import json, numpy as np\n"
            "from g2o_b200 import workloads as W\n"
            "from g2o_b200.binding import CudaSolver\n"
            "out = []\n"
            "for g, name in [(W.bal_synthetic(n_cameras=60, n_points=6000, n_obs=30000, seed=5, k_max=40, min_window=4), 'lm_fix9_3_cuda'), (W.sphere(nodes_per_level=16, laps=8), 'lm_var_cuda'), (W.ba_demo(), 'lm_fix6_3_cuda')]:\n"
            "    s = CudaSolver(g, name, device=0); s.initialize_optimization(); n, st = s.optimize(4)\n"
            "    out.append([n, [x['chi2'] for x in st], [x['iterations_linear_solver'] for x in st], [x['levenberg_iterations'] for x in st]])\n"
            "print(json.dumps(out))\n")
    outs = []
    for mode in ("fused", "split"):
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=root, timeout=300, env=dict(os.environ, G2OCU_PCG_TAIL=mode, G2OCU_PCG_FUSED_MAX="65536"))
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(json.loads(r.stdout.strip().splitlines()[-1]))
    for a, b in zip(*outs):
        assert a[0] == b[0] and a[2] == b[2] and a[3] == b[3], (a, b)
        assert all(abs(x - y) <= 1e-9 * abs(y) for x, y in zip(a[1], b[1])), (a[1], b[1])


@pytest.mark.parametrize("dim", [3, 6, 9])
def test_linear_solver_level_entry_point(dim):
    """g2ocu_linear_solve = LinearSolverPCG::solve (linear_solver_pcg.hpp:80-156) on a block matrix in the reference's block-column layout:
    the iterates stop at the reference's rule (r.M^-1 r <= 1e-6 of its initial value), so the residual is small in that norm and the
    solution agrees with a dense solve once the tolerance is tightened; a second solve with the same pattern reuses the flattened view."""
    import ctypes
    from g2o_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(dim)
    nb = 40
    dense = np.zeros((nb * dim, nb * dim))
    colptr, rowidx, vals = [0], [], []
    for c in range(nb):
        for r in range(c + 1):
            if r == c or rng.random() < 0.15:
                B = rng.normal(size=(dim, dim))
                if r == c:
                    B = B @ B.T + 3 * dim * np.eye(dim)
                dense[r * dim:(r + 1) * dim, c * dim:(c + 1) * dim] = B
                if r != c:
                    dense[c * dim:(c + 1) * dim, r * dim:(r + 1) * dim] = B.T
                rowidx.append(r); vals.append(B.ravel(order="F"))
        colptr.append(len(rowidx))
    dense += 2.0 * np.sum(np.abs(dense), axis=1).max() / 10 * np.eye(nb * dim)          # comfortably positive definite
    k = 0
    for c in range(nb):                                                               # the diagonal shift into the block list as well
        for j in range(colptr[c], colptr[c + 1]):
            if rowidx[j] == c:
                vals[j] = dense[c * dim:(c + 1) * dim, c * dim:(c + 1) * dim].ravel(order="F")
    colptr = np.asarray(colptr, dtype=np.int32); rowidx = np.asarray(rowidx, dtype=np.int32); vals = np.ascontiguousarray(np.concatenate(vals))
    h = ctypes.c_void_p(); cfg = _lib.Config(); L.g2ocu_default_config(ctypes.byref(cfg)); cfg.device = 0
    assert L.g2ocu_linear_create(ctypes.byref(cfg), ctypes.byref(h)) == 0
    try:
        assert L.g2ocu_linear_init(h) == 0
        assert L.g2ocu_linear_set_property(h, b"pcgTolerance", 1e-16) == 0              # LinearSolverPCG::setTolerance: far below the reference's 1e-6 so that x can be compared
        for trial in range(2):
            b = rng.normal(size=nb * dim); x = np.zeros(nb * dim)
            ok = ctypes.c_int32(); its = ctypes.c_int32()
            rc = L.g2ocu_linear_solve(h, nb, dim, colptr.ctypes.data_as(ctypes.c_void_p), rowidx.ctypes.data_as(ctypes.c_void_p), vals.ctypes.data_as(ctypes.c_void_p),
                                      b.ctypes.data_as(ctypes.c_void_p), x.ctypes.data_as(ctypes.c_void_p), ctypes.byref(ok), ctypes.byref(its))
            assert rc == 0, L.g2ocu_linear_last_error(h)
            assert ok.value == 1 and 0 < its.value < nb * dim
            xd = np.linalg.solve(dense, b)
            assert np.linalg.norm(x - xd) <= 1e-6 * np.linalg.norm(xd), (trial, its.value, np.linalg.norm(x - xd) / np.linalg.norm(xd))
            L.g2ocu_linear_init(h)
        bad = rowidx.copy(); bad[0] = 5                                               # a block below the diagonal is refused
        assert L.g2ocu_linear_solve(h, nb, dim, colptr.ctypes.data_as(ctypes.c_void_p), bad.ctypes.data_as(ctypes.c_void_p), vals.ctypes.data_as(ctypes.c_void_p),
                                    b.ctypes.data_as(ctypes.c_void_p), x.ctypes.data_as(ctypes.c_void_p), None, None) < 0
    finally:
        L.g2ocu_linear_destroy(h)


def _dense_hpp(s):
    """Hpp of the device (upper blocks, CCS order, column-major) as a dense symmetric matrix."""
    dims = s.get_i32("dims"); nb, n = int(dims[0]), int(dims[2]); P = n // nb
    colptr, rowidx, vals = s.get_i32("hpp_colptr"), s.get_i32("hpp_rowidx"), s.get_f64("hpp_values").reshape(-1, P, P)
    H = np.zeros((n, n))
    for c in range(nb):
        for k in range(colptr[c], colptr[c + 1]):
            r = int(rowidx[k]); blk = vals[k].T                        # stored column-major
            H[r * P:(r + 1) * P, c * P:(c + 1) * P] = blk
            H[c * P:(c + 1) * P, r * P:(r + 1) * P] = blk.T
    return H, P


@pytest.mark.parametrize("name", ["sphere", "bal_medium", "slam2d", "ba_pose_constraints"])
def test_compute_marginals_are_blocks_of_the_inverse_of_hpp(name):
    """g2ocu_compute_marginals (SparseOptimizer::computeMarginals, sparse_optimizer.cpp:594-596): every requested block against numpy's
    inverse of the same Hpp; all diagonal blocks at once (several column batches on the larger graphs), symmetric pairs, the damped matrix
    after setLambda; the error paths."""
    g, s, o = make(name)
    s.init(); s.build_structure(); s.compute_active_errors(); s.build_system()
    H, P = _dense_hpp(s); nb = H.shape[0] // P
    inv = np.linalg.inv(H); scale = float(np.max(np.abs(inv)))
    pairs = [(i, i) for i in range(nb)] + [(0, nb - 1), (nb - 1, 0), (nb // 2, nb // 3), (nb // 3, nb // 2)]
    got = s.compute_marginals(pairs)
    assert got is not None and len(got) == len(pairs)
    for (r, c), b in zip(pairs, got):
        assert np.max(np.abs(b - inv[r * P:(r + 1) * P, c * P:(c + 1) * P])) <= 1e-9 * scale, (name, r, c)
    assert o.algorithm_init() and o.build_structure()                   # ... and against the oracle's restatement on its own Hpp
    o.compute_active_errors(); o.build_system()
    for b, w in zip(got[-6:], o.compute_marginals(pairs[-6:])):
        assert b.shape == w.shape and np.max(np.abs(b - w)) <= 1e-8 * scale, name
    s.set_lambda(0.5)                                                  # the reference factorises Hpp as it stands: damped until restoreDiagonal
    inv2 = np.linalg.inv(H + 0.5 * np.eye(H.shape[0]))
    b = s.compute_marginals([(1, 1)])[0]
    assert np.max(np.abs(b - inv2[P:2 * P, P:2 * P])) <= 1e-9 * float(np.max(np.abs(inv2)))
    s.restore_diagonal()
    assert s.compute_marginals([]) == []
    with pytest.raises(Exception):
        s.compute_marginals([(0, nb)])


def test_compute_marginals_of_a_singular_system():
    """Not positive definite -> the bool of the reference's solvePattern is false (None here)."""
    g = W.sphere(nodes_per_level=8, laps=4, fix_first=False)          # gauge freedom: Hpp is singular
    s = CudaSolver(g, "gn_var_cuda", device=0); s.initialize_optimization(); s.init(); s.build_structure(); s.compute_active_errors(); s.build_system()
    s.set_lambda(-1.0)                                                 # make sure a pivot goes negative whatever the rounding does
    assert s.compute_marginals([(0, 0)]) is None


def test_compute_marginals_of_a_whole_system():
    """Points that are not marginalized: Hpp is the whole system over all vertices in id order (blocks of 3 and 2 on a 2-D SLAM graph).  Every
    diagonal block and a few rectangular ones against numpy's inverse of the dense matrix rebuilt from H v products in the reference's order."""
    g = W.slam2d(n_poses=60, n_landmarks=20, world_size=10.0, marginalize_landmarks=False)
    s = CudaSolver(g, "gn_var_cuda", device=0); s.initialize_optimization(); s.init(); s.build_structure(); s.compute_active_errors(); s.build_system()
    ends = s.get_i32("pose_block_indices").astype(int); n = int(ends[-1]); starts = np.concatenate([[0], ends[:-1]])
    H = np.stack([s.multiply_hessian(np.eye(n)[:, j]) for j in range(n)], axis=1)
    assert np.max(np.abs(H - H.T)) <= 1e-9 * np.max(np.abs(H))
    inv = np.linalg.inv(H); scale = float(np.max(np.abs(inv)))
    nb = len(ends)
    assert {int(e - b) for b, e in zip(starts, ends)} == {2, 3}
    pairs = [(i, i) for i in range(nb)] + [(0, nb - 1), (nb - 1, 0), (nb // 2, nb // 3), (1, nb - 2)]
    got = s.compute_marginals(pairs)
    assert got is not None and len(got) == len(pairs)
    for (r, c), b in zip(pairs, got):
        want = inv[starts[r]:ends[r], starts[c]:ends[c]]
        assert b.shape == want.shape and np.max(np.abs(b - want)) <= 1e-9 * scale, (r, c)


def test_gauss_newton_sphere():
    g = W.sphere(nodes_per_level=12, laps=6)
    s = CudaSolver(g, "gn_var_cuda", device=0); s.initialize_optimization()
    o = Oracle(g, "gn", "pcg"); o.initialize_optimization()
    n, st = s.optimize(4); no, sto = o.optimize(4)
    assert n == no
    for a, b in zip(st, sto):
        assert abs(a["chi2"] - b["chi2"]) <= 1e-6 * abs(b["chi2"])


def test_multiply_hessian_and_second_optimize_continues():
    g, s, o = make("bal_small")
    n, st = s.optimize(3)
    n2, st2 = s.optimize(2)            # a second optimize() continues from the current estimates (vertices keep state)
    assert st2[0]["chi2"] <= st[-1]["chi2"] * (1 + 1e-9)
    s.build_system()
    v = np.random.default_rng(0).normal(size=s.get_i32("dims")[2])
    hv = s.multiply_hessian(v)
    # reference: y = Hpp v with the upper blocks used twice (sparse_block_matrix.hpp:288-312)
    colptr, rowidx, vals = s.get_i32("hpp_colptr"), s.get_i32("hpp_rowidx"), s.get_f64("hpp_values").reshape(-1, 9, 9)
    y = np.zeros_like(v)
    for c in range(len(colptr) - 1):
        for k in range(colptr[c], colptr[c + 1]):
            r = rowidx[k]; B = vals[k].T      # column-major block
            y[r * 9:(r + 1) * 9] += B @ v[c * 9:(c + 1) * 9]
            if r != c:
                y[c * 9:(c + 1) * 9] += B.T @ v[r * 9:(r + 1) * 9]
    assert rel(hv, y) < 1e-12


DENSE = {
    "ba_demo": ("lm_fix6_3_cuda", lambda: W.ba_demo()),
    "bal_medium": ("lm_fix9_3_cuda", lambda: W.bal_synthetic(n_cameras=60, n_points=6000, n_obs=30000, seed=5, k_max=40, min_window=4)),
    "sphere": ("lm_var_cuda", lambda: W.sphere(nodes_per_level=16, laps=8)),
    "slam2d": ("lm_fix3_2_cuda", lambda: W.slam2d(n_poses=800, n_landmarks=200, world_size=30.0)),
}


@pytest.mark.parametrize("name", list(DENSE))
def test_dense_cholesky_trajectory(name):
    """LinearSolverDense semantics (linear_solver_dense.h:65-115) on the device: DMMA Cholesky of the dense (reduced) system."""
    solver, fn = DENSE[name]
    g = fn()
    s = CudaSolver(g, solver, linear="dense", device=0); s.initialize_optimization()
    o = Oracle(g, "lm", "dense"); assert o.initialize_optimization()
    n, st = s.optimize(6); no, sto = o.optimize(6)
    assert n == no and len(st) == len(sto)
    for i, (a, b) in enumerate(zip(st, sto)):
        tol = 1e-8 if i == 0 else 1e-6
        assert abs(a["chi2"] - b["chi2"]) <= tol * abs(b["chi2"]), (i, a["chi2"], b["chi2"])
        assert a["levenberg_iterations"] == int(b["levenbergIterations"]), i
        assert abs(a["lambda"] - b["lambda"]) <= 1e-6 * abs(b["lambda"]), i
    eo = o.estimates()
    assert np.max(np.abs(s.get_estimates() - eo) / (1.0 + np.abs(eo))) < 1e-6


def test_dense_cholesky_solution_and_failure():
    g = W.bal_synthetic(n_cameras=40, n_points=3000, n_obs=15000, seed=3, k_max=30, min_window=4)
    s = CudaSolver(g, "lm_fix9_3_cuda", linear="dense", device=0); s.initialize_optimization()
    p = CudaSolver(g, "lm_fix9_3_cuda", linear="pcg", device=0, pcg_tolerance=1e-14); p.initialize_optimization()
    for t in (s, p):
        t.init(); t.build_structure(); t.compute_active_errors(); t.build_system()
    lam = s.compute_lambda_init()
    s.set_lambda(lam); p.set_lambda(lam)
    assert s.solve() and p.solve()
    # direct factorisation against a PCG run to 1e-14: same linear system, same solution
    assert rel(s.x(), p.x()) < 1e-7
    s.restore_diagonal()
    s.set_lambda(-1e15)                      # indefinite system: the reference's LDLT reports !isPositive() and solve() returns false
    assert not s.solve()


# ---------------------------------------------------------------------------------------------------------------
# Powell's dogleg (optimization_algorithm_dogleg.cpp:56-197): same trajectory, trust region and step types as the oracle
def _bal_gauge_fixed():
    """BAL graph with two cameras held fixed: the 7-dof gauge is gone, the undamped Gauss-Newton system is positive definite."""
    g = W.bal_synthetic(n_cameras=60, n_points=6000, n_obs=30000, seed=5, k_max=40, min_window=4)
    g.v_fixed = np.array(g.v_fixed, dtype=np.uint8)
    g.v_fixed[np.flatnonzero(np.asarray(g.v_type) == G.VERTEX_CAM_BAL)[:2]] = 1
    return g


DOGLEG = {
    # name: (graph, linear solver, initialDelta, iterations)
    "sphere_dense": (lambda: W.sphere(nodes_per_level=16, laps=8), "dense", 1e4, 3),
    "sphere_small_delta": (lambda: W.sphere(nodes_per_level=16, laps=8), "dense", 2.0, 5),
    "slam2d_dense": (lambda: W.slam2d(n_poses=800, n_landmarks=200, world_size=30.0), "dense", 1e4, 5),
    "slam2d_small_delta": (lambda: W.slam2d(n_poses=800, n_landmarks=200, world_size=30.0), "dense", 1.0, 5),
    "bal_dense": (_bal_gauge_fixed, "dense", 1e4, 5),
    "bal_pcg": (_bal_gauge_fixed, "pcg", 1e4, 5),
}


@pytest.mark.parametrize("name", list(DOGLEG))
def test_dogleg_trajectory(name):
    fn, linear, delta0, iters = DOGLEG[name]
    g = fn()
    s = CudaSolver(g, "dl_var_cuda", linear=linear, device=0); s.set_property("doglegInitialDelta", delta0); s.initialize_optimization()
    o = Oracle(g, "dl", linear); o.set_dogleg_params(initial_delta=delta0); assert o.initialize_optimization()
    n, st = s.optimize(iters); no, sto = o.optimize(iters)
    assert n == no and len(st) == len(sto)
    loose = linear == "pcg"            # PCG stops at a relative residual of 1e-6: the undamped steps agree to that order only
    for i, (a, b) in enumerate(zip(st, sto)):
        tol = 1e-5 if loose else (1e-8 if i == 0 else 1e-6)
        assert abs(a["chi2"] - b["chi2"]) <= tol * abs(b["chi2"]), (i, a["chi2"], b["chi2"])
        assert a["result"] == int(b["result"])
    ds, do = s.dogleg_state(), o.dogleg_state()
    assert ds["last_step"] == do["last_step"] and ds["tries"] == do["tries"] and ds["was_pd"] == do["was_pd"], (ds, do)
    assert abs(ds["delta"] - do["delta"]) <= (1e-4 if loose else 1e-6) * do["delta"], (ds, do)
    eo = o.estimates()
    assert np.max(np.abs(s.get_estimates() - eo) / (1.0 + np.abs(eo))) < (1e-4 if loose else 1e-6)


def test_dogleg_step_norm_and_rejection():
    """BAL vertices add the increment: an accepted Descent / Dogleg step moves the estimates by exactly the trust-region radius."""
    g = _bal_gauge_fixed()
    for delta0 in (1e-4, 3e-2):
        s = CudaSolver(g, "dl_var_cuda", linear="dense", device=0); s.set_property("doglegInitialDelta", delta0); s.initialize_optimization()
        e0 = s.get_estimates().copy()
        n, st = s.optimize(1)
        d = s.dogleg_state()
        assert n == 1 and d["tries"] == 1 and d["last_step"] in (1, 3), d
        moved = np.linalg.norm(s.get_estimates() - e0)
        assert abs(moved - delta0) <= 1e-9 * delta0, (d, moved)
        assert d["delta"] >= delta0      # rho > 0.75 on a step this short: the region grows to 3 |hdl| (or stays)


def test_dogleg_on_a_rank_deficient_system_stays_finite():
    """No fixed vertex (7-dof gauge): the dense factorisation may fail and Dogleg damps (dogleg.cpp:117-135), or it succeeds with
    tiny pivots and the long Gauss-Newton step is cut to the trust region.  Either way chi2 must not increase."""
    g = W.bal_small()
    s = CudaSolver(g, "dl_var_cuda", linear="dense", device=0); s.initialize_optimization()
    s.compute_active_errors(); chi0 = s.active_robust_chi2()
    n, st = s.optimize(3)
    assert n >= 0
    chi = s.active_robust_chi2()
    assert np.isfinite(chi) and chi <= chi0 * (1 + 1e-12)
    d = s.dogleg_state()
    assert 1e-12 <= d["lambda"] <= 1e3


# ---------------------------------------------------------------------------------------------------------------
# Points that are not marginalized (`lm_var` on SLAM / BA graphs, BlockSolverX with two block sizes): the reference solves the
# whole system; vectors at the boundary are in its order (all vertices by id)
def _points_free(g):
    g.v_marginalized = np.zeros_like(g.v_marginalized)
    return g


FULL = {
    "slam2d": lambda: W.slam2d(n_poses=800, n_landmarks=200, world_size=30.0, marginalize_landmarks=False),
    "ba_demo": lambda: _points_free(W.ba_demo()),
    "bal_small": lambda: _points_free(W.bal_small()),
}


@pytest.mark.parametrize("name", list(FULL))
def test_full_system_blocks_and_solve(name):
    g = FULL[name]()
    s = CudaSolver(g, "lm_var_cuda", device=0); s.initialize_optimization()
    o = Oracle(g, "lm", "pcg"); assert o.initialize_optimization()
    s.init(); s.build_structure()
    assert o.algorithm_init() and o.build_structure() and not o.do_schur()
    assert np.array_equal(s.get_i32("dims"), o.get_i32("dims"))
    s.compute_active_errors(); o.compute_active_errors()
    assert rel(s.get_f64("errors"), o.get_f64("errors")) < 1e-11
    assert abs(s.active_robust_chi2() - o.active_robust_chi2()) <= 1e-11 * abs(o.active_robust_chi2())
    s.build_system(); o.build_system()
    assert rel(s.b(), o.get_f64("b")) < 1e-11
    lam_s, lam_o = s.compute_lambda_init(), o.compute_lambda_init()
    assert abs(lam_s - lam_o) <= 1e-11 * lam_o
    # H v over the whole system, with and without damping
    v = np.random.default_rng(3).normal(size=s.vector_size())
    assert rel(s.multiply_hessian(v), o.multiply_hessian(v)) < 1e-11
    s.set_lambda(lam_o); o.set_lambda(lam_o)
    assert rel(s.multiply_hessian(v), o.multiply_hessian(v)) < 1e-11
    assert s.solve() and o.solve()
    xs, xo = s.x(), o.get_f64("x")
    assert rel(xs, xo) < 1e-6            # PCG stops at a relative residual of 1e-6
    assert abs(s.compute_scale(lam_o) - o.compute_scale()) <= 1e-6 * abs(o.compute_scale())
    s.restore_diagonal(); o.restore_diagonal()
    s.push(); o.push()
    s.update(xo); o.update(xo)           # a step in the reference's vector order
    assert rel(s.get_estimates(), o.estimates()) < 1e-12
    s.compute_active_errors(); o.compute_active_errors()
    assert abs(s.active_robust_chi2() - o.active_robust_chi2()) <= 1e-10 * abs(o.active_robust_chi2())
    s.pop(); o.pop()
    assert rel(s.get_estimates(), o.estimates()) == 0.0


@pytest.mark.parametrize("linear", ["pcg", "dense"])
@pytest.mark.parametrize("name", list(FULL))
def test_full_system_lm_trajectory(name, linear):
    g = FULL[name]()
    s = CudaSolver(g, "lm_var_cuda", linear=linear, device=0); s.initialize_optimization()
    o = Oracle(g, "lm", linear); assert o.initialize_optimization()
    n, st = s.optimize(6); no, sto = o.optimize(6)
    assert n == no and len(st) == len(sto)
    for i, (a, b) in enumerate(zip(st, sto)):
        tol = 1e-8 if i == 0 else 1e-6
        assert abs(a["chi2"] - b["chi2"]) <= tol * abs(b["chi2"]), (i, a["chi2"], b["chi2"])
        assert a["levenberg_iterations"] == int(b["levenbergIterations"]), i
        assert abs(a["lambda"] - b["lambda"]) <= 1e-6 * abs(b["lambda"]), i
        assert a["hessian_pose_dimension"] == int(b["hessianPoseDimension"]) and a["hessian_landmark_dimension"] == 0
    eo = o.estimates()
    assert np.max(np.abs(s.get_estimates() - eo) / (1.0 + np.abs(eo))) < 1e-6


def test_full_system_gauss_newton_and_dogleg():
    g = FULL["slam2d"]()
    for name, alg, iters in (("gn_var_cuda", "gn", 3), ("dl_var_cuda", "dl", 4)):
        s = CudaSolver(g, name, linear="dense", device=0); s.initialize_optimization()
        o = Oracle(g, alg, "dense"); assert o.initialize_optimization()
        n, st = s.optimize(iters); no, sto = o.optimize(iters)
        assert n == no
        for i, (a, b) in enumerate(zip(st, sto)):
            assert abs(a["chi2"] - b["chi2"]) <= (1e-8 if i == 0 else 1e-6) * abs(b["chi2"]), (name, i, a["chi2"], b["chi2"])
        if alg == "dl":
            ds, do = s.dogleg_state(), o.dogleg_state()
            assert ds["last_step"] == do["last_step"] and ds["tries"] == do["tries"] and abs(ds["delta"] - do["delta"]) <= 1e-6 * do["delta"], (ds, do)
