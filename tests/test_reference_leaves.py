"""Leaf functions of the REAL reference, compiled from /root/reference into oracle/_ref/libg2o_ref_leaves.so (oracle/Makefile: the
translation units use Eigen only as a typed array and build against a declaration shim), against the oracle's restatements:
the nine robust kernels (g2o/core/robust_kernel_impl.cpp:50-181), dq/dR of EdgeSE3's Jacobian (g2o/types/slam3d/dquat2mat.cpp:35-85 and
the generated dquat2mat_maxima_generated.cpp), normalize_theta (g2o/stuff/misc.h:114-127).  These pin parts of the oracle that the
reference's own unit tests do not cover."""
import ctypes

import numpy as np
import pytest
from scipy.spatial.transform import Rotation

from g2o_b200 import graph as G
from oracle import oracle

REF = oracle.reference_leaves()
pytestmark = pytest.mark.skipif(REF is None, reason="oracle/_ref/libg2o_ref_leaves.so was not built (no reference tree at build time)")


def ref_robustify(name, delta, e2):
    rho = np.zeros(3)
    assert REF.ref_robustify(name.encode(), delta, e2, rho.ctypes.data_as(ctypes.c_void_p)) == 0, name
    return rho


@pytest.mark.parametrize("name", [n for n in G.KERNEL_BY_NAME if n])
def test_robust_kernels_match_the_reference_code(name):
    kind = G.KERNEL_BY_NAME[name]
    rng = np.random.default_rng(kind)
    for delta in (0.3, 1.0, 2.5):
        e2s = np.concatenate([[0.0, 1e-12, delta * delta, delta * delta * (1 + 1e-12), delta * delta * (1 - 1e-12)], rng.uniform(0, 4 * delta * delta, 40), rng.uniform(0, 400, 20)])
        for e2 in e2s:
            want, got = ref_robustify(name, delta, float(e2)), oracle.robustify(kind, delta, float(e2))
            assert np.allclose(got, want, rtol=2e-15, atol=0.0), (name, delta, e2, got, want)   # a few ulps: the oracle is built with FMA contraction


def test_unknown_kernel_name_is_unknown_to_the_reference_too():
    assert REF.ref_robustify(b"NoSuchKernel", 1.0, 1.0, np.zeros(3).ctypes.data_as(ctypes.c_void_p)) == -1


def test_dq_dR_matches_the_generated_reference_code():
    rng = np.random.default_rng(7)
    seen = set()
    mats = [Rotation.from_rotvec(v).as_matrix() for v in rng.normal(size=(200, 3)) * rng.uniform(0.05, 3.1, size=(200, 1))]
    mats += [Rotation.from_rotvec(np.pi * 0.999 * np.eye(3)[k]).as_matrix() for k in range(3)]      # trace < 0: the x / y / z branches
    for R in mats:
        R9 = np.asfortranarray(R).ravel(order="F").copy()
        want = np.zeros(27); REF.ref_dq_dR(R9.ctypes.data_as(ctypes.c_void_p), want.ctypes.data_as(ctypes.c_void_p))
        got = oracle.dq_dR(R)                 # 3 x 9
        tr = np.trace(R)
        seen.add(0 if tr > 0 else (1 if (R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]) else (2 if R[1, 1] > R[2, 2] else 3)))
        assert np.allclose(got, want.reshape(9, 3).T, rtol=1e-14, atol=1e-15), np.max(np.abs(got - want.reshape(9, 3).T))
    assert seen == {0, 1, 2, 3}


def test_normalize_theta_matches_the_reference_code():
    rng = np.random.default_rng(11)
    for th in np.concatenate([rng.uniform(-20, 20, 200), [np.pi, -np.pi, 3 * np.pi, -3 * np.pi, 0.0, 2 * np.pi]]):
        est, _ = oracle.vertex_oplus(G.VERTEX_SE2, [0.0, 0.0, 0.0], [0.0, 0.0, float(th)])
        assert est[2] == REF.ref_normalize_theta(float(th)), th


def test_sphere_noise_stream_matches_the_reference_sampler():
    """create_sphere draws its measurement noise from two default-seeded std::mt19937 engines through ONE static
    std::normal_distribution (g2o/stuff/sampler.cpp:31-45), whose saved value crosses between the engines: per edge three rotation draws
    (engine 1), then three translation draws (engine 0) (create_sphere.cpp:175,183).  g2o_b200/workloads.py restates libstdc++'s
    distribution and generate_canonical in Python; here the same call pattern runs through the reference's own sampleGaussian, in a fresh
    process because the static distribution keeps state."""
    import subprocess, sys, json
    n_edges = 500
    code = ("import ctypes, json, numpy as np\n"
            "from oracle import oracle\n"
            "L = oracle.reference_leaves()\n"
            f"which = np.array(([1, 1, 1, 0, 0, 0] * {n_edges}), dtype=np.int32); out = np.zeros(which.size)\n"
            "L.ref_sample_gaussian_two_engines(which.ctypes.data_as(ctypes.c_void_p), which.size, out.ctypes.data_as(ctypes.c_void_p))\n"
            "print(json.dumps([float.hex(float(v)) for v in out]))\n")
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=root, timeout=120)
    assert r.returncode == 0, r.stderr
    want = np.array([float.fromhex(h) for h in json.loads(r.stdout)])
    from g2o_b200.workloads import _StdNormalShared
    nd = _StdNormalShared(); gen_trans, gen_rot = nd.engine(), nd.engine()
    got = []
    for _ in range(n_edges):
        got += [nd(gen_rot) for _ in range(3)]
        got += [nd(gen_trans) for _ in range(3)]
    assert np.array_equal(np.array(got), want)      # bit for bit


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_se2_edges_and_oplus_match_the_reference_se2_class():
    """EdgeSE2 / EdgeSE2PointXY errors and VertexSE2::oplusImpl evaluated with the reference's own SE2 class (g2o/types/slam2d/se2.h: where the
    angle is normalised in operator*= and inverse()) against the oracle, including angles around +-pi where the normalisation matters."""
    rng = np.random.default_rng(5)
    worst = 0.0
    for k in range(400):
        big = 1.0 if k % 4 else 30.0        # some poses far outside [-pi, pi)
        x0 = np.array([*rng.normal(size=2) * 5, rng.uniform(-np.pi, np.pi) * big]); x1 = np.array([*rng.normal(size=2) * 5, rng.uniform(-np.pi, np.pi) * big])
        z = np.array([*rng.normal(size=2), rng.uniform(-np.pi, np.pi)])
        if k % 7 == 0:                      # relative rotation next to the +-pi cut
            x1[2] = x0[2] + z[2] + np.pi - 1e-9 * (k % 3 - 1)
        want = np.zeros(3); REF.ref_edge_se2_error(_p(x0), _p(x1), _p(z), _p(want))
        got = oracle.edge_error(G.EDGE_SE2, x0, x1, z)
        # the angle may sit on either side of the cut only if both are within rounding of it; compare on the circle
        d = got - want; d[2] = (d[2] + np.pi) % (2 * np.pi) - np.pi
        worst = max(worst, float(np.max(np.abs(d))))
        assert np.max(np.abs(d)) <= 1e-12 * (1 + np.max(np.abs(want))), (x0, x1, z, got, want)
        l, zl = rng.normal(size=2) * 5, rng.normal(size=2)
        want2 = np.zeros(2); REF.ref_edge_se2_pointxy_error(_p(x0), _p(l), _p(zl), _p(want2))
        assert np.allclose(oracle.edge_error(G.EDGE_SE2_POINT_XY, x0, l, zl), want2, rtol=1e-13, atol=1e-13)
        upd = np.array([*rng.normal(size=2), rng.uniform(-4, 4)])
        est = x0.copy(); REF.ref_vertex_se2_oplus(_p(est), _p(upd))
        got3, _ = oracle.vertex_oplus(G.VERTEX_SE2, x0, upd)
        assert np.allclose(got3, est, rtol=0, atol=1e-13), (x0, upd, got3, est)   # exact up to the FMA contraction the oracle is built with
    assert worst < 1e-12


def _rand_se3(rng, angle_scale=1.0):
    q = Rotation.from_rotvec(rng.normal(size=3) * angle_scale).as_quat()      # x y z w
    if q[3] < 0:
        q = -q
    return np.concatenate([rng.normal(size=3) * 3, q])


def test_se3quat_exp_log_edges_and_oplus_match_the_reference_class():
    """SE3Quat::exp / log / adj / product / inverse / map of the reference (g2o/types/slam3d/se3quat.h, compiled) under VertexSE3Expmap::oplusImpl,
    EdgeSE3Expmap::computeError + linearizeOplus and EdgeProjectXYZ2UV::computeError, against the oracle; both branches of exp (theta < 1e-5)
    and log (d > 0.99999) are visited."""
    rng = np.random.default_rng(9)
    for k in range(300):
        scale = [1.0, 1e-3, 1e-7, 2.5][k % 4]
        u = np.concatenate([rng.normal(size=3) * scale, rng.normal(size=3)])
        want = np.zeros(7); REF.ref_se3quat_exp(_p(u), _p(want))
        assert np.allclose(oracle.se3_exp(u), want, rtol=0, atol=1e-14), (u, oracle.se3_exp(u), want)
        T = _rand_se3(rng, scale)
        wl = np.zeros(6); REF.ref_se3quat_log(_p(T), _p(wl))
        assert np.allclose(oracle.se3_log(T), wl, rtol=1e-9, atol=1e-13), (T, oracle.se3_log(T), wl)     # acos near d = 1 amplifies rounding
        est = _rand_se3(rng); upd = np.concatenate([rng.normal(size=3) * scale, rng.normal(size=3) * 0.1])
        e2 = est.copy(); REF.ref_vertex_se3expmap_oplus(_p(e2), _p(upd))
        got, _ = oracle.vertex_oplus(G.VERTEX_SE3_EXPMAP, est, upd)
        assert np.allclose(got, e2, rtol=0, atol=1e-13), (est, upd, got, e2)
        x0, x1, z = _rand_se3(rng), _rand_se3(rng), _rand_se3(rng)
        if k % 3 == 0:
            # measurement consistent with the poses up to a tiny perturbation: the error lands in the small-angle branch of log
            R0, R1 = Rotation.from_quat(x0[3:]), Rotation.from_quat(x1[3:])
            Rz = R1 * R0.inv() * Rotation.from_rotvec(rng.normal(size=3) * 1e-4)
            q = Rz.as_quat(); q = -q if q[3] < 0 else q
            z = np.concatenate([x1[:3] - Rz.apply(x0[:3]) + rng.normal(size=3) * 1e-3, q])
        we, wJ0, wJ1 = np.zeros(6), np.zeros(36), np.zeros(36)
        REF.ref_edge_se3expmap(_p(x0), _p(x1), _p(z), _p(we), _p(wJ0), _p(wJ1))
        assert np.allclose(oracle.edge_error(G.EDGE_SE3_EXPMAP, x0, x1, z), we, rtol=1e-9, atol=1e-12), (x0, x1, z)
        J0, J1 = oracle.edge_jacobian(G.EDGE_SE3_EXPMAP, x0, x1, z)
        assert np.allclose(J0, wJ0.reshape(6, 6, order="F"), rtol=0, atol=1e-12) and np.allclose(J1, wJ1.reshape(6, 6, order="F"), rtol=0, atol=1e-12)
        X = rng.normal(size=3) + np.array([0, 0, 6.0]); Tc = _rand_se3(rng, 0.1); obs = rng.normal(size=2) * 100; prm = np.array([800.0, 320.0, 240.0])
        wp = np.zeros(2); REF.ref_edge_project_xyz2uv_error(_p(X), _p(Tc), _p(obs), _p(prm), _p(wp))
        assert np.allclose(oracle.edge_error(G.EDGE_PROJECT_XYZ2UV, X, Tc, obs, prm), wp, rtol=1e-13, atol=1e-11)


class RefPcg:
    """The reference's LinearSolverPCG<MatrixType> (g2o/solvers/pcg/linear_solver_pcg.hpp) over its SparseBlockMatrix, compiled; one object per optimisation."""

    def __init__(self, block_size):
        self.h = REF.ref_pcg_create(block_size)

    def __del__(self):
        REF.ref_pcg_destroy(self.h)

    def solve(self, block_indices, colptr, rowidx, values, b, tol=1e-6, max_iter=-1, absolute=True):
        bi, cp, ri = (np.ascontiguousarray(a, dtype=np.int32) for a in (block_indices, colptr, rowidx))
        vals, rhs = np.ascontiguousarray(values, dtype=np.float64), np.ascontiguousarray(b, dtype=np.float64)
        x = np.zeros(int(bi[-1])); it = ctypes.c_int(-1)
        ok = REF.ref_pcg_solve(self.h, len(bi), _p(bi), _p(cp), _p(ri), _p(vals), _p(rhs), _p(x), tol, max_iter, int(absolute), ctypes.byref(it))
        return bool(ok), x, it.value


PCG_CASES = {
    # name: (graph, block size handed to the reference solver; -1 = MatrixX)
    "sphere": (lambda: __import__("g2o_b200.workloads", fromlist=["x"]).sphere(nodes_per_level=8, laps=4), 6),
    "slam2d_schur": (lambda: __import__("g2o_b200.workloads", fromlist=["x"]).slam2d(n_poses=150, n_landmarks=50, world_size=14.0), 3),
    "ba_demo_schur": (lambda: __import__("g2o_b200.workloads", fromlist=["x"]).ba_demo(num_cameras=8, num_points=80), 6),
    "bal_schur": (lambda: __import__("g2o_b200.workloads", fromlist=["x"]).bal_small(), 9),
    "slam2d_points_free": (lambda: __import__("g2o_b200.workloads", fromlist=["x"]).slam2d(n_poses=150, n_landmarks=50, world_size=14.0, marginalize_landmarks=False), -1),
}


@pytest.mark.parametrize("name", list(PCG_CASES))
def test_pcg_matches_the_reference_solver(name):
    """The oracle's restated PCG against the reference's own LinearSolverPCG::solve on the same block matrix and right-hand side: iteration
    counts, solution, and the `_residual` carried into the next solve (three consecutive damped systems, as three LM trials would produce)."""
    fn, bs = PCG_CASES[name]
    g = fn()
    o = oracle.Oracle(g, "lm", "pcg")
    assert o.initialize_optimization() and o.algorithm_init() and o.build_structure()
    o.compute_active_errors(); o.build_system()
    lam = o.compute_lambda_init()
    ref = RefPcg(bs)
    schur = o.do_schur()
    for trial, scale in enumerate((1.0, 10.0, 0.1)):
        o.set_lambda(lam * scale)
        assert o.solve()
        x_o = o.get_f64("x"); res_o, it_o = o.get_f64("pcg_state")
        if schur:
            bi, cp, ri, vals, rhs = o.get_i32("pose_block_indices"), o.get_i32("hschur_colptr"), o.get_i32("hschur_rowidx"), o.get_f64("hschur_values"), o.get_f64("bschur")
        else:
            bi, cp, ri, vals, rhs = o.get_i32("pose_block_indices"), o.get_i32("hpp_colptr"), o.get_i32("hpp_rowidx"), o.get_f64("hpp_values"), o.get_f64("b")
        ok, x_r, it_r = ref.solve(bi, cp, ri, vals, rhs)
        assert ok
        assert it_r == int(it_o), (name, trial, it_r, it_o)
        n = len(x_r)
        assert np.max(np.abs(x_o[:n] - x_r)) <= 1e-9 * np.max(np.abs(x_r)), (name, trial, np.max(np.abs(x_o[:n] - x_r)))
        o.restore_diagonal()
    if not schur:   # BlockSolver::multiplyHessian = _Hpp->multiplySymmetricUpperTriangle on the same (undamped again) matrix as the oracle's
        ok, _, _ = ref.solve(o.get_i32("pose_block_indices"), o.get_i32("hpp_colptr"), o.get_i32("hpp_rowidx"), o.get_f64("hpp_values"), o.get_f64("b"), max_iter=0)
        v = np.random.default_rng(2).normal(size=len(x_r)); dest = np.zeros_like(v)
        REF.ref_multiply_symmetric_upper(ref.h, _p(dest), _p(v))
        mine = o.multiply_hessian(v)
        assert np.max(np.abs(mine - dest)) <= 1e-13 * np.max(np.abs(dest))
    # a tolerance-limited solve with the relative criterion, after LinearSolverPCG::init()
    REF.ref_pcg_init(ref.h)
    o2 = oracle.Oracle(g, "lm", "pcg"); o2.set_pcg_params(tol=1e-12, absolute=False)
    assert o2.initialize_optimization() and o2.algorithm_init() and o2.build_structure()
    o2.compute_active_errors(); o2.build_system(); o2.set_lambda(lam); assert o2.solve()
    names = ("hschur_colptr", "hschur_rowidx", "hschur_values", "bschur") if schur else ("hpp_colptr", "hpp_rowidx", "hpp_values", "b")
    ok, x_r, it_r = ref.solve(o2.get_i32("pose_block_indices"), o2.get_i32(names[0]), o2.get_i32(names[1]), o2.get_f64(names[2]), o2.get_f64(names[3]), tol=1e-12, absolute=False)
    assert it_r == int(o2.get_f64("pcg_state")[1])
    assert np.max(np.abs(o2.get_f64("x")[:len(x_r)] - x_r)) <= 1e-10 * np.max(np.abs(x_r))


# ---------------------------------------------------------------------------------------------------------------------------------------
# The Eigen stand-in against independent implementations.  The reference under oracle/_ref is compiled against oracle/eigen_shim, which
# restates the Eigen algorithms the path calls; the oracle restates them too.  These tests pin the stand-in's routines (exported one by one
# by oracle/shim_probe.cpp) to LAPACK (numpy) and to scipy's Rotation, so that an error common to both restatements cannot hide.
def _p(a):
    return a.ctypes.data_as(__import__("ctypes").c_void_p)


def test_stand_in_inverse_determinant_and_llt_against_lapack():
    L = oracle.shim_probe()
    rng = np.random.default_rng(11)
    for n in (2, 3, 6, 7, 9):
        for trial in range(20):
            A = rng.normal(size=(n, n)); S = A @ A.T + (0.1 + trial) * np.eye(n)          # general and SPD, well and less well conditioned
            for M in (A, S):
                Mf = np.asfortranarray(M); out = np.zeros((n, n), order="F")
                assert L.shim_inverse(n, _p(Mf), _p(out)) == 1
                ref = np.linalg.inv(M)
                assert np.max(np.abs(out - ref)) <= 1e-11 * np.linalg.cond(M) * np.max(np.abs(ref)), (n, trial)
                out2 = np.zeros((n, n), order="F"); L.shim_inverse_dynamic(n, _p(Mf), _p(out2))
                assert np.max(np.abs(out2 - ref)) <= 1e-11 * np.linalg.cond(M) * np.max(np.abs(ref)), (n, trial)
                if n in (2, 3, 6):
                    d = L.shim_determinant(n, _p(Mf)); dr = np.linalg.det(M)
                    assert abs(d - dr) <= 1e-11 * abs(dr) * np.linalg.cond(M), (n, trial)
            if n in (2, 3, 6):                                                             # BaseVertex::solveDirect: H.llt().solve(b)
                b = rng.normal(size=n); x = np.zeros(n); Sf = np.asfortranarray(S)
                assert L.shim_llt_solve(n, _p(Sf), _p(b), _p(x)) == 1
                xr = np.linalg.solve(S, b)
                assert np.max(np.abs(x - xr)) <= 1e-11 * np.linalg.cond(S) * np.max(np.abs(xr))
                bad = np.asfortranarray(S - (np.max(np.linalg.eigvalsh(S)) + 1.0) * np.eye(n))
                assert L.shim_llt_solve(n, _p(bad), _p(b), _p(x)) == 0                      # not positive definite: info() != Success


def test_stand_in_rotations_against_scipy():
    L = oracle.shim_probe()
    rng = np.random.default_rng(12)
    for trial in range(200):
        # all branches of Quaternion(Matrix3): positive trace and each of the three diagonal entries largest (angles near pi)
        rv = rng.normal(size=3); rv *= (rng.uniform(0, np.pi) if trial % 2 else np.pi - 1e-3 * rng.random()) / np.linalg.norm(rv)
        R = Rotation.from_rotvec(rv).as_matrix()
        q = np.zeros(4); L.shim_quat_from_R(_p(np.asfortranarray(R)), _p(q))
        qs = Rotation.from_matrix(R).as_quat()
        assert min(np.max(np.abs(q - qs)), np.max(np.abs(q + qs))) < 1e-12, trial            # same rotation, either sign
        Rb = np.zeros((3, 3), order="F"); L.shim_R_from_quat(_p(q), _p(Rb))
        assert np.max(np.abs(Rb - R)) < 1e-12
        q2 = Rotation.from_rotvec(rng.normal(size=3)).as_quat(); prod = np.zeros(4); L.shim_quat_mul(_p(q), _p(q2), _p(prod))
        ps = (Rotation.from_quat(q) * Rotation.from_quat(q2)).as_quat()
        assert min(np.max(np.abs(prod - ps)), np.max(np.abs(prod + ps))) < 1e-12
        v = rng.normal(size=3); rot = np.zeros(3); L.shim_quat_rotate(_p(q), _p(v), _p(rot))
        assert np.max(np.abs(rot - R @ v)) < 1e-12
        axis = rv / np.linalg.norm(rv); Ra = np.zeros((3, 3), order="F"); L.shim_angle_axis_R(float(np.linalg.norm(rv)), _p(axis), _p(Ra))
        assert np.max(np.abs(Ra - R)) < 1e-12
        # Isometry3: A^-1 B and A v (EdgeSE3's error is (Z^-1 (T0^-1 T1)), isometry3d_mappings / edge_se3.cpp)
        A = np.concatenate([R.ravel(order="F"), rng.normal(size=3)]); R2 = Rotation.from_quat(q2).as_matrix(); B = np.concatenate([R2.ravel(order="F"), rng.normal(size=3)])
        out = np.zeros(12); L.shim_iso_inverse_times(_p(A), _p(B), _p(out))
        assert np.max(np.abs(out[:9].reshape(3, 3, order="F") - R.T @ R2)) < 1e-12 and np.max(np.abs(out[9:] - R.T @ (B[9:] - A[9:]))) < 1e-12
        av = np.zeros(3); L.shim_iso_apply(_p(A), _p(v), _p(av))
        assert np.max(np.abs(av - (R @ v + A[9:]))) < 1e-12
        ang = float(rng.uniform(-np.pi, np.pi)); v2 = rng.normal(size=2); o2 = np.zeros(2); back = np.zeros(1)
        L.shim_rotation2d(ang, _p(v2), _p(o2), _p(back))
        c, s = np.cos(ang), np.sin(ang)
        assert np.max(np.abs(o2 - np.array([c * v2[0] - s * v2[1], s * v2[0] + c * v2[1]]))) < 1e-14 and abs(back[0] - ang) < 1e-14
