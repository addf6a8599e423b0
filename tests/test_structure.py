"""Bit-exact block index map: host half of the backend vs the oracle's restatement of
SparseOptimizer::initializeOptimization (sparse_optimizer.cpp:208-279,168-193) and BlockSolver::buildStructure
(block_solver.hpp:103-256).  Arrays compared are the ones listed in SURVEY.md Appendix B.  Runs without a GPU:
g2ocu_build_structure does the integer work first and only then needs a device."""
import numpy as np
import pytest

from g2o_b200 import workloads as W
from g2o_b200 import graph as G
from g2o_b200.binding import CudaSolver, G2oCudaError
from g2o_b200 import _lib
from oracle.oracle import Oracle

NAMES_ALWAYS = ["hessian_index", "active_vertices", "active_edges", "index_mapping", "dims", "pose_block_indices",
                "hpp_colptr", "hpp_rowidx", "edge_targets"]
NAMES_SCHUR = ["landmark_block_indices", "hpl_colptr", "hpl_rowidx", "hschur_colptr", "hschur_rowidx", "hschur_t_colptr",
               "hschur_t_rowidx"]


def host_structure(graph, name="lm_var_cuda", level=0):
    s = CudaSolver(graph, name)
    s.initialize_optimization(level)
    try:
        s.build_structure()
    except G2oCudaError as e:          # no GPU here: the host structure is complete, the device upload is not
        assert e.code == _lib.E_CUDA, str(e)
    return s


def check(graph, level=0):
    s = host_structure(graph, level=level)
    o = Oracle(graph)
    assert o.initialize_optimization(level)
    assert o.algorithm_init()
    assert o.build_structure()
    if o.do_schur():
        # the reference extends the Schur pattern by Hpp's blocks on its first solve (_Hpp->add(*_Hschur),
        # block_solver.hpp:333-335); the backend has them from the start
        o.compute_active_errors(); o.build_system(); o.set_lambda(1.0); o.solve(); o.restore_diagonal()
    for n in NAMES_ALWAYS + (NAMES_SCHUR if o.do_schur() else []):
        a, b = s.get_i32(n), o.get_i32(n)
        assert a.shape == b.shape and np.array_equal(a, b), n
    return s, o


def test_ba_demo_structure():
    check(W.ba_demo())
    check(W.ba_demo(edge_type=G.EDGE_PROJECT_XYZ2UV, robust_kernel=True))


def test_bal_structure():
    check(W.bal_small())
    check(W.bal_synthetic(n_cameras=40, n_points=3000, n_obs=14000, seed=3, k_max=30, min_window=4))


def test_sphere_structure():
    check(W.sphere(nodes_per_level=12, laps=6))


def test_slam2d_structure():
    check(W.slam2d(n_poses=600, n_landmarks=150, world_size=30.0))
    check(W.slam2d(n_poses=300, n_landmarks=400, world_size=60.0))      # many landmarks never observed -> inactive


def test_points_not_marginalized():
    """`lm_var` on graphs with poses and points and no marginalization (BlockSolverX, one Hpp with blocks of two sizes): the arrays of
    the reference's buildStructure, and the permutation between its vector order and the internal [poses | points] layout."""
    graphs = [W.slam2d(n_poses=150, n_landmarks=50, world_size=14.0, marginalize_landmarks=False), W.ba_demo(num_cameras=6, num_points=40), W.bal_small()]
    for g in graphs:
        g.v_marginalized = np.zeros_like(g.v_marginalized)
        s, o = check(g)
        assert not o.do_schur()
        nposes, npoints, sp, sl = s.get_i32("internal_dims")
        perm = s.get_i32("full_system_permutation")
        assert sorted(perm.tolist()) == list(range(sp + sl)) and int(s.get_i32("dims")[2]) == sp + sl
        # poses keep their relative order inside the pose part, points inside the point part
        pose_part, point_part = perm[perm < sp], perm[perm >= sp]
        assert np.all(np.diff(pose_part) > 0) and np.all(np.diff(point_part) > 0)


def test_fixed_and_levels():
    g = W.bal_small()
    g.v_fixed[2] = 1                      # a fixed camera: its edges keep only the landmark diagonal
    g.v_fixed[g.meta["n_cameras"] + 5] = 1   # a fixed point
    g.e_level[::7] = 1                    # level-1 edges are inactive at level 0 but still shape the Schur pattern
    check(g)
    check(g, level=1)
    check(g, level=-1)


def test_parallel_edges_share_a_block():
    g = W.bal_small()
    dup = slice(0, 20)
    g2 = G.Graph(v_id=g.v_id, v_type=g.v_type, v_fixed=g.v_fixed, v_marginalized=g.v_marginalized, v_estimate=g.v_estimate,
                 e_type=np.concatenate([g.e_type, g.e_type[dup]]), e_v0=np.concatenate([g.e_v0, g.e_v0[dup]]),
                 e_v1=np.concatenate([g.e_v1, g.e_v1[dup]]), e_measurement=np.concatenate([g.e_measurement, g.e_measurement[:40]]),
                 e_information=np.concatenate([g.e_information, g.e_information[:80]]))
    check(g2)


def test_rejections():
    g = W.bal_small()
    bad = g.copy(); bad.e_type = bad.e_type.copy(); bad.e_type[0] = 9
    with pytest.raises(ValueError):
        bad.validate()
    s = CudaSolver(solver_name="lm_fix9_3_cuda")
    cg = bad.as_c()
    import ctypes
    rc = s._L.g2ocu_set_graph(s._h, ctypes.byref(cg))
    assert rc == _lib.E_UNSUPPORTED and b"unsupported edge type" in s._L.g2ocu_last_error(s._h)
    # some points marginalized and some not: two block sizes inside the pose block next to a Schur complement is rejected, not mishandled
    g3 = W.bal_small(); g3.v_marginalized = g3.v_marginalized.copy()
    g3.v_marginalized[np.flatnonzero(g3.v_marginalized)[::2]] = 0
    s3 = CudaSolver(g3, "lm_var_cuda"); s3.initialize_optimization()
    with pytest.raises(G2oCudaError) as ei:
        s3.build_structure()
    assert ei.value.code == _lib.E_UNSUPPORTED
    # empty graph
    s4 = CudaSolver(solver_name="lm_var_cuda")
    with pytest.raises(G2oCudaError):
        s4.initialize_optimization()


def test_library_exports_every_declared_symbol():
    import re, os
    hdr = open(os.path.join(os.path.dirname(_lib.__file__), "..", "include", "g2ocu.h")).read()
    declared = set(re.findall(r"\b(g2ocu_[a-z0-9_]+)\s*\(", hdr)) - {"g2ocu_allreduce_fn"}
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert declared == set(_lib.EXPORTS)
    assert L.g2ocu_version() == 1


def test_every_algorithm_fails_loudly_without_a_device():
    """No CPU fallback: on a box without a GPU the iteration entry point reports G2OCU_E_CUDA for GN, LM and Dogleg alike
    (on a GPU box this test has nothing to check)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    g = W.sphere(nodes_per_level=6, laps=3)
    for name in ("gn_var_cuda", "lm_var_cuda", "dl_var_cuda"):
        s = CudaSolver(g, name); s.initialize_optimization()
        with pytest.raises(G2oCudaError) as ei:
            s.optimize(2)
        assert ei.value.code == _lib.E_CUDA, str(ei.value)
    s = CudaSolver(g, "lm_var_cuda"); s.initialize_optimization(); s.init()
    import ctypes
    rc = s._L.g2ocu_solver_iteration(s._h, 7, 0, None)      # unknown algorithm code
    assert rc < 0


def test_sharded_handles_reject_what_they_cannot_do():
    """Dogleg and graphs whose points are not marginalized run on single-GPU handles only; a sharded handle says so instead of
    computing something else (checked before any device work, so this runs without a GPU)."""
    hook = lambda buf, count, op, stream: 0
    g = W.slam2d(n_poses=60, n_landmarks=20, world_size=10.0, marginalize_landmarks=False)
    s = CudaSolver(g, "lm_var_cuda"); s.set_shard(0, 2, hook); s.initialize_optimization()
    with pytest.raises(G2oCudaError) as ei:
        s.build_structure()
    assert ei.value.code == _lib.E_UNSUPPORTED and "not marginalized" in str(ei.value)
    g = W.sphere(nodes_per_level=6, laps=3)
    s = CudaSolver(g, "dl_var_cuda"); s.set_shard(1, 2, hook); s.initialize_optimization()
    with pytest.raises(G2oCudaError) as ei:
        s.optimize(1)
    assert ei.value.code == _lib.E_UNSUPPORTED and "Dogleg" in str(ei.value)
