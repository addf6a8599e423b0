"""The reference's `g2o` command line re-created over the host mirror (g2o_b200/host/g2o_cli.cpp): .g2o / BAL text I/O on the CPU,
and on the GPU the same answers as the Python binding for the same graph."""
import os
import re
import subprocess

import numpy as np
import pytest

from g2o_b200 import workloads as W

LIBDIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "g2o_b200", "lib")
CLI = os.path.join(LIBDIR, "g2o_cuda")


def _numbers(path):
    out = []
    for line in open(path):
        tok = line.split()
        out.append((tok[0], [float(t) for t in tok[1:]]))
    return out


def test_g2o_files_round_trip_through_the_reader_and_writer(tmp_path):
    assert os.path.exists(CLI), "run __graft_entry__.build()"
    for name, g in [("sphere", W.sphere(nodes_per_level=8, laps=4)), ("slam2d", W.slam2d(n_poses=120, n_landmarks=40, world_size=12.0))]:
        a, b = str(tmp_path / f"{name}.g2o"), str(tmp_path / f"{name}_resaved.g2o")
        W.write_g2o(g, a)
        r = subprocess.run([CLI, "-summary", "-o", b, a], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stderr
        m = re.match(r"vertices (\d+) edges (\d+) fixed (\d+)", r.stdout)
        assert m and int(m.group(1)) == g.n_vertices and int(m.group(2)) == g.n_edges and int(m.group(3)) == int(g.v_fixed.sum())
        # FIX lines are written right after their vertex by both writers; everything else must agree to rounding of the quaternion round trip
        xa, xb = _numbers(a), _numbers(b)
        assert [t for t, _ in xa] == [t for t, _ in xb]
        for (_, va), (_, vb) in zip(xa, xb):
            assert np.allclose(va, vb, rtol=1e-13, atol=1e-13)


def test_unknown_tags_are_skipped_like_the_reference(tmp_path):
    p = tmp_path / "mixed.g2o"
    p.write_text("# comment\nVERTEX_SE2 0 0 0 0\nVERTEX_SE2 1 1 0 0\nFIX 0\nVERTEX_FANCY 7 1 2 3\nEDGE_SE2 0 1 1 0 0 500 0 0 500 0 5000\n")
    r = subprocess.run([CLI, "-summary", str(p)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "vertices 2 edges 1 fixed 1" in r.stdout and "skipped_lines 1" in r.stdout
    assert "unknown type: VERTEX_FANCY" in r.stderr


def _cli_chi2(args):
    r = subprocess.run([CLI] + args, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    m = re.search(r"iterations (\d+) chi2 (\S+) robust_chi2 (\S+)", r.stdout)
    return int(m.group(1)), float(m.group(2)), float(m.group(3))


@pytest.mark.gpu
def test_cli_matches_the_binding(tmp_path):
    from g2o_b200.binding import CudaSolver
    cases = [("sphere", W.sphere(nodes_per_level=10, laps=5), "lm_var_cuda", []),
             ("slam2d", W.slam2d(n_poses=400, n_landmarks=100, world_size=20.0), "lm_fix3_2_cuda", ["-robustKernel", "Huber"]),
             # lm_var does not ask for marginalization (g2o.cpp:318-331): poses and points stay in one system
             ("slam2d_free", W.slam2d(n_poses=400, n_landmarks=100, world_size=20.0, marginalize_landmarks=False), "lm_var_cuda", ["-robustKernel", "Huber"])]
    for name, g, solver, extra in cases:
        path = str(tmp_path / f"{name}.g2o")
        W.write_g2o(g, path)
        n, chi2, rchi2 = _cli_chi2(["-i", "6", "-solver", solver, "-o", str(tmp_path / "out.g2o")] + extra + [path])
        s = CudaSolver(g, solver, device=0); s.initialize_optimization()
        nb, st = s.optimize(6)
        assert n == nb
        assert abs(rchi2 - st[-1]["chi2"]) <= 1e-7 * abs(st[-1]["chi2"]), (name, rchi2, st[-1]["chi2"])
        assert os.path.getsize(str(tmp_path / "out.g2o")) > 0
    g = W.bal_synthetic(n_cameras=20, n_points=800, n_obs=4000, seed=2, k_max=12, min_window=4)
    path = str(tmp_path / "problem.txt")
    W.write_bal(g, path)
    n, chi2, rchi2 = _cli_chi2(["-bal", "-i", "5", "-solver", "lm_fix9_3_cuda", "-robustKernel", "Huber", path])
    s = CudaSolver(g, "lm_fix9_3_cuda", device=0); s.initialize_optimization()
    nb, st = s.optimize(5)
    assert n == nb and abs(rchi2 - st[-1]["chi2"]) <= 1e-7 * abs(st[-1]["chi2"])


def test_expmap_and_camera_parameter_tags(tmp_path):
    """VERTEX_SE3:EXPMAP files hold camera-to-world poses (types_six_dof_expmap.cpp:92-112): the reader inverts, the writer inverts back;
    EDGE_PROJECT_XYZ2UV:EXPMAP refers to a PARAMS_CAMERAPARAMETERS line (types_six_dof_expmap.cpp:46,190-215)."""
    a, b = tmp_path / "ba.g2o", tmp_path / "ba_resaved.g2o"
    q = np.array([0.1, -0.2, 0.05, 0.0]); q[3] = np.sqrt(1 - np.sum(q[:3] ** 2))
    a.write_text("PARAMS_CAMERAPARAMETERS 0 1000 320 240 0\n"
                 f"VERTEX_SE3:EXPMAP 0 0.5 -0.25 0.125 {float(q[0])!r} {float(q[1])!r} {float(q[2])!r} {float(q[3])!r}\n"
                 "VERTEX_SE3:EXPMAP 1 0 0 0 0 0 0 1\nFIX 1\n"
                 "VERTEX_XYZ 2 0.3 0.1 4.0\n"
                 "EDGE_PROJECT_XYZ2UV:EXPMAP 2 0 0 331.5 250.25 1 0 1\n"
                 "EDGE_PROJECT_XYZ2UV:EXPMAP 2 1 0 395 265 2 0.5 2\n")
    r = subprocess.run([CLI, "-summary", "-o", str(b), str(a)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    assert "vertices 3 edges 2 fixed 1 dimensions 3 6" in r.stdout
    xa, xb = _numbers(str(a)), _numbers(str(b))
    key = lambda rec: (rec[0], rec[1][:1])
    assert sorted(t for t, _ in xa) == sorted(t for t, _ in xb)
    for (ta, va), (tb, vb) in zip(sorted(xa, key=key), sorted(xb, key=key)):
        assert ta == tb and np.allclose(va, vb, rtol=1e-13, atol=1e-13), (ta, va, vb)
    # an edge that names an unknown parameter id is an error, not a silent default
    bad = tmp_path / "bad.g2o"
    bad.write_text("VERTEX_SE3:EXPMAP 0 0 0 0 0 0 0 1\nVERTEX_XYZ 1 0 0 1\nEDGE_PROJECT_XYZ2UV:EXPMAP 1 0 7 1 1 1 0 1\n")
    r = subprocess.run([CLI, "-summary", str(bad)], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "PARAMS_CAMERAPARAMETERS" in r.stderr


@pytest.mark.gpu
def test_cli_runs_dogleg(tmp_path):
    """`-solver dl_var_cuda` with Dogleg's own property names (optimization_algorithm_dogleg.cpp:44-47) and its verbose line (:199-207)."""
    from g2o_b200.binding import CudaSolver
    g = W.sphere(nodes_per_level=10, laps=5)
    path = str(tmp_path / "sphere.g2o")
    W.write_g2o(g, path)
    r = subprocess.run([CLI, "-v", "-i", "4", "-solver", "dl_var_cuda", "-solverProperties", "initialDelta=50,maxTrialsAfterFailure=20", path],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "Delta=" in r.stderr and "step=" in r.stderr and "tries=" in r.stderr
    m = re.search(r"iterations (\d+) chi2 (\S+) robust_chi2 (\S+)", r.stdout)
    s = CudaSolver(g, "dl_var_cuda", device=0); s.initialize_optimization(); s.compute_active_errors()
    assert int(m.group(1)) > 0 and float(m.group(3)) < s.active_robust_chi2()


@pytest.mark.gpu
def test_cli_compute_marginals(tmp_path):
    """`-computeMarginals` (g2o.cpp:581-608): per active vertex the blocks (h, h) and (h - 1, h) of the inverse of the system matrix on stderr;
    the printed diagonal block of the last vertex against the binding's answer for the same graph and iterations."""
    from g2o_b200.binding import CudaSolver
    g = W.sphere(nodes_per_level=8, laps=4)
    path = str(tmp_path / "sphere.g2o")
    W.write_g2o(g, path)
    r = subprocess.run([CLI, "-i", "3", "-solver", "gn_var_cuda", "-computeMarginals", path], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    s = CudaSolver(g, "gn_var_cuda", device=0); s.initialize_optimization(); s.optimize(3)
    nb = int(s.get_i32("dims")[0])
    assert r.stderr.count("inv block :") == 2 * nb - 1
    want = s.compute_marginals([(nb - 1, nb - 1)])[0]
    text = r.stderr[r.stderr.index(f"inv block :{nb - 1}, {nb - 1}"):].splitlines()[1:7]
    got = np.array([[float(x) for x in line.split()] for line in text])
    assert got.shape == (6, 6) and np.max(np.abs(got - want)) <= 1e-5 * np.max(np.abs(want))
