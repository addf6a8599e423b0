"""Generates tests/golden/*.json with the CPU oracle (run here, in the build container):

    python tests/golden/make_golden.py          (adds the cases missing from the file; --all regenerates every case)

The reference holds no fixtures for this path.  The vectors are written by the oracle; tests/test_golden.py checks that the
REAL reference (oracle/_ref/libg2o_ref_core.so: g2o/core + BlockSolver + LM / GN + LinearSolverPCG + the types compiled from /root/reference)
reproduces every PCG case of them - so the GPU tests, which only read this file, compare the CUDA path with what the reference itself
produces, without needing the reference or the oracle on the GPU box.  Inputs are the seeded synthetic graphs of g2o_b200/workloads.py."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from g2o_b200 import graph as G  # noqa: E402
from g2o_b200 import workloads as W  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

CASES = {
    "ba_demo_15x300": (lambda: W.ba_demo(), "lm", "pcg"),
    "ba_demo_huber_outliers": (lambda: W.ba_demo(edge_type=G.EDGE_PROJECT_XYZ2UV, robust_kernel=True, outlier_ratio=0.05), "lm", "pcg"),
    "bal_small": (lambda: W.bal_small(), "lm", "pcg"),
    "sphere_16x8": (lambda: W.sphere(nodes_per_level=16, laps=8), "lm", "pcg"),
    "slam2d_800": (lambda: W.slam2d(n_poses=800, n_landmarks=200, world_size=30.0), "lm", "pcg"),
    "sphere_gn": (lambda: W.sphere(nodes_per_level=12, laps=6), "gn", "pcg"),
    # poses and points in one system (nothing marginalized, BlockSolverX with two block sizes), PCG over the whole matrix
    "slam2d_points_free": (lambda: W.slam2d(n_poses=800, n_landmarks=200, world_size=30.0, marginalize_landmarks=False), "lm", "pcg"),
    # Powell's dogleg with an exact linear solver
    "slam2d_dogleg": (lambda: W.slam2d(n_poses=800, n_landmarks=200, world_size=30.0), "dl", "dense"),
    # round 2: VertexSE3Expmap / EdgeSE3Expmap as a pose graph and next to projection edges in one BA; a BAL graph under the Cauchy kernel
    "sphere_expmap": (lambda: W.sphere_expmap(nodes_per_level=16, laps=8), "lm", "pcg"),
    "ba_pose_constraints": (lambda: W.ba_demo_with_pose_constraints(), "lm", "pcg"),
    "bal_cauchy": (lambda: _with_kernel(W.bal_synthetic(n_cameras=24, n_points=1200, n_obs=6000, seed=4, k_max=16, min_window=4, outlier_fraction=0.05), G.KERNEL_CAUCHY, 1.5), "lm", "pcg"),
}


def _with_kernel(g, kind, delta):
    g.e_kernel = np.full(g.n_edges, kind, dtype=np.int32); g.e_kernel_delta = np.full(g.n_edges, float(delta)); return g



def main():
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lm_trajectories.json")
    out = json.load(open(path)) if os.path.exists(path) and "--all" not in sys.argv else {}   # default: only add the cases that are missing
    for name, (fn, alg, lin) in CASES.items():
        if name in out:
            continue
        g = fn()
        o = Oracle(g, alg, lin); o.initialize_optimization(); o.algorithm_init(); o.build_structure()
        o.compute_active_errors(); o.build_system()
        rec = {"n_vertices": g.n_vertices, "n_edges": g.n_edges, "chi2_0": o.active_robust_chi2(), "plain_chi2_0": o.active_chi2(),
               "dims": o.get_i32("dims").tolist(), "b_head": o.get_f64("b")[:24].tolist(), "b_sum": float(np.sum(o.get_f64("b"))),
               "hpp_sum": float(np.sum(o.get_f64("hpp_values"))), "lambda_init": o.compute_lambda_init(),
               "hschur_nnz_blocks": 0,
               "structure_checksum": int(np.sum(o.get_i32("hessian_index").astype(np.int64) * (np.arange(g.n_vertices) % 97 + 1)))}
        if o.do_schur():   # the reference adds Hpp's blocks to the Schur pattern on its first solve (block_solver.hpp:333-335)
            o.set_lambda(1.0); o.solve(); o.restore_diagonal()
            rec["hschur_nnz_blocks"] = int(o.get_i32("hschur_colptr")[-1])
        o2 = Oracle(g, alg, lin); o2.initialize_optimization()
        n, st = o2.optimize(6)
        rec.update({"iterations": n, "chi2": [s["chi2"] for s in st], "lambda": [s["lambda"] for s in st],
                    "trials": [int(s["levenbergIterations"]) for s in st], "estimate_head": o2.estimates()[:32].tolist(),
                    "algorithm": alg})
        out[name] = rec
        print(name, rec["chi2_0"], rec["chi2"][-1])
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
