"""Generates tests/golden/marginals.json with the CPU oracle (run here, in the build container):

    python tests/golden/make_golden_marginals.py

Blocks of the inverse of Hpp (SparseOptimizer::computeMarginals, sparse_optimizer.cpp:594-596) at the initial estimates of seeded synthetic graphs:
what `Oracle.compute_marginals` restates.  tests/test_golden.py checks that the REAL reference (compiled from /root/reference, LinearSolverCSparse's
solvePattern + MarginalCovarianceCholesky) reproduces the file on the CPU, and that the CUDA path reproduces it on the GPU box - where neither the
reference nor the oracle is needed for this check."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from g2o_b200 import workloads as W  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

# name -> (graph, reference block solver with CSparse, CUDA solver name)
CASES = {
    "sphere_8x4": (lambda: W.sphere(nodes_per_level=8, laps=4), "var_csparse", "gn_var_cuda"),
    "slam2d_chain_120": (lambda: W.slam2d(n_poses=120, n_landmarks=0, world_size=12.0), "3_2_csparse", "gn_fix3_2_cuda"),
    "slam2d_points_in_the_system": (lambda: W.slam2d(n_poses=80, n_landmarks=25, world_size=10.0, marginalize_landmarks=False), "var_csparse", "gn_var_cuda"),
}


def pairs_of(n):
    rng = np.random.default_rng(n)
    return [(0, 0), (n - 1, n - 1), (0, n - 1), (n // 2, n // 3)] + [(int(a), int(b)) for a, b in rng.integers(0, n, size=(4, 2))]


def main():
    out = {}
    for name, (fn, _, _) in CASES.items():
        g = fn()
        o = Oracle(g, "gn", "pcg"); o.initialize_optimization(); o.algorithm_init(); o.build_structure(); o.compute_active_errors(); o.build_system()
        n = len(o.get_i32("pose_block_indices"))
        pairs = pairs_of(n)
        blocks = o.compute_marginals(pairs)
        out[name] = {"n_blocks": n, "pairs": pairs, "blocks": [b.ravel(order="F").tolist() for b in blocks], "shapes": [list(b.shape) for b in blocks]}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "marginals.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=0)
    print("wrote", path, {k: len(v["pairs"]) for k, v in out.items()})


if __name__ == "__main__":
    main()
