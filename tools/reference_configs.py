"""CPU only: the compiled reference (oracle/_ref/libg2o_ref_core.so) against the oracle on BASELINE.json's configurations at full size
(C3 is in profiles/r01_bench_reference_real_g2o_container_8cores.json).  Writes one JSON object per configuration.

    python tools/reference_configs.py > profiles/r01_reference_vs_oracle_configs.json
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from g2o_b200 import workloads as W  # noqa: E402
from oracle import oracle  # noqa: E402

CONFIGS = [
    ("C1 ba_demo 15 x 300, BlockSolver_6_3 + LinearSolverCSparse (BASELINE configs[0])", lambda: W.ba_demo(), "6_3_csparse", "csparse_block", 10),
    ("C1 ba_demo 15 x 300, BlockSolver_6_3 + LinearSolverPCG", lambda: W.ba_demo(), "6_3", "pcg", 10),
    ("C2 create_sphere 100 x 100 (10 000 VertexSE3 / 39 698 EdgeSE3), lm_var = BlockSolverX + CSparse (BASELINE configs[1])", lambda: W.sphere(), "var_csparse", "csparse", 8),
    ("C2 create_sphere 100 x 100, BlockSolverX + LinearSolverPCG", lambda: W.sphere(), "var", "pcg", 8),
    ("C5-shaped slam2d, 20 000 poses / 4 000 landmarks, Huber, BlockSolver_3_2 + PCG", lambda: W.slam2d(n_poses=20000, n_landmarks=4000, world_size=110.0), "3_2", "pcg", 8),
]


def main():
    out = []
    for desc, fn, bs, olin, iters in CONFIGS:
        g = fn()
        t0 = time.perf_counter(); ref = oracle.ReferenceG2o(g, "lm", bs, threads=1); ok = ref.initialize_optimization(); n_r, st_r = ref.optimize(iters); t_ref = time.perf_counter() - t0
        t0 = time.perf_counter(); o = oracle.Oracle(g, "lm", olin); o.initialize_optimization(); n_o, st_o = o.optimize(iters); t_orc = time.perf_counter() - t0
        rel = [abs(a["chi2"] - b["chi2"]) / b["chi2"] for a, b in zip(st_o, st_r)]
        rec = {"config": desc, "vertices": g.n_vertices, "edges": g.n_edges, "iterations": [n_r, n_o],
               "index_map_identical": bool(np.array_equal(ref.hessian_index(), o.get_i32("hessian_index"))),
               "chi2_reference": [s["chi2"] for s in st_r], "chi2_oracle": [s["chi2"] for s in st_o], "chi2_max_relative_difference": max(rel) if rel else None,
               "lm_trials_reference": [int(s["levenbergIterations"]) for s in st_r], "lm_trials_oracle": [int(s["levenbergIterations"]) for s in st_o],
               "pcg_iterations_reference": [int(s["iterationsLinearSolver"]) for s in st_r], "pcg_iterations_oracle": [int(s["iterationsLinearSolver"]) for s in st_o],
               "lambda_reference": ref.current_lambda(), "lambda_oracle": st_o[-1]["lambda"] if st_o else None,
               "estimates_max_abs_difference": float(np.max(np.abs(ref.estimates() - o.estimates()))),
               "seconds": {"reference_1_thread_incl_graph_construction": round(t_ref, 2), "oracle_1_thread": round(t_orc, 2)}}
        out.append(rec)
        print(desc, "max rel chi2 diff", rec["chi2_max_relative_difference"], "trials equal", rec["lm_trials_reference"] == rec["lm_trials_oracle"],
              "pcg equal", rec["pcg_iterations_reference"] == rec["pcg_iterations_oracle"], file=sys.stderr, flush=True)
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
