"""Every configuration BASELINE.json names, through the CUDA backend on one GPU: LM iterations/s (device-timed phases summed from
CUDA events, wall clock for the total), chi2 trajectory, and - where the CPU oracle finishes in seconds - parity against it.
C3 is the bench.py workload; the others are parity / coverage cases.  Output: one JSON object (stdout)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from g2o_b200 import graph as G
from g2o_b200 import workloads as W
from g2o_b200.binding import CudaSolver

which = sys.argv[1].split(",") if len(sys.argv) > 1 else ["C1", "C2", "C3", "C4", "C5"]
CASES = {
    "C1": ("ba_demo 15 cameras / 300 points, EdgeSE3ProjectXYZ, BlockSolver_6_3", lambda: W.ba_demo(), "lm_fix6_3_cuda", 10, True),
    "C2": ("create_sphere 10 000 VertexSE3 / 39 698 EdgeSE3, lm_var", lambda: W.sphere(nodes_per_level=100, laps=100), "lm_var_cuda", 10, True),
    "C3": ("BAL-Venice-shaped 1778 / 993 923 / 5 001 946, Huber, BlockSolver<9,3>", lambda: W.bal_venice(), "lm_fix9_3_cuda", 8, False),
    "C4": ("BAL-shaped 10 000 / 4 000 000 / 20 000 000, Huber, BlockSolver<9,3>", lambda: W.bal_large(), "lm_fix9_3_cuda", 6, False),
    "C5": ("simulator2d-shaped 100 000 poses / 20 000 landmarks, Huber, Schur on landmarks (lm_fix3_2)", lambda: W.slam2d(), "lm_fix3_2_cuda", 8, False),
}
out = {}
for key in which:
    desc, fn, solver, iters, with_oracle = CASES[key]
    t0 = time.perf_counter(); g = fn(); t_gen = time.perf_counter() - t0
    s = CudaSolver(g, solver, device=0)
    t0 = time.perf_counter(); s.initialize_optimization(); s.init(); s.build_structure(); t_struct = time.perf_counter() - t0
    st = []
    s.solver_iteration(0); s.set_estimates(g.v_estimate); s.init()     # warm-up iteration (allocations, first-touch), then restart
    t0 = time.perf_counter()
    for i in range(iters):
        st.append(s.solver_iteration(i))
    dt = time.perf_counter() - t0
    rec = {"config": desc, "solver": solver, "vertices": int(g.n_vertices), "edges": int(g.n_edges), "iterations": iters, "lm_iterations_per_s": iters / dt,
           "ms_per_iteration": 1e3 * dt / iters, "structure_build_s": t_struct, "generate_s": t_gen,
           "chi2": [x["chi2"] for x in st], "trials": [x["levenberg_iterations"] for x in st], "linear_iterations": [x["iterations_linear_solver"] for x in st],
           "dims": [int(v) for v in s.get_i32("dims")]}
    if with_oracle:
        from oracle.oracle import Oracle
        o = Oracle(g, "lm", "pcg"); o.initialize_optimization()
        t0 = time.perf_counter(); n, ost = o.optimize(iters); rec["oracle_s"] = time.perf_counter() - t0
        rec["oracle_chi2"] = [x["chi2"] for x in ost]
        rec["max_rel_chi2_diff_vs_oracle"] = float(max(abs(a["chi2"] - b["chi2"]) / max(abs(b["chi2"]), 1e-300) for a, b in zip(st, ost)))
        eo = o.estimates()
        rec["max_rel_estimate_diff_vs_oracle"] = float(np.max(np.abs(s.get_estimates() - eo) / (1 + np.abs(eo))))
    out[key] = rec
    print(key, json.dumps(rec), file=sys.stderr, flush=True)
    del s
print(json.dumps(out))
