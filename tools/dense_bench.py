"""Dense FP64 Cholesky of the C3 reduced camera system (n = 16002): time and DMMA-pipe utilisation of the factorisation.
Not the bench metric (PCG is the north-star solver at this size); reported under profiles/ as the a11 'dense' row's measurement."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from g2o_b200 import workloads as W
from g2o_b200.binding import CudaSolver

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
g = W.bal_venice() if scale == 1.0 else W.bal_synthetic(n_cameras=int(1778 * scale ** 0.5), n_points=int(993923 * scale), n_obs=int(5001946 * scale), k_max=min(500, int(1778 * scale ** 0.5)))
s = CudaSolver(g, "lm_dense9_3_cuda", device=0)
s.initialize_optimization(); s.init()
stats = [s.solver_iteration(i) for i in range(3)]
s.reset_counters()
stats += [s.solver_iteration(i) for i in range(3, 5)]
n = int(s.get_i32("dims")[2])
sec, _, calls = s.phase_time("dense_cholesky")
asec, _, _ = s.phase_time("dense_assemble")
flops = n ** 3 / 3.0 + 2.0 * n * n
peak = json.load(open(os.path.join(ROOT, "profiles", "fp64_peaks_b200.json")))["dmma_tflops"]
print(json.dumps({"n": n, "calls": calls, "dense_cholesky_ms": 1e3 * sec / calls, "dense_assemble_ms": 1e3 * asec / calls, "tflops": flops / (sec / calls) / 1e12,
                  "frac_of_fp64_tensor_peak": flops / (sec / calls) / 1e12 / peak, "chi2": [st["chi2"] for st in stats], "trials": [st["levenberg_iterations"] for st in stats]}))
