"""Small end-to-end cases for compute-sanitizer (one tool per gpurun call):
   compute-sanitizer --tool memcheck python tools/sanitizer_case.py
Covers the BAL path (Schur tile kernel on the FP64 tensor pipe, segmented short tracks, one-launch CG tail), a pose graph and a 2-D SLAM graph."""
import os, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
from g2o_b200 import workloads as W
from g2o_b200.binding import CudaSolver
cases = [(W.bal_synthetic(n_cameras=60, n_points=6000, n_obs=30000, seed=5, k_max=40, min_window=4), "lm_fix9_3_cuda"),
         (W.ba_demo(), "lm_fix6_3_cuda"), (W.sphere(nodes_per_level=12, laps=6), "lm_var_cuda"),
         (W.slam2d(n_poses=300, n_landmarks=80, world_size=20.0), "lm_fix3_2_cuda")]
for g, name in cases:
    s = CudaSolver(g, name, device=0)
    s.initialize_optimization()
    n, st = s.optimize(2)
    print(name, n, [round(x["chi2"], 4) for x in st], flush=True)
print("SANITIZER_CASE_OK")
