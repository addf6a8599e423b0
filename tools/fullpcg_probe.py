import time, sys, numpy as np
sys.path.insert(0, "/root/repo")
from g2o_b200 import workloads as W
from g2o_b200.binding import CudaSolver
g = W.slam2d(n_poses=20000, n_landmarks=4000, world_size=110.0, marginalize_landmarks=False)
for timing in (0.0, 1.0):
    s = CudaSolver(g, "lm_var_cuda", device=0); s.initialize_optimization(); s.init(); s.build_structure()
    s.set_property("kernelTiming", timing)
    s.reset_counters()
    t=time.time(); n, st = s.optimize(6); dt=time.time()-t
    its=sum(x["iterations_linear_solver"] for x in st)
    print("kernelTiming", timing, "optimize(6):", round(dt,3), "s; PCG iterations", its, "launches", s.launch_count())
    for ph in ("linear_solver","pcg_spmv","pcg_vec","build","errors"):
        print("   ", ph, s.phase_time(ph))
