"""Developer helper for ncu: a few LM iterations of a BASELINE workload through the C ABI (python tools/profile_step.py [c3] [iterations])."""
import os, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
import bench
from g2o_b200.binding import CudaSolver
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
g, desc, solver = bench.workload(name, 1.0)
s = CudaSolver(g, solver, device=0)
s.initialize_optimization()
n, st = s.optimize(iters)
print(desc, n, [round(x["chi2"], 3) for x in st], "launches", s.launch_count())
