"""Developer helper: wall time of the one-time set-up of a workload (set_graph, initializeOptimization, buildStructure incl. the device lists) next to
the LM iterations that follow (python tools/setup_time.py [c3])."""
import os, sys, time
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
os.environ.setdefault("G2OCU_TRACE", "1")
import bench
from g2o_b200.binding import CudaSolver
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
g, desc, solver = bench.workload(name, 1.0)
for rep in range(2):
    t0 = time.time(); s = CudaSolver(g, solver, device=0); t1 = time.time()
    s.initialize_optimization(); t2 = time.time()
    s.init(); s.build_structure(); t3 = time.time()
    n, st = s.optimize(5); t4 = time.time()
    print(f"rep {rep}: create + set_graph {t1 - t0:.3f} s, initialize_optimization {t2 - t1:.3f} s, build_structure {t3 - t2:.3f} s, optimize(5) {t4 - t3:.3f} s", flush=True)
