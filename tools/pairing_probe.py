"""Developer probe: the real g2o + CUDA plugin pairing, case by case, with full output (tests/test_reference_core.py holds the test)."""
import ctypes, os, sys, traceback
import numpy as np
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
from oracle import oracle
from g2o_b200 import workloads as W
oracle.reference_core()
ctypes.CDLL(os.path.join(root, "oracle", "_ref", "libg2o_solver_cuda.so"))
cases = [(W.sphere(nodes_per_level=10, laps=5), 'var', 'lm_var_cuda'), (W.ba_demo(num_cameras=8, num_points=80), '6_3', 'lm_fix6_3_cuda'),
         (W.slam2d(n_poses=200, n_landmarks=60, world_size=16.0), '3_2', 'lm_fix3_2_cuda')]
for g, bs, name in cases:
    try:
        cpu = oracle.ReferenceG2o(g, 'lm', bs, threads=1); assert cpu.initialize_optimization(); n1, s1 = cpu.optimize(5)
        gpu = oracle.ReferenceG2o(g, 'factory', name); assert gpu.initialize_optimization(); n2, s2 = gpu.optimize(5)
        print(name, "n", n1, n2)
        for i, (a, b) in enumerate(zip(s2, s1)):
            print("  it", i, "gpu", a['chi2'], "cpu", b['chi2'], "rel", abs(a['chi2'] - b['chi2']) / b['chi2'])
        e1, e2 = cpu.estimates(), gpu.estimates()
        print("  est diff", np.max(np.abs(e1 - e2) / (1 + np.abs(e1))))
    except Exception:
        traceback.print_exc()
    sys.stdout.flush()
