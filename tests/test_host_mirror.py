"""The C++ host-side mirror of the reference's plugin interface (g2o_b200/host): registration symbols (CPU) and the
reference's LM/GN unit tests re-expressed in C++ against the CUDA backend (GPU)."""
import ctypes
import os
import subprocess

import pytest

LIBDIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "g2o_b200", "lib")


def test_solver_library_registers_like_a_g2o_plugin():
    so = os.path.join(LIBDIR, "libg2o_solver_cuda.so")
    assert os.path.exists(so), "run __graft_entry__.build()"
    assert "_solver_" in os.path.basename(so)            # the glob g2o uses to find solver plugins (g2o_common.cpp:82)
    L = ctypes.CDLL(so)
    assert hasattr(L, "g2o_optimization_library_cuda")   # G2O_REGISTER_OPTIMIZATION_LIBRARY(cuda)
    for name in ["gn_var_cuda", "lm_var_cuda", "gn_fix3_2_cuda", "lm_fix3_2_cuda", "gn_fix6_3_cuda", "lm_fix6_3_cuda",
                 "gn_fix7_3_cuda", "lm_fix7_3_cuda", "gn_fix9_3_cuda", "lm_fix9_3_cuda",
                 "gn_dense_cuda", "lm_dense_cuda", "gn_dense3_2_cuda", "lm_dense3_2_cuda", "gn_dense6_3_cuda", "lm_dense6_3_cuda",
                 "gn_dense7_3_cuda", "lm_dense7_3_cuda", "gn_dense9_3_cuda", "lm_dense9_3_cuda", "dl_var_cuda"]:
        assert hasattr(L, "g2o_optimization_algorithm_" + name), name


@pytest.mark.gpu
def test_reference_unit_tests_in_cpp():
    exe = os.path.join(LIBDIR, "host_tests")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "HOST_TESTS_OK" in out.stdout
