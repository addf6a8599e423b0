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
                 "gn_fix9_3_cuda", "lm_fix9_3_cuda",
                 "gn_dense_cuda", "lm_dense_cuda", "gn_dense3_2_cuda", "lm_dense3_2_cuda", "gn_dense6_3_cuda", "lm_dense6_3_cuda",
                 "gn_dense9_3_cuda", "lm_dense9_3_cuda", "dl_var_cuda"]:
        assert hasattr(L, "g2o_optimization_algorithm_" + name), name
    assert not hasattr(L, "g2o_optimization_algorithm_lm_fix7_3_cuda")      # no sim3 types in the backend: the name is not offered


@pytest.mark.gpu
def test_reference_unit_tests_in_cpp():
    exe = os.path.join(LIBDIR, "host_tests")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "HOST_TESTS_OK" in out.stdout


def test_batch_statistics_line_has_the_reference_format(tmp_path):
    """`g2o -stats file` writes one `operator<<(G2OBatchStatistics)` line per iteration: "name= value\\t " for 22 fields in a fixed
    order (core/batch_stats.cpp:48-83); scripts that parse those files must keep working with the mirror's stream operator."""
    src = tmp_path / "stats.cpp"
    src.write_text('#include <iostream>\n#include "g2o_mirror.hpp"\n'
                   'int main() { g2o::G2OBatchStatistics s; std::cout << s << std::endl; s.iteration = 3; s.numVertices = 7; s.numEdges = 9; s.chi2 = 1.5;\n'
                   '  s.levenbergIterations = 2; s.iterationsLinearSolver = 11; s.hessianPoseDimension = 12; s.hessianLandmarkDimension = 30; s.hessianDimension = 42; std::cout << s << std::endl; }\n')
    exe = tmp_path / "stats"
    host = os.path.join(os.path.dirname(LIBDIR), "host")
    subprocess.run(["g++", "-std=c++17", "-I", host, str(src), "-o", str(exe), "-L", LIBDIR, "-lg2o_solver_cuda", "-lg2ocu", f"-Wl,-rpath,{LIBDIR}"], check=True)
    lines = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines()
    order = ["iteration", "numVertices", "numEdges", "chi2", "timeLinearSolution", "iterationsLinearSolver", "timeQrDecomposition", "timeResiduals",
             "timeLinearize", "timeQuadraticForm", "timeSchurComplement", "timeSymbolicDecomposition", "timeNumericDecomposition", "timeUpdate",
             "timeIteration", "levenbergIterations", "timeLinearSolver", "hessianDimension", "hessianPoseDimension", "hessianLandmarkDimension",
             "choleskyNNZ", "timeMarginals"]
    for line, want in zip(lines, [{"iteration": "-1"}, {"iteration": "3", "numVertices": "7", "numEdges": "9", "chi2": "1.5", "levenbergIterations": "2",
                                                        "iterationsLinearSolver": "11", "hessianDimension": "42", "hessianPoseDimension": "12", "hessianLandmarkDimension": "30"}]):
        fields = [f for f in line.split("\t ") if f]
        assert [f.split("= ")[0] for f in fields] == order
        got = dict(f.split("= ") for f in fields)
        for k in order:
            assert got[k] == want.get(k, "0"), (k, got[k])
