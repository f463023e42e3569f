"""The C++ host mirror compiles against include/zkmsm.h and links libzkmsm.so (CPU only: compile + link, no run)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_wrapper_compiles_and_links(tmp_path):
    src = tmp_path / "t.cpp"
    src.write_text('#include "zkvm_b200/cpp/ristretto_msm.hpp"\n'
                   "int main(int argc, char**) { if (argc > 100) { zkvm_b200::Context c(0); zkvm_b200::PointTable t(c);"
                   " std::vector<zkvm_b200::Scalar> s; zkvm_b200::RistrettoPoint::vartime_multiscalar_mul(c, s, t);"
                   " zkvm_b200::RistrettoPoint::optional_multiscalar_mul(c, s, {}); } return 0; }\n")
    exe = tmp_path / "t"
    subprocess.check_call(["g++", "-std=c++17", "-I", ROOT, str(src), "-o", str(exe), "-L", os.path.join(ROOT, "zkvm_b200"),
                           "-l:libzkmsm.so", f"-Wl,-rpath,{os.path.join(ROOT, 'zkvm_b200')}"])
    subprocess.check_call([str(exe)])
