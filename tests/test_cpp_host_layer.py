"""include/acm.hpp (the C++ host mirror of the reference's trait surface): compiles against the
header and links libacm.so on CPU; on the GPU box the program runs the reference's unit tests
restated in tests/cpp/host_test.cpp."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "apex_camera_models_b200", "lib")


def _build(tmp_path):
    exe = str(tmp_path / "host_test")
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "host_test.cpp"), "-o", exe, "-L", LIBDIR, "-lacm", f"-Wl,-rpath,{LIBDIR}"]
    subprocess.run(cmd, check=True)
    return exe


def test_cpp_host_layer_compiles_and_fails_loudly_without_gpu(tmp_path):
    import torch
    exe = _build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by the gpu-marked test")
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True)
    assert r.returncode != 0 and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_cpp_host_layer_runs_reference_unit_tests(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "HOST_TEST_OK" in r.stdout
