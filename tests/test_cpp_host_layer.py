"""include/acm.hpp (the C++ host mirror of the reference's trait surface): compiles against the
header and links libacm.so on CPU; on the GPU box the program runs the reference's unit tests
restated in tests/cpp/host_test.cpp."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "apex_camera_models_b200", "lib")


def _build(tmp_path, name="host_test"):
    exe = str(tmp_path / name)
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-pthread", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", name + ".cpp"), "-o", exe, "-L", LIBDIR, "-lacm", f"-Wl,-rpath,{LIBDIR}"]
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


def test_single_process_multi_gpu_test_compiles_and_fails_loudly_without_gpu(tmp_path):
    import torch
    exe = _build(tmp_path, "multi_gpu_test")
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by the gpu-marked test")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode != 0 and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_single_process_multi_gpu_and_scalar_host_path(tmp_path):
    """acm_comm_init_all + acm_*_multi from one host thread (2 contexts in one process) against the same pipeline
    on one GPU, plus the scalar acm_project_host path.  The multi-GPU half reports MULTI_GPU_SKIPPED on a 1-GPU box."""
    exe = _build(tmp_path, "multi_gpu_test")
    ndev = os.environ.get("ACM_MULTI_NDEV", "2")   # 2 contexts by default; ACM_MULTI_NDEV=8 drives a whole box from one thread
    r = subprocess.run([exe, ndev], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "MULTI_TEST_OK" in r.stdout
    print(r.stdout)
