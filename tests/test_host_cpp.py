"""The C++ host mirror (solverforge_b200/host/solverforge_gpu.hpp): compiled against include/sfgpu.h
and linked with libsfgpu.so; replay logic checked on CPU, full parity on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "_build", "host_test")


def _build():
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    src = os.path.join(ROOT, "tests", "cpp", "host_test.cpp")
    deps = [src, os.path.join(ROOT, "solverforge_b200", "host", "solverforge_gpu.hpp"),
            os.path.join(ROOT, "include", "sfgpu.h")]
    if os.path.exists(BIN) and all(os.path.getmtime(BIN) >= os.path.getmtime(d) for d in deps):
        return
    lib_dir = os.path.join(ROOT, "solverforge_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Wno-unused-function", src, "-o", BIN,
                           f"-L{lib_dir}", "-lsfgpu", f"-Wl,-rpath,{lib_dir}"])


def test_host_replay_matches_oracle():
    _build()
    out = subprocess.run([BIN, "replay"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr


@pytest.mark.gpu
def test_host_cpp_gpu_parity():
    _build()
    out = subprocess.run([BIN, "gpu"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    # every section ran: graph colouring + tabu, projected roster, ten CVRP local-search phases
    assert int(out.stdout.split("HOST TEST OK")[1]) > 5000, out.stdout
