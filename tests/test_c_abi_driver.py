"""The C ABI without Python in the loop: tests/c_abi_driver.c is compiled with gcc against include/blangpt.h and
libblangpt.so and run as its own process (the calls a JNI / Panama shim makes, java/blang/engines/gpu/GpuPT.java)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_DIR = os.path.join(ROOT, "blangsdk_b200")


def _build(tmp_path):
    exe = str(tmp_path / "c_abi_driver")
    subprocess.check_call(["gcc", "-std=c11", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_abi_driver.c"), "-o", exe, "-L", LIB_DIR, "-lblangpt",
                           "-Wl,-rpath," + LIB_DIR, "-lm"])
    return exe


def test_c_driver_compiles_and_fails_loudly_without_a_gpu(tmp_path):
    """header and library are usable from plain C; with no CUDA device create() refuses (no CPU fallback)"""
    if not os.path.exists(os.path.join(LIB_DIR, "libblangpt.so")):
        pytest.skip("libblangpt.so not built")
    exe = _build(tmp_path)
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present: covered by the gpu test below")
    except ImportError:
        pass
    p = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert p.returncode == 3, (p.returncode, p.stdout, p.stderr)
    assert "no CPU fallback" in p.stderr


@pytest.mark.gpu
def test_c_driver_runs_on_the_gpu(tmp_path):
    exe = _build(tmp_path)
    p = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, (p.stdout, p.stderr)
    assert "C ABI driver: OK" in p.stdout
