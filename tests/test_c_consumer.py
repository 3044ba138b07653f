"""The C ABI from plain C (tests/c_consumer/consumer.c compiled with gcc against include/reach.h and linked to
libreach_cuda.so): no ctypes, no Python between the caller and the library."""
import os
import subprocess

import pytest

from reach_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.dirname(_lib.LIB_PATH)


def build(tmp_path):
    exe = str(tmp_path / "consumer")
    cmd = ["gcc", "-std=c99", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c_consumer", "consumer.c"), "-L", PKG, "-lreach_cuda",
           "-Wl,-rpath," + PKG, "-lm", "-o", exe]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


def test_header_compiles_as_c99_and_library_links(tmp_path):
    exe = build(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "reach_version" in out.stdout


@pytest.mark.gpu
def test_c_consumer_runs_lgg09_glgm06_and_multi(tmp_path):
    exe = build(tmp_path)
    out = subprocess.run([exe, "--run"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.strip().endswith("OK")
