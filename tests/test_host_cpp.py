"""The C++ host mirror of the crate's interface (host/cpp/twenty_first_b200.hpp): the reference is compiled
code, so next to the Python mirror there is a compiled one above the same C ABI.  Its parity program
(host/cpp/test_host.cpp) reads like the reference's tests and uses the CPU oracle as the checker."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "host", "cpp")


def _binary():
    exe = os.path.join(CPP, "test_host")
    import oracle

    oracle.build()
    subprocess.run(["make", "-C", CPP], check=True, capture_output=True)
    return exe


def test_cpp_host_mirror_links_and_maps_panics_without_a_gpu():
    """links against libtf21.so; ntt on 12 elements panics with the reference's message before any CUDA call"""
    r = subprocess.run([_binary(), "--link-only"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "power of two" in r.stdout


@pytest.mark.gpu
def test_cpp_host_mirror_parity_program():
    r = subprocess.run([_binary()], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "10 tests, 0 failed" in r.stdout, r.stdout
