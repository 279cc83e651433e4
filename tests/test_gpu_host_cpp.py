"""GPU: the C++ host-side mirror of the reference's compute package (host/compute.hpp) against the
hand-derived known answers, through the C ABI only."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_host_mirror_selftest():
    exe = os.path.join(ROOT, "go-vectorsearch_b200", "build", "host_selftest")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "go-vectorsearch_b200", "host")])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "selftest: ok" in r.stdout


def test_cpp_one_process_drives_several_devices():
    """host/selftest_sharded.cpp: one host process, vs_sharded_* over two stripes (two GPUs when the box has them)."""
    exe = os.path.join(ROOT, "go-vectorsearch_b200", "build", "host_selftest_sharded")
    assert os.path.exists(exe), "run __graft_entry__.build() first"
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "sharded selftest: ok" in r.stdout
