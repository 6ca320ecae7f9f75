"""The compiled-language host side (include/raiko_kzg.hpp) replaying the reference's own unit
tests (eip4844.rs:147-214) against libraiko_kzg.so."""
import os
import subprocess

import pytest

from kzg_testlib import ROOT, SETUP

pytestmark = pytest.mark.gpu


def test_cpp_host_mirror_replays_reference_tests(golden, tmp_path):
    exe = str(tmp_path / "test_eip4844_cpp")
    libdir = os.path.join(ROOT, "raiko_b200")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "native", "test_eip4844_cpp.cpp"),
                           "-L" + libdir, "-lraiko_kzg", "-Wl,-rpath," + libdir])
    out = subprocess.run([exe, SETUP], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "CPP_HOST_OK" in out.stdout, out.stdout + out.stderr
    mod64 = [c for c in golden["cases"] if c["name"] == "C3_mod64"][0]
    assert out.stdout.split()[-1] == mod64["commitment"]
