"""powf restatement (580-raytracer_b200/csrc/powf_glibc.cuh) against the host libm on the CPU:
the same header the device compiles, built with g++ (tests/native/powf_host_check.cpp)."""
import os
import subprocess

from conftest import ROOT


def test_powf_restatement_matches_libm(tmp_path):
    exe = str(tmp_path / "powf_check")
    subprocess.check_call(["g++", "-O2", "-mfma", "-o", exe, os.path.join(ROOT, "tests", "native", "powf_host_check.cpp"), "-lm"])
    for seed in (1, 580):
        n, bad = subprocess.check_output([exe, "4000000", str(seed)], text=True).split()
        assert int(n) == 4000000 and int(bad) == 0
