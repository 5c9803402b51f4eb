"""The XU-free conversion helpers of csrc/pcf_device.cuh (f2d_exact, narrow_f32, voxel_axis, transform_point) are
__host__ __device__: compile them for the host and bit-compare with the plain C conversions / floor(a / res)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not on PATH")
def test_conversion_helpers_match_plain_c():
    out = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    exe = os.path.join(out, "host_kat")
    subprocess.run(["nvcc", "-std=c++17", "-O2", "-Xcompiler", "-ffp-contract=off", "--expt-relaxed-constexpr",
                    os.path.join(ROOT, "tests", "host_kat.cu"), "-o", exe], check=True, capture_output=True)
    r = subprocess.run([exe, "4000000"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
