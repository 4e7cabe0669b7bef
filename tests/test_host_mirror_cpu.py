"""Host-only half of the C++ mirror (stonefish_b200/host/StonefishHydro.hpp): the Bullet entry points instantiated with the
reference's own vendored Bullet headers, field constructors against Bullet's arithmetic, the fill/sum helpers. No GPU: the
program never creates a HydroEngine (it links libsfh.so only because the header's inline methods reference its symbols)."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference/3rdparty"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "LinearMath")), reason="reference checkout not mounted")
def test_cpp_mirror_bullet_entry_points_and_host_helpers():
    lib = os.path.join(ROOT, "stonefish_b200", "libsfh.so")
    assert os.path.exists(lib), "build first: make -C stonefish_b200/csrc"
    os.makedirs(os.path.join(HERE, "_build"), exist_ok=True)
    exe = os.path.join(HERE, "_build", "host_frombullet")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-DBT_USE_DOUBLE_PRECISION", "-I" + REF, "-I" + os.path.join(REF, "LinearMath"),
                           "-o", exe, os.path.join(HERE, "host_frombullet.cpp"), "-L" + os.path.dirname(lib), "-lsfh",
                           "-Wl,-rpath," + os.path.dirname(lib)])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stdout + out.stderr
