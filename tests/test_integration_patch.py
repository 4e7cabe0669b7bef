"""SURVEY.md §8 row N1 as an artefact: integration/stonefish_v1.6.0.patch is a unified diff against the Stonefish v1.6.0 tree (two new
files — sf::HydroGPU — and anchored edits of nine reference files) that routes the recomputation of all solids' hydrodynamic forces
through libsfh.so. Where the reference checkout is mounted: the committed patch equals a freshly generated one, applies cleanly to the
reference's files, and the file it adds compiles against the reference's own definitions (oracle/adapter_harness.cpp)."""
import os
import shutil
import subprocess
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
PATCH = os.path.join(ROOT, "integration", "stonefish_v1.6.0.patch")
HAVE_REF = os.path.isdir(os.path.join(REF, "Library", "src"))


def test_patch_is_committed_and_touches_the_call_sites():
    text = open(PATCH, newline="").read()
    for f in ("Library/src/core/SimulationManager.cpp", "Library/src/core/HydroGPU.cpp", "Library/include/core/HydroGPU.h",
              "Library/include/entities/SolidEntity.h", "Library/include/entities/solids/Compound.h", "Library/include/graphics/OpenGLRealOcean.h",
              "CMakeLists.txt"):
        assert f"+++ b/{f}" in text, f
    assert "simManager->hydroGPU->ComputeHydrodynamicForces(simManager->ocean, solids);" in text
    # the new file in the patch IS the file the adapter harness compiles
    src = open(os.path.join(ROOT, "integration", "files", "Library", "src", "core", "HydroGPU.cpp")).read()
    added = "".join(ln[1:] for ln in text.split("+++ b/Library/src/core/HydroGPU.cpp")[1].split("\ndiff -ruN")[0].splitlines(keepends=True)[2:]
                    if ln.startswith("+"))
    assert added.replace("\r\n", "\n").strip() == src.strip()


@pytest.mark.skipif(not HAVE_REF, reason="reference checkout not mounted")
def test_patch_is_current_and_applies_to_the_reference_tree():
    subprocess.check_call([sys.executable, os.path.join(ROOT, "integration", "make_patch.py"), "--check"])
    tmp = tempfile.mkdtemp(prefix="sfh_apply_")
    try:
        touched = sorted({ln.split("\t")[0][6:].strip() for ln in open(PATCH, newline="").read().splitlines() if ln.startswith("--- a/")})
        for f in touched:
            if os.path.exists(os.path.join(REF, f)):
                os.makedirs(os.path.dirname(os.path.join(tmp, f)), exist_ok=True)
                shutil.copyfile(os.path.join(REF, f), os.path.join(tmp, f))
        r = subprocess.run(["patch", "-p1", "--dry-run", "--binary", "-i", PATCH], cwd=tmp, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        r = subprocess.run(["patch", "-p1", "--binary", "-i", PATCH], cwd=tmp, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        sm = open(os.path.join(tmp, "Library", "src", "core", "SimulationManager.cpp"), newline="").read()
        assert sm.count("recomputeHere") == 4 and sm.count("sim->atmosphere->ApplyFluidForces(world, co, recompute);") == 1  # the atmosphere loop is untouched
        assert os.path.exists(os.path.join(tmp, "Library", "src", "core", "HydroGPU.cpp"))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


@pytest.mark.skipif(not HAVE_REF, reason="reference checkout not mounted")
def test_adapter_compiles_next_to_the_reference_text():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "stonefish_b200", "csrc")], stdout=subprocess.DEVNULL)
    out = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "adapter"], capture_output=True, text=True)
    assert out.returncode == 0 and "built oracle/_ref/libadapter.so" in out.stdout, out.stdout + out.stderr
    syms = subprocess.run(["nm", "-D", "--defined-only", os.path.join(ROOT, "oracle", "_ref", "libadapter.so")], capture_output=True, text=True).stdout
    for s in ("adp_new", "adp_add_solid", "adp_add_compound", "adp_set_state", "adp_step", "_ZN2sf8HydroGPU25ComputeHydrodynamicForcesEPNS_5OceanERKSt6vectorIPNS_11SolidEntityESaIS5_EE"):
        assert s in syms, s
