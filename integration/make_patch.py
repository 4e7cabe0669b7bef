#!/usr/bin/env python
"""Build integration/stonefish_v1.6.0.patch: a unified diff (apply with `patch -p1` or `git apply` in the root of a Stonefish
v1.6.0 checkout) that routes the recomputation of the hydrodynamic forces of all solids through libsfh.so.

The patch = the two new files under integration/files/ + the anchored edits listed below of nine reference files. It is generated
— run where the reference checkout is mounted (/root/reference, read-only: everything happens in a temporary copy) — and
committed, so that the repository carries the artefact and tests/test_integration_patch.py can check that it still applies.
    python integration/make_patch.py [--check]      (--check: fail if the committed patch differs from a fresh one)
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("STONEFISH_REF", "/root/reference")
PATCH = os.path.join(HERE, "stonefish_v1.6.0.patch")

# (file, anchor text that occurs exactly once, "before" | "after", text to insert)
EDITS = [
    ("CMakeLists.txt", 'option(EMBED_RESOURCES "Embed internal resources in the library executable" OFF)\n', "after",
     'option(WITH_B200_HYDRO "Evaluate the geometry-based hydrodynamics on a CUDA device (stonefish_b200: libsfh.so)" OFF)\n'
     'set(SFH_ROOT "" CACHE PATH "Root of the stonefish_b200 repository (include/sfh.h, stonefish_b200/host, stonefish_b200/libsfh.so)")\n'),
    ("CMakeLists.txt", "set(BULLET_FLAGS BT_EULER_DEFAULT_ZYX BT_USE_DOUBLE_PRECISION)\n", "after",
     "if(WITH_B200_HYDRO)\n"
     "    # HydroGPU (Library/src/core/HydroGPU.cpp) talks to the CUDA library through its C ABI and the header-only C++ mirror\n"
     "    include_directories(${SFH_ROOT}/include ${SFH_ROOT}/stonefish_b200/host)\n"
     "    list(APPEND LIBRARIES ${SFH_ROOT}/stonefish_b200/libsfh.so)\n"
     "    list(APPEND BULLET_FLAGS STONEFISH_B200_HYDRO)\n"
     "endif()\n"),
    ("Library/include/core/SimulationManager.h", "    class NameManager;\n", "before", "    class HydroGPU;\n"),
    ("Library/include/core/SimulationManager.h", "        Atmosphere* atmosphere;\n", "after",
     "        HydroGPU* hydroGPU; // Batched hydrodynamics on a CUDA device (nullptr: the thread pool computes them)\n"),
    ("Library/src/core/SimulationManager.cpp", '#include "core/NameManager.h"\n', "before", '#include "core/HydroGPU.h"\n'),
    ("Library/src/core/SimulationManager.cpp", "    atmosphere = nullptr;\n    trackball = nullptr;\n", "after",
     "    hydroGPU = nullptr;\n"
     "#ifdef STONEFISH_B200_HYDRO\n"
     "    try { hydroGPU = new HydroGPU(0); }\n"
     "    catch(const std::exception& e) { cWarning(\"Hydrodynamics stay on the CPU: %s\", e.what()); }\n"
     "#endif\n"),
    ("Library/src/core/SimulationManager.cpp", "    DestroyScenario();\n    if(atmosphere != nullptr) delete atmosphere;\n", "after",
     "#ifdef STONEFISH_B200_HYDRO\n"
     "    if(hydroGPU != nullptr) delete hydroGPU;\n"
     "#endif\n"),
    ("Library/src/core/SimulationManager.cpp", "void SimulationManager::DestroyScenario()\n{\n", "after",
     "#ifdef STONEFISH_B200_HYDRO\n"
     "    if(hydroGPU != nullptr) hydroGPU->Clear(); //Physics meshes die with the entities\n"
     "#endif\n"),
    ("Library/src/core/SimulationManager.cpp",
     "        btBroadphasePairArray& pairArray = simManager->ocean->getGhost()->getOverlappingPairCache()->getOverlappingPairArray();\n"
     "        int numPairs = pairArray.size();\n", "after",
     "        bool solidsOnGPU = false;\n"
     "#ifdef STONEFISH_B200_HYDRO\n"
     "        if(recompute && numPairs > 0 && simManager->hydroGPU != nullptr)\n"
     "        {\n"
     "            //All solids that overlap the ocean in ONE step on the device; below their forces are only applied\n"
     "            std::vector<SolidEntity*> solids;\n"
     "            for(int h=0; h<numPairs; ++h)\n"
     "            {\n"
     "                const btBroadphasePair& pair = pairArray[h];\n"
     "                btBroadphasePair* colPair = world->getPairCache()->findPair(pair.m_pProxy0, pair.m_pProxy1);\n"
     "                if (!colPair)\n"
     "                    continue;\n"
     "                btCollisionObject* candidate1 = (btCollisionObject*)colPair->m_pProxy0->m_clientObject;\n"
     "                btCollisionObject* candidate2 = (btCollisionObject*)colPair->m_pProxy1->m_clientObject;\n"
     "                SolidEntity* solid = HydroGPU::SolidFromCollisionObject(candidate1 == simManager->ocean->getGhost() ? candidate2 : candidate1);\n"
     "                if(solid != nullptr) solids.push_back(solid);\n"
     "            }\n"
     "            simManager->hydroGPU->ComputeHydrodynamicForces(simManager->ocean, solids);\n"
     "            solidsOnGPU = true;\n"
     "        }\n"
     "#endif\n"),
    # the two calls of the ocean loop (the atmosphere loop above them has the same text with `atmosphere`)
    ("Library/src/core/SimulationManager.cpp",
     "                btCollisionObject* co = candidate1 == simManager->ocean->getGhost() ? candidate2 : candidate1;\n", "after",
     "                //Solids computed on the device only need ApplyHydrodynamicForces; cables are still recomputed here\n"
     "                bool recomputeHere = recompute;\n"
     "#ifdef STONEFISH_B200_HYDRO\n"
     "                if(solidsOnGPU && HydroGPU::SolidFromCollisionObject(co) != nullptr) recomputeHere = false;\n"
     "#endif\n"),
    ("Library/include/entities/SolidEntity.h", "        friend class BeltJoint;\n", "after", "        friend class HydroGPU;\n"),
    ("Library/include/entities/solids/Compound.h", "    private:\n        std::vector<CompoundPart> parts; //Parts of the compound solid\n", "before",
     "        friend class HydroGPU;\n"),
    ("Library/include/entities/forcefields/Ocean.h", "    private:\n        Fluid liquid;\n", "before", "        friend class HydroGPU;\n"),
    ("Library/include/entities/forcefields/Jet.h", "    private:\n        Vector3 c, n;\n", "before", "        friend class HydroGPU;\n"),
    ("Library/include/entities/forcefields/Pipe.h", "    private:\n        Vector3 p1, n;\n", "before", "        friend class HydroGPU;\n"),
    ("Library/include/graphics/OpenGLOcean.h", "        //! A method calculating Henyey-Greenstein scattering factor.\n        glm::vec3 getLightScattering();\n", "after",
     "\n        //! A method returning the parameters of the wave simulation (FFT size, grid sizes).\n"
     "        const OceanParams& getOceanParams() const { return params; }\n"),
    ("Library/include/graphics/OpenGLRealOcean.h",
     "        //! A method that generates the stencil mask.\n        /*!\n         \\param view a pointer to the active view\n         */\n"
     "        void DrawUnderwaterMask(OpenGLView* view) override;\n", "after",
     "\n        //! A method returning the wave data last read back from the GPU (4 layers of fftSize x fftSize RGBA32F texels).\n"
     "        const GLfloat* getFFTData() const { return fftData; }\n"),
]
# plain replacements (text that occurs exactly once after the insertions above)
REPLACE = [
    ("Library/src/core/SimulationManager.cpp",
     "                        sim->ocean->ApplyFluidForces(world, co, recompute); \n                    }, simManager, world, co, recompute);\n"
     "                else\n                    simManager->ocean->ApplyFluidForces(world, co, recompute);\n",
     "                        sim->ocean->ApplyFluidForces(world, co, recompute); \n                    }, simManager, world, co, recomputeHere);\n"
     "                else\n                    simManager->ocean->ApplyFluidForces(world, co, recomputeHere);\n"),
]
NEW_FILES = ["Library/include/core/HydroGPU.h", "Library/src/core/HydroGPU.cpp"]


def build(ref=REF):
    tmp = tempfile.mkdtemp(prefix="sfh_patch_")
    try:
        touched = sorted({e[0] for e in EDITS} | {r[0] for r in REPLACE})
        for side in ("a", "b"):
            for f in touched:
                dst = os.path.join(tmp, side, f)
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                shutil.copyfile(os.path.join(ref, f), dst)
        for f, anchor, where, text in EDITS:
            path = os.path.join(tmp, "b", f)
            s = open(path, newline="").read()
            crlf = "\r\n" in s
            if crlf:  # some reference files have DOS line ends: edit and write them that way
                anchor, text = anchor.replace("\n", "\r\n"), text.replace("\n", "\r\n")
            assert s.count(anchor) == 1, f"{f}: anchor occurs {s.count(anchor)} times: {anchor!r}"
            s = s.replace(anchor, anchor + text if where == "after" else text + anchor)
            open(path, "w", newline="").write(s)
        for f, old, new in REPLACE:
            path = os.path.join(tmp, "b", f)
            s = open(path, newline="").read()
            if "\r\n" in s:
                old, new = old.replace("\n", "\r\n"), new.replace("\n", "\r\n")
            assert s.count(old) == 1, f"{f}: text occurs {s.count(old)} times: {old!r}"
            open(path, "w", newline="").write(s.replace(old, new))
        for f in NEW_FILES:
            dst = os.path.join(tmp, "b", f)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(os.path.join(HERE, "files", f), dst)
        r = subprocess.run(["diff", "-ruN", "a", "b"], cwd=tmp, capture_output=True)  # bytes: two of the files have DOS line ends
        assert r.returncode in (0, 1), r.stderr
        out = []
        for ln in r.stdout.splitlines(keepends=True):
            # drop diff's timestamps: the patch should not change when it is regenerated
            if ln.startswith((b"--- a/", b"+++ b/", b"--- b/", b"+++ a/")):
                ln = ln.split(b"\t")[0] + b"\n"
            out.append(ln)
        return b"".join(out)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    text = build()
    if "--check" in sys.argv:
        sys.exit(0 if open(PATCH, "rb").read() == text else "integration/stonefish_v1.6.0.patch is stale: run python integration/make_patch.py")
    open(PATCH, "wb").write(text)
    print("wrote", PATCH, len(text.splitlines()), "lines")
