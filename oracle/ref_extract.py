#!/usr/bin/env python
"""Pull the DEFINITIONS of the hot-path functions out of the reference's source files, where they lie, into ONE generated
include file under oracle/_ref/ (git-ignored build output, deleted again by the Makefile once compiled). Nothing of the
reference is stored in the repository: this script only knows file names and function signatures.

usage: python ref_extract.py <reference root> <output .inc>

A definition runs from the line that starts with the signature to the next line that is exactly `}` in column 0 (the
reference's style). TEST INFRASTRUCTURE ONLY."""
import os
import sys

WANTED = [  # (file under Library/src — or Library/include when it starts with "include/" —, signature prefixes at column 0)
    ("include/utils/SystemUtil.hpp", ["inline long lrandom(", "inline float frandom("]),
    ("graphics/OpenGLOcean.cpp", ["int OpenGLOcean::bitReverse(", "void OpenGLOcean::computeWeight(", "GLfloat* OpenGLOcean::ComputeButterflyLookupTable(",
                                  "float OpenGLOcean::sqr(", "float OpenGLOcean::omega(", "float OpenGLOcean::spectrum(",
                                  "void OpenGLOcean::GetSpectrumSample(", "void OpenGLOcean::GenerateWavesSpectrum("]),
    ("graphics/OpenGLDataStructs.cpp", ["glm::mat4 glMatrixFromTransform(", "glm::vec3 glVectorFromVector("]),
    ("entities/forcefields/Uniform.cpp", ["Uniform::Uniform(", "void Uniform::setVelocity(", "Vector3 Uniform::GetVelocityAtPoint("]),
    ("entities/forcefields/Jet.cpp", ["Jet::Jet(", "void Jet::setOutletVelocity(", "Vector3 Jet::GetVelocityAtPoint("]),
    ("entities/forcefields/Pipe.cpp", ["Pipe::Pipe(", "void Pipe::setInletVelocity(", "Vector3 Pipe::GetVelocityAtPoint("]),
    ("graphics/OpenGLRealOcean.cpp", ["GLfloat OpenGLRealOcean::ComputeInterpolatedWaveData(", "GLfloat OpenGLRealOcean::ComputeWaveHeight("]),
    ("entities/forcefields/Ocean.cpp", ["float Ocean::GetDepth(const glm::vec3&", "Scalar Ocean::GetDepth(const Vector3&",
                                        "Vector3 Ocean::GetFluidVelocity(const Vector3&", "glm::vec3 Ocean::GetFluidVelocity(const glm::vec3&"]),
    ("entities/forcefields/Atmosphere.cpp", ["Vector3 Atmosphere::GetFluidVelocity(const Vector3&", "glm::vec3 Atmosphere::GetFluidVelocity(const glm::vec3&"]),
    ("entities/SolidEntity.cpp", ["BodyFluidPosition SolidEntity::CheckBodyFluidPosition(", "void SolidEntity::CorrectHydrodynamicForces(",
                                  "void SolidEntity::ComputeHydrodynamicForcesSurface(", "void SolidEntity::ComputeHydrodynamicForcesSubmerged(",
                                  "void SolidEntity::ComputeHydrodynamicForces(HydrodynamicsSettings", "void SolidEntity::ComputeAerodynamicForces(Atmosphere*",
                                  "void SolidEntity::ComputeAerodynamicForces(const Mesh*", "void SolidEntity::CorrectAerodynamicForces("]),
    ("entities/solids/Compound.cpp", ["void Compound::ComputeHydrodynamicForces(", "void Compound::ComputeAerodynamicForces("]),
    ("entities/CableEntity.cpp", ["Scalar CableEntity::circularSegmentArea(", "void CableEntity::ComputeHydrodynamicForces("]),
]


def main():
    root, out = sys.argv[1], sys.argv[2]
    chunks, missing = [], []
    for rel, sigs in WANTED:
        path = os.path.join(root, "Library", rel) if rel.startswith("include/") else os.path.join(root, "Library", "src", rel)
        lines = open(path, encoding="utf-8", errors="replace").read().splitlines()
        for sig in sigs:
            start = next((i for i, l in enumerate(lines) if l.startswith(sig)), None)
            if start is None:
                missing.append((rel, sig))
                continue
            end = next(i for i in range(start + 1, len(lines)) if lines[i] == "}")
            chunks.append(f"// ---- {rel}:{start + 1}-{end + 1}\n#line {start + 1} \"{path}\"\n" + "\n".join(lines[start:end + 1]) + "\n")
    if missing:
        sys.exit(f"ref_extract: definitions not found: {missing}")
    with open(out, "w") as f:
        f.write("// GENERATED from the reference sources by oracle/ref_extract.py — build intermediate, never committed\n")
        f.write("\n".join(chunks))
    print(f"ref_extract: {len(chunks)} definitions -> {out}")


if __name__ == "__main__":
    main()
