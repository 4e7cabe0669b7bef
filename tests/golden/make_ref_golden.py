#!/usr/bin/env python
"""Write tests/golden/ref_golden.json: outputs of the reference's OWN hot-path definitions — its source text compiled here by
`make -C oracle refsrc` into oracle/_ref/librefsrc.so (real Bullet, a stand-in for the GLM primitives; see
oracle/ref_harness.cpp) — on the deterministic cases of tests/golden/ref_cases.py. Every number is stored as a C99 hex float, so
tests/test_oracle_refsource.py can hold the oracle to it bit for bit on machines where the reference is not mounted.
Run where /root/reference exists:  python tests/golden/make_ref_golden.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyrefsrc as rs  # noqa: E402
from tests.golden import ref_cases as rc  # noqa: E402


def hx(v):
    return [float(x).hex() for x in np.asarray(v, dtype=np.float64).reshape(-1)]


def info_rec(r):
    return {"vec": {k: hx(r[k]) for k in rc.VEC}, "Swet": float(r["Swet"]).hex(), "Vsub": float(r["Vsub"]).hex(),
            "position": int(r["position"]), "wrote": [int(r["wrote_buoyancy"]), int(r["wrote_damping"])]}


out = {"doc": __doc__, "bodies": [], "depth": [], "velocity": [], "wind": [], "aero": [], "compound": [], "compound_aero": None, "cables": []}
for k in range(len(rc.BODY_SCENES)):
    sc = rc.body_scene(k)
    fl = rs.ocean_fluid(sc.ocean)
    out["bodies"].append([info_rec(rs.body(sc.meshes[int(sc.body_mesh[i])], fl, sc.states[i:i + 1])) for i in range(sc.n_bodies)])

oc = rc.field_ocean(True)
fl = rs.ocean_fluid(oc)
for p in rc.query_points():
    out["depth"].append(float(rs.depth(fl, p)).hex())
    out["velocity"].append(hx(rs.velocity_f(fl, p)))
afl = rs.fluid_struct(rc.wind("jetpipe").wind, 1.2)
for p in rc.query_points():
    out["wind"].append(hx(rs.wind_f(afl, p)))

meshes_ = rc.AERO_MESHES()
for kind, dispatch, mi, st, typ, params in rc.aero_cases():
    atm = rc.wind(kind)
    F, T = rs.aero(meshes_[mi], rs.fluid_struct(atm.wind, atm.density), st, dispatch, typ, params)
    out["aero"].append({"Fda": hx(F), "Tda": hx(T)})

for depth, waves, flags in rc.COMPOUND_CASES:
    ocean, ms, part_mesh, parts, external, origins, o2cs, comp, T_O = rc.compound_case(depth, waves, flags)
    out["compound"].append(info_rec(rs.compound(ms, part_mesh, parts, external, origins, o2cs, rs.ocean_fluid(ocean), comp, T_O)))
atm, ms, part_mesh, pa, external, origins, o2cs, comp, T_O, kinds, params = rc.compound_aero_case()
r = rs.compound(ms, part_mesh, pa, external, origins, o2cs, rs.fluid_struct(atm.wind, atm.density), comp, T_O, aero_types=kinds, aero_params=params)
out["compound_aero"] = {"Fda": hx(r["Fdq"]), "Tda": hx(r["Tdq"])}

for ocean, radius, flags, pos, vel in rc.cable_cases():
    out["cables"].append(hx(rs.cable(rs.ocean_fluid(ocean), radius, flags, pos, vel)))

# the wave producer's host half: 2 x 1 MB of spectrum and the butterfly table, kept as SHA-256 of their bytes
import hashlib  # noqa: E402

from stonefish_b200.hydro import wave_params  # noqa: E402

s12, s34, bf = rs.wave_tables(wave_params(1.0))
out["wave_tables"] = {"sea_state": 1.0, "seed": 1234, "spectrum12": hashlib.sha256(s12.tobytes()).hexdigest(),
                      "spectrum34": hashlib.sha256(s34.tobytes()).hexdigest(), "butterfly": hashlib.sha256(bf.tobytes()).hexdigest(),
                      "spectrum12_first_nonzero": [float(x).hex() for x in s12[np.nonzero(s12)[0][:8]]]}

# the reference's own meshes (Tests/Data/*.obj): positions and faces as a fixture, the reference text's outputs on them
from stonefish_b200 import meshes as M  # noqa: E402

real = {n: M.load_obj(f"/root/reference/Tests/Data/{n}.obj") for n in rc.REAL_MESH_NAMES}
np.savez_compressed(rc.REAL_MESH_FIXTURE, **{f"{n}/{k}": (m.positions if k == "v" else m.faces) for n, m in real.items() for k in ("v", "f")})
real = rc.load_real_meshes()  # what the tests will see
out["real_meshes"] = []
for sc in rc.real_mesh_scenes(real):
    fl = rs.ocean_fluid(sc.ocean)
    out["real_meshes"].append({"scene": sc.name, "bodies": [info_rec(rs.body(sc.meshes[int(sc.body_mesh[i])], fl, sc.states[i:i + 1]))
                                                             for i in range(sc.n_bodies)]})

path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_golden.json")
json.dump(out, open(path, "w"), indent=0)
print("wrote", path, {k: (len(v) if isinstance(v, list) else 1) for k, v in out.items() if k != "doc"})
print("wrote", rc.REAL_MESH_FIXTURE, os.path.getsize(rc.REAL_MESH_FIXTURE), "bytes")
