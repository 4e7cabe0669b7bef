"""SURVEY.md §8 row N1, end to end: the FluidDynamicsTest scene (FluidDynamicsTestManager.cpp:93-161, its real OBJ hulls, compounds as
sf::Compound objects) stepped 200 times through the PATCHED call site — sf::HydroGPU of integration/stonefish_v1.6.0.patch, compiled
unchanged in oracle/adapter_harness.cpp, libsfh.so underneath — and, from the same states, through the reference's own
ComputeHydrodynamicForces text. The bodies are integrated with the GPU arm's wrench (a plain semi-implicit Euler: Bullet is out of
scope), so they fall, cross the surface and sink while the two arms are compared at every step."""
import numpy as np
import pytest

from stonefish_b200 import meshes as M
from stonefish_b200 import scenes
from stonefish_b200._abi import SFH_F_BUOYANT, SFH_POS_AUTO
from stonefish_b200.hydro import Ocean, Uniform, make_transform, new_body_states, quat_to_basis, transform_mul
from tests import adapter
from tests.golden import ref_cases as rc
from tests.parity import VEC_FIELDS
from tests.test_trajectory_gpu import quat_mul

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not adapter.available(), reason="oracle/_ref/libadapter.so not built (needs the reference checkout)")]
SENTINEL = 123.0


def build(real):
    ocean = Ocean(1000.0, 9.81)
    ocean.AddVelocityField(Uniform((0.2, -0.1, 0.05)))
    ocean.EnableCurrents()
    meshes = [M.icosphere(0.5, 3), M.box(1.0, 0.5, 0.2, 3), real["hull_hydro"], real["sphere_R=1"]]
    sc = adapter.AdapterScene(ocean)
    ents = []
    for kind, parts, q, o, v, w in scenes.config1_layout():
        e = {"kind": kind, "parts": parts, "q": np.array(q, dtype=float), "pos": np.array(o, dtype=float), "v": np.array(v, dtype=float),
             "w": np.array(w, dtype=float)}
        st = new_body_states(1)
        st["mode"], st["flags"] = SFH_POS_AUTO, SFH_F_BUOYANT
        st["Cd"][0], st["Cf"][0] = (1.05, 0.9, 1.2), (0.05, 0.08, 0.1)
        st["P_CB"][0] = (0.01, -0.02, 0.03)
        if kind == "solid":
            m = meshes[parts[0][0]]
            st["volume"][0], st["surface"][0] = M.signed_volume(m), M.surface_area(m)
            e["mass"] = 1000.0 * st["volume"][0] * (0.7 if len(ents) % 2 else 1.3)   # some float, some sink
            e["state"] = st
            fill(e, meshes)
            e["id"] = sc.add_solid(m, e["state"])
        else:
            pm = [meshes[mi] for mi, _ in parts]
            st["volume"][0], st["surface"][0] = sum(M.signed_volume(m) for m in pm), sum(M.surface_area(m) for m in pm)
            e["mass"] = 1000.0 * st["volume"][0] * 1.1
            e["state"] = st
            ps = new_body_states(len(parts))
            for i in range(len(parts)):
                ps["Cd"][i], ps["Cf"][i], ps["flags"][i] = (1.0 + 0.1 * i, 0.9, 1.1), (0.05, 0.06, 0.07), SFH_F_BUOYANT
            fill(e, meshes)
            origins = np.array([make_transform(None, po) for _, po in parts])
            o2c = np.array([make_transform() for _ in parts])
            e["id"] = sc.add_compound(pm, ps, [1] * len(parts), origins, o2c, e["state"][0], e["T"])
        ents.append(e)
    return ocean, sc, ents, meshes


def fill(e, meshes):
    """pose -> the state record (T_CG = T_O, T_C of a plain solid = T_CG; AABB of the entity's meshes in the world)"""
    T = make_transform(quat_to_basis(e["q"]), e["pos"])
    e["T"] = T
    st = e["state"]
    st["T_CG"][0], st["T_C"][0], st["T_O_basis"][0] = T, T, T[:9]
    st["v"][0], st["omega"][0] = e["v"], e["w"]
    lo, hi = np.full(3, 1e30), np.full(3, -1e30)
    for mi, po in e["parts"]:
        a, b = scenes.world_aabb(meshes[mi], transform_mul(T, make_transform(None, po)))
        lo, hi = np.minimum(lo, a), np.maximum(hi, b)
    st["aabb_min"][0], st["aabb_max"][0] = lo, hi


def test_fluid_dynamics_scene_through_the_patched_call_site():
    ocean, sc, ents, meshes = build(rc.load_real_meshes())
    dt, steps = 0.005, 200
    seen = set()
    worst = 0.0
    for _ in range(steps):
        for e in ents:
            sc.set_state(e["id"], e["state"], e["T"] if e["kind"] == "compound" else None)
        g, r = sc.step()
        for k, e in enumerate(ents):
            seen.add(int(r["position"][k]))
            # members the reference leaves untouched are untouched by the adapter too (both arms started from the same sentinel)
            for f in VEC_FIELDS + ("Swet", "Vsub"):
                assert np.array_equal(np.asarray(g[f][k]) == SENTINEL, np.asarray(r[f][k]) == SENTINEL), (k, f, g[f][k], r[f][k])
            live = lambda x: np.where(np.asarray(x) == SENTINEL, 0.0, np.asarray(x))  # noqa: E731
            F = sum(np.linalg.norm(live(r[f][k])) for f in ("Fb", "Fdq", "Fdf"))
            T = sum(np.linalg.norm(live(r[f][k])) for f in ("Tb", "Tdq", "Tdf"))
            for f in VEC_FIELDS:
                scale = max(F if f.startswith("F") else max(T, F), 1e-12)
                err = np.abs(live(g[f][k]) - live(r[f][k])).max() / scale
                worst = max(worst, err)
                assert err <= 2e-4, (k, e["kind"], f, g[f][k], r[f][k])   # the reference's own fp32 summation noise
            for f in ("Swet", "Vsub"):
                assert abs(live(g[f][k]) - live(r[f][k])) <= 2e-4 * max(abs(live(r[f][k])), 1e-12), (k, f)
            # ApplyHydrodynamicForces (SolidEntity.cpp:1958-1962) + weight, with the GPU arm's members
            Fh = live(g["Fb"][k]) + live(g["Fdq"][k]) + live(g["Fdf"][k]) + np.array([0.0, 0.0, ocean.gravity * e["mass"]])
            Th = live(g["Tb"][k]) + live(g["Tdq"][k]) + live(g["Tdf"][k])
            inertia = e["mass"] * np.array([0.1, 0.2, 0.25])
            R = quat_to_basis(e["q"]).reshape(3, 3)
            e["v"] = e["v"] + dt * Fh / e["mass"]
            e["w"] = e["w"] + dt * (R @ ((R.T @ Th) / inertia))
            e["pos"] = e["pos"] + dt * e["v"]
            q = e["q"] + 0.5 * dt * quat_mul(np.array([e["w"][0], e["w"][1], e["w"][2], 0.0]), e["q"])
            e["q"] = q / np.linalg.norm(q)
            fill(e, meshes)
    sc.close()
    assert seen == {0, 1, 2}          # bodies were outside, inside and across the surface along the way
    assert worst > 0.0                 # and the two arms are not the same code
