"""The oracle pinned against the reference's OWN source text.

tests/golden/ref_golden.json holds the outputs of the reference's definitions of the hot path — SolidEntity's integrators,
dispatcher and corrections, Compound, CableEntity, Ocean, the wave lookup, Uniform/Jet/Pipe, Atmosphere — compiled here from
the text where it lies (oracle/ref_extract.py + oracle/ref_harness.cpp, real Bullet, a stand-in for the GLM primitives only) on
the deterministic cases of tests/golden/ref_cases.py. The oracle, in the reference's own accumulation mode (fp32 sequential
sums), must reproduce every number bit for bit. Where the reference checkout is mounted the library is rebuilt and compared
live on more cases."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from stonefish_b200 import scenes
from tests.golden import ref_cases as rc

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_golden.json")))
HAVE_REF = os.path.isdir("/root/reference/Library/src") and os.path.isdir("/root/reference/3rdparty/LinearMath")


def unhex(v):
    return np.array([float.fromhex(x) for x in v])


def check_info(got, rec, where):
    for k in rc.VEC:
        assert np.array_equal(np.asarray(got[k]), unhex(rec["vec"][k])), (where, k, got[k], unhex(rec["vec"][k]))
    assert float(got["Swet"]) == float.fromhex(rec["Swet"]) and float(got["Vsub"]) == float.fromhex(rec["Vsub"]), where
    assert int(got["position"]) == rec["position"], where
    assert [int(got["wrote_buoyancy"]), int(got["wrote_damping"])] == rec["wrote"], where


def test_bodies_every_mode_and_flag(oracle):
    """288+ bodies: raw integrators and the dispatcher, every position, waves, Uniform/Jet/Pipe, stale-output semantics."""
    n = 0
    for k, recs in enumerate(G["bodies"]):
        sc = rc.body_scene(k)
        osc = oracle.OracleScene(sc.meshes, sc.body_mesh, sc.ocean, sc.states)
        assert len(recs) == sc.n_bodies
        for i, rec in enumerate(recs):
            check_info(osc.body(i, oracle.ORC_REF_F32_SEQ), rec, (k, i))
            n += 1
    assert n >= 100
    assert {r["position"] for recs in G["bodies"] for r in recs} == {0, 1, 2}


def test_depth_and_fluid_velocity_points(oracle):
    oc = rc.field_ocean(True)
    o, w = oc.to_struct(), oc.waves_struct()
    L = oracle.lib()
    for p, d, v in zip(rc.query_points(), G["depth"], G["velocity"]):
        assert L.orc_get_depth(C.byref(o), C.byref(w), float(p[0]), float(p[1]), float(p[2])) == float.fromhex(d)
        out = (C.c_float * 3)()
        L.orc_fluid_velocity_f(C.byref(o), (C.c_float * 3)(*[float(x) for x in p]), out)
        assert np.array_equal(np.array(out, dtype=np.float64), unhex(v), equal_nan=True)  # the jet outlet centre is 0/0 in the reference
    oracle._aero_lib()
    a = rc.wind("jetpipe").to_struct()
    for p, v in zip(rc.query_points(), G["wind"]):
        out = (C.c_float * 3)()
        L.orc_wind_velocity_f(C.byref(a), (C.c_float * 3)(*[float(x) for x in p]), out)
        assert np.array_equal(np.array(out, dtype=np.float64), unhex(v), equal_nan=True)


def test_aerodynamics_static_and_member(oracle):
    meshes_ = rc.AERO_MESHES()
    for (kind, dispatch, mi, st, typ, params), rec in zip(rc.aero_cases(), G["aero"]):
        r = oracle.aero(meshes_, [mi], rc.wind(kind), st, acc_mode=oracle.ORC_REF_F32_SEQ)
        assert np.array_equal(r["Fdq"][0], unhex(rec["Fda"])) and np.array_equal(r["Tdq"][0], unhex(rec["Tda"])), (kind, dispatch, typ)
    assert any(np.any(unhex(r["Fda"]) != 0) for r in G["aero"])


def test_compound_hydrodynamics_and_aerodynamics(oracle):
    for (depth, waves, flags), rec in zip(rc.COMPOUND_CASES, G["compound"]):
        ocean, ms, part_mesh, parts, external, origins, o2cs, comp, T_O = rc.compound_case(depth, waves, flags)
        check_info(oracle.compound(ms, part_mesh, parts, external, ocean, comp, acc_mode=oracle.ORC_REF_F32_SEQ), rec, (depth, waves, flags))
    atm, ms, part_mesh, pa, external, origins, o2cs, comp, T_O, kinds, params = rc.compound_aero_case()
    r = oracle.compound_aero(ms, part_mesh, pa, external, atm, comp, acc_mode=oracle.ORC_REF_F32_SEQ)
    assert np.array_equal(r["Fdq"], unhex(G["compound_aero"]["Fda"])) and np.array_equal(r["Tdq"], unhex(G["compound_aero"]["Tda"]))


def test_cables_in_waves_and_currents(oracle):
    for (ocean, radius, flags, pos, vel), rec in zip(rc.cable_cases(), G["cables"]):
        assert np.array_equal(oracle.cable(ocean, radius, flags, pos, vel).reshape(-1), unhex(rec))


def test_wave_producer_host_tables(oracle):
    """OpenGLOcean::GenerateWavesSpectrum (spectrum, LCG phases) and ComputeButterflyLookupTable from the reference's text: the
    oracle's restatement gives the same 2 x 262144 + 8192 floats, bit for bit."""
    import hashlib

    from stonefish_b200.hydro import wave_params

    rec = G["wave_tables"]
    w = oracle.WaveOracle(wave_params(rec["sea_state"], rec["seed"]))
    assert hashlib.sha256(w.spectrum12.tobytes()).hexdigest() == rec["spectrum12"]
    assert hashlib.sha256(w.spectrum34.tobytes()).hexdigest() == rec["spectrum34"]
    assert hashlib.sha256(w.butterfly.tobytes()).hexdigest() == rec["butterfly"]
    assert [float(x).hex() for x in w.spectrum12[np.nonzero(w.spectrum12)[0][:8]]] == rec["spectrum12_first_nonzero"]


@pytest.mark.skipif(not HAVE_REF, reason="reference checkout not mounted")
def test_live_against_the_compiled_reference_text(oracle):
    """Rebuild oracle/_ref/librefsrc.so from the reference's sources and compare on scenes that are NOT in the golden file,
    incl. a 50k-face hull per position."""
    from oracle import pyrefsrc as rs

    oracle.build()
    assert rs.available()
    for seed, waves, cur in ((31, True, "jetpipe"), (32, False, "uniform")):
        sc = scenes.mixed_small(seed, waves, cur)
        fl = rs.ocean_fluid(sc.ocean)
        osc = oracle.OracleScene(sc.meshes, sc.body_mesh, sc.ocean, sc.states)
        for i in range(sc.n_bodies):
            r, o = rs.body(sc.meshes[int(sc.body_mesh[i])], fl, sc.states[i:i + 1]), osc.body(i, oracle.ORC_REF_F32_SEQ)
            for k in rc.VEC + ("Swet", "Vsub", "position", "wrote_buoyancy", "wrote_damping"):
                assert np.array_equal(r[k], o[k]), (seed, i, k)
    for sc in (scenes.config2(100, 51), scenes.config3(100, 51, True), scenes.config3(100, 51, False), scenes.config4(2, 250, 101)):
        fl = rs.ocean_fluid(sc.ocean)
        osc = oracle.OracleScene(sc.meshes, sc.body_mesh, sc.ocean, sc.states)
        for i in range(sc.n_bodies):
            r, o = rs.body(sc.meshes[int(sc.body_mesh[i])], fl, sc.states[i:i + 1]), osc.body(i, oracle.ORC_REF_F32_SEQ)
            for k in rc.VEC + ("Swet", "Vsub"):
                assert np.array_equal(r[k], o[k]), (sc.name, i, k)
    # the wave producer's tables for another sea state (the spectrum only once per process: the reference's seed is a static)
    from stonefish_b200.hydro import wave_params

    p = wave_params(2.5)
    s12, s34, bf = rs.wave_tables(p)
    w = oracle.WaveOracle(p)
    assert np.array_equal(s12.view(np.uint32), w.spectrum12.view(np.uint32)) and np.array_equal(s34.view(np.uint32), w.spectrum34.view(np.uint32))
    assert np.array_equal(bf.view(np.uint32), w.butterfly.view(np.uint32))
    # glMatrixFromTransform
    T = sc.states["T_C"][0]
    got, want = (C.c_float * 16)(), (C.c_float * 16)()
    rs.lib().refx_gl_matrix((C.c_double * 12)(*T), got)
    oracle.lib().orc_gl_matrix((C.c_double * 12)(*T), want)
    assert list(got) == list(want)


def test_the_references_own_meshes(oracle):
    """Tests/Data hull_hydro, sphere_R=1, boat, base_link_hydro and torus (fixture tests/golden/ref_meshes.npz): the FluidDynamicsTest
    snapshot with its real hulls, the FloatingTest boat crossing flat and wavy seas at three drafts, every mesh under and across the
    surface in Uniform + Jet + Pipe currents — the oracle reproduces the reference text's outputs bit for bit on all of them."""
    real = rc.load_real_meshes()
    assert [real[n].n_faces for n in rc.REAL_MESH_NAMES] == [2880, 960, 6339, 240, 2304]
    scs = rc.real_mesh_scenes(real)
    assert [s.name for s in scs] == [g["scene"] for g in G["real_meshes"]]
    seen = set()
    for sc, g in zip(scs, G["real_meshes"]):
        osc = oracle.OracleScene(sc.meshes, sc.body_mesh, sc.ocean, sc.states)
        assert len(g["bodies"]) == sc.n_bodies
        for i, rec in enumerate(g["bodies"]):
            check_info(osc.body(i, oracle.ORC_REF_F32_SEQ), rec, (sc.name, i))
            seen.add(rec["position"])
    assert seen == {0, 1, 2}
