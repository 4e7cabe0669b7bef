"""The C++ host mirror (stonefish_b200/host/StonefishHydro.hpp: sfb::Ocean, sfb::Jet/Pipe/Uniform, sfb::HydroEngine::
ComputeHydrodynamicForces) driven by the demo executable on a FluidDynamicsTest-style scene, against the CPU oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

from stonefish_b200 import Jet, Pipe, Uniform, scenes
from stonefish_b200.hydro import INFO_DTYPE
from tests.parity import REL_TOL, max_rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "stonefish_b200", "host", "demo_fluid_dynamics")


def write_scene(path, sc):
    with open(path, "wb") as f:
        f.write(struct.pack("<IIIddII", 0x53484653, len(sc.meshes), sc.n_bodies, sc.ocean.density, sc.ocean.gravity,
                            1 if sc.ocean.currentsEnabled else 0, len(sc.ocean.currents)))
        for c in sc.ocean.currents:
            if isinstance(c, Uniform):
                t, a = 0, list(c.v) + [0.0] * 7
            elif isinstance(c, Jet):
                t, a = 1, list(c.c) + list(c.direction) + [c.r, c.vout, 0.0, 0.0]
            else:
                assert isinstance(c, Pipe)
                t, a = 2, list(c.p1) + list(c.p2) + [c.r1, c.r2, c.vin, c.gamma]
            f.write(struct.pack("<II10d", t, 1 if c.isEnabled() else 0, *a))
        for m in sc.meshes:
            f.write(struct.pack("<III", m.n_verts, m.stride_bytes, m.n_faces))
            f.write(m.vertex_data.tobytes())
            f.write(m.faces.tobytes())
        st = sc.states
        for k in range(sc.n_bodies):
            f.write(struct.pack("<II", int(sc.body_mesh[k]), 1 if st["flags"][k] & 4 else 0))
            T_O = np.concatenate([st["T_O_basis"][k], np.zeros(3)])
            for arr in (st["T_CG"][k], st["T_C"][k], T_O, st["v"][k], st["omega"][k], st["aabb_min"][k], st["aabb_max"][k], st["Cd"][k],
                        st["Cf"][k], [st["volume"][k], st["surface"][k]], st["P_CB"][k]):
                f.write(np.asarray(arr, dtype=np.float64).tobytes())


def read_results(path, n):
    rec = np.dtype([("d", np.float64, 20), ("u", np.uint32, 13)])
    raw = np.fromfile(path, dtype=rec, count=n)
    out = np.zeros(n, dtype=INFO_DTYPE)
    for i, k in enumerate(("Fb", "Tb", "Fdq", "Tdq", "Fdf", "Tdf")):
        out[k] = raw["d"][:, 3 * i:3 * i + 3]
    out["Swet"], out["Vsub"] = raw["d"][:, 18], raw["d"][:, 19]
    out["position"], out["n_faces"] = raw["u"][:, 0], raw["u"][:, 1]
    out["class_count"] = raw["u"][:, 2:10]
    out["n_degenerate"], out["n_quadratic"], out["n_skin"] = raw["u"][:, 10], raw["u"][:, 11], raw["u"][:, 12]
    return out


def test_cpp_host_mirror_runs_the_fluid_dynamics_scene(tmp_path, oracle):
    assert os.path.exists(DEMO), "build the host mirror demo first (make -C stonefish_b200/host)"
    sc = scenes.config1()
    sc.ocean.AddVelocityField(Uniform((0.2, -0.1, 0.0)))
    sc.ocean.AddVelocityField(Jet((0.0, -3.0, 1.0), (0.1, 1.0, 0.05), 0.3, 5.0))
    sc.ocean.AddVelocityField(Pipe((-6.0, 0.5, 1.5), (6.0, -0.5, 2.0), 2.0, 3.0, 2.0, 2.0))
    sc.ocean.EnableCurrents()
    scene, res = str(tmp_path / "scene.bin"), str(tmp_path / "out.bin")
    write_scene(scene, sc)
    txt = subprocess.check_output([DEMO, scene, res, "5"], text=True)
    assert txt.startswith("Hydrodynamics time: ")  # the reference's only live output (FluidDynamicsTestManager.cpp:166)
    got = read_results(res, sc.n_bodies)
    ref, _ = oracle.OracleScene(sc.meshes, sc.body_mesh, sc.ocean, sc.states).step(1, n_threads=4)
    for k in ("class_count", "n_degenerate", "n_faces", "position"):
        assert np.array_equal(got[k], ref[k]), k
    assert np.abs(got["n_quadratic"].astype(int) - ref["n_quadratic"].astype(int)).max() <= 1  # exp/pow 1-2 ulp, see DESIGN.md
    err, where = max_rel_err(got, ref)
    assert err <= REL_TOL, (err, where)
    assert set(np.unique(got["position"])) == {0, 1, 2}


def test_cpp_host_mirror_aerodynamics(tmp_path, oracle):
    """sfb::Atmosphere + HydroEngine::ComputeAerodynamicForces (incl. the proxy vectors of CorrectAerodynamicForces built in
    C++) on the same scene, against the oracle fed through the Python mirror's aero_proxy."""
    from stonefish_b200._abi import SFH_F_DISPATCH
    from stonefish_b200.hydro import Atmosphere, aero_proxy

    assert os.path.exists(DEMO), "build the host mirror demo first (make -C stonefish_b200/host)"
    sc = scenes.config1()
    fields = [Uniform((4.0, -1.0, 0.0)), Jet((0.0, -3.0, 1.0), (0.1, 1.0, 0.05), 0.3, 9.0),
              Pipe((-6.0, 0.5, 1.5), (6.0, -0.5, 2.0), 2.0, 3.0, 5.0, 2.0)]
    fields[1].setEnabled(False)  # wind ignores it (Atmosphere.cpp:146-152)
    atm = Atmosphere(1.2)
    for f in fields:
        sc.ocean.AddVelocityField(f)
        atm.AddVelocityField(f)
    scene, res, ares = str(tmp_path / "scene.bin"), str(tmp_path / "out.bin"), str(tmp_path / "aero.bin")
    write_scene(scene, sc)
    subprocess.check_output([DEMO, scene, res, "2", ares], text=True)
    raw = np.fromfile(ares, dtype=np.float64).reshape(sc.n_bodies, 6)
    st = sc.states.copy()
    for i in range(sc.n_bodies):
        cd, flag = aero_proxy(i % 4, st["Cd"][i])
        st["Cd"][i] = cd
        st["flags"][i] = SFH_F_DISPATCH | flag
    ref = oracle.aero(sc.meshes, sc.body_mesh, atm, st)
    got = np.zeros(sc.n_bodies, dtype=INFO_DTYPE)
    got["Fdq"], got["Tdq"] = raw[:, :3], raw[:, 3:]
    err, where = max_rel_err(got, ref)
    assert err <= REL_TOL, (err, where)
    assert np.linalg.norm(ref["Fdq"], axis=1).min() > 0


def test_cpp_host_mirror_generated_waves(tmp_path, ctx):
    """sfb::HydroEngine::InitWaves / SimulateWaves + Ocean::useGeneratedWaves in C++ give the step the Python mirror gives
    on the same generated field (same library underneath: bitwise)."""
    from stonefish_b200.hydro import wave_params

    assert os.path.exists(DEMO), "build the host mirror demo first (make -C stonefish_b200/host)"
    sc = scenes.config1()
    scene = str(tmp_path / "scene.bin")
    outs = [str(tmp_path / n) for n in ("out.bin", "aero.bin", "waves.bin", "cable.bin")]
    write_scene(scene, sc)
    subprocess.check_output([DEMO, scene, outs[0], "2", outs[1], "1.0", outs[2], outs[3]], text=True)  # sea state 1: decimetre waves
    raw = np.fromfile(outs[2], dtype=np.float64).reshape(sc.n_bodies, 8)
    ctx.wavegen_init(wave_params(1.0))
    ctx.wavegen_simulate(1.5)
    sc.ocean.useGeneratedWaves(True)
    st = sc.bind(ctx)
    st["mode"] = 3  # SFH_POS_AUTO, flags as HydroEngine::ComputeHydrodynamicForces sets them
    st["flags"] = (st["flags"] & 4) | 8 | 1 | 2
    got = ctx.compute(sc.ocean, st)
    crossing = got["position"] == 2
    assert crossing.any()
    assert np.array_equal(raw[crossing, 0:3], got["Fb"][crossing]) and np.array_equal(raw[crossing, 6], got["Swet"][crossing])
    flat = read_results(outs[0], sc.n_bodies)
    assert not np.array_equal(flat["Fb"][crossing], got["Fb"][crossing])  # the waves changed the answer
    # sfb::HydroEngine::ComputeCableForces on the same generated sea: the demo's 40-node cable through the Python mirror
    s = np.arange(40) / 39.0
    pos = np.stack([-2.0 + 4.0 * s, -1.0 + 2.0 * s, -0.5 + 2.0 * s], axis=1)
    vel = np.stack([0.1 * s, np.full(40, -0.2), 0.05 * np.arange(40)], axis=1)
    cab = ctx.compute_cables(sc.ocean, [(0, 40, 0.02, 7)], pos, vel)
    assert np.array_equal(np.fromfile(outs[3], dtype=np.float64).reshape(40, 9), cab) and np.abs(cab[:, 2]).max() > 0


@pytest.mark.parametrize("contexts", [2, 3])
def test_cpp_host_mirror_several_contexts(tmp_path, oracle, contexts):
    """sfb::HydroEngineMulti: one process, one host thread and one sfh context per shard, bodies block-partitioned (shard_range).
    With one GPU the shards are contexts on the same device: what is checked is the host side — partition, concurrent steps, results
    back in the caller's bodies, counters equal and sums equal to 1e-12 against ONE context, and against the oracle."""
    import json

    assert os.path.exists(DEMO), "build the host mirror demo first (make -C stonefish_b200/host)"
    sc = scenes.config1()                      # 25 bodies: outside, inside and across the surface
    sc.ocean.AddVelocityField(Uniform((0.2, -0.1, 0.0)))
    sc.ocean.EnableCurrents()
    scene, res = str(tmp_path / "scene.bin"), str(tmp_path / "out.bin")
    write_scene(scene, sc)
    txt = subprocess.check_output([DEMO, scene, res, "4", "--contexts", str(contexts)], text=True)
    rec = json.loads(txt.strip().splitlines()[-1])
    assert rec["scaling"] == "strong" and rec["contexts"] == contexts and rec["bodies_total"] == sc.n_bodies
    assert rec["counts_equal"] is True and rec["max_rel_diff_vs_one_context"] <= 1e-12
    got = read_results(res, sc.n_bodies)
    ref, _ = oracle.OracleScene(sc.meshes, sc.body_mesh, sc.ocean, sc.states).step(1, n_threads=4)
    for k in ("class_count", "n_degenerate", "n_quadratic", "n_skin", "n_faces", "position"):
        assert np.array_equal(got[k], ref[k]), k
    err, where = max_rel_err(got, ref)
    assert err <= REL_TOL, (err, where)
