"""The deterministic cases behind tests/golden/ref_golden.json: every input is rebuilt from code (scenes / seeds), so the golden
file only has to hold the OUTPUTS of the reference's own compiled source text (oracle/_ref/librefsrc.so). Shared by
make_ref_golden.py (writes the file where the reference is mounted) and tests/test_oracle_refsource.py (checks the oracle
against it everywhere, and against the library itself where it exists)."""
import numpy as np

from stonefish_b200 import meshes as M
from stonefish_b200 import scenes
from stonefish_b200._abi import SFH_F_BUOYANCY, SFH_F_BUOYANT, SFH_F_DAMPING, SFH_F_DISPATCH, SFH_POS_AUTO
from stonefish_b200.hydro import (Atmosphere, GeometryApproxType, Jet, Ocean, Pipe, Uniform, aero_proxy, make_transform, new_body_states,
                                  quat_to_basis, transform_mul)

BODY_SCENES = [(7, False, "none"), (7, True, "uniform"), (7, True, "jetpipe"), (8, False, "jetpipe"), (21, True, "none")]
VEC = ("Fb", "Tb", "Fdq", "Tdq", "Fdf", "Tdf")


def body_scene(k):
    seed, waves, cur = BODY_SCENES[k]
    return scenes.mixed_small(seed, waves, cur)


def field_ocean(waves=True):
    oc = Ocean(1000.0, 9.81)
    if waves:
        oc.setWaves(scenes.wave_field())
    oc.AddVelocityField(Uniform((0.2, -0.1, 0.0)))
    oc.AddVelocityField(Jet((0.0, -3.0, 1.0), (0.1, 1.0, 0.05), 0.3, 5.0))
    oc.AddVelocityField(Pipe((-6.0, 0.5, 1.5), (6.0, -0.5, 2.0), 2.0, 3.0, 2.0, 2.0))
    dis = Uniform((9.0, 9.0, 9.0))
    dis.setEnabled(False)
    oc.AddVelocityField(dis)
    oc.EnableCurrents()
    return oc


def query_points():
    rng = np.random.default_rng(1)
    pts = [np.float32(rng.uniform(-8, 8, 3)) if i % 3 else np.float32(rng.uniform(-2000, 2000, 3)) for i in range(400)]
    pts += [np.float32(p) for p in ([0, 0, 0], [0.5 * 893 / 256, 0, 1], [-0.0, 101.0, -1], [893.0, 893.0, 0.25], [0.0, -3.0, 1.0], [0.3, -2.0, 1.05])]
    return np.array(pts, dtype=np.float32)


def wind(kind):
    atm = Atmosphere(1.2)
    if kind in ("uniform", "jetpipe"):
        atm.AddVelocityField(Uniform((6.0, -2.0, 0.5)))
        off = Uniform((-1.0, 0.5, 0.0))
        off.setEnabled(False)  # Atmosphere::GetFluidVelocity does not look at it
        atm.AddVelocityField(off)
    if kind == "jetpipe":
        atm.AddVelocityField(Jet((0.0, -6.0, -1.0), (0.1, 1.0, 0.05), 0.5, 12.0))
        atm.AddVelocityField(Pipe((-8.0, 0.5, -1.5), (8.0, -0.5, -2.0), 2.0, 3.0, 4.0, 2.0))
    return atm


AERO_MESHES = lambda: [M.icosphere(0.5, 3), M.box(1.0, 0.5, 0.2, 3), M.uv_ellipsoid(1.2, 0.15, 0.15, 60, 41)]  # noqa: E731


def aero_cases():
    """(wind kind, dispatch, mesh index, state, approx type, params) — state carries the CorrectAerodynamicForces inputs in the sfh
    form (Cd vector, SFH_F_AERO_CD), type/params are what the reference's member function is given."""
    out = []
    q = scenes.random_unit_quats(11, 12)
    rng = np.random.default_rng(11)
    for kind in ("none", "uniform", "jetpipe"):
        for dispatch in (False, True):
            for i in range(6):
                st = new_body_states(1)
                T = make_transform(quat_to_basis(q[i]), rng.uniform(-4, 4, 3))
                st["T_CG"][0] = T
                st["T_C"][0] = transform_mul(T, make_transform(None, rng.uniform(-0.05, 0.05, 3)))
                st["v"][0], st["omega"][0] = rng.uniform(-3, 3, 3), rng.uniform(-1, 1, 3)
                st["T_O_basis"][0] = quat_to_basis(q[(i + 5) % 12]).reshape(-1)
                typ, params = i % 4, rng.uniform(0.1, 2.0, 3)
                if dispatch:
                    cd, flag = aero_proxy(typ, params)
                    st["Cd"][0], st["flags"][0] = cd, SFH_F_DISPATCH | flag
                out.append((kind, dispatch, i % 3, st, typ, params))
    return out


COMPOUND_MESHES = lambda: [M.icosphere(0.5, 2), M.box(1.0, 0.5, 0.2, 2), M.uv_ellipsoid(0.9, 0.18, 0.18, 24, 13)]  # noqa: E731
COMPOUND_LAYOUT = [(0, (0.0, 0.0, 0.0), 1, 1), (1, (1.2, 0.0, 0.0), 1, 1), (2, (0.0, 0.9, -1.5), 1, 1), (1, (-1.0, 0.0, 0.1), 1, 0),
                   (0, (0.3, 0.3, 0.2), 0, 1), (1, (0.0, -0.8, 0.0), 0, 0)]
COMPOUND_CASES = [(d, w, f) for d, w in ((0.05, False), (0.1, True), (4.0, False), (-6.0, False), (0.3, True))
                  for f in (SFH_F_DAMPING | SFH_F_BUOYANCY | SFH_F_BUOYANT, SFH_F_DAMPING, SFH_F_BUOYANCY | SFH_F_BUOYANT, SFH_F_DAMPING | SFH_F_BUOYANCY)]


def compound_case(depth, waves, flags):
    """Inputs of one Compound::ComputeHydrodynamicForces call: (ocean, meshes, part_mesh, parts, external, origins, o2c, comp, T_O).
    parts[i].T_C / T_O_basis are the products the reference forms itself (T_O * origin_i * O2C_i), made with transform_mul, which is
    pinned bit-exact against btTransform::operator* (tests/test_oracle_golden.py)."""
    meshes_ = COMPOUND_MESHES()
    ocean = Ocean(1000.0, 9.81)
    if waves:
        ocean.setWaves(scenes.wave_field())
    ocean.AddVelocityField(Uniform((0.3, -0.2, 0.05)))
    ocean.EnableCurrents()
    T_CG = make_transform(quat_to_basis(scenes.euler_zyx_quat(0.4, 0.1, -0.2)), (1.0, -2.0, depth))
    T_O = transform_mul(T_CG, make_transform(None, (0.05, 0.0, -0.02)))
    parts = new_body_states(len(COMPOUND_LAYOUT))
    external, part_mesh, origins, o2cs = [], [], [], []
    lo, hi = np.full(3, 1e30), np.full(3, -1e30)
    for i, (mi, org, ext, buoy) in enumerate(COMPOUND_LAYOUT):
        origin = make_transform(quat_to_basis(scenes.euler_zyx_quat(0.1 * i, 0, 0)), org)
        o2c = make_transform(None, (0.01, 0.0, 0.0))
        T_O_part = transform_mul(T_O, origin)
        parts["T_C"][i] = transform_mul(T_O_part, o2c)
        parts["T_O_basis"][i] = T_O_part[:9]
        parts["Cd"][i], parts["Cf"][i] = (1.0 + 0.1 * i, 0.9, 1.1), (0.05, 0.06 + 0.01 * i, 0.07)
        parts["flags"][i] = SFH_F_BUOYANT if buoy else 0
        a, b = scenes.world_aabb(meshes_[mi], parts["T_C"][i])
        lo, hi = np.minimum(lo, a), np.maximum(hi, b)
        external.append(ext)
        part_mesh.append(mi)
        origins.append(origin)
        o2cs.append(o2c)
    comp = new_body_states(1)[0]
    comp["T_CG"], comp["v"], comp["omega"] = T_CG, (0.5, 0.1, -0.2), (0.05, 0.02, 0.1)
    comp["mode"], comp["flags"] = SFH_POS_AUTO, flags
    comp["aabb_min"], comp["aabb_max"] = lo, hi
    comp["volume"], comp["surface"], comp["P_CB"] = 1.234, 5.678, (0.02, 0.0, -0.03)
    return ocean, meshes_, part_mesh, parts, external, np.array(origins), np.array(o2cs), comp, T_O


def compound_aero_case():
    ocean, meshes_, part_mesh, parts, external, origins, o2cs, comp, T_O = compound_case(0.3, False, SFH_F_DAMPING)
    kinds = [GeometryApproxType.SPHERE, GeometryApproxType.CYLINDER, GeometryApproxType.ELLIPSOID, GeometryApproxType.CYLINDER,
             GeometryApproxType.AUTO, GeometryApproxType.ELLIPSOID]
    pa, params = new_body_states(len(COMPOUND_LAYOUT)), []
    for i in range(len(COMPOUND_LAYOUT)):
        pa["T_C"][i] = parts["T_C"][i]
        prm = (0.9, 0.18 + 0.05 * i, 0.3)
        params.append(prm)
        cd, flag = aero_proxy(kinds[i], prm)
        pa["Cd"][i], pa["flags"][i] = cd, SFH_F_DISPATCH | flag
        pa["T_O_basis"][i] = quat_to_basis(scenes.euler_zyx_quat(0.0, 0.2 * i, 0.1)).reshape(-1)
    return wind("jetpipe"), meshes_, part_mesh, pa, external, origins, o2cs, comp, T_O, kinds, params


def cable_cases():
    """(ocean, radius, flags, pos, vel): cables in waves with Uniform + Jet + Pipe currents (flat-sea cables are in bt_golden.json)."""
    oc = field_ocean(True)
    rng = np.random.default_rng(5)
    out = []
    for n, r, flags in ((120, 0.01, 7), (2, 0.05, 7), (65, 0.03, 3), (40, 0.02, 5)):
        s = np.linspace(0, 1, n)
        line = np.stack([-4 + 8 * s + 0.05 * rng.standard_normal(n), -3 + 5 * s, -0.6 + 2.5 * s + 0.3 * np.sin(9 * s)], axis=1)
        out.append((oc, r, flags, line, rng.uniform(-0.8, 0.8, (n, 3))))
    return out


# ------------------------------------------------------------------------------------------------ the reference's own meshes
# Tests/Data/*.obj of the reference checkout: irregular valence, seams and slivers that UV ellipsoids do not have. Their vertex
# positions and faces are kept as a fixture (tests/golden/ref_meshes.npz, written by make_ref_golden.py where the reference is
# mounted) so that the GPU box, which has no reference checkout, runs the CUDA path on them.
REAL_MESH_NAMES = ("hull_hydro", "sphere_R=1", "boat", "base_link_hydro", "torus_R=1_r=025")
REAL_MESH_FIXTURE = __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "ref_meshes.npz")


def load_real_meshes():
    from stonefish_b200.hydro import Mesh

    z = np.load(REAL_MESH_FIXTURE)
    return {n: Mesh(z[n + "/v"], z[n + "/f"]) for n in REAL_MESH_NAMES}


def real_mesh_scenes(real):
    """Scenes on the reference's meshes: (a) the FluidDynamicsTest snapshot with its real hulls (FluidDynamicsTestManager.cpp:93-161),
    (b) the FloatingTest boat (FloatingTestManager.cpp:54-89) crossing a flat and a wavy sea at three drafts, raw integrator and
    dispatcher, (c) every mesh under water and crossing the surface in Uniform + Jet + Pipe currents."""
    from stonefish_b200.scenes import CG2C, FULL, RAW, Scene, _fill_body, config1, euler_zyx_quat, wave_field
    from stonefish_b200._abi import SFH_POS_CROSSING, SFH_POS_INSIDE

    out = [config1(real)]
    boat = real["boat"]
    for waves in (False, True):
        ocean = Ocean(1000.0, 9.81)
        if waves:
            ocean.setWaves(wave_field())
        ocean.AddVelocityField(Uniform((0.2, -0.1, 0.0)))
        ocean.EnableCurrents()
        st = new_body_states(6)
        for k in range(6):
            draft = (-0.05, 0.08, 0.2)[k % 3]  # CG above / at / below the mean surface (z down)
            q = euler_zyx_quat(0.4 * k, 0.03 * (k - 2), 0.06 * (k - 3))
            T_CG = make_transform(quat_to_basis(q), (3.0 * k - 7.0, 2.0 - 0.7 * k, draft))
            _fill_body(st, k, boat, T_CG, transform_mul(T_CG, CG2C), (0.9, 0.1 * k, 0.05), (0.02, 0.05, -0.1 + 0.03 * k),
                       SFH_POS_CROSSING if k < 3 else SFH_POS_AUTO, RAW if k < 3 else FULL, Cd=(1.0, 0.9, 1.1), Cf=(0.05, 0.06, 0.07))
        out.append(Scene("boat_" + ("waves" if waves else "flat"), [boat], np.zeros(6, dtype=np.int64), ocean, st))
    names = list(REAL_MESH_NAMES)
    ocean = field_ocean(True)
    st = new_body_states(2 * len(names))
    body_mesh = np.zeros(2 * len(names), dtype=np.int64)
    quats = scenes.random_unit_quats(77, 2 * len(names))
    for k in range(2 * len(names)):
        mi = k % len(names)
        body_mesh[k] = mi
        under = k < len(names)
        # under water inside the pipe / at the jet's outlet, then crossing the wavy surface above them
        org = (-4.0 + 2.0 * mi, 0.2 * mi - 0.4, 1.6 + 0.1 * mi) if under else (-3.0 + 1.5 * mi, -2.5 + 0.5 * mi, 0.05 * mi - 0.1)
        T_CG = make_transform(quat_to_basis(quats[k]), org)
        _fill_body(st, k, real[names[mi]], T_CG, transform_mul(T_CG, CG2C), (0.3, -0.2, 0.1 * mi), (0.1, -0.05, 0.02 * k),
                   SFH_POS_INSIDE if under else SFH_POS_CROSSING, FULL if k % 2 else RAW, Cd=(1.1, 0.8, 0.9), Cf=(0.1, 0.05, 0.08))
    out.append(Scene("real_meshes_jetpipe", [real[n] for n in names], body_mesh, ocean, st))
    return out
