"""The BASELINE.json configs as concrete, deterministic synthetic inputs.

A scene = (list of Mesh, mesh index per body, Ocean, body-state array). Scenes hold inputs only; the same arrays
are handed to the CUDA library (HydroContext) and, in tests/bench, to the CPU oracle.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from . import meshes as M
from ._abi import SFH_F_BUOYANCY, SFH_F_BUOYANT, SFH_F_DAMPING, SFH_F_DISPATCH, SFH_POS_AUTO, SFH_POS_CROSSING, SFH_POS_INSIDE
from .hydro import Jet, Mesh, Ocean, Pipe, Uniform, make_transform, new_body_states, quat_to_basis, transform_mul


@dataclass
class Scene:
    name: str
    meshes: list[Mesh]
    body_mesh: np.ndarray  # index into `meshes` per body
    ocean: Ocean
    states: np.ndarray     # BODY_STATE_DTYPE; mesh_id is filled by `bind`
    notes: dict = field(default_factory=dict)

    @property
    def n_bodies(self) -> int:
        return int(self.states.shape[0])

    def faces_per_step(self) -> int:
        return int(sum(self.meshes[i].n_faces for i in self.body_mesh))

    def algorithmic_bytes(self) -> int:
        """sum over bodies of 12*F + 12*V, + 256+208 per body, + 1 MiB wave layer when waves are on."""
        b = sum(12 * self.meshes[i].n_faces + 12 * self.meshes[i].n_verts for i in self.body_mesh)
        b += (256 + 208) * self.n_bodies
        if self.ocean.hasWaves():
            b += 1048576
        return int(b)

    def bind(self, ctx) -> np.ndarray:
        """Upload every mesh once to `ctx` and return a state array whose mesh_id fields are device mesh ids."""
        ids = np.array([ctx.upload_mesh(m) for m in self.meshes], dtype=np.uint32)
        st = self.states.copy()
        st["mesh_id"] = ids[self.body_mesh]
        return st


def euler_zyx_quat(yaw: float, pitch: float, roll: float):
    """btQuaternion(yaw, pitch, roll) with BT_EULER_DEFAULT_ZYX = setEulerZYX (btQuaternion.h:119-135)."""
    hy, hp, hr = yaw * 0.5, pitch * 0.5, roll * 0.5
    cy, sy, cp, sp, cr, sr = math.cos(hy), math.sin(hy), math.cos(hp), math.sin(hp), math.cos(hr), math.sin(hr)
    return (sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy)


def random_unit_quats(seed: int, n: int) -> np.ndarray:
    u = M.splitmix_uniform(seed, 3 * n).reshape(n, 3)
    a, b = np.sqrt(1.0 - u[:, 0]), np.sqrt(u[:, 0])
    t1, t2 = 2.0 * np.pi * u[:, 1], 2.0 * np.pi * u[:, 2]
    return np.stack([a * np.sin(t1), a * np.cos(t1), b * np.sin(t2), b * np.cos(t2)], axis=1)


def world_aabb(mesh: Mesh, T_C: np.ndarray):
    p = mesh.positions.astype(np.float64) @ T_C[:9].reshape(3, 3).T + T_C[9:12]
    return p.min(axis=0), p.max(axis=0)


def wave_field(seed: int = 0x5EA5, n: int = 256, layers: int = 4) -> np.ndarray:
    """Deterministic periodic stand-in for OpenGLRealOcean::fftData, float32 [layers, N, N, 4] (SURVEY.md 8d cfg 3)."""
    j, i = np.meshgrid(np.arange(n, dtype=np.float64), np.arange(n, dtype=np.float64), indexing="ij")
    noise = (M.splitmix_uniform(seed, layers * n * n * 4).reshape(layers, n, n, 4) * 2.0 - 1.0) * 0.01
    f = noise.copy()
    f[0, :, :, 0] += 0.4 * np.sin(2 * np.pi * (3 * i + 2 * j) / n) + 0.1 * np.sin(2 * np.pi * (11 * i - 7 * j) / n)
    f[0, :, :, 1] += 0.05 * np.sin(2 * np.pi * (5 * i + 9 * j) / n)
    return f.astype(np.float32)


def _fill_body(st, k, mesh, T_CG, T_C, v, omega, mode, flags, Cd=(1.0, 1.0, 1.0), Cf=(0.1, 0.1, 0.1), T_O=None):
    st["mode"][k] = mode
    st["flags"][k] = flags
    st["T_CG"][k] = T_CG
    st["T_C"][k] = T_C
    st["v"][k] = v
    st["omega"][k] = omega
    lo, hi = world_aabb(mesh, T_C)
    st["aabb_min"][k], st["aabb_max"][k] = lo, hi
    st["T_O_basis"][k] = (T_CG if T_O is None else T_O)[:9]
    st["Cd"][k], st["Cf"][k] = Cd, Cf
    st["volume"][k] = M.signed_volume(mesh)
    st["surface"][k] = M.surface_area(mesh)
    st["P_CB"][k] = (0.01, -0.02, 0.03)


RAW = SFH_F_DAMPING | SFH_F_BUOYANCY
FULL = RAW | SFH_F_BUOYANT | SFH_F_DISPATCH
CG2C = make_transform(None, (-0.1, 0.0, 0.02))  # mesh frame offset from the CG frame


def config2(segments: int = 400, stacks: int = 251) -> Scene:
    """Single 200k-triangle AUV hull, fully submerged, uniform current (kernel K1)."""
    mesh = M.uv_ellipsoid(2.5, 0.3, 0.3, segments, stacks)
    ocean = Ocean(1000.0, 9.81)
    ocean.AddVelocityField(Uniform((0.5, -0.3, 0.0)))
    ocean.EnableCurrents()
    q = random_unit_quats(0xA11CE, 1)[0]
    T_CG = make_transform(quat_to_basis(q), (3.0, -2.0, 20.0))
    st = new_body_states(1)
    _fill_body(st, 0, mesh, T_CG, transform_mul(T_CG, CG2C), (1.0, 0.2, -0.1), (0.05, -0.1, 0.2), SFH_POS_INSIDE, RAW)
    return Scene("cfg2_submerged_hull", [mesh], np.zeros(1, dtype=np.int64), ocean, st, {"faces": mesh.n_faces})


def config3(segments: int = 400, stacks: int = 251, waves: bool = True, mesh: Mesh | None = None) -> Scene:
    """Partially submerged vessel with wave-height lookup and surface clipping (kernel K2)."""
    mesh = mesh or M.uv_ellipsoid(4.0, 1.0, 0.6, segments, stacks)
    ocean = Ocean(1000.0, 9.81)
    if waves:
        ocean.setWaves(wave_field())
    q = euler_zyx_quat(0.3, math.radians(2.0), math.radians(5.0))
    T_CG = make_transform(quat_to_basis(q), (12.0, -7.0, 0.1))
    st = new_body_states(1)
    _fill_body(st, 0, mesh, T_CG, transform_mul(T_CG, CG2C), (0.8, 0.1, 0.05), (0.02, 0.05, -0.1), SFH_POS_CROSSING, RAW)
    return Scene("cfg3_floating_vessel" + ("_waves" if waves else "_flat"), [mesh], np.zeros(1, dtype=np.int64), ocean, st)


def config4(n_instances: int = 4096, segments: int = 250, stacks: int = 101, unique_meshes: bool = False, first: int = 0) -> Scene:
    """Independent AUV instances (LearningTest-style batch), 50k triangles each; instance ids first..first+n-1."""
    base = M.uv_ellipsoid(1.2, 0.15, 0.15, segments, stacks)
    ids = np.arange(first, first + n_instances)
    if unique_meshes:
        meshes = [M.jitter(base, 0xB200 + int(i)) for i in ids]
        body_mesh = np.arange(n_instances, dtype=np.int64)
    else:
        meshes, body_mesh = [base], np.zeros(n_instances, dtype=np.int64)
    ocean = Ocean(1000.0, 9.81)
    ocean.AddVelocityField(Uniform((0.3, 0.1, -0.05)))
    ocean.EnableCurrents()
    st = new_body_states(n_instances)
    vol, surf = M.signed_volume(base), M.surface_area(base)
    for k, i in enumerate(ids):
        u = M.splitmix_uniform(0xB200 + int(i), 12)
        q = random_unit_quats(0xB200 + int(i) + 77, 1)[0]
        T_CG = make_transform(quat_to_basis(q), (-50 + 100 * u[0], -50 + 100 * u[1], 5 + 45 * u[2]))
        st["mode"][k], st["flags"][k] = SFH_POS_INSIDE, RAW
        st["T_CG"][k], st["T_C"][k] = T_CG, transform_mul(T_CG, CG2C)
        st["v"][k] = -2 + 4 * u[3:6]
        st["omega"][k] = -0.5 + u[6:9]
        st["T_O_basis"][k] = T_CG[:9]
        st["Cd"][k], st["Cf"][k] = (1.0, 0.8, 0.8), (0.1, 0.1, 0.1)
        st["volume"][k], st["surface"][k] = vol, surf
        org = T_CG[9:12]
        st["aabb_min"][k], st["aabb_max"][k] = org - 1.4, org + 1.4
    return Scene(f"cfg4_auv_batch_{'unique' if unique_meshes else 'shared'}", meshes, body_mesh, ocean, st)


def config5(n_vessels: int = 128, segments: int = 1000, stacks: int = 501, unique_meshes: bool = False, first: int = 0) -> Scene:
    """Vessels with 1M-triangle hulls crossing a wavy surface under Jet + Pipe currents."""
    base = M.uv_ellipsoid(4.0, 1.0, 0.6, segments, stacks)
    ids = np.arange(first, first + n_vessels)
    if unique_meshes:
        meshes = [M.jitter(base, 0xC0FFEE + int(i)) for i in ids]
        body_mesh = np.arange(n_vessels, dtype=np.int64)
    else:
        meshes, body_mesh = [base], np.zeros(n_vessels, dtype=np.int64)
    ocean = Ocean(1000.0, 9.81)
    ocean.setWaves(wave_field())
    ocean.AddVelocityField(Jet((0.0, 0.0, 1.0), (0.0, 1.0, 0.0), 0.3, 5.0))  # UnderwaterTestManager.cpp:117
    ocean.AddVelocityField(Pipe((-20.0, 0.0, 2.0), (20.0, 0.0, 2.0), 3.0, 5.0, 2.0, 2.0))
    ocean.EnableCurrents()
    st = new_body_states(n_vessels)
    for k, i in enumerate(ids):
        u = M.splitmix_uniform(0xC0FFEE + int(i), 12)
        q = euler_zyx_quat(2 * math.pi * u[0], math.radians(-4 + 8 * u[1]), math.radians(-10 + 20 * u[2]))
        T_CG = make_transform(quat_to_basis(q), (-15 + 30 * u[3], -6 + 12 * u[4], -0.2 + 0.5 * u[5]))
        st["mode"][k], st["flags"][k] = SFH_POS_CROSSING, RAW
        st["T_CG"][k], st["T_C"][k] = T_CG, transform_mul(T_CG, CG2C)
        st["v"][k] = -1 + 2 * u[6:9]
        st["omega"][k] = -0.2 + 0.4 * u[9:12]
        st["T_O_basis"][k] = T_CG[:9]
    return Scene(f"cfg5_vessels_{'unique' if unique_meshes else 'shared'}", meshes, body_mesh, ocean, st)


def config1_layout():
    """The entities of the FluidDynamicsTest snapshot: (kind, [(mesh index, part offset)], quaternion, origin, v, omega) with
    mesh indices 0 sphere, 1 box, 2 hull_hydro, 3 sphere_R=1 (FluidDynamicsTestManager.cpp:93-161)."""
    I = (0.0, 0.0, 0.0, 1.0)
    Z = (0.0, 0.0, 0.0)
    out = [("solid", [(0, Z)], I, (0, 0, 0), (0, 0, 0.4), Z),                                       # Sphere, crossing the surface
           ("solid", [(1, Z)], I, (2, 0, 0), (0.1, 0, 0.3), (0.0, 0.2, 0.0)),                        # Box, crossing
           ("compound", [(0, (0, 0, 0)), (0, (2.0, 0, 0))], euler_zyx_quat(0.0, 0.5, 0.0), (6, 0, -5), Z, Z),   # in the air
           ("solid", [(2, Z)], euler_zyx_quat(0.0, 0.5, 0.0), (8, 0, -5), Z, Z),                     # Hull, in the air
           ("solid", [(3, Z)], euler_zyx_quat(0.0, 0.5, 0.0), (11, 0, -0.3), (0, 0, 1.0), Z),        # SphereR1M, dropped to the surface
           ("solid", [(1, Z)], I, (4, 0, 0), (0, 0, 0.2), Z),                                        # Featherstone link
           ("solid", [(1, Z)], I, (-7, 0, 1), (0.6, 0, 0), Z),                                       # Hull1 (thruster driven)
           ("solid", [(1, Z)], euler_zyx_quat(math.pi / 2, 0, 0), (-5, 2, 1), (0, 0.6, 0), Z)]       # Hull2
    for i in range(8):
        yaw = math.pi / 4 * i
        out.append(("compound", [(1, (0, 0.5, 0)), (1, (0, -0.5, 0))], euler_zyx_quat(yaw, 0, 0), (0, 0, 2 + 2 * i),
                    (0.8 * math.cos(yaw), 0.8 * math.sin(yaw), 0.0), (0, 0, 0.05)))
    return out


def config1(real_meshes: dict | None = None) -> Scene:
    """FluidDynamicsTest scene (Tests/FluidDynamicsTest/FluidDynamicsTestManager.cpp:81-161) as a static snapshot:
    flat sea, SUBMERGED-mode buoyant bodies at their initial poses, full dispatcher (AUTO position from the AABB,
    analytic INSIDE buoyancy, coefficient correction). Compounds enter as one body per external part sharing T_CG.
    `real_meshes`: {"hull_hydro": Mesh, "sphere_R=1": Mesh} — the reference's own OBJ hulls (:111-114), e.g. from the fixture
    tests/golden/ref_meshes.npz; without them synthetic closed meshes with the same face counts (2880 / 960) stand in."""
    sphere = M.icosphere(0.5, 3)                 # Sphere.cpp:63 -> 1280 faces
    boxm = M.box(1.0, 0.5, 0.2, 3)               # Box.cpp:72   -> 768 faces
    real_meshes = real_meshes or {}
    hull = real_meshes.get("hull_hydro") or M.uv_ellipsoid(0.9, 0.18, 0.18, 40, 37)   # 2880 faces
    sph1 = real_meshes.get("sphere_R=1") or M.uv_ellipsoid(1.0, 1.0, 1.0, 24, 21)     # 960 faces
    meshes = [sphere, boxm, hull, sph1]
    ocean = Ocean(1000.0, 9.81)
    bodies = []  # (mesh idx, T_CG, T_C, v, omega)
    for kind, parts, q, o, v, w in config1_layout():
        t = make_transform(quat_to_basis(q), o)
        for mi, po in parts:  # a plain solid is one part at the origin of its own frame
            bodies.append((mi, t, transform_mul(t, make_transform(None, po)) if kind == "compound" else t, v, w))
    st = new_body_states(len(bodies))
    body_mesh = np.zeros(len(bodies), dtype=np.int64)
    for k, (mi, tcg, tc, v, w) in enumerate(bodies):
        body_mesh[k] = mi
        _fill_body(st, k, meshes[mi], tcg, tc, v, w, SFH_POS_AUTO, FULL, Cd=(1.05, 0.9, 1.2), Cf=(0.05, 0.08, 0.1))
    return Scene("cfg1_fluid_dynamics_test", meshes, body_mesh, ocean, st)


def mixed_small(seed: int = 7, waves: bool = True, currents: str = "jetpipe") -> Scene:
    """Small randomised scene that reaches every branch: all positions, all clipping classes, both flag settings,
    Uniform/Jet/Pipe, a degenerate-face mesh, strided (44-byte) vertices. Used by the parity tests."""
    a = M.uv_ellipsoid(1.5, 0.4, 0.3, 48, 25, stride_floats=6)    # Vertex{pos,normal} stride 24
    b = M.icosphere(0.6, 2, stride_floats=11)                     # TexturableVertex stride 44
    c = M.box(1.0, 0.5, 0.2, 2)
    deg_faces = np.concatenate([c.faces, [[0, 0, 1], [2, 2, 2], [3, 4, 3]]]).astype(np.uint32)
    d = Mesh(c.positions.copy(), deg_faces)                        # with degenerate faces
    meshes = [a, b, c, d]
    ocean = Ocean(1025.0, 9.81)
    if waves:
        ocean.setWaves(wave_field())
    if currents != "none":
        ocean.AddVelocityField(Uniform((0.2, -0.1, 0.05)))
        if currents == "jetpipe":
            ocean.AddVelocityField(Jet((0.0, -3.0, 1.0), (0.1, 1.0, 0.05), 0.3, 5.0))
            ocean.AddVelocityField(Pipe((-6.0, 0.5, 1.5), (6.0, -0.5, 2.0), 2.0, 3.0, 2.0, 2.0))
            off = Uniform((9.0, 9.0, 9.0))
            off.setEnabled(False)
            ocean.AddVelocityField(off)
        ocean.EnableCurrents()
    n = 24
    st = new_body_states(n)
    body_mesh = np.zeros(n, dtype=np.int64)
    u = M.splitmix_uniform(seed, n * 16).reshape(n, 16)
    quats = random_unit_quats(seed + 1, n)
    for k in range(n):
        mi = k % 4
        body_mesh[k] = mi
        depth = [-0.2, 0.05, 0.3, 3.0, -4.0, 0.0][k % 6] + 0.1 * u[k, 0]
        T_CG = make_transform(quat_to_basis(quats[k]), (-4 + 8 * u[k, 1], -4 + 8 * u[k, 2], depth))
        mode = [SFH_POS_AUTO, SFH_POS_CROSSING, SFH_POS_INSIDE, SFH_POS_AUTO][(k // 4) % 4]
        flags = [FULL, RAW, SFH_F_DAMPING, SFH_F_BUOYANCY, FULL & ~SFH_F_BUOYANT, SFH_F_DISPATCH | SFH_F_BUOYANT | SFH_F_BUOYANCY][k % 6]
        _fill_body(st, k, meshes[mi], T_CG, transform_mul(T_CG, CG2C), -1 + 2 * u[k, 3:6], -0.5 + u[k, 6:9], mode, flags,
                   Cd=0.5 + u[k, 9:12], Cf=0.05 + 0.1 * u[k, 12:15], T_O=make_transform(quat_to_basis(quats[(k + 3) % n])))
    return Scene(f"mixed_small_{seed}", meshes, body_mesh, ocean, st)
