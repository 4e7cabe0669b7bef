"""Closed loop (the question row N1 of SURVEY.md §8 asks: does parity survive feedback?): a floating body is integrated for 200
steps with the hydrodynamic wrench taken each step from (a) the CUDA path through the C ABI, (b) the oracle summing in fp64
(the parity contract) and (c) the oracle in the reference's own fp32 sequential sums. The integrator is a plain semi-implicit
Euler written here — Bullet is out of scope — so this checks error GROWTH, not Bullet: (a) and (b) must stay together to 1e-9,
and (a) must stay as close to (c) as the reference's rounding noise allows."""
import numpy as np
import pytest

from stonefish_b200 import meshes as M
from stonefish_b200 import scenes
from stonefish_b200._abi import SFH_F_BUOYANCY, SFH_F_BUOYANT, SFH_F_DAMPING, SFH_F_DISPATCH, SFH_POS_AUTO
from stonefish_b200.hydro import Ocean, Uniform, make_transform, new_body_states, quat_to_basis

pytestmark = pytest.mark.gpu


def quat_mul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz])


def simulate(wrench, mesh, ocean, steps, dt):
    mass, inertia = 60.0, np.array([4.0, 9.0, 11.0])
    pos, q = np.array([0.3, -0.2, 0.05]), np.array(scenes.euler_zyx_quat(0.3, 0.1, -0.15))
    v, w = np.array([0.4, 0.0, 0.2]), np.array([0.1, -0.05, 0.2])
    g = np.array([0.0, 0.0, ocean.gravity * mass])
    traj = []
    for _ in range(steps):
        R = quat_to_basis(q)
        T = make_transform(R, pos)
        st = new_body_states(1)
        st["T_CG"][0], st["T_C"][0], st["T_O_basis"][0] = T, T, T[:9]
        st["v"][0], st["omega"][0] = v, w
        st["mode"][0] = SFH_POS_AUTO
        st["flags"][0] = SFH_F_DISPATCH | SFH_F_DAMPING | SFH_F_BUOYANCY | SFH_F_BUOYANT
        lo, hi = scenes.world_aabb(mesh, T)
        st["aabb_min"][0], st["aabb_max"][0] = lo, hi
        st["Cd"][0], st["Cf"][0] = (1.0, 1.2, 0.9), (0.05, 0.05, 0.08)
        st["volume"][0], st["surface"][0] = M.signed_volume(mesh), M.surface_area(mesh)
        info = wrench(st)
        F = info["Fb"] + info["Fdq"] + info["Fdf"] + g   # ApplyHydrodynamicForces (SolidEntity.cpp:1958-1962) + weight
        Tq = info["Tb"] + info["Tdq"] + info["Tdf"]
        v = v + dt * F / mass
        Rm = R.reshape(3, 3)
        w = w + dt * (Rm @ ((Rm.T @ Tq) / inertia))
        pos = pos + dt * v
        q = q + 0.5 * dt * quat_mul(np.array([w[0], w[1], w[2], 0.0]), q)
        q /= np.linalg.norm(q)
        traj.append(np.concatenate([pos, q, v, w]))
    return np.array(traj)


@pytest.mark.parametrize("waves", [False, True])
def test_closed_loop_trajectories_stay_together(ctx, oracle, waves):
    mesh = M.box(1.2, 0.8, 0.4, 3)
    ocean = Ocean(1000.0, 9.81)
    if waves:
        ocean.setWaves(scenes.wave_field())
    ocean.AddVelocityField(Uniform((0.2, 0.1, 0.0)))
    ocean.EnableCurrents()
    mid = ctx.upload_mesh(mesh)

    def gpu(st):
        st = st.copy()
        st["mesh_id"] = mid
        return ctx.compute(ocean, st)[0]

    def cpu(mode):
        return lambda st: oracle.OracleScene([mesh], [0], ocean, st).body(0, mode)

    steps, dt = 200, 0.005
    a = simulate(gpu, mesh, ocean, steps, dt)
    b = simulate(cpu(oracle.ORC_F64_ACC), mesh, ocean, steps, dt)
    c = simulate(cpu(oracle.ORC_REF_F32_SEQ), mesh, ocean, steps, dt)
    scale = np.abs(b).max(axis=0) + 1e-12
    assert (np.abs(a - b) / scale).max() <= 1e-9          # the parity contract survives 200 steps of feedback
    assert (np.abs(a - c) / scale).max() <= 1e-3          # the reference's own fp32 summation noise, amplified by the loop
    assert np.abs(b[-1, :3] - b[0, :3]).max() > 0.05       # the body really moved
    ctx.release_mesh(mid)
