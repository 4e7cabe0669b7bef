#!/usr/bin/env python
"""A floating hull on a generated sea with a wind blowing and a tether hanging from it — every entry point of the library in
one place (needs a B200; there is no CPU fallback).

    python examples/floating_vessel.py

What happens per simulated step, and what it replaces in Stonefish:
  ctx.wavegen_simulate(t)        OpenGLOcean::Simulate + the PBO read-back           (waves without OpenGL)
  ctx.compute(ocean, states)     the Ocean::ApplyFluidForces loop over solids        (buoyancy, form drag, skin friction)
  ctx.compute_aero(atm, states)  the Atmosphere::ApplyFluidForces loop               (air drag on the same mesh)
  ctx.compute_cables(...)        the CABLE branch of Ocean::ApplyFluidForces         (per-node forces of a tether)
Bullet would integrate the wrench; here a few lines of semi-implicit Euler stand in for it."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stonefish_b200 import HydroContext, meshes, scenes  # noqa: E402
from stonefish_b200._abi import SFH_F_BUOYANCY, SFH_F_BUOYANT, SFH_F_DAMPING, SFH_F_DISPATCH, SFH_POS_AUTO  # noqa: E402
from stonefish_b200.hydro import (Atmosphere, GeometryApproxType, Ocean, Uniform, aero_proxy, make_transform, new_body_states,  # noqa: E402
                                  quat_to_basis, wave_params)


def main():
    hull = meshes.uv_ellipsoid(2.0, 0.6, 0.4, 200, 101)          # 40k triangles
    mass, inertia = 0.55 * 1000.0 * meshes.signed_volume(hull), np.array([120.0, 900.0, 950.0])
    with HydroContext(0) as ctx:
        mesh_id = ctx.upload_mesh(hull)                           # once: Morton-ordered chunks in HBM
        ctx.wavegen_init(wave_params(sea_state=1.0))              # once: spectrum + butterfly table
        ocean = Ocean(1025.0, 9.81)
        ocean.useGeneratedWaves(True)
        ocean.AddVelocityField(Uniform((0.3, 0.0, 0.0)))
        ocean.EnableCurrents()
        wind = Atmosphere(1.2)
        wind.AddVelocityField(Uniform((8.0, 2.0, 0.0)))

        pos, q = np.array([0.0, 0.0, 0.1]), np.array(scenes.euler_zyx_quat(0.2, 0.0, 0.05))
        v, w = np.zeros(3), np.zeros(3)
        dt, cd, flag = 0.002, *aero_proxy(GeometryApproxType.ELLIPSOID, (2.0, 0.6, 0.4))
        for step in range(500):
            t = step * dt
            ctx.wavegen_simulate(t)
            R = quat_to_basis(q)
            T = make_transform(R, pos)
            st = new_body_states(1)
            st["mesh_id"], st["mode"] = mesh_id, SFH_POS_AUTO
            st["flags"] = SFH_F_DISPATCH | SFH_F_DAMPING | SFH_F_BUOYANCY | SFH_F_BUOYANT
            st["T_CG"][0], st["T_C"][0], st["T_O_basis"][0] = T, T, T[:9]
            st["v"][0], st["omega"][0] = v, w
            st["aabb_min"][0], st["aabb_max"][0] = scenes.world_aabb(hull, T)
            st["Cd"][0], st["Cf"][0] = (1.0, 1.1, 1.2), (0.05, 0.05, 0.05)
            st["volume"][0], st["surface"][0] = meshes.signed_volume(hull), meshes.surface_area(hull)
            hydro = ctx.compute(ocean, st)[0]

            air = st.copy()
            air["flags"] = SFH_F_DISPATCH | flag
            air["Cd"][0] = cd
            aero = ctx.compute_aero(wind, air)[0]                 # Fda, Tda arrive in Fdq, Tdq

            s = np.linspace(0.0, 1.0, 60)[:, None]                # a tether from the keel to a point 6 m down-current
            nodes = pos + np.array([0.0, 0.0, 0.4]) + s * np.array([3.0, 0.0, 6.0])
            tether = ctx.compute_cables(ocean, [(0, 60, 0.01, 7)], nodes, np.tile(v, (60, 1)))

            F = hydro["Fb"] + hydro["Fdq"] + hydro["Fdf"] + aero["Fdq"] + np.array([0.0, 0.0, 9.81 * mass])
            Tq = hydro["Tb"] + hydro["Tdq"] + hydro["Tdf"] + aero["Tdq"]
            v = v + dt * F / mass
            Rm = R.reshape(3, 3)
            w = w + dt * (Rm @ ((Rm.T @ Tq) / inertia))
            pos = pos + dt * v
            dq = 0.5 * dt * np.array([w[0] * q[3] + w[1] * q[2] - w[2] * q[1], -w[0] * q[2] + w[1] * q[3] + w[2] * q[0],
                                      w[0] * q[1] - w[1] * q[0] + w[2] * q[3], -w[0] * q[0] - w[1] * q[1] - w[2] * q[2]])
            q = (q + dq) / np.linalg.norm(q + dq)
            if step % 100 == 0:
                print(f"t={t:5.2f} s  heave z={pos[2]:+.3f} m  Vsub={hydro['Vsub']:.3f} m^3  wetted {hydro['Swet']:.2f} m^2  "
                      f"|air drag|={np.linalg.norm(aero['Fdq']):.1f} N  tether pull={np.linalg.norm(tether[:, :9].reshape(60, 3, 3).sum(axis=(0, 1))):.1f} N")


if __name__ == "__main__":
    main()
