"""Parity of the aerodynamic twin — sfh_aero_step through the C ABI against the oracle's restatement
of SolidEntity::ComputeAerodynamicForces / CorrectAerodynamicForces / Compound::ComputeAerodynamicForces.
Bar as for the hydrodynamic path: per-face codes and counters bit-exact, sums within 1e-9 of the force norm."""
import numpy as np
import pytest

from stonefish_b200 import meshes as M
from stonefish_b200 import scenes
from stonefish_b200._abi import SFH_F_DISPATCH, SFH_FACE_DEGENERATE, SFH_FACE_QUADRATIC
from stonefish_b200.hydro import (Atmosphere, GeometryApproxType, Jet, Mesh, Pipe, Uniform, aero_proxy, make_transform,
                                  new_body_states, quat_to_basis, sum_compound_aero, transform_mul)
from tests.parity import REL_TOL, assert_counts_equal, max_rel_err

pytestmark = pytest.mark.gpu


def _atmosphere(kind):
    atm = Atmosphere(1.2)
    if kind in ("uniform", "jetpipe"):
        atm.AddVelocityField(Uniform((6.0, -2.0, 0.5)))
        atm.AddVelocityField(Uniform((-1.0, 0.5, 0.0)))
    if kind == "jetpipe":
        atm.AddVelocityField(Jet((0.0, -6.0, -1.0), (0.1, 1.0, 0.05), 0.5, 12.0))
        atm.AddVelocityField(Pipe((-8.0, 0.5, -1.5), (8.0, -0.5, -2.0), 2.0, 3.0, 4.0, 2.0))
    return atm


def _bodies(seed, n, meshes_, dispatch):
    rng = np.random.default_rng(seed)
    st = new_body_states(n)
    body_mesh = rng.integers(0, len(meshes_), n)
    q = scenes.random_unit_quats(seed, n)
    for i in range(n):
        T = make_transform(quat_to_basis(q[i]), rng.uniform(-4, 4, 3))
        st["T_CG"][i] = T
        st["T_C"][i] = transform_mul(T, make_transform(None, rng.uniform(-0.05, 0.05, 3)))
        st["v"][i] = rng.uniform(-3, 3, 3)
        st["omega"][i] = rng.uniform(-1, 1, 3)
        st["mode"][i] = i % 4  # ignored by the aerodynamic path
        if dispatch:
            kind = i % 4
            cd, flag = aero_proxy(kind, rng.uniform(0.1, 2.0, 3))
            st["Cd"][i] = cd
            st["flags"][i] = SFH_F_DISPATCH | flag
            st["T_O_basis"][i] = quat_to_basis(q[(i + 1) % n]).reshape(-1)
    return st, body_mesh


@pytest.mark.parametrize("wind", ["none", "uniform", "jetpipe"])
@pytest.mark.parametrize("dispatch", [False, True])
def test_aero_bit_exact_codes_and_forces(ctx, oracle, wind, dispatch):
    meshes_ = [M.icosphere(0.5, 3), M.box(1.0, 0.5, 0.2, 3), M.uv_ellipsoid(1.2, 0.15, 0.15, 60, 41),
               M.jitter(M.uv_ellipsoid(0.8, 0.3, 0.2, 37, 23), 5, 1e-3)]
    atm = _atmosphere(wind)
    st, body_mesh = _bodies(11, 24, meshes_, dispatch)
    ids = [ctx.upload_mesh(m) for m in meshes_]
    st["mesh_id"] = [ids[int(k)] for k in body_mesh]
    n_faces = int(sum(meshes_[int(k)].n_faces for k in body_mesh))
    gi, gc = ctx.compute_aero_debug(atm, st, n_faces)
    ri, rc = oracle.aero(meshes_, body_mesh, atm, st, codes=True)
    assert_counts_equal(gi, ri)
    if wind == "jetpipe":  # exp/pow: CUDA vs glibc differ by <= 1-2 ulp
        assert np.array_equal(gc & 15, rc & 15) and (gc != rc).sum() <= 1
    else:
        assert np.array_equal(gc, rc)
    err, where = max_rel_err(gi, ri)
    assert err <= REL_TOL, (err, where)
    assert gi["n_quadratic"].sum() > 0 and (gi["position"] == 0).all() and (gi["wrote_damping"] == 1).all()
    assert not gi["Fdf"].any() and not gi["Fb"].any() and not gi["Swet"].any()
    # the non-debug entry gives the same numbers
    g2 = ctx.compute_aero(atm, st)
    for k in ("Fdq", "Tdq", "n_quadratic", "n_degenerate", "n_faces"):
        assert np.array_equal(g2[k], gi[k])
    for i in ids:
        ctx.release_mesh(i)


def test_aero_known_answer_and_reference_noise(ctx, oracle):
    """Box in a uniform wind: 0.5 * rho * 3 * U * A_face (the reference's vn.length() == 3), and the reference's own fp32
    sequential sums differ from the fp64 sums of the same terms only by their rounding noise."""
    m = M.box(1.0, 0.5, 0.2, 3)
    big = M.uv_ellipsoid(2.5, 0.3, 0.3, 400, 251)
    atm = Atmosphere(1.2)
    atm.AddVelocityField(Uniform((4.0, 0.0, 0.0)))
    ids = [ctx.upload_mesh(m), ctx.upload_mesh(big)]
    st = new_body_states(2)
    st["T_CG"][:] = make_transform()
    st["T_C"][:] = make_transform()
    st["mesh_id"] = ids
    st["omega"][1] = (0.0, 0.3, 0.1)
    gi = ctx.compute_aero(atm, st)
    assert abs(gi["Fdq"][0][0] - 0.5 * 1.2 * 3.0 * 4.0 * 0.1) < 1e-6
    assert gi["n_quadratic"][0] == m.n_faces // 6
    r32 = oracle.aero([m, big], [0, 1], atm, st, acc_mode=oracle.ORC_REF_F32_SEQ)
    r64 = oracle.aero([m, big], [0, 1], atm, st)
    assert_counts_equal(gi, r64)
    assert_counts_equal(gi, r32)
    assert max_rel_err(gi, r64)[0] <= REL_TOL
    assert max_rel_err(gi, r32)[0] <= 1e-4
    for i in ids:
        ctx.release_mesh(i)


def test_aero_degenerate_faces_and_soup(ctx, oracle):
    base = M.uv_ellipsoid(1.0, 0.4, 0.3, 40, 21)
    pos = base.positions.copy()
    faces = base.faces.copy()
    faces[5] = (faces[5][0], faces[5][0], faces[5][1])  # zero-area faces
    faces[300] = (faces[300][2], faces[300][2], faces[300][2])
    soup = Mesh(pos[faces[:700]].reshape(-1, 3), np.arange(2100, dtype=np.uint32).reshape(-1, 3))
    meshes_ = [Mesh(pos, faces), soup]
    atm = _atmosphere("uniform")
    st, _ = _bodies(3, 2, meshes_, True)
    ids = [ctx.upload_mesh(x) for x in meshes_]
    st["mesh_id"] = ids
    gi, gc = ctx.compute_aero_debug(atm, st, meshes_[0].n_faces + meshes_[1].n_faces)
    ri, rc = oracle.aero(meshes_, [0, 1], atm, st, codes=True)
    assert_counts_equal(gi, ri)
    assert np.array_equal(gc, rc)
    assert gi["n_degenerate"][0] == 2 and (gc[5] & SFH_FACE_DEGENERATE) and not (gc[5] & SFH_FACE_QUADRATIC)
    assert max_rel_err(gi, ri)[0] <= REL_TOL
    for i in ids:
        ctx.release_mesh(i)


def test_compound_aero_parts_sum_like_the_reference(ctx, oracle):
    """Compound::ComputeAerodynamicForces (Compound.cpp:439-464): external parts as bodies with the compound's T_CG and
    velocities, each corrected with its own proxy on the device, summed on the host in part order."""
    meshes_ = [M.icosphere(0.5, 2), M.box(1.0, 0.5, 0.2, 2), M.uv_ellipsoid(0.9, 0.18, 0.18, 24, 13)]
    atm = _atmosphere("jetpipe")
    T_CG = make_transform(quat_to_basis(scenes.euler_zyx_quat(0.4, 0.1, -0.2)), (1.0, -2.0, -3.0))
    T_O = transform_mul(T_CG, make_transform(None, (0.05, 0.0, -0.02)))
    layout = [(0, (0.0, 0.0, 0.0), 1, GeometryApproxType.SPHERE), (1, (1.2, 0.0, 0.0), 1, GeometryApproxType.CYLINDER),
              (2, (0.0, 0.9, -1.5), 1, GeometryApproxType.ELLIPSOID), (1, (-1.0, 0.0, 0.1), 0, GeometryApproxType.CYLINDER),
              (0, (0.3, 0.3, 0.2), 1, GeometryApproxType.AUTO)]
    parts = new_body_states(len(layout))
    external, part_mesh = [], []
    for i, (mi, org, ext, kind) in enumerate(layout):
        T_O_part = transform_mul(T_O, make_transform(quat_to_basis(scenes.euler_zyx_quat(0.1 * i, 0.05, 0)), org))
        parts["T_C"][i] = transform_mul(T_O_part, make_transform(None, (0.01, 0.0, 0.0)))
        cd, flag = aero_proxy(kind, (0.9, 0.18, 0.3))
        parts["Cd"][i] = cd
        parts["flags"][i] = SFH_F_DISPATCH | flag
        parts["T_O_basis"][i] = quat_to_basis(scenes.euler_zyx_quat(0.0, 0.2 * i, 0.1)).reshape(-1)
        external.append(ext)
        part_mesh.append(mi)
    comp = new_body_states(1)[0]
    comp["T_CG"], comp["v"], comp["omega"] = T_CG, (2.5, 0.1, -0.2), (0.05, 0.02, 0.3)
    st = parts.copy()
    st["T_CG"][:] = comp["T_CG"]
    st["v"][:] = comp["v"]
    st["omega"][:] = comp["omega"]
    ids = [ctx.upload_mesh(m) for m in meshes_]
    st["mesh_id"] = [ids[mi] for mi in part_mesh]
    got = sum_compound_aero(ctx.compute_aero(atm, st), external)
    ref = np.atleast_1d(oracle.compound_aero(meshes_, part_mesh, parts, external, atm, comp))
    for k in ("n_faces", "n_degenerate", "n_quadratic", "wrote_damping"):
        assert np.array_equal(got[k], ref[k]), k
    assert max_rel_err(got, ref)[0] <= REL_TOL
    assert np.linalg.norm(got["Fdq"]) > 0
    for i in ids:
        ctx.release_mesh(i)


def test_aero_rejects_bad_arguments(ctx):
    from stonefish_b200.hydro import SfhError

    st = new_body_states(1)
    st["mesh_id"] = 123456
    with pytest.raises(SfhError):
        ctx.compute_aero(Atmosphere(1.2), st)
    assert ctx.compute_aero(Atmosphere(1.2), new_body_states(0)).shape == (0,)
