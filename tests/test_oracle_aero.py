"""The aerodynamic twin in the oracle: CorrectAerodynamicForces pinned bit-for-bit against the
reference's own Bullet (tests/golden/bt_golden.json), and the static integrator against cases that can be done by hand —
including the two quirks the restatement keeps (`vn.length()` is glm's component count 3, wind ignores isEnabled())."""
import ctypes as C
import json
import os

import numpy as np

from stonefish_b200 import _abi, meshes
from stonefish_b200.hydro import (Atmosphere, GeometryApproxType, Jet, Uniform, aero_proxy, make_transform, new_body_states,
                                  quat_to_basis)

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "bt_golden.json")))
D = C.c_double


def unhex(v):
    return [float.fromhex(x) for x in v]


def test_correct_aerodynamic_forces_bit_exact(oracle):
    oracle._aero_lib()
    fn = oracle.lib().orc_correct_aero
    kinds = set()
    for rec in G["correct_aero"]:
        cd, flag = aero_proxy(rec["type"], unhex(rec["params"]))
        F, T = [(D * 3)(*unhex(v)) for v in rec["in"]]
        fn(F, T, (D * 3)(*cd), (D * 9)(*unhex(rec["B"])), 1 if flag else 0)
        assert list(F) == unhex(rec["out"][0]) and list(T) == unhex(rec["out"][1]), rec
        kinds.add(rec["type"])
    assert kinds == {0, 1, 2, 3}


def test_ellipsoid_proxy_vector_matches_bullet():
    for rec in G["ellipsoid_cd"]:
        cd, flag = aero_proxy(GeometryApproxType.ELLIPSOID, unhex(rec["params"]))
        assert cd.tolist() == unhex(rec["cd"]) and flag == _abi.SFH_F_AERO_CD
    assert aero_proxy(GeometryApproxType.CYLINDER)[0].tolist() == [0.5, 0.5, 1.0]
    assert aero_proxy(GeometryApproxType.SPHERE)[1] == 0 and aero_proxy(GeometryApproxType.AUTO)[1] == 0


def _state(T=None, v=(0, 0, 0), w=(0, 0, 0), flags=0):
    st = new_body_states(1)
    T = make_transform() if T is None else T
    st["T_CG"][0] = T
    st["T_C"][0] = T
    st["v"][0] = v
    st["omega"][0] = w
    st["flags"][0] = flags
    return st


def test_box_in_uniform_wind_known_answer(oracle):
    """Only the windward face (outward normal against the wind) contributes, and with the reference's `vn.length()` = 3
    the term is 3 * (vc.n) n * A: Fda = 0.5 * rho * 3 * U * A_face along the wind — LINEAR in U."""
    m = meshes.box(1.0, 0.5, 0.2, 3)
    rho = 1.2
    for U in (2.0, 4.0):
        atm = Atmosphere(rho)
        atm.AddVelocityField(Uniform((U, 0.0, 0.0)))
        out, codes = oracle.aero([m], [0], atm, _state(), codes=True)
        want = 0.5 * rho * 3.0 * U * (0.5 * 0.2)
        assert abs(out["Fdq"][0][0] - want) < 1e-5 * want
        assert abs(out["Fdq"][0][1]) < 1e-6 and abs(out["Fdq"][0][2]) < 1e-6
        assert np.linalg.norm(out["Tdq"][0]) < 1e-6  # symmetric face about the CG
        assert out["n_quadratic"][0] == m.n_faces // 6  # the triangles of the -x face
        assert int(out["n_quadratic"][0]) == int(np.count_nonzero(codes & _abi.SFH_FACE_QUADRATIC))
        assert out["n_faces"][0] == m.n_faces and out["wrote_damping"][0] == 1
        assert np.all(out["Fdf"] == 0) and np.all(out["Fb"] == 0)


def test_moving_body_in_still_air_equals_wind_on_still_body(oracle):
    """vc = wind - v: a body moving at -U through still air sees the same relative flow as wind +U on a resting body."""
    m = meshes.icosphere(0.5, 2)
    atm_w = Atmosphere(1.2)
    atm_w.AddVelocityField(Uniform((3.0, -1.0, 0.5)))
    a = oracle.aero([m], [0], atm_w, _state())
    b = oracle.aero([m], [0], Atmosphere(1.2), _state(v=(-3.0, 1.0, -0.5)))
    assert np.allclose(a["Fdq"], b["Fdq"], rtol=1e-6, atol=1e-9) and a["n_quadratic"][0] == b["n_quadratic"][0]


def test_wind_ignores_the_enabled_flag(oracle):
    """Atmosphere::GetFluidVelocity has no isEnabled() test (Atmosphere.cpp:146-152), unlike Ocean::GetFluidVelocity."""
    m = meshes.icosphere(0.5, 1)
    f = Uniform((5.0, 0.0, 0.0))
    atm = Atmosphere(1.2)
    atm.AddVelocityField(f)
    on = oracle.aero([m], [0], atm, _state())
    f.setEnabled(False)
    off = oracle.aero([m], [0], atm, _state())
    assert on["Fdq"].tolist() == off["Fdq"].tolist() and on["Fdq"][0][0] > 0


def test_f32_sequential_and_f64_sums_agree(oracle):
    m = meshes.uv_ellipsoid(1.2, 0.15, 0.15, 60, 41)
    atm = Atmosphere(1.2)
    atm.AddVelocityField(Uniform((1.0, 2.0, -0.5)))
    atm.AddVelocityField(Jet((0.0, -3.0, 0.0), (0.0, 1.0, 0.0), 0.3, 8.0))
    T = make_transform(quat_to_basis((0.2, -0.1, 0.3, 0.9)), (0.1, 0.2, -0.3))
    st = _state(T, v=(0.3, 0.0, 0.1), w=(0.2, -0.4, 0.1))
    a = oracle.aero([m], [0], atm, st, acc_mode=oracle.ORC_REF_F32_SEQ)
    b = oracle.aero([m], [0], atm, st, acc_mode=oracle.ORC_F64_ACC)
    assert a["n_quadratic"][0] == b["n_quadratic"][0] > 0
    assert np.allclose(a["Fdq"], b["Fdq"], rtol=2e-5, atol=1e-6) and np.allclose(a["Tdq"], b["Tdq"], rtol=2e-4, atol=1e-5)


def test_dispatch_applies_the_proxy_correction(oracle):
    m = meshes.icosphere(0.5, 2)
    atm = Atmosphere(1.2)
    atm.AddVelocityField(Uniform((0.0, 0.0, 4.0)))
    raw = oracle.aero([m], [0], atm, _state())
    st = _state(flags=_abi.SFH_F_DISPATCH)
    st["T_O_basis"][0] = np.eye(3).reshape(-1)
    sph = oracle.aero([m], [0], atm, st)  # AUTO/SPHERE: factor 1
    assert sph["Fdq"].tolist() == raw["Fdq"].tolist()
    cd, flag = aero_proxy(GeometryApproxType.CYLINDER)
    st["Cd"][0] = cd
    st["flags"][0] |= flag
    cyl = oracle.aero([m], [0], atm, st)  # force along z of the proxy: factor |Cd.z| = 1 up to the tiny x/y components
    assert np.allclose(cyl["Fdq"], raw["Fdq"], rtol=1e-4)
    st["T_O_basis"][0] = quat_to_basis((0.0, np.sin(np.pi / 4), 0.0, np.cos(np.pi / 4))).reshape(-1)  # proxy z -> world x
    cyl90 = oracle.aero([m], [0], atm, st)
    assert np.allclose(cyl90["Fdq"], 0.5 * raw["Fdq"], rtol=1e-4)
