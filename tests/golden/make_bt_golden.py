#!/usr/bin/env python
"""Generate tests/golden/bt_golden.json from the reference's OWN vendored Bullet LinearMath (oracle/_ref/libbtcheck.so,
built by `make -C oracle ref` from /root/reference/3rdparty/LinearMath — only possible where the reference checkout is
mounted). The vectors pin the fp64 half of the hot path (current fields, CorrectHydrodynamicForces, analytic buoyancy
tails, getOpenGLMatrix narrowing, btTransform/btQuaternion helpers); every value is stored as a C99 hex float so the
comparison in tests/test_oracle_golden.py is bit-exact. Run:  python tests/golden/make_bt_golden.py"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from stonefish_b200 import _abi, meshes  # noqa: E402

lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libbtcheck.so"))
assert lib.btc_scalar_size() == 8, "reference Bullet must be built with BT_USE_DOUBLE_PRECISION"
D = C.c_double


def arr(n):
    return (D * n)()


def hx(v):
    return [float(x).hex() for x in v]


def fhx(v):
    return [float(np.float32(x)).hex() for x in v]


def current(kind, a, n, r1=0.0, r2=0.0, l=0.0, vel=0.0, gamma=0.0):
    c = _abi.sfh_current()
    c.type, c.enabled = kind, 1
    c.a[:], c.n[:] = list(a), list(n)
    c.r1, c.r2, c.l, c.vel, c.gamma = r1, r2, l, vel, gamma
    return c


out = {"doc": __doc__, "currents": [], "axes": [], "correct": [], "gl_matrix": [], "inside_buoyancy": [], "flat_buoyancy": [],
       "euler": [], "quat": [], "mul": []}
u = meshes.splitmix_uniform(0xB17, 4000)
k = 0


def take(n, lo=-1.0, hi=1.0):
    global k
    v = lo + (hi - lo) * u[k:k + n]
    k += n
    return [float(x) for x in v]


# --- constructor-time axes (Jet.cpp:31-37, Pipe.cpp:31-40)
jet_n, pipe_n, pipe_l = arr(3), arr(3), D()
lib.btc_jet_axis((D * 3)(0.1, 1.0, 0.05), jet_n)
lib.btc_pipe_axis((D * 3)(-6.0, 0.5, 1.5), (D * 3)(6.0, -0.5, 2.0), pipe_n, C.byref(pipe_l))
out["axes"] = {"jet_dir": hx([0.1, 1.0, 0.05]), "jet_n": hx(jet_n), "pipe_p1": hx([-6.0, 0.5, 1.5]), "pipe_p2": hx([6.0, -0.5, 2.0]),
               "pipe_n": hx(pipe_n), "pipe_l": pipe_l.value.hex()}

# --- currents at points that hit every branch (inside/outside cone, behind outlet, beyond pipe ends, on the axis)
fields = {"uniform": current(0, (0.5, -0.3, 0.0), (0, 0, 0)),
          "jet": current(1, (0.0, -3.0, 1.0), list(jet_n), r1=0.3, vel=5.0),
          "pipe": current(2, (-6.0, 0.5, 1.5), list(pipe_n), r1=2.0, r2=3.0, l=pipe_l.value, vel=2.0, gamma=2.0),
          "pipe_frac": current(2, (-6.0, 0.5, 1.5), list(pipe_n), r1=1.0, r2=0.5, l=pipe_l.value, vel=-1.5, gamma=0.7)}
for name, c in fields.items():
    pts = [take(3, -8, 8) for _ in range(60)] + [[0.0, -3.0, 1.0], [0.3, -2.0, 1.05], [0.0, -5.0, 1.0], [0.0, 0.0, 1.2], [-6.0, 0.5, 1.5],
                                               [6.0, -0.5, 2.0], [0.0, 0.0, 1.75], [7.0, -0.6, 2.1]]
    if c.type != 0:  # points inside the cone / the pipe: axis point + small radial offset
        base, axis = np.array(list(c.a)), np.array(list(c.n))
        span = c.l if c.type == 2 else 6.0
        for _ in range(60):
            t, off = take(1, 0.05, 1.0)[0] * span, np.array(take(3, -0.6, 0.6))
            pts.append([float(x) for x in base + t * axis + off])
    for p in pts:
        o = arr(3)
        lib.btc_current_velocity(C.byref(c), (D * 3)(*p), o)
        out["currents"].append({"field": name, "p": hx(p), "v": hx(o)})
out["fields"] = {n: {"type": c.type, "a": hx(c.a), "n": hx(c.n), "r1": c.r1.hex(), "r2": c.r2.hex(), "l": c.l.hex(), "vel": c.vel.hex(),
                     "gamma": c.gamma.hex()} for n, c in fields.items()}

# --- transforms
for _ in range(12):
    y, p, r = take(3, -3.1, 3.1)
    org = take(3, -20, 20)
    o = arr(12)
    lib.btc_transform_from_euler(D(y), D(p), D(r), (D * 3)(*org), o)
    out["euler"].append({"ypr": hx([y, p, r]), "origin": hx(org), "T": hx(o)})
Ts = []
for _ in range(12):
    q = np.array(take(4))
    q /= np.linalg.norm(q)
    org = take(3, -20, 20)
    o = arr(12)
    lib.btc_transform_from_quat((D * 4)(*q), (D * 3)(*org), o)
    out["quat"].append({"q": hx(q), "origin": hx(org), "T": hx(o)})
    Ts.append(list(o))
for i in range(0, 12, 2):
    o = arr(12)
    lib.btc_transform_mul((D * 12)(*Ts[i]), (D * 12)(*Ts[i + 1]), o)
    out["mul"].append({"A": hx(Ts[i]), "B": hx(Ts[i + 1]), "AB": hx(o)})

# --- glMatrixFromTransform narrowing
for T in Ts[:6]:
    m = (C.c_float * 16)()
    lib.btc_gl_matrix((D * 12)(*T), m)
    out["gl_matrix"].append({"T": hx(T), "M": fhx(m)})

# --- CorrectHydrodynamicForces, incl. a zero vector (safeNormalize fallback)
for i in range(10):
    vecs = [take(3, -50, 50) for _ in range(4)]
    if i == 3:
        vecs[1] = [0.0, 0.0, 0.0]
    Cd, Cf = take(3, 0.2, 1.5), take(3, 0.01, 0.2)
    B = Ts[i][:9]
    bufs = [(D * 3)(*v) for v in vecs]
    lib.btc_correct(D(1025.0), *bufs, (D * 3)(*Cd), (D * 3)(*Cf), (D * 9)(*B))
    out["correct"].append({"density": (1025.0).hex(), "in": [hx(v) for v in vecs], "Cd": hx(Cd), "Cf": hx(Cf), "B": hx(B),
                           "out": [hx(b) for b in bufs]})

# --- analytic buoyancy tails
for i in range(8):
    pcb, vol = take(3, -0.2, 0.2), take(1, 0.01, 3.0)[0]
    Fb, Tb = arr(3), arr(3)
    lib.btc_inside_buoyancy((D * 12)(*Ts[i]), (D * 3)(*pcb), D(vol), D(1000.0), D(9.81), Fb, Tb)
    out["inside_buoyancy"].append({"T_CG": hx(Ts[i]), "P_CB": hx(pcb), "volume": vol.hex(), "Fb": hx(Fb), "Tb": hx(Tb)})
    cb = take(3, -20, 20)
    lib.btc_flat_buoyancy((D * 12)(*Ts[i]), (D * 3)(*cb), D(vol), D(1000.0), D(9.81), Fb, Tb)
    out["flat_buoyancy"].append({"T_CG": hx(Ts[i]), "CB": hx(cb), "Vsub": vol.hex(), "Fb": hx(Fb), "Tb": hx(Tb)})

# --- CorrectAerodynamicForces for every GeometryApproxType (SolidEntity.cpp:1925-1956), incl. a zero force and the
# ELLIPSOID coefficient vector on its own (for the host mirrors)
out["correct_aero"], out["ellipsoid_cd"] = [], []
for i in range(12):
    F, T = take(3, -30, 30), take(3, -30, 30)
    if i == 5:
        F = [0.0, 0.0, 0.0]
    kind = i % 4
    params = take(3, 0.05, 2.0)
    B = Ts[i][:9]
    bF, bT = (D * 3)(*F), (D * 3)(*T)
    lib.btc_correct_aero(bF, bT, C.c_int(kind), (D * 3)(*params), (D * 9)(*B))
    out["correct_aero"].append({"type": kind, "params": hx(params), "B": hx(B), "in": [hx(F), hx(T)], "out": [hx(bF), hx(bT)]})
    cd = arr(3)
    lib.btc_ellipsoid_cd((D * 3)(*params), cd)
    out["ellipsoid_cd"].append({"params": hx(params), "cd": hx(cd)})
# --- CableEntity::ComputeHydrodynamicForces on a flat sea with a uniform current + a jet; node lines that reach every branch
# of the submerged-volume case analysis (CableEntity.cpp:271-316)
out["cables"] = []
cab_ocean = _abi.sfh_ocean()
cab_ocean.density, cab_ocean.gravity_z, cab_ocean.currents_enabled, cab_ocean.n_currents = 1025.0, 9.81, 1, 2
cab_ocean.currents[0] = fields["uniform"]
cab_ocean.currents[1] = fields["jet"]
r = 0.05
lines = {
    "deep_inclined": [(0.1 * i, 0.05 * i, 2.0 + 0.07 * i) for i in range(6)],
    "dry": [(0.1 * i, 0.0, -1.0 - 0.05 * i) for i in range(4)],
    "piercing": [(0.2 * i, -0.1 * i, -0.5 + 0.3 * i) for i in range(7)],
    "horizontal_awash": [(0.3 * i, 0.1 * i, 0.02) for i in range(5)],
    "vertical": [(1.0, 2.0, -0.3 + 0.25 * i) for i in range(5)],
    "both_caps_awash": [(0.5 * i, 0.0, -0.01 + 0.012 * i) for i in range(4)],
    "second_cap_awash": [(0.0, 0.0, -0.2), (0.4, 0.1, 0.03), (0.8, 0.2, 0.031)],
    "first_cap_awash": [(0.0, 0.0, -0.02), (0.3, 0.2, 0.3), (0.5, 0.2, 0.9)],
    "in_the_jet": [(0.02 * i, -2.8 + 0.3 * i, 1.0 + 0.01 * i) for i in range(6)],
}
for name, nodes in lines.items():
    for flags in (7, 1, 6):
        pos = [c for nd in nodes for c in nd]
        vel = take(3 * len(nodes), -1.5, 1.5)
        res = arr(9 * len(nodes))
        lib.btc_cable(C.byref(cab_ocean), C.c_uint32(len(nodes)), D(r), C.c_uint32(flags), (D * len(pos))(*pos), (D * len(vel))(*vel), res)
        out["cables"].append({"name": name, "radius": r.hex(), "flags": flags, "pos": hx(pos), "vel": hx(vel), "out": hx(res)})
out["cable_ocean"] = {"density": (1025.0).hex(), "gravity_z": (9.81).hex(), "fields": ["uniform", "jet"]}
assert k < len(u)

path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bt_golden.json")
json.dump(out, open(path, "w"), indent=0)
print("wrote", path, {k: len(v) for k, v in out.items() if isinstance(v, list)})
