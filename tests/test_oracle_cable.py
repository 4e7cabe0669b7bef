"""Cables in the oracle: CableEntity::ComputeHydrodynamicForces pinned bit-for-bit against the same
statements evaluated with the reference's own Bullet LinearMath (tests/golden/bt_golden.json:cables), plus hand-checkable cases."""
import json
import math
import os

import numpy as np

from stonefish_b200.hydro import Jet, Ocean, Uniform

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "bt_golden.json")))


def unhex(v):
    return [float.fromhex(x) for x in v]


def golden_ocean():
    o = Ocean(float.fromhex(G["cable_ocean"]["density"]), float.fromhex(G["cable_ocean"]["gravity_z"]))
    fu, fj = G["fields"]["uniform"], G["fields"]["jet"]
    o.AddVelocityField(Uniform(unhex(fu["a"])))
    jet = Jet(unhex(fj["a"]), unhex(fj["n"]), float.fromhex(fj["r1"]), float.fromhex(fj["vel"]))
    jet.n = np.array(unhex(fj["n"]))  # the stored axis is already normalised: keep its exact bits
    o.AddVelocityField(jet)
    o.EnableCurrents()
    return o


def test_cable_forces_bit_exact_against_bullet(oracle):
    o = golden_ocean()
    names = set()
    for rec in G["cables"]:
        got = oracle.cable(o, float.fromhex(rec["radius"]), rec["flags"], unhex(rec["pos"]), unhex(rec["vel"]))
        want = np.array(unhex(rec["out"])).reshape(-1, 9)
        assert np.array_equal(got, want), rec["name"]
        names.add(rec["name"])
        if rec["flags"] == 1:
            assert not got[:, :3].any()  # no buoyancy without reallisticBuoyancy / phy_.buoyancy
        if rec["flags"] == 6:
            assert not got[:, 3:].any()
    assert len(names) == 9


def test_submerged_straight_cable_displaces_its_cylinder(oracle):
    """Fully submerged cable: every segment displaces pi r^2 L; nodes share it half and half, end nodes are doubled
    (CableEntity.cpp:363-371 — the reference's "correction", kept)."""
    r, n, L = 0.03, 6, 0.5
    pos = np.array([(L * i, 0.0, 5.0) for i in range(n)])
    f = oracle.cable(Ocean(1000.0, 9.81), r, 7, pos, np.zeros((n, 3)))
    seg = math.pi * r * r * L * 1000.0 * 9.81
    assert np.allclose(f[1:-1, 2], -seg, rtol=1e-14)  # two half segments
    assert np.allclose(f[[0, -1], 2], -seg, rtol=1e-14)  # one half segment, doubled
    assert not f[:, 3:].any()  # at rest in still water


def test_circular_segment_area(oracle):
    oracle.cable(Ocean(), 0.1, 0, np.zeros((2, 3)), np.zeros((2, 3)))  # binds the symbols
    area = oracle.lib().orc_circular_segment_area
    r = 0.2
    assert area(r, -0.1) == 0.0 and area(r, 0.5) == math.pi * r * r
    assert abs(area(r, r) - math.pi * r * r / 2) < 1e-15
    assert abs(area(r, 0.05) + area(r, 2 * r - 0.05) - math.pi * r * r) < 1e-15


def test_drag_opposes_relative_motion(oracle):
    n = 5
    pos = np.array([(0.4 * i, 0.0, 3.0) for i in range(n)])
    vel = np.tile([0.0, 0.2, 0.0], (n, 1))  # moving across the cable axis: f1 = |segmentVec . relV| = 0
    f = oracle.cable(Ocean(1000.0, 9.81), 0.02, 1, pos, vel)
    assert (f[:, 4] < 0).all() and (f[:, 7] < 0).all() and not f[:, [3, 5, 6, 8]].any()
    o = Ocean(1000.0, 9.81)
    o.AddVelocityField(Uniform((0.0, 0.2, 0.0)))
    o.EnableCurrents()
    assert not oracle.cable(o, 0.02, 1, pos, vel)[:, 3:].any()  # drifting with the current
