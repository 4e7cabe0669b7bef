"""Parity of the cable path (sfh_cable_step) against the oracle's restatement of
CableEntity::ComputeHydrodynamicForces and against the vectors produced with the reference's own Bullet.
fp64 throughout; the geometry goes through atan/cos/sin/acos whose last bit differs between CUDA and glibc, so the bar is
1e-12 of the cable's largest force component, stated here; everything else (branch selection, zeros) is exact."""
import json
import os

import numpy as np
import pytest

from stonefish_b200 import scenes
from stonefish_b200.hydro import Jet, Ocean, Pipe, SfhError, Uniform, wave_params
from tests.test_oracle_cable import G, golden_ocean, unhex

pytestmark = pytest.mark.gpu
TOL = 1e-12


def close(got, ref):
    scale = max(np.abs(ref).max(), 1e-300)
    assert np.abs(got - ref).max() <= TOL * scale, np.abs(got - ref).max() / scale
    assert np.array_equal(got == 0, ref == 0)  # dry segments / disabled terms are exactly zero on both sides


def test_golden_cables_in_one_batch(ctx):
    o = golden_ocean()
    cables, pos, vel, want = [], [], [], []
    for rec in G["cables"]:
        p = np.array(unhex(rec["pos"])).reshape(-1, 3)
        cables.append((sum(len(q) for q in pos), len(p), float.fromhex(rec["radius"]), rec["flags"]))
        pos.append(p)
        vel.append(np.array(unhex(rec["vel"])).reshape(-1, 3))
        want.append(np.array(unhex(rec["out"])).reshape(-1, 9))
    got = ctx.compute_cables(o, cables, np.concatenate(pos), np.concatenate(vel))
    for (first, n, _, _), w, rec in zip(cables, want, G["cables"]):
        close(got[first:first + n], w)


def test_long_cables_in_waves_and_currents(ctx, oracle):
    ocean = Ocean(1025.0, 9.81)
    ocean.setWaves(scenes.wave_field())
    ocean.AddVelocityField(Uniform((0.3, -0.1, 0.02)))
    ocean.AddVelocityField(Jet((0.0, -3.0, 1.0), (0.1, 1.0, 0.05), 0.3, 5.0))
    ocean.AddVelocityField(Pipe((-6.0, 0.5, 1.5), (6.0, -0.5, 2.0), 2.0, 3.0, 2.0, 2.0))
    ocean.EnableCurrents()
    rng = np.random.default_rng(5)
    cables, pos, vel = [], [], []
    first = 7  # some nodes belong to no cable: they must come back as zeros
    pos.append(rng.uniform(-1, 1, (first, 3)))
    vel.append(rng.uniform(-1, 1, (first, 3)))
    for n, r, flags in ((300, 0.01, 7), (2, 0.05, 7), (129, 0.03, 3), (64, 0.02, 5)):
        s = np.linspace(0, 1, n)
        line = np.stack([-4 + 8 * s + 0.05 * rng.standard_normal(n), -3 + 5 * s, -0.6 + 2.5 * s + 0.3 * np.sin(9 * s)], axis=1)
        cables.append((first, n, r, flags))
        pos.append(line)
        vel.append(rng.uniform(-0.8, 0.8, (n, 3)))
        first += n
    P, V = np.concatenate(pos), np.concatenate(vel)
    got = ctx.compute_cables(ocean, cables, P, V)
    assert not got[:7].any()
    for f, n, r, flags in cables:
        ref = oracle.cable(ocean, r, flags, P[f:f + n], V[f:f + n])
        close(got[f:f + n], ref)
        assert np.abs(ref).max() > 0
    # the generated wave field works for cables too
    ctx.wavegen_init(wave_params(1.0))
    ctx.wavegen_simulate(3.0)
    fft = ctx.wavegen_download()
    gen = Ocean(1025.0, 9.81)
    gen.useGeneratedWaves(True)
    a = ctx.compute_cables(gen, cables[:1], P, V)
    up = Ocean(1025.0, 9.81)
    up.setWaves(fft)
    assert np.array_equal(a, ctx.compute_cables(up, cables[:1], P, V))
    f, n, r, flags = cables[0]
    close(a[f:f + n], oracle.cable(up, r, flags, P[f:f + n], V[f:f + n]))


def test_cable_argument_checks(ctx):
    o = Ocean()
    z = np.zeros((4, 3))
    with pytest.raises(SfhError):
        ctx.compute_cables(o, [(0, 1, 0.01, 7)], z, z)  # one node
    with pytest.raises(SfhError):
        ctx.compute_cables(o, [(2, 3, 0.01, 7)], z, z)  # runs past the node array
    with pytest.raises(SfhError):
        ctx.compute_cables(o, [(0, 4, 0.0, 7)], z, z)  # radius
    assert not ctx.compute_cables(o, [], z, z).any()
