"""Rows a3-a8 of the scope table in isolation: Ocean::GetDepth (wave lookup) and Ocean::GetFluidVelocity
(Uniform/Jet/Pipe) evaluated on the device for batches of points, against the oracle point by point."""
import ctypes as C

import numpy as np
import pytest

from stonefish_b200 import Jet, Ocean, Pipe, Uniform, meshes, scenes

pytestmark = pytest.mark.gpu


def _points(seed, n, span):
    u = meshes.splitmix_uniform(seed, 3 * n).reshape(n, 3)
    return ((2 * u - 1) * span).astype(np.float32)


def _oracle_depth(oracle, ocean, pts):
    o, w = ocean.to_struct(), ocean.waves_struct()
    f = oracle.lib().orc_get_depth
    return np.array([f(C.byref(o), C.byref(w), float(p[0]), float(p[1]), float(p[2])) for p in pts], dtype=np.float32)


@pytest.mark.parametrize("span", [3.0, 60.0, 2500.0])
def test_get_depth_waves_bit_exact(ctx, oracle, span):
    ocean = Ocean()
    ocean.setWaves(scenes.wave_field())
    pts = _points(101 + int(span), 20000, np.array([span, span, 2.0]))
    # texel centres, texel edges, the origin, exact grid multiples and tiny negatives (i0f -> 1.0f fix-up path)
    special = []
    for g in (893.0, 101.0):
        for i in (0, 1, 127, 255, 256, -1, -256):
            special += [((i + 0.5) / 256 * g, (i + 0.5) / 256 * g, 0.1), (i / 256 * g, -i / 256 * g, 0.0), (i * g, -i * g, 0.2)]
    special += [(0, 0, 0), (-1e-7, -1e-7, 0), (1e-7, -1e-9, 0), (-1e-12, 1e-12, 0.5), (1.744140625, 0.197265625, 0)]
    pts = np.concatenate([pts, np.array(special, dtype=np.float32)])
    got = ctx.GetDepth(ocean, pts)
    ref = _oracle_depth(oracle, ocean, pts)
    bad = np.nonzero(got.view(np.uint32) != ref.view(np.uint32))[0]
    assert bad.size == 0, (bad[:5], pts[bad[:5]], got[bad[:5]], ref[bad[:5]])


def test_get_depth_flat_is_z(ctx):
    pts = _points(5, 1000, 50.0)
    assert np.array_equal(ctx.GetDepth(Ocean(), pts), pts[:, 2])


def test_wave_lookup_known_answers(ctx):
    """At a texel centre the GL-style bilinear filter returns that texel: depth = z + ch0[j,i] + ch1[j',i']."""
    field = np.zeros((4, 256, 256, 4), np.float32)
    field[0, :, :, 0] = np.arange(256 * 256, dtype=np.float32).reshape(256, 256) / 1024.0
    ocean = Ocean()
    ocean.setWaves(field, 256, (256.0, 256.0, 21.0, 11.0))  # 1 m per texel on both used layers
    ij = np.array([(3, 5), (0, 0), (255, 255), (17, 200)])
    pts = np.array([(i + 0.5, j + 0.5, 1.0) for i, j in ij], dtype=np.float32)
    got = ctx.GetDepth(ocean, pts)
    want = np.array([1.0 + field[0, j, i, 0] for i, j in ij], dtype=np.float32)  # z - (0 - h0 - 0)
    assert np.array_equal(got, want)
    # periodic wrap: one whole grid away reads the same texel
    assert np.array_equal(ctx.GetDepth(ocean, pts + np.float32([256.0, -256.0, 0.0])), want)


def _ocean_with(*fields):
    o = Ocean()
    for f in fields:
        o.AddVelocityField(f)
    o.EnableCurrents()
    return o


def _oracle_velocity(oracle, ocean, pts):
    o = ocean.to_struct()
    f = oracle.lib().orc_fluid_velocity_f
    out = np.zeros((len(pts), 3), np.float32)
    for k, p in enumerate(pts):
        pin = (C.c_float * 3)(*[float(t) for t in p])
        pout = (C.c_float * 3)()
        f(C.byref(o), pin, pout)
        out[k] = pout[:]
    return out


def test_fluid_velocity_uniform_and_disabled(ctx, oracle):
    pts = _points(9, 2000, 30.0)
    a, b = Uniform((0.5, -0.3, 0.0)), Uniform((0.1, 0.2, 0.3))
    off = Uniform((7, 7, 7))
    off.setEnabled(False)
    ocean = _ocean_with(a, off, b)
    got = ctx.GetFluidVelocity(ocean, pts)
    assert np.array_equal(got, _oracle_velocity(oracle, ocean, pts))
    assert np.array_equal(got[0], np.float32(np.array([0.5, -0.3, 0.0]) + np.array([0.1, 0.2, 0.3])))
    ocean.DisableCurrents()
    assert not ctx.GetFluidVelocity(ocean, pts).any()


def test_fluid_velocity_jet_pipe(ctx, oracle):
    jet = Jet((0.0, 0.0, 1.0), (0.0, 1.0, 0.0), 0.3, 5.0)  # UnderwaterTestManager.cpp:117
    pipe = Pipe((-20.0, 0.0, 2.0), (20.0, 0.0, 2.0), 3.0, 5.0, 2.0, 2.0)
    ocean = _ocean_with(Uniform((1.0, 0.0, 0.0)), jet, pipe)
    u = meshes.splitmix_uniform(77, 3 * 30000).reshape(-1, 3)
    pts = np.stack([-25 + 50 * u[:, 0], -2 + 12 * u[:, 1], -4 + 10 * u[:, 2]], axis=1).astype(np.float32)
    # axis points, outlet plane (t == 0), behind the outlet, pipe ends
    pts = np.concatenate([pts, np.float32([(0, 0, 1), (0, 1, 1), (0, -1, 1), (0.05, 0, 1), (-20, 0, 2), (20, 0, 2), (0, 0, 2), (25, 0, 2)])])
    got = ctx.GetFluidVelocity(ocean, pts)
    ref = _oracle_velocity(oracle, ocean, pts)
    inside = (np.abs(ref - np.float32([1, 0, 0])).max(axis=1) > 0).sum()
    assert inside > 1000  # the sample really exercises both fields
    # exp/pow: CUDA <= 1-2 ulp vs glibc; after the fp64 -> fp32 narrowing at most a handful of 1-ulp differences
    both_nan = np.isnan(got) & np.isnan(ref)
    diff = (got.view(np.int32).astype(np.int64) - ref.view(np.int32).astype(np.int64))
    diff[both_nan] = 0
    assert np.abs(diff).max() <= 1
    assert (diff != 0).sum() <= 3


def test_non_finite_coordinates_stay_in_bounds(ctx):
    """NaN / Inf / huge coordinates: the reference's (int) cast is undefined there; here the lookup must simply stay inside
    the wave planes (cvt.rzi turns NaN into index 0) and return something without faulting."""
    from stonefish_b200 import scenes
    from stonefish_b200.hydro import Ocean

    ocean = Ocean()
    ocean.setWaves(scenes.wave_field())
    pts = np.array([[np.nan, 0, 0], [0, np.nan, 1], [np.inf, 1, 0], [1, -np.inf, 0], [3e38, -3e38, 0], [1e30, 1e-30, 2], [0, 0, np.nan]], dtype=np.float32)
    d = ctx.GetDepth(ocean, pts)
    assert d.shape == (7,)
    assert np.isfinite(d[4]) and np.isfinite(d[5])
    # the context is still healthy
    assert np.isfinite(ctx.GetDepth(ocean, np.zeros((1, 3), np.float32)))[0]
