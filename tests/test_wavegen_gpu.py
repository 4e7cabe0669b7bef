"""Parity of the device wave-field producer (sfh_wavegen_*) against the oracle's restatement of
OpenGLOcean::GenerateWavesSpectrum / ComputeButterflyLookupTable / Simulate and the three shaders.
Bar: spectrum bit-exact; at t = 0 (cos = 1, sin = 0 in every implementation) the whole fftData bit-exact, i.e. the 16 FFT passes
and the h(k,t) algebra reproduce the oracle to the last bit; at t > 0 the only difference is sinf/cosf (CUDA vs glibc, <= 2 ulp),
bounded at 2e-6 of each layer's amplitude. The field then drives the hydrodynamic step like an uploaded fftData."""
import numpy as np
import pytest

from stonefish_b200 import HydroContext, scenes
from stonefish_b200.hydro import SfhError, wave_params
from tests.parity import REL_TOL, assert_counts_equal, max_rel_err

pytestmark = pytest.mark.gpu
TOL_TRIG = 2e-6


@pytest.mark.parametrize("state,seed", [(1.0, 1234), (3.0, 77)])
def test_spectrum_and_t0_field_bit_exact(ctx, oracle, state, seed):
    p = wave_params(state, seed)
    ctx.wavegen_init(p)
    w = oracle.WaveOracle(p)
    s12, s34 = ctx.wavegen_spectrum()
    assert np.array_equal(s12, w.spectrum12) and np.array_equal(s34, w.spectrum34)
    ctx.wavegen_simulate(0.0)
    got, ref = ctx.wavegen_download(), w.simulate(0.0)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    assert got[0].std() > 0 and not got[3].any()


@pytest.mark.parametrize("t", [0.016, 1.0, 37.5, 3600.0])
def test_field_at_time_t_within_trig_tolerance(ctx, oracle, t):
    p = wave_params(2.0)
    ctx.wavegen_init(p)
    w = oracle.WaveOracle(p)
    ctx.wavegen_simulate(t)
    got, ref = ctx.wavegen_download().astype(np.float64), w.simulate(t).astype(np.float64)
    for layer in range(3):
        err = np.abs(got[layer] - ref[layer]).max() / np.abs(ref[layer]).max()
        assert err <= TOL_TRIG, (layer, err)
    assert not got[3].any()
    # the field moves: it is not the t = 0 field
    assert np.abs(ref[0] - w.simulate(0.0)[0]).max() > 1e-3 * np.abs(ref[0]).max()


def test_other_fft_sizes(ctx, oracle):
    for passes in (5, 9):
        p = wave_params(1.5)
        p.passes = passes
        ctx.wavegen_init(p)
        w = oracle.WaveOracle(p)
        ctx.wavegen_simulate(0.0)
        assert np.array_equal(ctx.wavegen_download().view(np.uint32), w.simulate(0.0).view(np.uint32))
        ctx.wavegen_simulate(2.5)
        got, ref = ctx.wavegen_download().astype(np.float64), w.simulate(2.5).astype(np.float64)
        assert np.abs(got[:3] - ref[:3]).max() <= TOL_TRIG * np.abs(ref[0]).max()


def test_generated_field_drives_the_hydrodynamic_step(ctx, oracle):
    """A step on the device-resident generated field == a step on the same fftData uploaded from the host (bitwise), and both
    match the oracle evaluated on that fftData."""
    p = wave_params(2.0)
    ctx.wavegen_init(p)
    ctx.wavegen_simulate(12.25)
    fft = ctx.wavegen_download()
    sc = scenes.mixed_small(7, True, "uniform")
    states = sc.bind(ctx)
    sc.ocean.setWaves(None)
    sc.ocean.useGeneratedWaves(True)
    a = ctx.compute(sc.ocean, states)
    sc.ocean.useGeneratedWaves(False)
    sc.ocean.setWaves(fft)
    b = ctx.compute(sc.ocean, states)
    assert a.tobytes() == b.tobytes()
    ref, _ = oracle.OracleScene(sc.meshes, sc.body_mesh, sc.ocean, sc.states).step(1, n_threads=4)
    assert_counts_equal(a, ref)
    err, where = max_rel_err(a, ref)
    assert err <= REL_TOL, (err, where)
    assert (a["position"] == 2).any()
    # waves matter: the flat-sea answer differs
    sc.ocean.setWaves(None)
    flat = ctx.compute(sc.ocean, states)
    assert not np.array_equal(flat["Fb"], a["Fb"])


def test_wavegen_argument_checks():
    with HydroContext(0) as c:
        with pytest.raises(SfhError):
            c.wavegen_simulate(0.0)  # before init
        p = wave_params(1.0)
        p.passes = 11
        with pytest.raises(SfhError):
            c.wavegen_init(p)
        p = wave_params(1.0)
        p.grid_sizes[2] = 0.0
        with pytest.raises(SfhError):
            c.wavegen_init(p)
