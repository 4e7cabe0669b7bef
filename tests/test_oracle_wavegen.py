"""The wave-field producer in the oracle (oracle/wave_oracle.cpp): the restated butterfly passes are an
unnormalised inverse DFT of h(k,t), checked against numpy's FFT on an independent float64 model of oceanInit.frag; the
GL_RGBA16F storage emulation against numpy's float16; the spectrum's structural properties."""
import numpy as np
import pytest

from stonefish_b200.hydro import wave_params

N = 256


@pytest.fixture(scope="module")
def wave(oracle):
    return oracle.WaveOracle(wave_params(1.0))


def model_h(w, t):
    """float64 model of oceanInit.frag: h_g(k,t) for the four grids and their wave vectors."""
    p = w.params
    r2 = np.float64(np.float32(1.414213562))
    s12 = w.spectrum12.reshape(N, N, 4).astype(np.float64) * r2
    s34 = w.spectrum34.reshape(N, N, 4).astype(np.float64) * r2
    idx = (-np.arange(N)) % N
    n = np.where(np.arange(N) >= N // 2, np.arange(N) - N, np.arange(N)).astype(np.float64)
    X, Y = np.meshgrid(n, n)
    hs, ks = [], []
    for g, (S, c0) in enumerate([(s12, 0), (s12, 2), (s34, 0), (s34, 2)]):
        Sc = S[idx][:, idx]
        inv = 2 * np.pi / p.grid_sizes[g]
        kx, ky = X * inv, Y * inv
        K = np.hypot(kx, ky)
        om = np.sqrt(9.81 * K * (1 + K * K / 370.0 ** 2))
        c, s = np.cos(om * t), np.sin(om * t)
        hx = (S[..., c0] + Sc[..., c0]) * c - (S[..., c0 + 1] + Sc[..., c0 + 1]) * s
        hy = (S[..., c0] - Sc[..., c0]) * s + (S[..., c0 + 1] - Sc[..., c0 + 1]) * c
        hs.append(hx + 1j * hy)
        ks.append((kx, ky))
    return hs, ks


def test_half_storage_matches_numpy_float16(oracle, wave):
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(20000) * 10.0 ** rng.integers(-9, 5, 20000)).astype(np.float32)
    x = np.concatenate([x, np.array([0, -0.0, 65504, 65519.9, 65520, 1e-8, 5.96e-8, 2.98e-8, 2.9802322e-8, 6.1e-5, -6.1e-5, 8.9e-8], np.float32)])
    ours = np.array([oracle.lib().orc_half_round(float(v)) for v in x], dtype=np.float32)
    with np.errstate(over="ignore"):
        ref = x.astype(np.float16).astype(np.float32)
    assert np.array_equal(ours, ref)


def test_passes_are_an_inverse_dft_of_h(wave):
    t = 0.7
    f = wave.simulate(t, half_storage=False).astype(np.float64)
    hs, ks = model_h(wave, t)
    for layer, pairs in ((0, [(hs[0] + 1j * hs[1]), (hs[2] + 1j * hs[3])]),
                         (1, [(1j * ks[0][0] - ks[0][1]) * hs[0], (1j * ks[1][0] - ks[1][1]) * hs[1]]),
                         (2, [(1j * ks[2][0] - ks[2][1]) * hs[2], (1j * ks[3][0] - ks[3][1]) * hs[3]])):
        for q, Z in enumerate(pairs):
            want = np.fft.ifft2(Z) * N * N  # w = exp(+2 pi i k / N), no 1/N: Tessendorf's sum, eq. 19
            got = f[layer, ..., 2 * q] + 1j * f[layer, ..., 2 * q + 1]
            assert np.abs(got - want).max() <= 2e-6 * np.abs(want).max(), (layer, q)
    assert not f[3].any()  # the reference never renders into layer 3


def test_heights_are_real_fields_packed_two_per_texel(wave):
    """h_g(-k) = conj(h_g(k)) by construction, so each grid's height field is real and channel 0 / 1 of layer 0 ARE the heights
    of grids 1 / 2 — what Ocean::GetDepth reads (OpenGLRealOcean.cpp:325-331)."""
    hs, _ = model_h(wave, 1.3)
    for h in hs:
        z = np.fft.ifft2(h) * N * N
        assert np.abs(z.imag).max() <= 1e-9 * max(np.abs(z.real).max(), 1e-30)
    f = wave.simulate(1.3, half_storage=False)
    z1 = (np.fft.ifft2(hs[0]) * N * N).real
    assert np.abs(f[0, ..., 0] - z1).max() <= 2e-6 * np.abs(z1).max()
    assert 0.02 < f[0, ..., 0].std() < 0.5  # sea state 1: decimetre waves


def test_half_precision_textures_cost_about_1e_3(wave):
    full, half = wave.simulate(0.7, half_storage=False), wave.simulate(0.7, half_storage=True)
    rel = np.abs(full[0] - half[0]).max() / np.abs(full[0]).max()
    assert 1e-5 < rel < 3e-3


def test_spectrum_structure(oracle, wave):
    s12 = wave.spectrum12.reshape(N, N, 4)
    s34 = wave.spectrum34.reshape(N, N, 4)
    # one-sided (propagate): kx < 0 carries nothing
    assert not s12[:, N // 2:, :].any() and not s34[:, N // 2:, :].any()
    # grid g+1 starts where grid g ends: samples with |kx|,|ky| below pi*N/L_g are zero (OpenGLOcean.cpp:1011-1015)
    dk2, kmin2 = np.float32(2 * np.pi / 101.0), np.float32(np.pi * N / 893.0)
    n = np.where(np.arange(N) >= N // 2, np.arange(N) - N, np.arange(N)).astype(np.float32)
    inside = (np.abs(n[None, :] * dk2) < kmin2) & (np.abs(n[:, None] * dk2) < kmin2)
    assert not s12[..., 2][inside].any() and s12[..., 2][~inside & (n[None, :] >= 0)].any()
    assert s12[0, 0, 0] == 0 and s12[0, 0, 1] == 0  # k = 0
    # deterministic in the seed, different for another seed
    again = oracle.WaveOracle(wave_params(1.0))
    other = oracle.WaveOracle(wave_params(1.0, seed=99))
    assert np.array_equal(again.spectrum12, wave.spectrum12) and not np.array_equal(other.spectrum12, wave.spectrum12)
    # amplitude |h0| does not depend on the random phase
    assert np.allclose(np.hypot(other.spectrum12[0::2], other.spectrum12[1::2]), np.hypot(wave.spectrum12[0::2], wave.spectrum12[1::2]),
                       rtol=1e-5, atol=1e-12)
    # rougher sea, bigger waves
    rough = oracle.WaveOracle(wave_params(3.0))
    assert rough.simulate(0.0)[0, ..., 0].std() > 3 * wave.simulate(0.0)[0, ..., 0].std()


def test_butterfly_table(oracle, wave):
    bt = wave.butterfly.reshape(8, N, 4)
    # texel-centre coordinates, exact in half precision
    j = bt[..., :2] * N - 0.5
    assert np.array_equal(j, np.round(j)) and j.min() == 0 and j.max() == N - 1
    h = np.array([oracle.lib().orc_half_round(float(v)) for v in bt[..., :2].reshape(-1)], dtype=np.float32)
    assert np.array_equal(h, bt[..., :2].reshape(-1))
    # pass 0 reads bit-reversed pairs with w = +-1
    assert np.array_equal(np.abs(bt[0, :, 2]), np.ones(N)) and not bt[0, :, 3].any()
    rev = np.array([int(f"{i:08b}"[::-1], 2) for i in range(N)])
    assert np.array_equal(j[0, 0::2, 0], rev[0::2]) and np.array_equal(j[0, 0::2, 1], rev[1::2])
    # last pass: w_k = exp(2 pi i k / N)
    k = np.arange(N // 2)
    assert np.allclose(bt[7, :N // 2, 2], np.cos(2 * np.pi * k / N), atol=1e-7) and np.allclose(bt[7, :N // 2, 3], np.sin(2 * np.pi * k / N), atol=1e-7)
