/*
 * wave_oracle.cpp — CPU restatement of the reference's wave-field PRODUCER, the part that fills
 * OpenGLRealOcean::fftData: spectrum generation (Library/src/graphics/OpenGLOcean.cpp:929-1047), butterfly table
 * (:857-927), the h(k,t) pass (Library/shaders/oceanInit.frag) and the 2 x 8 FFT passes (oceanFFTX.frag / oceanFFTY.frag)
 * as OpenGLOcean::Simulate sequences them (:455-534).
 *
 * TEST INFRASTRUCTURE ONLY (see hydro_oracle.h). PARITY, two halves:
 *   - host half (spectrum, LCG phases, butterfly table): PINNED bit for bit against the reference's own definitions compiled from
 *     its source text (oracle/ref_harness.cpp: OpenGLOcean::GenerateWavesSpectrum / spectrum / GetSpectrumSample /
 *     ComputeButterflyLookupTable, lrandom / frandom; tests/test_oracle_refsource.py::test_wave_producer_host_tables). Unqualified
 *     sqrt/pow/log/cos... on float arguments resolve to the float overloads there as here (Bullet's btScalar.h includes <math.h>).
 *   - shader half (oceanInit.frag, oceanFFTX/Y.frag): UNPINNED, and necessarily so: the reference evaluates it on whatever GPU
 *     runs OpenGL, where sin/cos precision and a*b+c contraction are implementation-defined (GLSL 3.30 §4.5), and it stores
 *     spectrum and butterfly table in GL_RGBA16F textures (OpenGLOcean.cpp:84-85,419-422) whose float->half rounding is also the
 *     driver's. This file fixes those choices to: round-to-nearest-even halves, libm sinf/cosf, no contraction.
 */
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "hydro_oracle.h"

namespace {

/* utils/SystemUtil.hpp:208-218 */
inline long lrandom(long* seed)
{
    *seed = (*seed * 1103515245 + 12345) & 0x7FFFFFFF;
    return *seed;
}
inline float frandom(long* seed)
{
    long r = lrandom(seed) >> (31 - 24);
    return r / (float)(1 << 24);
}

inline float sqr(float x) { return x * x; }

struct Spec {
    sfh_wave_params p;
    float omega(float k) const { return std::sqrt(9.81 * k * (1.0 + sqr(k / p.km))); } /* :948-951 */

    /* OpenGLOcean::spectrum (:954-1003), omnispectrum = false */
    float spectrum(float kx, float ky) const
    {
        float U10 = p.wind;
        float Omega = p.omega;
        float k = std::sqrt(kx * kx + ky * ky);
        float c = omega(k) / k;
        float kp = 9.81 * sqr(Omega / U10);
        float cp = omega(kp) / kp;
        float z0 = 3.7e-5 * sqr(U10) / 9.81 * std::pow(U10 / cp, 0.9f);
        float u_star = 0.41 * U10 / std::log(10.0 / z0);
        float Lpm = std::exp(-5.0 / 4.0 * sqr(kp / k));
        float gamma = Omega < 1.0 ? 1.7 : 1.7 + 6.0 * std::log(Omega);
        float sigma = 0.08 * (1.0 + 4.0 / std::pow(Omega, 3.0f));
        float Gamma = std::exp(-1.0 / (2.0 * sqr(sigma)) * sqr(std::sqrt(k / kp) - 1.0));
        float Jp = std::pow(gamma, Gamma);
        float Fp = Lpm * Jp * std::exp(-Omega / std::sqrt(10.0) * (std::sqrt(k / kp) - 1.0));
        float alphap = 0.006 * std::sqrt(Omega);
        float Bl = 0.5 * alphap * cp / c * Fp;
        float alpham = 0.01 * (u_star < p.cm ? 1.0 + std::log(u_star / p.cm) : 1.0 + 3.0 * std::log(u_star / p.cm));
        float Fm = std::exp(-0.25 * sqr(k / p.km - 1.0));
        float Bh = 0.5 * alpham * p.cm / c * Fm;
        Bh *= Lpm;
        float a0 = std::log(2.0) / 4.0;
        float ap = 4.0;
        float am = 0.13 * u_star / p.cm;
        float Delta = std::tanh(a0 + ap * std::pow(c / cp, 2.5f) + am * std::pow(p.cm / c, 2.5f));
        float phi = std::atan2(ky, kx);
        if (p.propagate) {
            if (kx < 0.0) return 0.0;
            Bl *= 2.0;
            Bh *= 2.0;
        }
        float tweak = 1.0f; /* :1000-1001 */
        return p.A * (Bl + Bh) * (1.0 + Delta * std::cos(2.0 * phi)) / (2.0 * M_PI * sqr(sqr(k))) * tweak;
    }

    /* OpenGLOcean::GetSpectrumSample (:1006-1022) */
    void sample(int i, int j, float lengthScale, float kMin, long* seed, float* result) const
    {
        float dk = 2.0 * M_PI / lengthScale;
        float kx = i * dk;
        float ky = j * dk;
        if (fabsf(kx) < kMin && fabsf(ky) < kMin) {
            result[0] = 0.0;
            result[1] = 0.0;
        } else {
            float S = spectrum(kx, ky);
            float h = sqrtf(S / 2.0) * dk;
            float phi = frandom(seed) * 2.0 * M_PI;
            result[0] = h * std::cos(phi);
            result[1] = h * std::sin(phi);
        }
    }
};

/* OpenGLOcean::bitReverse (:857-871) */
int bit_reverse(int i, int N)
{
    int j = i, M = N, Sum = 0, W = 1;
    M = M / 2;
    while (M != 0) {
        j = (i & M) > M - 1;
        Sum += j * W;
        W *= 2;
        M = M / 2;
    }
    return Sum;
}

} // namespace

extern "C" {

/* GL_RGBA16F storage of a float (round to nearest even, subnormal halves kept, overflow to infinity) read back as float */
float orc_half_round(float x)
{
    uint32_t u;
    std::memcpy(&u, &x, 4);
    const uint32_t sign = u & 0x80000000u;
    uint32_t a = u & 0x7fffffffu;
    float r;
    if (a >= 0x7f800000u) { /* inf / nan */
        r = x;
        return r;
    }
    if (a >= 0x477ff000u) { /* >= 65520: rounds to half infinity */
        uint32_t inf = sign | 0x7f800000u;
        std::memcpy(&r, &inf, 4);
        return r;
    }
    if (a < 0x38800000u) { /* below the smallest normal half 2^-14: quantum 2^-24 */
        float m;
        std::memcpy(&m, &a, 4);
        /* adding 2^-1 forces rounding at the 2^-24 position (binary32 has 24 bits): (m + 0.5f) - 0.5f, RNE */
        volatile float t = m + 0.5f;
        float q = t - 0.5f;
        uint32_t qb;
        std::memcpy(&qb, &q, 4);
        qb |= sign;
        std::memcpy(&r, &qb, 4);
        return r;
    }
    /* normal half: keep 10 mantissa bits, RNE on the 13 dropped bits */
    const uint32_t lsb = (a >> 13) & 1u;
    a += 0x0fffu + lsb;
    a &= ~0x1fffu;
    a |= sign;
    std::memcpy(&r, &a, 4);
    return r;
}

/* OpenGLOcean::GenerateWavesSpectrum (:1025-1047): spectrum12 / spectrum34, fftSize^2 * 4 floats each (fp32, as generated) */
void orc_wave_spectrum(const sfh_wave_params* p, float* spectrum12, float* spectrum34)
{
    Spec S{*p};
    const int N = 1 << p->passes;
    long seed = (long)p->seed;
    const float* g = p->grid_sizes;
    for (int y = 0; y < N; ++y) {
        for (int x = 0; x < N; ++x) {
            int offset = 4 * (x + y * N);
            int i = x >= N / 2 ? x - N : x;
            int j = y >= N / 2 ? y - N : y;
            S.sample(i, j, g[0], M_PI / g[0], &seed, spectrum12 + offset);
            S.sample(i, j, g[1], M_PI * N / g[0], &seed, spectrum12 + offset + 2);
            S.sample(i, j, g[2], M_PI * N / g[1], &seed, spectrum34 + offset);
            S.sample(i, j, g[3], M_PI * N / g[2], &seed, spectrum34 + offset + 2);
        }
    }
}

/* OpenGLOcean::ComputeButterflyLookupTable / computeWeight (:873-927): size * passes * 4 floats */
void orc_wave_butterfly(uint32_t size_, uint32_t passes, float* data)
{
    const int size = (int)size_;
    for (unsigned int i = 0; i < passes; ++i) {
        int nBlocks = (int)powf(2.0, float(passes - 1 - (int)i));
        int nHInputs = (int)powf(2.0, float(i));
        for (int j = 0; j < nBlocks; ++j) {
            for (int k = 0; k < nHInputs; ++k) {
                int i1 = j * nHInputs * 2 + k, i2 = j * nHInputs * 2 + nHInputs + k, j1, j2;
                if (i == 0) {
                    j1 = bit_reverse(i1, size);
                    j2 = bit_reverse(i2, size);
                } else {
                    j1 = i1;
                    j2 = i2;
                }
                float wr = cosl(2.f * M_PI * (k * nBlocks) / float(size));
                float wi = sinl(2.f * M_PI * (k * nBlocks) / float(size));
                int offset1 = 4 * (i1 + i * size);
                data[offset1 + 0] = (j1 + 0.5) / size;
                data[offset1 + 1] = (j2 + 0.5) / size;
                data[offset1 + 2] = wr;
                data[offset1 + 3] = wi;
                int offset2 = 4 * (i2 + i * size);
                data[offset2 + 0] = (j1 + 0.5) / size;
                data[offset2 + 1] = (j2 + 0.5) / size;
                data[offset2 + 2] = -wr;
                data[offset2 + 3] = -wi;
            }
        }
    }
}

/* OpenGLOcean::Simulate (:455-534) at time t: oceanInit.frag into layers 0..2, then passes FFT-X and passes FFT-Y over the
 * three layers, ping-ponging between two texture arrays; the result is what glGetTexImage reads into fftData
 * (OpenGLRealOcean.cpp:349-354): [4 layers][N][N][4] floats, layer 3 never written (zero here).
 * half_storage != 0: spectrum and butterfly table go through GL_RGBA16F as in the reference; 0: kept in fp32 (lets a test
 * compare the pass structure with a textbook DFT). */
void orc_wave_simulate(const sfh_wave_params* p, const float* spectrum12, const float* spectrum34, const float* butterfly, float t,
                       int half_storage, float* fft_data)
{
    const int N = 1 << p->passes, passes = (int)p->passes;
    const size_t L = (size_t)N * N * 4;
    auto tex = [&](const float* src, int x, int y, int c) {
        float v = src[4 * (x + (size_t)y * N) + c];
        return half_storage ? orc_half_round(v) : v;
    };
    std::vector<float> A(3 * L), B(3 * L);
    /* Simulate(): inverseGridSizes = 2.f*M_PI*fftSize/gridSizes[i], evaluated in double, stored in a glm::vec4 */
    float inv[4];
    for (int g = 0; g < 4; ++g) inv[g] = (float)(2.f * M_PI * (float)N / p->grid_sizes[g]);
    const float fftSize = (float)N;

    /* oceanInit.frag, one fragment per texel, texcoord at the pixel centre */
    for (int py = 0; py < N; ++py) {
        for (int px = 0; px < N; ++px) {
            const float tcx = (px + 0.5f) / fftSize, tcy = (py + 0.5f) / fftSize;
            const float stx = std::floor(tcx * fftSize) / fftSize, sty = std::floor(tcy * fftSize) / fftSize;
            const float x = tcx > 0.5f ? stx - 1.0f : stx;
            const float y = tcy > 0.5f ? sty - 1.0f : sty;
            /* NEAREST + repeat: texcoord -> texel (px,py); vec2(1.0 + 0.5/fftSize) - st -> texel ((N-px)%N, (N-py)%N) */
            const int cx = (N - px) % N, cy = (N - py) % N;
            float s[8], sc[8];
            for (int c = 0; c < 4; ++c) {
                s[c] = tex(spectrum12, px, py, c) * 1.414213562f;
                s[4 + c] = tex(spectrum34, px, py, c) * 1.414213562f;
                sc[c] = tex(spectrum12, cx, cy, c) * 1.414213562f;
                sc[4 + c] = tex(spectrum34, cx, cy, c) * 1.414213562f;
            }
            float h[4][2], kx[4], ky[4];
            for (int g = 0; g < 4; ++g) {
                kx[g] = x * inv[g];
                ky[g] = y * inv[g];
                const float K = std::sqrt(kx[g] * kx[g] + ky[g] * ky[g]);
                const float w = std::sqrt(9.81f * K * (1.0f + K * K / (370.0f * 370.0f)));
                const float c = cosf(w * t), sn = sinf(w * t);
                const float s0x = s[2 * g], s0y = s[2 * g + 1], s0cx = sc[2 * g], s0cy = sc[2 * g + 1];
                h[g][0] = (s0x + s0cx) * c - (s0y + s0cy) * sn;
                h[g][1] = (s0x - s0cx) * sn + (s0y - s0cy) * c;
            }
            float* o0 = &A[0 * L + 4 * (px + (size_t)py * N)];
            float* o1 = &A[1 * L + 4 * (px + (size_t)py * N)];
            float* o2 = &A[2 * L + 4 * (px + (size_t)py * N)];
            /* fragColor[0] = vec4(h1 + i(h2), h3 + i(h4)) with i(z) = (-z.y, z.x) */
            o0[0] = h[0][0] + -h[1][1]; o0[1] = h[0][1] + h[1][0];
            o0[2] = h[2][0] + -h[3][1]; o0[3] = h[2][1] + h[3][0];
            /* fragColor[1] = vec4(i(k1.x * h1) - (k1.y * h1), i(k2.x * h2) - (k2.y * h2)); fragColor[2]: grids 3, 4 */
            auto slope = [&](int g, float* o) {
                const float ax = kx[g] * h[g][0], ay = kx[g] * h[g][1];
                const float bx = ky[g] * h[g][0], by = ky[g] * h[g][1];
                o[0] = -ay - bx;
                o[1] = ax - by;
            };
            slope(0, o1); slope(1, o1 + 2);
            slope(2, o2); slope(3, o2 + 2);
        }
    }

    /* oceanFFTX.frag / oceanFFTY.frag: out = input1 + w * input2 on both complex pairs of a texel */
    auto bf = [&](int pass, int idx, int c) {
        float v = butterfly[4 * (idx + (size_t)pass * N) + c];
        return half_storage ? orc_half_round(v) : v;
    };
    std::vector<float>* src = &A;
    std::vector<float>* dst = &B;
    for (int pass = 0; pass < 2 * passes; ++pass) {
        const bool along_x = pass < passes;
        const int row = along_x ? pass : pass - passes;
        for (int layer = 0; layer < 3; ++layer) {
            for (int py = 0; py < N; ++py) {
                for (int px = 0; px < N; ++px) {
                    const int idx = along_x ? px : py;
                    /* data.xy are texel-centre coordinates (j + 0.5)/N, exact in half precision */
                    const int j1 = (int)(bf(row, idx, 0) * N), j2 = (int)(bf(row, idx, 1) * N);
                    const float wx = bf(row, idx, 2), wy = bf(row, idx, 3);
                    const float* in1 = along_x ? &(*src)[layer * L + 4 * (j1 + (size_t)py * N)] : &(*src)[layer * L + 4 * (px + (size_t)j1 * N)];
                    const float* in2 = along_x ? &(*src)[layer * L + 4 * (j2 + (size_t)py * N)] : &(*src)[layer * L + 4 * (px + (size_t)j2 * N)];
                    float* o = &(*dst)[layer * L + 4 * (px + (size_t)py * N)];
                    const float res1x = wx * in2[0] - wy * in2[1];
                    const float res1y = wy * in2[0] + wx * in2[1];
                    const float res2x = wx * in2[2] - wy * in2[3];
                    const float res2y = wy * in2[2] + wx * in2[3];
                    o[0] = in1[0] + res1x; o[1] = in1[1] + res1y; o[2] = in1[2] + res2x; o[3] = in1[3] + res2y;
                }
            }
        }
        std::swap(src, dst);
    }
    std::memcpy(fft_data, src->data(), 3 * L * sizeof(float));
    std::memset(fft_data + 3 * L, 0, L * sizeof(float));
}

} // extern "C"
