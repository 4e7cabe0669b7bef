// sfh_wavegen.cuh — the wave-field PRODUCER on CUDA: what OpenGLOcean / OpenGLRealOcean compute with
// fragment shaders and read back through a PBO (OpenGLOcean.cpp:455-534, OpenGLRealOcean.cpp:333-356) so that headless and
// batched runs have waves without OpenGL.
//
//   host, once   wavegen_spectrum   OpenGLOcean::GenerateWavesSpectrum (OpenGLOcean.cpp:929-1047): 4 nested grids x N^2 samples
//                wavegen_butterfly  ComputeButterflyLookupTable (:873-927)
//                                   both stored the way the reference's GL_RGBA16F textures hold them (:84-85,419-422):
//                                   rounded to half precision (nearest even), kept as floats on the device
//   device, per simulated time t
//     k_wave_rows<<<(N, 3), N>>>    oceanInit.frag for one row of one layer straight into shared memory, then the 8
//                                   oceanFFTX.frag passes ping-ponging in shared memory
//     k_wave_cols<<<(N, 3), N>>>    the 8 oceanFFTY.frag passes of one column; writes fftData [4][N][N][4] (layer 3 = 0)
//   then k_waves_planar (sfh_kernels.cuh) turns layer 0 into the planar height planes the per-face kernels read.
//
// Arithmetic: fp32, every operation rounded once and nothing fused, in the shader's expression order; sin/cos are CUDA's
// sinf/cosf (the reference's are whatever the OpenGL driver provides — see oracle/wave_oracle.cpp for what parity means here).
#pragma once

#include <cmath>
#include <cstdint>
#include <vector>

#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "../../include/sfh_types.h"

namespace sfh {

// ------------------------------------------------------------------------------------------------ host side
inline float half_storage(float x) { return __half2float(__float2half_rn(x)); } // what a GL_RGBA16F texel returns

// Elfouhaily-style directional spectrum of OpenGLOcean::spectrum (:954-1003). Everything that does not depend on the wave
// vector is evaluated once per parameter set (the reference re-evaluates it for every sample; same values).
struct WaveSpectrum {
    float A, km, cm, Omega, kp, cp, u_star, gamma, sigma, alphap, alpham, a0, am;
    bool propagate;

    static float sqr(float x) { return x * x; }
    float omega(float k) const { return std::sqrt(9.81 * k * (1.0 + sqr(k / km))); }

    explicit WaveSpectrum(const sfh_wave_params& p) : A(p.A), km(p.km), cm(p.cm), Omega(p.omega), propagate(p.propagate != 0)
    {
        const float U10 = p.wind;
        kp = 9.81 * sqr(Omega / U10);
        cp = omega(kp) / kp;
        const float z0 = 3.7e-5 * sqr(U10) / 9.81 * std::pow(U10 / cp, 0.9f);
        u_star = 0.41 * U10 / std::log(10.0 / z0);
        gamma = Omega < 1.0 ? 1.7 : 1.7 + 6.0 * std::log(Omega);
        sigma = 0.08 * (1.0 + 4.0 / std::pow(Omega, 3.0f));
        alphap = 0.006 * std::sqrt(Omega);
        alpham = 0.01 * (u_star < cm ? 1.0 + std::log(u_star / cm) : 1.0 + 3.0 * std::log(u_star / cm));
        a0 = std::log(2.0) / 4.0;
        am = 0.13 * u_star / cm;
    }

    float operator()(float kx, float ky) const
    {
        if (propagate && kx < 0.0) return 0.0; // the reference decides this after the k-dependent terms; they have no side effects
        const float k = std::sqrt(kx * kx + ky * ky);
        const float c = omega(k) / k;
        const float Lpm = std::exp(-5.0 / 4.0 * sqr(kp / k));
        const float Gamma = std::exp(-1.0 / (2.0 * sqr(sigma)) * sqr(std::sqrt(k / kp) - 1.0));
        const float Jp = std::pow(gamma, Gamma);
        const float Fp = Lpm * Jp * std::exp(-Omega / std::sqrt(10.0) * (std::sqrt(k / kp) - 1.0));
        float Bl = 0.5 * alphap * cp / c * Fp;
        const float Fm = std::exp(-0.25 * sqr(k / km - 1.0));
        float Bh = 0.5 * alpham * cm / c * Fm;
        Bh *= Lpm;
        const float ap = 4.0;
        const float Delta = std::tanh(a0 + ap * std::pow(c / cp, 2.5f) + am * std::pow(cm / c, 2.5f));
        const float phi = std::atan2(ky, kx);
        if (propagate) { Bl *= 2.0; Bh *= 2.0; }
        return A * (Bl + Bh) * (1.0 + Delta * std::cos(2.0 * phi)) / (2.0 * M_PI * sqr(sqr(k)));
    }
};

// spectrum12 / spectrum34 as generated (fp32, N*N*4 each); the random phases come from the reference's LCG in its order:
// texel by texel, grid 1..4, one draw per sample that is not inside the previous grid's band (SystemUtil.hpp:208-218)
inline void wavegen_spectrum(const sfh_wave_params& p, std::vector<float>& s12, std::vector<float>& s34)
{
    const int N = 1 << p.passes;
    const WaveSpectrum S(p);
    s12.assign((size_t)N * N * 4, 0.f);
    s34.assign((size_t)N * N * 4, 0.f);
    long seed = (long)p.seed;
    const float kmin[4] = {(float)(M_PI / p.grid_sizes[0]), (float)(M_PI * N / p.grid_sizes[0]), (float)(M_PI * N / p.grid_sizes[1]),
                           (float)(M_PI * N / p.grid_sizes[2])};
    float dk[4];
    for (int g = 0; g < 4; ++g) dk[g] = 2.0 * M_PI / p.grid_sizes[g];
    for (int y = 0; y < N; ++y)
        for (int x = 0; x < N; ++x) {
            const int i = x >= N / 2 ? x - N : x, j = y >= N / 2 ? y - N : y;
            for (int g = 0; g < 4; ++g) {
                float* out = (g < 2 ? s12.data() : s34.data()) + 4 * ((size_t)x + (size_t)y * N) + 2 * (g & 1);
                const float kx = i * dk[g], ky = j * dk[g];
                if (fabsf(kx) < kmin[g] && fabsf(ky) < kmin[g]) continue; // band already covered by the coarser grid: zero
                const float h = sqrtf(S(kx, ky) / 2.0) * dk[g];
                seed = (seed * 1103515245 + 12345) & 0x7FFFFFFF;
                const float phi = ((seed >> 7) / (float)(1 << 24)) * 2.0 * M_PI;
                out[0] = h * std::cos(phi);
                out[1] = h * std::sin(phi);
            }
        }
}

// Radix-2 decimation-in-time passes: for output index x of pass i, the two source indices and the twiddle (:873-927).
inline void wavegen_butterfly(uint32_t passes, std::vector<int2>& src, std::vector<float2>& w)
{
    const uint32_t N = 1u << passes;
    src.assign((size_t)N * passes, int2{0, 0});
    w.assign((size_t)N * passes, float2{0.f, 0.f});
    for (uint32_t i = 0; i < passes; ++i) {
        const uint32_t half = 1u << i, blocks = N >> (i + 1);
        for (uint32_t j = 0; j < blocks; ++j)
            for (uint32_t k = 0; k < half; ++k) {
                const uint32_t i1 = j * half * 2 + k, i2 = i1 + half;
                uint32_t j1 = i1, j2 = i2;
                if (i == 0) { // first pass reads in bit-reversed order
                    j1 = j2 = 0;
                    for (uint32_t b = 0; b < passes; ++b) {
                        j1 |= ((i1 >> b) & 1u) << (passes - 1 - b);
                        j2 |= ((i2 >> b) & 1u) << (passes - 1 - b);
                    }
                }
                const float wr = half_storage((float)cosl(2.f * M_PI * (int)(k * blocks) / float(N)));
                const float wi = half_storage((float)sinl(2.f * M_PI * (int)(k * blocks) / float(N)));
                src[i1 + (size_t)i * N] = int2{(int)j1, (int)j2};
                src[i2 + (size_t)i * N] = int2{(int)j1, (int)j2};
                w[i1 + (size_t)i * N] = float2{wr, wi};
                w[i2 + (size_t)i * N] = float2{-wr, -wi};
            }
    }
}

// ------------------------------------------------------------------------------------------------ device side
struct WaveGenDev {
    const float4* s12;  // N*N, half-rounded spectrum of grids 1,2 (re, im, re, im)
    const float4* s34;  // grids 3,4
    const int2* bf_src; // passes * N
    const float2* bf_w; // passes * N
    float4* tmp;        // 3 * N * N, after the X passes
    float4* fft;        // 4 * N * N, fftData
    int32_t N, passes;
    float inv_grid[4];  // 2*pi*N / gridSizes[g] (OpenGLOcean.cpp:472-475)
};

// out = input1 + w * input2 on both complex pairs of a texel (oceanFFTX.frag:37-47)
__device__ __forceinline__ float4 wave_butterfly(float4 in1, float4 in2, float2 w)
{
    const float r1x = __fsub_rn(__fmul_rn(w.x, in2.x), __fmul_rn(w.y, in2.y));
    const float r1y = __fadd_rn(__fmul_rn(w.y, in2.x), __fmul_rn(w.x, in2.y));
    const float r2x = __fsub_rn(__fmul_rn(w.x, in2.z), __fmul_rn(w.y, in2.w));
    const float r2y = __fadd_rn(__fmul_rn(w.y, in2.z), __fmul_rn(w.x, in2.w));
    return make_float4(__fadd_rn(in1.x, r1x), __fadd_rn(in1.y, r1y), __fadd_rn(in1.z, r2x), __fadd_rn(in1.w, r2y));
}

// `passes` butterfly passes over the N texels held by the block; returns the buffer (0/1) that holds the result
__device__ __forceinline__ int wave_fft_passes(const WaveGenDev& G, float4* buf0, float4* buf1, uint32_t idx)
{
    int cur = 0;
    for (int p = 0; p < G.passes; ++p) {
        __syncthreads();
        const int2 s = G.bf_src[idx + (size_t)p * G.N];
        const float2 w = G.bf_w[idx + (size_t)p * G.N];
        const float4* in = cur ? buf1 : buf0;
        float4* out = cur ? buf0 : buf1;
        out[idx] = wave_butterfly(in[s.x], in[s.y], w);
        cur ^= 1;
    }
    __syncthreads();
    return cur;
}

// h(k,t) of oceanInit.frag:38-44 for one grid
__device__ __forceinline__ float2 wave_h(float K, float t, float2 s0, float2 s0c)
{
    const float w = __fsqrt_rn(__fmul_rn(__fmul_rn(9.81f, K), __fadd_rn(1.0f, __fdiv_rn(__fmul_rn(K, K), 370.0f * 370.0f))));
    const float wt = __fmul_rn(w, t);
    const float c = cosf(wt), s = sinf(wt);
    return make_float2(__fsub_rn(__fmul_rn(__fadd_rn(s0.x, s0c.x), c), __fmul_rn(__fadd_rn(s0.y, s0c.y), s)),
                       __fadd_rn(__fmul_rn(__fsub_rn(s0.x, s0c.x), s), __fmul_rn(__fsub_rn(s0.y, s0c.y), c)));
}

// block (row py, layer): thread px evaluates oceanInit.frag's output `layer` for texel (px, py), then the X passes
__global__ void k_wave_rows(WaveGenDev G, float t)
{
    extern __shared__ float4 s_wave[]; // 2 * N
    const int N = G.N, px = threadIdx.x, py = blockIdx.x, layer = blockIdx.y;
    const float fN = (float)N;
    // st = floor(texcoord * fftSize) / fftSize = px / N; x = texcoord.x > 0.5 ? st.x - 1 : st.x
    const float x = px >= N / 2 ? __fsub_rn(__fdiv_rn((float)px, fN), 1.0f) : __fdiv_rn((float)px, fN);
    const float y = py >= N / 2 ? __fsub_rn(__fdiv_rn((float)py, fN), 1.0f) : __fdiv_rn((float)py, fN);
    // NEAREST + repeat: h0(k) at texel (px,py), its mirror at ((N-px)%N, (N-py)%N)
    const size_t here = (size_t)px + (size_t)py * N, mirror = (size_t)((N - px) % N) + (size_t)((N - py) % N) * N;
    const float4 sa = layer < 2 ? G.s12[here] : G.s34[here], sb = layer < 2 ? G.s12[mirror] : G.s34[mirror];
    const float4 sc12 = G.s12[mirror], s12 = G.s12[here];
    const float r2 = 1.414213562f;
    float4 out;
    if (layer == 0) { // vec4(h1 + i(h2), h3 + i(h4))
        const float4 s34 = G.s34[here], sc34 = G.s34[mirror];
        float2 h[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const float kx = __fmul_rn(x, G.inv_grid[g]), ky = __fmul_rn(y, G.inv_grid[g]);
            const float K = __fsqrt_rn(__fadd_rn(__fmul_rn(kx, kx), __fmul_rn(ky, ky)));
            const float4 a = g < 2 ? s12 : s34, b = g < 2 ? sc12 : sc34;
            const float2 s0 = (g & 1) ? make_float2(__fmul_rn(a.z, r2), __fmul_rn(a.w, r2)) : make_float2(__fmul_rn(a.x, r2), __fmul_rn(a.y, r2));
            const float2 s0c = (g & 1) ? make_float2(__fmul_rn(b.z, r2), __fmul_rn(b.w, r2)) : make_float2(__fmul_rn(b.x, r2), __fmul_rn(b.y, r2));
            h[g] = wave_h(K, t, s0, s0c);
        }
        out = make_float4(__fadd_rn(h[0].x, -h[1].y), __fadd_rn(h[0].y, h[1].x), __fadd_rn(h[2].x, -h[3].y), __fadd_rn(h[2].y, h[3].x));
    } else { // vec4(i(kA.x * hA) - (kA.y * hA), i(kB.x * hB) - (kB.y * hB)), (A,B) = grids (1,2) or (3,4)
        float o[4];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int g = 2 * (layer - 1) + q;
            const float kx = __fmul_rn(x, G.inv_grid[g]), ky = __fmul_rn(y, G.inv_grid[g]);
            const float K = __fsqrt_rn(__fadd_rn(__fmul_rn(kx, kx), __fmul_rn(ky, ky)));
            const float2 s0 = q ? make_float2(__fmul_rn(sa.z, r2), __fmul_rn(sa.w, r2)) : make_float2(__fmul_rn(sa.x, r2), __fmul_rn(sa.y, r2));
            const float2 s0c = q ? make_float2(__fmul_rn(sb.z, r2), __fmul_rn(sb.w, r2)) : make_float2(__fmul_rn(sb.x, r2), __fmul_rn(sb.y, r2));
            const float2 h = wave_h(K, t, s0, s0c);
            o[2 * q] = __fsub_rn(-__fmul_rn(kx, h.y), __fmul_rn(ky, h.x));
            o[2 * q + 1] = __fsub_rn(__fmul_rn(kx, h.x), __fmul_rn(ky, h.y));
        }
        out = make_float4(o[0], o[1], o[2], o[3]);
    }
    float4* buf0 = s_wave;
    float4* buf1 = s_wave + N;
    buf0[px] = out;
    const int cur = wave_fft_passes(G, buf0, buf1, (uint32_t)px);
    G.tmp[(size_t)layer * N * N + here] = (cur ? buf1 : buf0)[px];
}

// block (column px, layer): the Y passes; also clears layer 3, which the reference never renders into
__global__ void k_wave_cols(WaveGenDev G)
{
    extern __shared__ float4 s_wave[];
    const int N = G.N, py = threadIdx.x, px = blockIdx.x, layer = blockIdx.y;
    const size_t here = (size_t)px + (size_t)py * N;
    float4* buf0 = s_wave;
    float4* buf1 = s_wave + N;
    buf0[py] = G.tmp[(size_t)layer * N * N + here];
    const int cur = wave_fft_passes(G, buf0, buf1, (uint32_t)py);
    G.fft[(size_t)layer * N * N + here] = (cur ? buf1 : buf0)[py];
    if (layer == 0) G.fft[(size_t)3 * N * N + here] = make_float4(0.f, 0.f, 0.f, 0.f);
}

} // namespace sfh
