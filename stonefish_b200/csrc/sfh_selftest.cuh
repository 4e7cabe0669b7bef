// sfh_selftest.cuh — on-device proofs/checks of the hand-written IEEE sequences in sfh_packed.cuh against the
// hardware's correctly rounded div.rn.f32 / sqrt.rn.f32 and of the packed (f32x2) operators against the scalar
// unfused ones. Driven by sfh_selftest() (include/sfh.h) from tests/test_arith_gpu.py.
#pragma once

#include "sfh_packed.cuh"

namespace sfh {

__device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ bool same_bits(float a, float b)
{
    return __float_as_uint(a) == __float_as_uint(b) || (a != a && b != b);
}

// which = 0: ieee_sqrt1 over EVERY float in [2^-101, FLT_MAX]                      (n ignored)
// which = 1: ieee_div3_1 over EVERY float that is zero or has |x| >= 2^-100, finite  (n ignored)
// which = 2: ieee_div_r1 on n pseudo-random (a, b): b in [2^-20, 2^64], a in {0} U +-[2^-36, 2b]
// which = 4: ieee_div_r1 with UNRELATED magnitudes (coordinate / grid size in the wave lookup): b in [2^-20, 2^40],
//            a in +-[2^-36, 2^64], n pseudo-random pairs
// which = 3: packed operators vs scalar unfused operators on n pseudo-random operand pairs,
//            including ieee_sqrt2 / ieee_div3_2 / ieee_div_r2 against their scalar twins
__global__ void k_selftest(uint32_t which, uint64_t n, unsigned long long* mismatches, unsigned long long* tested)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t bad = 0, cnt = 0, first_bad = 0; // first_bad: an operand (index) of this thread that failed
    if (which == 0) {
        for (uint64_t u = 0x0d000000ull + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; u <= 0x7f7fffffull; u += stride) {
            const float x = __uint_as_float((uint32_t)u);
            if (!same_bits(ieee_sqrt1(x), __fsqrt_rn(x))) { ++bad; first_bad = u; }
            ++cnt;
        }
    } else if (which == 1) {
        for (uint64_t u = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; u < 0x100000000ull; u += stride) {
            const uint32_t m = (uint32_t)u & 0x7fffffffu;
            if (m >= 0x7f800000u || (m != 0 && m < 0x0d800000u)) continue;
            const float x = __uint_as_float((uint32_t)u);
            const float got = ieee_div3_1(x), want = __fdiv_rn(x, 3.f);
            // -0 / 3 is -0 in IEEE and +0 here (fma(+0, c, -0)); the sign of a zero never reaches a result
            if (!(same_bits(got, want) || (got == 0.f && want == 0.f))) { ++bad; first_bad = u; }
            ++cnt;
        }
    } else if (which == 2) {
        for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
            const uint64_t h = mix64(i), g = mix64(h ^ 0x5851F42D4C957F2Dull);
            // b: exponent uniform in [-20, 63], mantissa random with adversarial patterns every 8th sample
            uint32_t mb = (uint32_t)h & 0x7fffffu, ma = (uint32_t)(h >> 23) & 0x7fffffu;
            const uint32_t pat = (uint32_t)(g >> 40) & 7u;
            if (pat == 0) mb = 0x7fffffu; else if (pat == 1) mb = 0; else if (pat == 2) mb = 0x7ffffeu;
            if (((g >> 43) & 7u) == 0) ma = 0x7fffffu; else if (((g >> 43) & 7u) == 1) ma = 0;
            const int eb = -20 + (int)((g & 0xffff) % 84);
            const float b = __uint_as_float(((uint32_t)(eb + 127) << 23) | mb);
            // a: zero with probability 1/64, else exponent uniform in [-36, eb+1] clipped to |a| <= 2b
            float a = 0.f;
            if (((g >> 16) & 63u) != 0) {
                const int span = eb + 1 + 36 + 1;
                const int ea = -36 + (int)(((g >> 22) & 0xffff) % span);
                a = __uint_as_float(((uint32_t)(g >> 63) << 31) | ((uint32_t)(ea + 127) << 23) | ma);
                if (fabsf(a) > 2.f * b) a = __uint_as_float(((uint32_t)(g >> 63) << 31) | (__float_as_uint(b) & 0x7fffffffu));
            }
            const float got = ieee_div_r1(a, b, refined_rcp1(b));
            const float want = __fdiv_rn(a, b);
            // the shared form may return +0 where IEEE gives -0 (documented; zero sign never reaches a result)
            if (!(same_bits(got, want) || (got == 0.f && want == 0.f))) { ++bad; first_bad = i; }
            ++cnt;
        }
    } else if (which == 4) {
        for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
            const uint64_t h = mix64(i ^ 0x1234567ull), g = mix64(h ^ 0x5851F42D4C957F2Dull);
            uint32_t mb = (uint32_t)h & 0x7fffffu, ma = (uint32_t)(h >> 23) & 0x7fffffu;
            const uint32_t pat = (uint32_t)(g >> 40) & 7u;
            if (pat == 0) mb = 0x7fffffu; else if (pat == 1) mb = 0; else if (pat == 2) mb = 0x7ffffeu;
            if (((g >> 43) & 7u) == 0) ma = 0x7fffffu; else if (((g >> 43) & 7u) == 1) ma = 0;
            const int eb = -20 + (int)((g & 0xffff) % 61), ea = -36 + (int)(((g >> 22) & 0xffff) % 101);
            const float b = __uint_as_float(((uint32_t)(eb + 127) << 23) | mb);
            const float a = __uint_as_float(((uint32_t)(g >> 63) << 31) | ((uint32_t)(ea + 127) << 23) | ma);
            const float got = ieee_div_r1(a, b, refined_rcp1(b));
            const float want = __fdiv_rn(a, b);
            if (!same_bits(got, want)) { ++bad; first_bad = i; }
            ++cnt;
        }
    } else {
        for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
            const uint64_t h = mix64(i ^ 0xA5A5A5A5ull), g = mix64(h);
            auto rnd = [&](uint64_t bits, int emin, int emax) {
                const int e = emin + (int)((bits >> 24) % (uint64_t)(emax - emin + 1));
                return __uint_as_float(((uint32_t)(bits >> 63) << 31) | ((uint32_t)(e + 127) << 23) | ((uint32_t)bits & 0x7fffffu));
            };
            const float a0 = rnd(h, -20, 20), a1 = rnd(mix64(h + 1), -20, 20), b0 = rnd(g, -20, 20), b1 = rnd(mix64(g + 1), -20, 20);
            const float c0 = rnd(mix64(h + 2), -20, 20), c1 = rnd(mix64(g + 2), -20, 20);
            // always the packed f32x2 carrier, whatever the build's default pair type is
            const PairX A = PairX::make(a0, a1), Bv = PairX::make(b0, b1), Cv = PairX::make(c0, c1);
            const PairX s = A + Bv, d = A - Bv, m = A * Bv;
            // the patterns ptxas would like to contract: product feeding a sum / a difference
            const PairX ms = A * Bv + Cv, md = A * Bv - Cv * A, dt = (A * A + Bv * Bv) + Cv * Cv;
            bool ok = same_bits(lo(s), __fadd_rn(a0, b0)) && same_bits(hi(s), __fadd_rn(a1, b1));
            ok &= same_bits(lo(d), __fsub_rn(a0, b0)) && same_bits(hi(d), __fsub_rn(a1, b1));
            ok &= same_bits(lo(m), __fmul_rn(a0, b0)) && same_bits(hi(m), __fmul_rn(a1, b1));
            ok &= same_bits(lo(ms), __fadd_rn(__fmul_rn(a0, b0), c0)) && same_bits(hi(ms), __fadd_rn(__fmul_rn(a1, b1), c1));
            ok &= same_bits(lo(md), __fsub_rn(__fmul_rn(a0, b0), __fmul_rn(c0, a0))) && same_bits(hi(md), __fsub_rn(__fmul_rn(a1, b1), __fmul_rn(c1, a1)));
            ok &= same_bits(lo(dt), __fadd_rn(__fadd_rn(__fmul_rn(a0, a0), __fmul_rn(b0, b0)), __fmul_rn(c0, c0)));
            ok &= same_bits(hi(dt), __fadd_rn(__fadd_rn(__fmul_rn(a1, a1), __fmul_rn(b1, b1)), __fmul_rn(c1, c1)));
            const PairX x2 = PairX::make(fabsf(a0), fabsf(a1));
            const PairX sq = ieee_sqrt2(x2), d3 = ieee_div3_2(A);
            ok &= same_bits(lo(sq), __fsqrt_rn(fabsf(a0))) && same_bits(hi(sq), __fsqrt_rn(fabsf(a1)));
            ok &= same_bits(lo(d3), __fdiv_rn(a0, 3.f)) && same_bits(hi(d3), __fdiv_rn(a1, 3.f));
            const PairX len = PairX::make(fabsf(b0), fabsf(b1)), nlen = neg(len);
            const PairX q = ieee_div_r2(A, nlen, refined_rcp2(len, nlen));
            if (fabsf(a0) <= 2.f * fabsf(b0) && fabsf(a1) <= 2.f * fabsf(b1))
                ok &= same_bits(lo(q), __fdiv_rn(a0, fabsf(b0))) && same_bits(hi(q), __fdiv_rn(a1, fabsf(b1)));
            if (!ok) { ++bad; first_bad = i; }
            ++cnt;
        }
    }
    if (bad) atomicMax(mismatches + 2, (unsigned long long)first_bad | (1ull << 63));
    bad = __reduce_add_sync(0xffffffffu, (uint32_t)bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(mismatches, (unsigned long long)bad);
    // cnt can exceed 32 bits per warp only for absurd n; accumulate per thread
    atomicAdd(tested, (unsigned long long)cnt);
}

} // namespace sfh
