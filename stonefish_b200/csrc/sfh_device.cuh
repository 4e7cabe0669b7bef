// sfh_device.cuh — device-side arithmetic of the hydrodynamics hot path (sm_100a).
//
// Every fp32/fp64 operation that feeds a branch or a sum goes through an explicit round-to-nearest
// intrinsic (__fmul_rn, __fadd_rn, __fdiv_rn, __fsqrt_rn, __dmul_rn, ...): nvcc never contracts those
// into FMAs, so the per-face terms are bit-identical to the reference's x86-64 build, which has no FMA
// in its baseline ISA (hard part 1). Denormals are kept (no -ftz).
#pragma once

#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/sfh_types.h"
#include "mesh_chunker.hpp"

namespace sfh {

// ---------------------------------------------------------------------------------------------
// device-side data
// ---------------------------------------------------------------------------------------------
struct MeshDev {
    const uint8_t* chunks;     // packed chunk stream (mesh_chunker.hpp)
    const uint32_t* chunk_off; // n_chunks + 1 offsets into `chunks`, 16-byte units
    const uint32_t* orig_face; // n_chunks * SFH_CH_FACES (debug only)
    const ChunkBounds* bounds; // n_chunks boxes + stream positions (the chunk classifier of the surface path)
    uint32_t n_chunks;
    uint32_t n_faces;
};

#define SFH_MESH_TOMBSTONE 0xffffffffu /* MeshDev::n_faces of a released mesh id */

enum { KIND_NONE = 0, KIND_SUBMERGED = 1, KIND_SURFACE = 2 };
enum { BF_DAMPING = 1u, BF_BUOYANCY = 2u };

// Per-body record prepared by the prologue kernel; everything the per-face kernels need, already narrowed to
// fp32 exactly like SolidEntity.cpp:1325-1334 / 1740-1746 (glMatrixFromTransform, glVectorFromVector).
struct BodyDev {
    float R[9];      // T_C basis, row-major: world = (R[0]*x + R[1]*y) + (R[2]*z + t[0]) ...
    float t[3];      // T_C origin
    float p[3];      // T_CG origin (torque reference point)
    float v[3];
    float omega[3];
    float cur[3];    // Ocean::GetFluidVelocity result when it is position independent (no Jet/Pipe enabled)
    uint32_t flags;  // BF_*
    uint32_t kind;   // KIND_*
    uint32_t n_chunks;
    uint32_t n_faces;
    const uint8_t* chunks;     // 16 bytes read together by the chunk prefetcher:
    uint32_t first_bytes;      //   chunks pointer + size of chunk 0
    uint32_t pad;
    const uint32_t* chunk_off;
    const uint32_t* orig_face;
    const ChunkBounds* bounds;
    uint64_t pad2;
};
static_assert(sizeof(BodyDev) % 16 == 0 && offsetof(BodyDev, chunks) % 16 == 0, "the chunk prefetcher reads {chunks, first_bytes} as one uint4");

// Ocean parameters staged on the device once per step.
struct OceanDev {
    double density, gravity_z;
    uint32_t has_waves, currents_enabled, n_currents, cur_general; // cur_general: some enabled current is Jet/Pipe
    sfh_current currents[SFH_MAX_CURRENTS];
};

struct WavesDev {
    const float* ch0; // planar copy of fftData channel 0 (grid 893 m), index j*N+i, N*N+N+1 entries
    const float* ch1; // channel 1 (grid 101 m)
    int32_t N;
    float gx, gy;     // gridSizes.x, gridSizes.y
    const uint32_t* hmax; // device: float bits of max|ch0|, max|ch1|; their sum bounds |wave height| anywhere
};

// Partial sums written by one thread block for one body.
#define SFH_NACC_SUB 12
#define SFH_NACC_SURF 23
struct PartialSub {
    double a[SFH_NACC_SUB]; // Fdq[3] Tdq[3] Fdf[3] Tdf[3]
    uint32_t n_degenerate, n_quadratic, n_skin, pad;
};
struct PartialSurf {
    double a[SFH_NACC_SURF]; // Fdq Tdq Fdf Tdf | Fb[3] Tb[3] | Swet Vsub | CBsub[3]
    double pad0;
    uint32_t cls[8];         // cls[0] is not counted: class 0 = the body's faces that are in no other class
    uint32_t n_degenerate, n_quadratic, n_skin, pad;
};

// Programmatic dependent launch (griddepcontrol): the kernels of a step are launched with programmatic stream serialization, so
// the next kernel's blocks are scheduled — launch latency, shared-memory set-up — while its predecessor still runs. pdl_wait()
// returns when the predecessor grid has COMPLETED and its writes are visible; every kernel of the chain executes it on every path
// before it reads anything a kernel of the same step wrote (and before it exits: completion of a grid that never waited would say
// nothing about ITS predecessor). Without the launch attribute both are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// fp32 chain: glm::vec3 semantics, one rounding per operator, nothing fused
// ---------------------------------------------------------------------------------------------
struct F3 {
    float x, y, z;
};
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float fsqrt(float a) { return __fsqrt_rn(a); }

__device__ __forceinline__ F3 operator+(F3 a, F3 b) { return {fadd(a.x, b.x), fadd(a.y, b.y), fadd(a.z, b.z)}; }
__device__ __forceinline__ F3 operator-(F3 a, F3 b) { return {fsub(a.x, b.x), fsub(a.y, b.y), fsub(a.z, b.z)}; }
__device__ __forceinline__ F3 operator-(F3 a) { return {-a.x, -a.y, -a.z}; }
__device__ __forceinline__ F3 operator*(F3 a, float s) { return {fmul(a.x, s), fmul(a.y, s), fmul(a.z, s)}; }
__device__ __forceinline__ F3 operator/(F3 a, float s) { return {fdiv(a.x, s), fdiv(a.y, s), fdiv(a.z, s)}; }
__device__ __forceinline__ float dot(F3 a, F3 b) { return fadd(fadd(fmul(a.x, b.x), fmul(a.y, b.y)), fmul(a.z, b.z)); }
__device__ __forceinline__ float length2(F3 a) { return dot(a, a); }
__device__ __forceinline__ F3 cross(F3 x, F3 y)
{
    return {fsub(fmul(x.y, y.z), fmul(y.y, x.z)), fsub(fmul(x.z, y.x), fmul(y.z, x.x)), fsub(fmul(x.x, y.y), fmul(y.x, x.y))};
}

// glm::vec3(TC * glm::vec4(p, 1.f)) = (m0*x + m1*y) + (m2*z + m3*1)
__device__ __forceinline__ F3 xform(const float* R, const float* t, float x, float y, float z)
{
    F3 r;
    r.x = fadd(fadd(fmul(R[0], x), fmul(R[1], y)), fadd(fmul(R[2], z), t[0]));
    r.y = fadd(fadd(fmul(R[3], x), fmul(R[4], y)), fadd(fmul(R[5], z), t[1]));
    r.z = fadd(fadd(fmul(R[6], x), fmul(R[7], y)), fadd(fmul(R[8], z), t[2]));
    return r;
}

// ---------------------------------------------------------------------------------------------
// Scalar IEEE sequences (see sfh_packed.cuh for the two-lane twins and sfh_selftest for their validation): correctly
// rounded results, identical to __fsqrt_rn / __fdiv_rn inside the stated domains, in fewer dispatched instructions.
// ---------------------------------------------------------------------------------------------
#define SFH_THIRD 0.3333333432674407958984375f /* RN(1/3) = 0x3eaaaaab */
// sqrt: x in [2^-101, FLT_MAX]
__device__ __forceinline__ float ieee_sqrt1(float x)
{
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    const float s = __fmul_rn(x, r), h = __fmul_rn(r, 0.5f);
    const float e = __fmaf_rn(-s, s, x);
    return __fmaf_rn(e, h, s);
}

// x / 3.f: x == 0 or 2^-100 <= |x| <= FLT_MAX
__device__ __forceinline__ float ieee_div3_1(float x)
{
    const float q = __fmul_rn(x, SFH_THIRD);
    const float rem = __fmaf_rn(-3.f, q, x);
    return __fmaf_rn(rem, SFH_THIRD, q);
}

// a / b through a refined reciprocal r of b, shared by several numerators: b in [2^-20, 2^64] and a == 0 or |a| in
// [2^-36, 2b] (selftest 2), or b in [2^-20, 2^40] and |a| in [2^-36, 2^64) (selftest 4)
__device__ __forceinline__ float refined_rcp1(float b)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
    const float e = __fmaf_rn(-b, r, 1.f);
    return __fmaf_rn(r, e, r);
}
__device__ __forceinline__ float ieee_div_r1(float a, float b, float r)
{
    const float q = __fmul_rn(a, r);
    const float rem = __fmaf_rn(-b, q, a);
    return __fmaf_rn(r, rem, q);
}

// ---- domain guards. key(x) = 2*|x|_bits - 1 (mod 2^32): 0xffffffff for +-0, monotone in |x| otherwise, so
// "some operand is nonzero and smaller than a threshold" is one unsigned min and one compare.
__device__ __forceinline__ uint32_t tiny_key(float x)
{
    const uint32_t b = __float_as_uint(x);
    return b + b - 1u;
}
#define SFH_TINY_KEY 0x1affffffu /* 2*bits(2^-100) - 1 */

#define SFH_NUM_KEY 0x5affffffu  /* 2*bits(2^-36) - 1 */
#define SFH_KEY_SPAN 0x63ffffffu /* keys of |a| in [2^-36, 2^64) are SFH_NUM_KEY + [0, SFH_KEY_SPAN] */

// x / 3.f and a / b with the fast forms inside their proven domains, the hardware division outside
__device__ __forceinline__ float div3(float x) { return tiny_key(x) >= SFH_TINY_KEY ? ieee_div3_1(x) : __fdiv_rn(x, 3.f); }
__device__ __forceinline__ float div_wide(float a, float b, float r, bool b_ok)
{
    return (b_ok && tiny_key(a) - SFH_NUM_KEY <= SFH_KEY_SPAN) ? ieee_div_r1(a, b, r) : __fdiv_rn(a, b);
}

// ---------------------------------------------------------------------------------------------
// fp64 pieces: Bullet btVector3 operator order (3rdparty/LinearMath/btVector3.h), unfused
// ---------------------------------------------------------------------------------------------
struct D3 {
    double x, y, z;
};
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ D3 operator-(D3 a, D3 b) { return {dsub(a.x, b.x), dsub(a.y, b.y), dsub(a.z, b.z)}; }
__device__ __forceinline__ D3 operator+(D3 a, D3 b) { return {dadd(a.x, b.x), dadd(a.y, b.y), dadd(a.z, b.z)}; }
__device__ __forceinline__ D3 operator*(D3 v, double s) { return {dmul(v.x, s), dmul(v.y, s), dmul(v.z, s)}; }
__device__ __forceinline__ double ddot(D3 a, D3 b) { return dadd(dadd(dmul(a.x, b.x), dmul(a.y, b.y)), dmul(a.z, b.z)); }
__device__ __forceinline__ D3 dcross(D3 a, D3 v)
{
    return {dsub(dmul(a.y, v.z), dmul(a.z, v.y)), dsub(dmul(a.z, v.x), dmul(a.x, v.z)), dsub(dmul(a.x, v.y), dmul(a.y, v.x))};
}
__device__ __forceinline__ double dnorm(D3 a) { return __dsqrt_rn(ddot(a, a)); }

// a[0..5] += (double){x0..x5}: the fp32 per-face terms enter the fp64 sums exactly
__device__ __forceinline__ void acc6(double* a, float x0, float x1, float x2, float x3, float x4, float x5)
{
    a[0] = dadd(a[0], (double)x0); a[1] = dadd(a[1], (double)x1); a[2] = dadd(a[2], (double)x2);
    a[3] = dadd(a[3], (double)x3); a[4] = dadd(a[4], (double)x4); a[5] = dadd(a[5], (double)x5);
}

// The same under a per-lane condition. SFH_ACC_VARIANT 1: the conversions run unconditionally — pinned with asm volatile so they
// are not sunk back into the conditional block — which lets ptxas predicate the DADDs (no BSSY/BRA/BSYNC per block) and spread the
// XU-pipe F2Fs over the surrounding fp32 work: k_submerged 0.4454 -> 0.4300 ms per 51.2 M faces. 0: a divergent
// block per condition.
#ifndef SFH_ACC_VARIANT
#define SFH_ACC_VARIANT 1
#endif
__device__ __forceinline__ double cvt_pinned(float x)
{
    double d;
    asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d) : "f"(x));
    return d;
}
__device__ __forceinline__ void acc6_if(bool p, double* a, float x0, float x1, float x2, float x3, float x4, float x5)
{
#if SFH_ACC_VARIANT == 1
    const double d0 = cvt_pinned(x0), d1 = cvt_pinned(x1), d2 = cvt_pinned(x2), d3 = cvt_pinned(x3), d4 = cvt_pinned(x4), d5 = cvt_pinned(x5);
    if (p) {
        a[0] = dadd(a[0], d0); a[1] = dadd(a[1], d1); a[2] = dadd(a[2], d2);
        a[3] = dadd(a[3], d3); a[4] = dadd(a[4], d4); a[5] = dadd(a[5], d5);
    }
#else
    if (p) acc6(a, x0, x1, x2, x3, x4, x5);
#endif
}

__device__ __forceinline__ void acc3_if(bool p, double* a, float x0, float x1, float x2)
{
#if SFH_ACC_VARIANT == 1
    const double d0 = cvt_pinned(x0), d1 = cvt_pinned(x1), d2 = cvt_pinned(x2);
    if (p) { a[0] = dadd(a[0], d0); a[1] = dadd(a[1], d1); a[2] = dadd(a[2], d2); }
#else
    if (p) { a[0] = dadd(a[0], (double)x0); a[1] = dadd(a[1], (double)x1); a[2] = dadd(a[2], (double)x2); }
#endif
}
__device__ __forceinline__ void acc1_if(bool p, double& a, float x)
{
#if SFH_ACC_VARIANT == 1
    const double d = cvt_pinned(x);
    if (p) a = dadd(a, d);
#else
    if (p) a = dadd(a, (double)x);
#endif
}

// Uniform/Jet/Pipe::GetVelocityAtPoint (Uniform.cpp:54-57, Jet.cpp:54-75, Pipe.cpp:57-78).
// exp/pow are CUDA's (<= 1-2 ulp) where the reference calls glibc's: see DESIGN.md "fp64 transcendentals".
__device__ __forceinline__ D3 current_velocity(const sfh_current& c, D3 p)
{
    D3 r = {0.0, 0.0, 0.0};
    if (c.type == SFH_CURRENT_UNIFORM) {
        r = D3{c.a[0], c.a[1], c.a[2]};
    } else if (c.type == SFH_CURRENT_JET) {
        D3 cc = {c.a[0], c.a[1], c.a[2]}, n = {c.n[0], c.n[1], c.n[2]};
        const double rr = c.r1, vout = c.vel;
        D3 cp = p - cc;
        double d = dnorm(dcross(cp, n));
        double t = ddot(cp, n);
        if (!(t < 0.0)) {
            double spread = dadd(t, dmul(5.0, rr));
            double r_ = dmul(0.2, spread); // Scalar(1)/Scalar(5) folds to the double nearest 0.2
            if (!(d >= r_)) {
                double s = dmul(ddiv(dmul(10.0, rr), spread), vout);
                D3 vmax = n * s;
                double f = exp(ddiv(dmul(dmul(-50.0, d), d), dmul(t, t)));
                r = vmax * f;
            }
        }
    } else {
        D3 p1 = {c.a[0], c.a[1], c.a[2]}, n = {c.n[0], c.n[1], c.n[2]};
        const double l = c.l, r1 = c.r1, r2 = c.r2, vin = c.vel, gamma = c.gamma;
        D3 p1p = p - p1;
        double d = dnorm(dcross(p1p, n));
        double t = ddot(p1p, n);
        if (!(t < 0.0 || t > l)) {
            double rad = dadd(r1, ddiv(dmul(dsub(r2, r1), t), l));
            if (!(d >= rad)) {
                D3 v = n * dmul(ddiv(r1, rad), vin);
                // btPow = libm pow. For the two profiles scenes use (gamma = 1 linear, gamma = 2 parabolic) the power is spelled
                // out: exact for 1, and for 2 within 1 ulp of glibc's pow (equal in 99.9 % of arguments; CUDA's pow is within 2 ulp) —
                // inside the fp64-transcendental tolerance of DESIGN.md 5.4 and 3x cheaper per face inside the pipe.
                const double base = dsub(1.0, ddiv(d, rad));
                double f = gamma == 2.0 ? dmul(base, base) : (gamma == 1.0 ? base : pow(base, gamma));
                r = v * f;
            }
        }
    }
    return r;
}

// Ocean::GetFluidVelocity(glm::vec3) (Ocean.cpp:183-201): widen, sum enabled currents in list order, narrow.
__device__ __forceinline__ F3 fluid_velocity_general(const sfh_current* cur, uint32_t n_cur, F3 fc)
{
    D3 p = {(double)fc.x, (double)fc.y, (double)fc.z};
    D3 fv = {0.0, 0.0, 0.0};
    for (uint32_t i = 0; i < n_cur; ++i)
        if (cur[i].enabled) fv = fv + current_velocity(cur[i], p);
    return F3{(float)fv.x, (float)fv.y, (float)fv.z};
}

// fp32 pre-test of a Jet / Pipe field against the bounding sphere of a CHUNK: most chunks of a scene lie far outside the
// cone / the pipe, where the field is exactly zero (Jet.cpp:62-67, Pipe.cpp:66-71). The test evaluates the field geometry
// at the sphere centre in fp32 with an error budget and only answers "certainly zero everywhere in the sphere"; fields that
// may reach the chunk are evaluated in fp64 per face as before, so results are unchanged.
struct CurGuard {
    float a[3], n[3]; // centre / first point and axis, rounded to fp32
    float r_hi, l_hi; // Jet: r; Pipe: max(r1, r2), and l — each rounded up
    float a_abs;      // |a.x| + |a.y| + |a.z|
    uint32_t type;
};
__device__ __forceinline__ CurGuard make_cur_guard(const sfh_current& c)
{
    CurGuard g;
    for (int k = 0; k < 3; ++k) { g.a[k] = (float)c.a[k]; g.n[k] = (float)c.n[k]; }
    const double r = c.type == SFH_CURRENT_PIPE ? fmax(c.r1, c.r2) : c.r1;
    g.r_hi = __double2float_ru(r);
    g.l_hi = __double2float_ru(c.l);
    g.a_abs = fabsf(g.a[0]) + fabsf(g.a[1]) + fabsf(g.a[2]);
    g.type = c.type;
    return g;
}
// true: the field is certainly zero at every point within `rho` of p. With t = (q-a).n and d = |(q-a) x n| (|n| = 1), moving q
// by at most rho moves t and d by at most rho; the fp32 evaluation at p is off by far less than E = 2e-5 (|p-a|_1 + |a|_1)
// (operand rounding 6e-8 |a|, a dozen roundings at 6e-8 of |p-a|, |n| = 1 to 1e-7).
__device__ __forceinline__ bool cur_certainly_zero(const CurGuard& g, F3 p, float rho)
{
    if (g.type == SFH_CURRENT_UNIFORM) return false;
    const float x = p.x - g.a[0], y = p.y - g.a[1], z = p.z - g.a[2];
    const float E = 2e-5f * (fabsf(x) + fabsf(y) + fabsf(z) + g.a_abs) + 1e-30f + rho;
    const float t = x * g.n[0] + y * g.n[1] + z * g.n[2];
    if (t < -E) return true; // behind the outlet / before the inlet
    const float cx = y * g.n[2] - z * g.n[1], cy = z * g.n[0] - x * g.n[2], cz = x * g.n[1] - y * g.n[0];
    const float d = sqrtf(cx * cx + cy * cy + cz * cz);
    if (g.type == SFH_CURRENT_JET) return d - E > 0.2f * (t + E) + g.r_hi; // r_ = (t + 5 r) / 5 <= 0.2 (t + E) + r_hi
    return t > g.l_hi + E || d - E > g.r_hi; // beyond the outlet, or outside the widest section
}
// Warp-collective: bit i set = Jet / Pipe field i may be non-zero somewhere in the sphere (centre p, radius rho); bits of
// Uniform fields stay clear (their sum is position-independent and precomputed by k_prologue). n_cur <= 64.
__device__ __forceinline__ uint64_t chunk_current_mask(const CurGuard* guard, uint32_t n_cur, F3 p, float rho)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t hi = min(lane + 32u, SFH_MAX_CURRENTS - 1u);
    const bool lo_may = lane < n_cur && guard[lane].type != SFH_CURRENT_UNIFORM && !cur_certainly_zero(guard[lane], p, rho);
    const bool hi_may = lane + 32u < n_cur && guard[hi].type != SFH_CURRENT_UNIFORM && !cur_certainly_zero(guard[hi], p, rho);
    return (uint64_t)__ballot_sync(0xffffffffu, lo_may) | ((uint64_t)__ballot_sync(0xffffffffu, hi_may) << 32);
}

// Ocean::GetFluidVelocity(glm::vec3) with the chunk mask: Uniform fields and the Jet / Pipe fields that can reach the chunk,
// in list order. Leaving out a field whose value is (+0, +0, +0) cannot change the sum, signed zeros included: the sum starts at
// +0 (Ocean.cpp:185) and under round-to-nearest a running sum that started at +0 is never -0 (+0 + -0 = +0, and an exact
// cancellation gives +0), so the skipped `x + (+0)` is always the identity — a Uniform field with a -0 component included.
__device__ __forceinline__ F3 fluid_velocity_masked(const sfh_current* cur, uint64_t may, uint32_t n_cur, F3 fc)
{
    const D3 p = {(double)fc.x, (double)fc.y, (double)fc.z};
    D3 fv = {0.0, 0.0, 0.0};
    for (uint32_t i = 0; i < n_cur; ++i)
        if (cur[i].enabled && (cur[i].type == SFH_CURRENT_UNIFORM || ((may >> i) & 1ull))) fv = fv + current_velocity(cur[i], p);
    return F3{(float)fv.x, (float)fv.y, (float)fv.z};
}

// Per-kernel constants of the wave lookup: the reference divides every coordinate by the grid size and recomputes
// 0.5 / fftSize on every call (OpenGLRealOcean.cpp:283,325-326); same values, computed once.
struct WaveConst {
    float fN, half; // (float)N, 0.5f / N (IEEE)
    float rgx, rgy; // refined reciprocals of gridSizes.x / .y for div_wide
    bool gx_ok, gy_ok;
};
__device__ __forceinline__ WaveConst wave_const(const WavesDev& w)
{
    WaveConst c;
    c.fN = (float)w.N;
    c.half = fdiv(0.5f, c.fN);
    c.rgx = refined_rcp1(w.gx);
    c.rgy = refined_rcp1(w.gy);
    c.gx_ok = w.gx >= 0x1p-20f && w.gx <= 0x1p40f;
    c.gy_ok = w.gy >= 0x1p-20f && w.gy <= 0x1p40f;
    return c;
}

// OpenGLRealOcean::ComputeInterpolatedWaveData (OpenGLRealOcean.cpp:280-319) on a planar channel copy.
// SFH_NOINLINE_WAVES: one copy of the lookup per kernel instead of one per call site (eight in k_surface_lists). Measured: the wet
// role 9 % slower at full size (the two channels' loads no longer overlap), small-step latency unchanged — not the default.
#ifdef SFH_NOINLINE_WAVES
#define SFH_WAVE_INLINE __noinline__
#else
#define SFH_WAVE_INLINE __forceinline__
#endif
__device__ SFH_WAVE_INLINE float wave_bilinear(const float* __restrict__ ch, int N, const WaveConst& c, float x, float y)
{
    const float fN = c.fN, half = c.half;
    // modf(v).fractional = v - trunc(v), exact for finite v. (std::modf returns -0 where this gives +0 — for negative whole
    // numbers — but a zero weight or index offset is the same number with either sign; nothing below looks at the sign.)
    auto frac = [](float v) { return fsub(v, truncf(v)); };
    float i0f = frac(fsub(x, half));
    float j0f = frac(fsub(y, half));
    if (i0f < 0.f) i0f = fsub(1.f, fabsf(i0f));
    if (j0f < 0.f) j0f = fsub(1.f, fabsf(j0f));
    int i0 = (int)truncf(fmul(i0f, fN));
    int j0 = (int)truncf(fmul(j0f, fN));
    float i1f = frac(fadd(x, half));
    float j1f = frac(fadd(y, half));
    if (i1f < 0.f) i1f = fsub(1.f, fabsf(i1f));
    if (j1f < 0.f) j1f = fsub(1.f, fabsf(j1f));
    int i1 = (int)truncf(fmul(i1f, fN));
    int j1 = (int)truncf(fmul(j1f, fN));
    float alpha = frac(fmul(i0f, fN));
    float beta = frac(fmul(j0f, fN));
    // i0f..j1f lie in [0, 1] for finite coordinates, so i, j lie in [0, N] and the indices in [0, N*N+N]: the planes hold that
    // many texels (+1). A NaN/Inf coordinate makes the fractions NaN, and cvt.rzi.s32.f32 turns NaN into 0 (the reference's
    // (int) cast of a NaN is undefined behaviour): still in range, no clamp needed.
    const int k00 = j0 * N + i0, k01 = j0 * N + i1, k10 = j1 * N + i0, k11 = j1 * N + i1;
    const float t0 = __ldg(ch + k00), t1 = __ldg(ch + k01), t2 = __ldg(ch + k10), t3 = __ldg(ch + k11);
    const float oma = fsub(1.f, alpha), omb = fsub(1.f, beta);
    float h = fmul(fmul(oma, omb), t0);
    h = fadd(h, fmul(fmul(alpha, omb), t1));
    h = fadd(h, fmul(fmul(oma, beta), t2));
    h = fadd(h, fmul(fmul(alpha, beta), t3));
    return h;
}

// OpenGLRealOcean::ComputeWaveHeight (OpenGLRealOcean.cpp:321-331)
__device__ __forceinline__ float wave_height(const WavesDev& w, const WaveConst& c, float x, float y)
{
    float z = 0.f;
    z = fsub(z, wave_bilinear(w.ch0, w.N, c, div_wide(x, w.gx, c.rgx, c.gx_ok), div_wide(y, w.gx, c.rgx, c.gx_ok)));
    z = fsub(z, wave_bilinear(w.ch1, w.N, c, div_wide(x, w.gy, c.rgy, c.gy_ok), div_wide(y, w.gy, c.rgy, c.gy_ok)));
    return z;
}

// Ocean::GetDepth(glm::vec3) (Ocean.cpp:149-168)
template <bool WAVES>
__device__ __forceinline__ float get_depth(const WavesDev& w, const WaveConst& c, float x, float y, float z)
{
    if (WAVES) return fsub(z, wave_height(w, c, x, y));
    return z;
}
template <bool WAVES>
__device__ __forceinline__ float get_depth(const WavesDev& w, float x, float y, float z)
{
    if (WAVES) return fsub(z, wave_height(w, wave_const(w), x, y));
    return z;
}

} // namespace sfh
