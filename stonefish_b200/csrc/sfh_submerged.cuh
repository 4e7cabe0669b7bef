// sfh_submerged.cuh — k_submerged: SolidEntity::ComputeHydrodynamicForcesSubmerged (SolidEntity.cpp:1749-1791)
// over (body, chunk) work items, two faces per lane side by side (two independent fp32 streams; FMUL2/FFMA2 with
// -DSFH_PACKED_F32X2), and its aerodynamic twin (AERO).
//
// Per chunk ONE WARP
//   1. transforms the chunk's <=256 vertices ONCE into its private shared-memory buffer (reference: once per incident
//      face, SolidEntity.cpp:1755-1757 — identical bits),
//   2. has the NEXT chunk (of this or of the next body) already in flight as one bulk async copy (sfh_chunkpipe.cuh),
//   3. evaluates faces base+lane and base+lane+32 side by side (P2 = one value per face), 4 such pairs per lane per
//      chunk, with the IEEE divisions and square roots written out (sfh_packed.cuh),
//   4. adds the fp32 per-face terms into per-lane fp64 accumulators under predicates (acc6_if).
// Steps 1 and 3-4 are straight-line code: no warp vote, no divergent block, one rare branch (operands outside the proven
// range of the written-out divisions) — ptxas schedules one large block per face pair, which is worth 9 %.
// Accumulators are reduced per (warp, body) and written as one partial record (see sfh_kernels.cuh).
#pragma once

#include "sfh_chunkpipe.cuh"
#include "sfh_packed.cuh"

namespace sfh {

// fn / len for three numerators sharing one reciprocal (len >= 1e-6 here, and |fn_i| <= ~len). The shared form
// is proven for numerators that are zero or >= 2^-36 in magnitude (quotient zero or >= 2^-100, remainder exact).
// ((a + b) + c) / 3.f uses ieee_div3_2, proven for sums that are zero or >= 2^-100.
// Anything else — never seen on real meshes — is redone with the hardware division by the caller.
__device__ __forceinline__ uint32_t min_key(V3P a)
{
    return min(min(min(tiny_key(lo(a.x)), tiny_key(lo(a.y))), tiny_key(lo(a.z))),
               min(min(tiny_key(hi(a.x)), tiny_key(hi(a.y))), tiny_key(hi(a.z))));
}
__device__ __forceinline__ V3P normalize_fast(V3P fn, P2 len)
{
    const P2 nlen = neg(len);
    const P2 r = refined_rcp2(len, nlen);
    return {ieee_div_r2(fn.x, nlen, r), ieee_div_r2(fn.y, nlen, r), ieee_div_r2(fn.z, nlen, r)};
}
__device__ __forceinline__ void normalize_centroid_slow(V3P fn, P2 len, V3P s, V3P& fn1, V3P& fc)
{
    fn1.x = pk(__fdiv_rn(lo(fn.x), lo(len)), __fdiv_rn(hi(fn.x), hi(len)));
    fn1.y = pk(__fdiv_rn(lo(fn.y), lo(len)), __fdiv_rn(hi(fn.y), hi(len)));
    fn1.z = pk(__fdiv_rn(lo(fn.z), lo(len)), __fdiv_rn(hi(fn.z), hi(len)));
    fc.x = pk(__fdiv_rn(lo(s.x), 3.f), __fdiv_rn(hi(s.x), 3.f));
    fc.y = pk(__fdiv_rn(lo(s.y), 3.f), __fdiv_rn(hi(s.y), 3.f));
    fc.z = pk(__fdiv_rn(lo(s.z), 3.f), __fdiv_rn(hi(s.z), 3.f));
}

// The scalar per-face pieces of a pair — the fp64 current evaluation, the wave lookups — run as a two-trip loop that is NOT
// unrolled: one copy of ~300 (currents) / ~200 (waves) instructions instead of two. The face kernels are bound by instruction
// issue AND by instruction fetch (k_surface_wet under waves and Jet + Pipe: 163 KB of SASS, `no_instruction` 1.0 warps per issue
// cycle), so code that is executed twice should exist once.
__device__ __forceinline__ V3P fluid_velocity_pair(const sfh_current* s_cur, uint64_t cur_may, uint32_t n_cur, V3P fc)
{
    F3 out[2];
#pragma unroll 1
    for (int f = 0; f < 2; ++f) {
        const F3 c = f ? F3{hi(fc.x), hi(fc.y), hi(fc.z)} : F3{lo(fc.x), lo(fc.y), lo(fc.z)};
        const F3 r = fluid_velocity_masked(s_cur, cur_may, n_cur, c);
        if (f) out[1] = r; else out[0] = r;
    }
    return V3P{pk(out[0].x, out[1].x), pk(out[0].y, out[1].y), pk(out[0].z, out[1].z)};
}
template <bool WAVES>
__device__ __forceinline__ P2 depth_pair(const WavesDev& wv, const WaveConst& wc, V3P fc)
{
    float out[2];
#pragma unroll 1
    for (int f = 0; f < 2; ++f) {
        const F3 c = f ? F3{hi(fc.x), hi(fc.y), hi(fc.z)} : F3{lo(fc.x), lo(fc.y), lo(fc.z)};
        const float d = get_depth<WAVES>(wv, wc, c.x, c.y, c.z);
        if (f) out[1] = d; else out[0] = d;
    }
    return pk(out[0], out[1]);
}

struct BodyP2 { // per-body constants broadcast to both lanes
    V3P p, v, omega, cur;
    float R[9], t[3]; // T_C, for the per-chunk vertex transform
};

// Damping block for a pair of faces (SolidEntity.cpp:1771-1790 / 1661-1683). mA/mB: the face exists and is not
// degenerate. Returns SFH_FACE_QUADRATIC / SFH_FACE_SKIN bits for A in bits 0..7 and for B in bits 8..15.
template <int CUR>
__device__ __forceinline__ uint32_t damping_pair(const BodyP2& B, const sfh_current* s_cur, uint64_t cur_may, uint32_t n_cur, V3P fc,
                                                 V3P fn1, P2 A, bool mA, bool mB, double* acc, uint32_t& epack)
{
    V3P fluid = B.cur; // sum of the enabled Uniform fields; complete when no Jet / Pipe field reaches the chunk
    if (CUR && cur_may) fluid = fluid_velocity_pair(s_cur, cur_may, n_cur, fc);
    const V3P r = fc - B.p;
    const V3P vc = fluid - (B.v + cross(B.omega, r));
    const P2 vc_n = dot(vc, fn1);
    const V3P vn = fn1 * vc_n;
    const V3P vt = vc - vn;
    uint32_t code = 0;

    const bool qA = mA && lo(vc_n) < -1e-12f, qB = mB && hi(vc_n) < -1e-12f;
    { // no warp vote around the block: straight-line code schedules better than the skipped work costs (0.418 -> 0.413 ms)
        // vc_n < -1e-12 implies |vc|^2 > 1e-24 > 2^-101: inside ieee_sqrt2's range for every lane that is used
        const P2 vmag2 = dot(vc, vc);
        const V3P q = ((vc * ieee_sqrt2(vmag2)) * neg(vc_n)) * A;
        const V3P tq = cross(r, q);
        acc6_if(qA, acc, lo(q.x), lo(q.y), lo(q.z), lo(tq.x), lo(tq.y), lo(tq.z));
        acc6_if(qB, acc, hi(q.x), hi(q.y), hi(q.z), hi(tq.x), hi(tq.y), hi(tq.z));
        epack += (qA ? 1u << 10 : 0u) + (qB ? 1u << 10 : 0u);
        code |= (qA ? (uint32_t)SFH_FACE_QUADRATIC : 0u) | (qB ? (uint32_t)SFH_FACE_QUADRATIC << 8 : 0u);
    }
    const P2 vt2 = dot(vt, vt);
    const bool sA = mA && lo(vt2) > 1e-9f, sB = mB && hi(vt2) > 1e-9f;
    { // nearly every face has tangential flow: no warp vote
        const V3P s = vt * A;
        const V3P ts = cross(r, s);
        acc6_if(sA, acc + 6, lo(s.x), lo(s.y), lo(s.z), lo(ts.x), lo(ts.y), lo(ts.z));
        acc6_if(sB, acc + 6, hi(s.x), hi(s.y), hi(s.z), hi(ts.x), hi(ts.y), hi(ts.z));
        epack += (sA ? 1u << 20 : 0u) + (sB ? 1u << 20 : 0u);
        code |= (sA ? (uint32_t)SFH_FACE_SKIN : 0u) | (sB ? (uint32_t)SFH_FACE_SKIN << 8 : 0u);
    }
    return code;
}

#ifdef SFH_SOL
// SPEED-OF-LIGHT PROBE (never the product: results are not the reference's). The mandated fp32 chain of a face pair and nothing
// else: no fp32->fp64 conversions, no fp64 adds, no event counters, no range guards, no predicates — the twelve sums are kept
// alive by plain fp32 adds. Its time is the measured ceiling of ANY bit-exact k_submerged with this data path
// (profiles/sol_k_submerged_r02.txt); build with -DSFH_SOL (tools/sol.sh), run with bench.py --no-validate.
__device__ __forceinline__ void sol_pair(const BodyP2& B, V3P fc, V3P fn1, P2 A, float* accf)
{
    const V3P r = fc - B.p;
    const V3P vc = B.cur - (B.v + cross(B.omega, r));
    const P2 vc_n = dot(vc, fn1);
    const V3P vn = fn1 * vc_n;
    const V3P vt = vc - vn;
    const P2 vmag2 = dot(vc, vc);
    const V3P q = ((vc * ieee_sqrt2(vmag2)) * neg(vc_n)) * A;
    const V3P tq = cross(r, q);
    const V3P s = vt * A;
    const V3P ts = cross(r, s);
    const P2 t[12] = {q.x, q.y, q.z, tq.x, tq.y, tq.z, s.x, s.y, s.z, ts.x, ts.y, ts.z};
#pragma unroll
    for (int i = 0; i < 12; ++i) accf[i] = __fadd_rn(__fadd_rn(accf[i], lo(t[i])), hi(t[i]));
}
#endif

// Air-drag block for a pair of faces: SolidEntity::ComputeAerodynamicForces (SolidEntity.cpp:1909-1919). The reference
// tests dot(fn1, vn), not dot(vc, fn1), and scales by vn.length() — glm's COMPONENT COUNT, 3 — through
// glm/gtx/scalar_multiplication.hpp (vec3 * int -> vec3 *= 3.f); both are kept. acc[0..5] = Fda Tda.
template <int CUR>
__device__ __forceinline__ uint32_t aero_pair(const BodyP2& B, const sfh_current* s_cur, uint64_t cur_may, uint32_t n_cur, V3P fc,
                                              V3P fn1, P2 A, bool mA, bool mB, double* acc, uint32_t& epack)
{
    V3P fluid = B.cur; // sum of the enabled Uniform fields; complete when no Jet / Pipe field reaches the chunk
    if (CUR && cur_may) fluid = fluid_velocity_pair(s_cur, cur_may, n_cur, fc);
    const V3P r = fc - B.p;
    const V3P vc = fluid - (B.v + cross(B.omega, r));
    const V3P vn = fn1 * dot(vc, fn1);
    const P2 t = dot(fn1, vn);
    uint32_t code = 0;
    const bool qA = mA && lo(t) < -1e-12f, qB = mB && hi(t) < -1e-12f;
    { // straight-line, as in damping_pair
        const V3P q = (vn * pk1(3.f)) * A;
        const V3P tq = cross(r, q);
        acc6_if(qA, acc, lo(q.x), lo(q.y), lo(q.z), lo(tq.x), lo(tq.y), lo(tq.z));
        acc6_if(qB, acc, hi(q.x), hi(q.y), hi(q.z), hi(tq.x), hi(tq.y), hi(tq.z));
        epack += (qA ? 1u << 10 : 0u) + (qB ? 1u << 10 : 0u);
        code |= (qA ? (uint32_t)SFH_FACE_QUADRATIC : 0u) | (qB ? (uint32_t)SFH_FACE_QUADRATIC << 8 : 0u);
    }
    return code;
}

// State of the block-wide walk over work items, shared by k_submerged and k_surface.
// Thread 0 alone tracks where the staged chunk came from so it can chain the next copy.
struct ItemWalk {
    uint32_t w, w1, b, b_begin, b_end;
    int s;               // staging buffer holding the current chunk
    uint32_t parity;     // bit s = phase parity of staging buffer s
    const uint8_t* src;  // thread 0: global address of the current chunk
    uint32_t bytes;      // thread 0: its size
};

// thread 0: start the copy of the chunk of work item `w` of body `b` (random access through chunk_off)
__device__ __forceinline__ void walk_issue_first(const PipeAddr& pipe, ItemWalk& it, const BodyDev& body)
{
    const uint32_t c = it.w - it.b_begin;
    const uint32_t o0 = __ldg(body.chunk_off + c), o1 = __ldg(body.chunk_off + c + 1);
    it.src = body.chunks + (size_t)o0 * 16u;
    it.bytes = (o1 - o0) * 16u;
    pipe_issue(pipe, it.s, it.src, it.bytes);
}

// thread 0, right after the block barrier that retires the previous chunk: chain the copy of item w+1 into the
// other buffer — the following chunk of this body (size from the staged header) or chunk 0 of the next body
// that has work (pointer and size read as one 16-byte load from its BodyDev record).
__device__ __forceinline__ void walk_issue_next(const PipeAddr& pipe, ItemWalk& it, const StepParams& P, uint32_t next_bytes,
                                                const uint8_t*& nsrc, uint32_t& nbytes)
{
    const uint32_t wn = it.w + 1;
    if (wn >= it.w1) return;
    if (wn < it.b_end) {
        nsrc = it.src + it.bytes;
        nbytes = next_bytes;
    } else {
        uint32_t nb = it.b + 1;
        while (__ldg(P.item_off + nb + 1) <= wn) ++nb;
        const uint4 rec = __ldg(reinterpret_cast<const uint4*>(&P.bodies[nb].chunks));
        nsrc = reinterpret_cast<const uint8_t*>(((unsigned long long)rec.y << 32) | rec.x);
        nbytes = rec.z;
    }
    pipe_issue(pipe, it.s ^ 1, nsrc, nbytes);
}

// ---------------------------------------------------------------------------------------------------------------
// Work decomposition: one WARP per stream of work items. Each warp owns a private chunk pipeline and vertex buffer,
// evaluates 4 face pairs per lane per chunk and never meets a block barrier (the first version gave one 128-thread
// block per stream: two block barriers per chunk and 1 pair per thread, 6 % slower — DESIGN.md §7).
// ---------------------------------------------------------------------------------------------------------------
#define SFH_SUB_WPB 2u /* warps (streams) per block */
#ifndef SFH_SUB_MIN_BLOCKS
#define SFH_SUB_MIN_BLOCKS 8
#endif
#define SFH_SUB_THREADS (32u * SFH_SUB_WPB)
#ifndef SFH_SUB_UNROLL
#define SFH_SUB_UNROLL 1 /* face pairs per lane interleaved in the face loop */
#endif
constexpr int kSubUnroll = SFH_SUB_UNROLL;

struct WarpStage { // shared memory private to one warp
    ChunkPipe pipe;
    float4 v[SFH_CH_VERTS];
    BodyDev body;
};
struct WarpStageSurf : WarpStage { // the surface kernel's clip role: which vertex slots a sub-item's faces refer to
    uint32_t need[8];
};

__device__ __forceinline__ PipeAddr warp_pipe_init(ChunkPipe& p, uint32_t lane)
{
    PipeAddr a;
    a.bar0 = smem_addr(&p.bar[0]);
    a.buf0 = smem_addr(&p.buf[0][0]);
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a.bar0));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a.bar0 + 8u));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    return a;
}

__device__ __forceinline__ void warp_load_body(BodyDev* sb, const BodyDev* __restrict__ src, uint32_t lane)
{
    const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
    uint32_t* d = reinterpret_cast<uint32_t*>(sb);
    for (uint32_t i = lane; i < sizeof(BodyDev) / 4; i += 32) d[i] = __ldg(s + i);
    __syncwarp();
}

// AERO: the same face loop with the air-drag block (SolidEntity::ComputeAerodynamicForces, SolidEntity.cpp:1886-1919)
template <int CUR, bool DEBUG, bool AERO = false>
__global__ void __launch_bounds__(SFH_SUB_THREADS, SFH_SUB_MIN_BLOCKS) k_submerged(StepParams P)
{
    __shared__ WarpStage s_stage[SFH_SUB_WPB];
    __shared__ sfh_current s_cur[CUR ? SFH_MAX_CURRENTS : 1];
    __shared__ CurGuard s_guard[CUR ? SFH_MAX_CURRENTS : 1];

    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t n_cur = CUR ? P.ocean->n_currents : 0;
    if constexpr (CUR != 0) {
        stage_currents(s_cur, s_guard, P.ocean, n_cur);
        __syncthreads(); // the only block barrier of the kernel, before any warp can leave
    }
    pdl_trigger();
    pdl_wait(); // the prologue's work lists and body records
    const uint64_t W = P.item_off[P.n_bodies], G = min((uint64_t)gridDim.x * SFH_SUB_WPB, W);
    const uint32_t wg = blockIdx.x * SFH_SUB_WPB + warp;
    if (wg >= G) return;
    WarpStage& S = s_stage[warp];
    BodyDev& s_body = S.body;
    float4* const s_v = S.v;

    ItemWalk it;
    it.w = (uint32_t)((uint64_t)wg * W / G);
    it.w1 = (uint32_t)(((uint64_t)wg + 1) * W / G);
    it.s = 0;
    it.parity = 0;

    double acc[SFH_NACC_SUB];
#pragma unroll
    for (int i = 0; i < SFH_NACC_SUB; ++i) acc[i] = 0.0;
#ifdef SFH_SOL
    float accf[SFH_NACC_SUB];
#pragma unroll
    for (int i = 0; i < SFH_NACC_SUB; ++i) accf[i] = 0.f;
#endif
    uint32_t epack = 0, since_spill = 0;
    uint32_t cnt0 = 0, cnt1 = 0, cnt2 = 0; // degenerate, quadratic, skin of the current body (warp-uniform)

    it.b = body_of_item(P.item_off, P.n_bodies, it.w);
    it.b_begin = P.item_off[it.b];
    it.b_end = P.item_off[it.b + 1];
    warp_load_body(&s_body, P.bodies + it.b, lane);
    const PipeAddr pa = warp_pipe_init(S.pipe, lane);
    if (lane == 0) walk_issue_first(pa, it, s_body);

    BodyP2 B;
    float rnorm = 0.f; // >= the 2-norm of the linear part of T_C (Frobenius norm)
    auto fetch_regs = [&]() {
        if (CUR) {
            float fro = 0.f;
            for (int k = 0; k < 9; ++k) fro += s_body.R[k] * s_body.R[k];
            rnorm = sqrtf(fro) * 1.000001f;
        }
        B.p = V3P{pk1(s_body.p[0]), pk1(s_body.p[1]), pk1(s_body.p[2])};
        B.v = V3P{pk1(s_body.v[0]), pk1(s_body.v[1]), pk1(s_body.v[2])};
        B.omega = V3P{pk1(s_body.omega[0]), pk1(s_body.omega[1]), pk1(s_body.omega[2])};
        B.cur = V3P{pk1(s_body.cur[0]), pk1(s_body.cur[1]), pk1(s_body.cur[2])};
    };
    auto spill = [&]() { // 3 x 10-bit per-lane event counters -> warp-uniform totals
        cnt0 += __reduce_add_sync(SFH_FULL, epack & 1023u);
        cnt1 += __reduce_add_sync(SFH_FULL, (epack >> 10) & 1023u);
        cnt2 += __reduce_add_sync(SFH_FULL, (epack >> 20) & 1023u);
        epack = 0;
        since_spill = 0;
    };
    auto flush = [&](uint32_t body) { // one partial record per (warp, body)
        spill();
#ifdef SFH_SOL
#pragma unroll
        for (int i = 0; i < SFH_NACC_SUB; ++i) { acc[i] = (double)accf[i]; accf[i] = 0.f; }
#endif
        PartialSub* rec = reinterpret_cast<PartialSub*>(P.partials) + (wg + body);
#ifdef SFH_SUB_FOLD_FLUSH /* the folded flush of the surface roles (warp_sum_store): 20 KB less code here, but the face loop around
                             it is scheduled differently and cfg 4 loses 2.2 % (127.2 against 130.0 G, A/B on one box) */
        warp_sum_store<SFH_NACC_SUB>(acc, rec->a, lane);
#else
#pragma unroll
        for (int i = 0; i < SFH_NACC_SUB; ++i) {
            const double v = warp_sum(acc[i]);
            if (lane == 0) rec->a[i] = v;
            acc[i] = 0.0;
        }
#endif
        if (lane == 0) { rec->n_degenerate = cnt0; rec->n_quadratic = cnt1; rec->n_skin = cnt2; }
        cnt0 = cnt1 = cnt2 = 0;
    };
    fetch_regs();

    for (;;) {
        pipe_wait(pa, it.s, (it.parity >> it.s) & 1u);
        it.parity ^= 1u << it.s;
        const uint8_t* buf = S.pipe.buf[it.s];
        const ChunkView cv = chunk_view(buf);

        __syncwarp(); // every lane is done with the previous chunk: s_v and the other staging buffer are free
        const uint8_t* nsrc = nullptr;
        uint32_t nbytes = 0;
        if (lane == 0) walk_issue_next(pa, it, P, cv.next_bytes, nsrc, nbytes);

        { // each vertex transformed once per chunk (SolidEntity.cpp:1755-1757). Slots 0..191 in one straight-line block without a
          // per-lane test: T_C is read from shared memory once, every address is a constant offset, nothing diverges. Slots past
          // the chunk's last vertex read other bytes of the staging buffer (x/y/z[i], i < 256, always lie inside it) and fill
          // vertex slots no face refers to. The groups above 192 (one chunk in five on a Morton-ordered hull) under
          // warp-uniform tests.
            auto xf = [&](uint32_t k) {
                const uint32_t i = lane + 32u * k;
                const F3 pw = xform(s_body.R, s_body.t, cv.x[i], cv.y[i], cv.z[i]);
                s_v[i] = make_float4(pw.x, pw.y, pw.z, 0.f);
            };
#pragma unroll
            for (uint32_t k = 0; k < 6u; ++k) xf(k);
            if (cv.nv > 192u) {
                xf(6u);
                if (cv.nv > 224u) xf(7u);
            }
        }
        __syncwarp();

        // which Jet / Pipe fields can reach this chunk at all (every face centroid lies inside the chunk's bounding sphere)
        uint64_t cur_may = 0;
        if (CUR) {
            const F3 cw = xform(s_body.R, s_body.t, cv.sphere.x, cv.sphere.y, cv.sphere.z);
            const float reach = fabsf(s_body.t[0]) + fabsf(s_body.t[1]) + fabsf(s_body.t[2]) + fabsf(cv.sphere.x) + fabsf(cv.sphere.y) +
                                fabsf(cv.sphere.z) + cv.sphere.w;
            cur_may = chunk_current_mask(s_guard, n_cur, cw, rnorm * cv.sphere.w + (4e-6f * reach + 1e-6f));
        }

        const uint32_t nf = cv.nf;
#pragma unroll kSubUnroll // 1; unrolling by 2 (127 registers) changes nothing: 0.3994 vs 0.3995 ms
        for (uint32_t base = 0; base < nf; base += 64u) { // faces base+lane and base+lane+32 side by side
            const uint32_t iA = base + lane, iB = iA + 32u;
            const bool vA = iA < nf, vB = iB < nf;
            const uint32_t fA = vA ? cv.faces[iA] : 0u, fB = vB ? cv.faces[iB] : 0u;
            // byte k of the packed face word -> vertex slot (PRMT) -> float4 in s_v (LEA)
            const float4 a1 = s_v[__byte_perm(fA, 0, 0x4440)], a2 = s_v[__byte_perm(fA, 0, 0x4441)], a3 = s_v[__byte_perm(fA, 0, 0x4442)];
            const float4 b1 = s_v[__byte_perm(fB, 0, 0x4440)], b2 = s_v[__byte_perm(fB, 0, 0x4441)], b3 = s_v[__byte_perm(fB, 0, 0x4442)];
            const V3P p1 = {pk(a1.x, b1.x), pk(a1.y, b1.y), pk(a1.z, b1.z)};
            const V3P p2 = {pk(a2.x, b2.x), pk(a2.y, b2.y), pk(a2.z, b2.z)};
            const V3P p3 = {pk(a3.x, b3.x), pk(a3.y, b3.y), pk(a3.z, b3.z)};
            const V3P fv1 = p2 - p1;
            const V3P fv2 = p3 - p1;
            const V3P fn = cross(fv1, fv2);
            const P2 len2 = dot(fn, fn);
#ifdef SFH_SOL
            {
                const P2 len = ieee_sqrt2(len2);
                const V3P csum = (p1 + p2) + p3;
                sol_pair(B, V3P{ieee_div3_2(csum.x), ieee_div3_2(csum.y), ieee_div3_2(csum.z)}, normalize_fast(fn, len), len * pk1(0.5f), accf);
                continue;
            }
#endif
            const bool dA = lo(len2) < 1e-12f, dB = hi(len2) < 1e-12f; // SolidEntity.cpp:1764
            const bool mA = vA && !dA, mB = vB && !dB;
            epack += (vA && dA ? 1u : 0u) + (vB && dB ? 1u : 0u);
            uint32_t code;
            {
                // len2 >= 1e-12 > 2^-101 on every lane that is used (unused lanes compute garbage that is masked out)
                const P2 len = ieee_sqrt2(len2);
                const P2 A = len * pk1(0.5f); // len / 2.f, exact scaling
                const V3P csum = (p1 + p2) + p3;
                // Range guard of the shared-reciprocal / multiply-by-1/3 forms, cheap test first: every |fn_i| >= 2^-36
                // (its square, already needed for len2, >= 2^-72) and every |csum_i| >= 2^-100. Faces with an exactly
                // zero component (axis-aligned geometry) fail it and take the exact integer test, which accepts zeros.
                const P2 sx = fn.x * fn.x, sy = fn.y * fn.y, sz = fn.z * fn.z;
                const float m_fn = fminf(fminf(fminf(lo(sx), lo(sy)), lo(sz)), fminf(fminf(hi(sx), hi(sy)), hi(sz)));
                const float m_cs = fminf(fminf(fminf(fabsf(lo(csum.x)), fabsf(lo(csum.y))), fabsf(lo(csum.z))),
                                         fminf(fminf(fabsf(hi(csum.x)), fabsf(hi(csum.y))), fabsf(hi(csum.z))));
                const bool cheap_ok = (m_fn >= 0x1p-72f) & (m_cs >= 0x1p-100f);
                // One straight-line block: the proven forms run for every pair; the rare pair that fails the cheap test takes the
                // exact integer test and, outside the proven range, redoes the six quotients with the hardware division.
                V3P fn1 = normalize_fast(fn, len);
                V3P fc = {ieee_div3_2(csum.x), ieee_div3_2(csum.y), ieee_div3_2(csum.z)};
                if (!cheap_ok && !((min_key(fn) >= SFH_NUM_KEY) & (min_key(csum) >= SFH_TINY_KEY)))
                    normalize_centroid_slow(fn, len, csum, fn1, fc);
                code = AERO ? aero_pair<CUR>(B, s_cur, cur_may, n_cur, fc, fn1, A, mA, mB, acc, epack)
                            : damping_pair<CUR>(B, s_cur, cur_may, n_cur, fc, fn1, A, mA, mB, acc, epack);
            }
            if (DEBUG) {
                const uint32_t* of = s_body.orig_face + (size_t)(it.w - it.b_begin) * SFH_CH_FACES;
                if (vA) P.codes[P.code_base[it.b] + of[iA]] = (uint8_t)(7u | (dA ? SFH_FACE_DEGENERATE : 0u) | (code & 255u));
                if (vB) P.codes[P.code_base[it.b] + of[iB]] = (uint8_t)(7u | (dB ? SFH_FACE_DEGENERATE : 0u) | (code >> 8));
            }
        }
        if (++since_spill == SFH_SPILL_EVERY) spill(); // 8 faces per lane per chunk: the 10-bit fields stay below 1023

        if (++it.w >= it.w1) break;
        it.s ^= 1;
        it.src = nsrc;
        it.bytes = nbytes;
        if (it.w >= it.b_end) {
            flush(it.b);
            do { ++it.b; it.b_end = P.item_off[it.b + 1]; } while (it.w >= it.b_end);
            it.b_begin = P.item_off[it.b];
            __syncwarp();
            warp_load_body(&s_body, P.bodies + it.b, lane);
            fetch_regs();
        }
    }
    flush(it.b);
}


} // namespace sfh
