// sfh_kernels.cuh — the four kernels of one hydrodynamics step (sm_100a).
//
//   k_prologue   1 thread/body  per-body: CheckBodyFluidPosition (a9), dispatcher gates (a10), fp64->fp32
//                               narrowing of transforms/velocities, position-independent current sum,
//   k_scan_items 1 block        exclusive scan of the per-body chunk counts -> two work lists (submerged / surface)
//   k_submerged  persistent     per-face ComputeHydrodynamicForcesSubmerged (a1) over (body, chunk) items, one stream per WARP
//   k_surface    persistent     per-face ComputeHydrodynamicForcesSurface  (a2) incl. waves (a3/a4)
//   k_epilogue   1 warp/body    fixed-order reduction of the per-block partials, finalisation (A.5),
//                               INSIDE analytic buoyancy and CorrectHydrodynamicForces (a11)
//
// Work decomposition: the work list is the concatenation of every active body's chunks; stream `blk` of G streams
// (a stream = one block of k_surface, one warp of k_submerged) owns the contiguous item range [blk*W/G, (blk+1)*W/G).
// A stream keeps fp64 partial sums in registers and writes ONE record per (stream, body) it touched at index blk + body — those indices are unique
// because block ranges are contiguous and monotone in the body index — so no atomics, no zero-fill, and the
// epilogue reduces each body's records in a fixed order: results are bitwise reproducible for a given G.
#pragma once

#include "sfh_device.cuh"

namespace sfh {

#ifndef SFH_BLOCK
#define SFH_BLOCK 128 /* threads of a k_surface block: 128 or 64 */
#endif
#define SFH_SURF_FPT (256 / SFH_BLOCK) /* faces (and vertices) of a chunk per k_surface thread */
#define SFH_FULL 0xffffffffu
#define SFH_SPILL_EVERY 100u /* chunks between counter spills: 8 faces/lane/chunk keeps k_submerged's 10-bit fields < 1023 */
#define SFH_SURF_SPILL_EVERY (250u / SFH_SURF_FPT) /* ... and SFH_SURF_FPT faces/thread/chunk k_surface's 8-bit fields < 255 */

struct StepParams {
    const BodyDev* bodies;
    const uint32_t* item_off; // n_bodies + 1, exclusive scan of chunk counts of this kernel's kind
    uint32_t n_bodies;
    const OceanDev* ocean;
    WavesDev waves;
    void* partials;
    uint8_t* codes;            // debug: per-face codes, nullptr otherwise
    const uint64_t* code_base; // debug: first code index of each body
    uint32_t* seg_info;        // k_surface: [0] next segment to hand out (zeroed by k_scan_items), [1] items per segment
};

// k_surface hands out its work list in SEGMENTS of seg_info[1] consecutive items through an atomic counter: crossing bodies
// mix dry chunks (dropped after 32 bytes) with clipped and submerged ones, and a static split leaves SMs idle for a third
// of the kernel. At most SFH_SEG_PER_BLOCK segments per resident block, at least SFH_SEG_MIN items per segment.
#define SFH_SEG_PER_BLOCK 16u
#define SFH_SEG_MIN 1u

__device__ __forceinline__ uint32_t block_of_item(uint64_t w, uint64_t G, uint64_t W)
{
    return (uint32_t)(((w + 1) * G - 1) / W);
}

// first body whose item range contains w (item_off is non-decreasing, empty bodies are skipped)
__device__ __forceinline__ uint32_t body_of_item(const uint32_t* __restrict__ item_off, uint32_t n_bodies, uint32_t w)
{
    uint32_t lo = 0, hi = n_bodies; // invariant: item_off[lo] <= w < item_off[hi]
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (item_off[mid] <= w) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = dadd(v, __shfl_down_sync(SFH_FULL, v, off));
    return v;
}

template <int NACC>
__device__ __forceinline__ void block_reduce_store(double (&acc)[NACC], double (*red)[NACC], double* out)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
        double v = warp_sum(acc[i]);
        if (lane == 0) red[warp][i] = v;
        acc[i] = 0.0;
    }
    __syncthreads();
    if (threadIdx.x < NACC) {
        double s = red[0][threadIdx.x];
#pragma unroll
        for (int w = 1; w < SFH_BLOCK / 32; ++w) s = dadd(s, red[w][threadIdx.x]);
        out[threadIdx.x] = s;
    }
}

struct BodyRegs { // hot per-body constants
    F3 p, v, omega, cur;
};

// Damping block shared by both integrators (SolidEntity.cpp:1661-1683, 1771-1790). acc[0..11] = Fdq Tdq Fdf Tdf.
template <int CUR>
__device__ __forceinline__ uint32_t damping_terms(const BodyRegs& B, const sfh_current* s_cur, uint64_t cur_may, uint32_t n_cur, F3 fc,
                                                  F3 fn1, float A, double* acc, uint32_t& epack)
{
    uint32_t code = 0;
    // no Jet / Pipe field reaches this chunk: the sum of the enabled Uniform fields, narrowed once by k_prologue
    F3 fluid = (CUR && cur_may) ? fluid_velocity_masked(s_cur, cur_may, n_cur, fc) : B.cur;
    F3 r = fc - B.p;
    F3 vc = fluid - (B.v + cross(B.omega, r));
    float vc_n = dot(vc, fn1);
    F3 vn = fn1 * vc_n;
    F3 vt = vc - vn;
    if (vc_n < -1e-12f) {
        float vmag2 = length2(vc);
        F3 q = ((vc * ieee_sqrt1(vmag2)) * (-vc_n)) * A; // vc_n < -1e-12 puts |vc|^2 above 1e-24 > 2^-101
        F3 tq = cross(r, q);
        acc[0] = dadd(acc[0], (double)q.x);  acc[1] = dadd(acc[1], (double)q.y);  acc[2] = dadd(acc[2], (double)q.z);
        acc[3] = dadd(acc[3], (double)tq.x); acc[4] = dadd(acc[4], (double)tq.y); acc[5] = dadd(acc[5], (double)tq.z);
        epack += 1u << 10;
        code |= SFH_FACE_QUADRATIC;
    }
    float vmag2 = length2(vt);
    if (vmag2 > 1e-9f) {
        F3 s = vt * A;
        F3 ts = cross(r, s);
        acc[6] = dadd(acc[6], (double)s.x);  acc[7] = dadd(acc[7], (double)s.y);   acc[8] = dadd(acc[8], (double)s.z);
        acc[9] = dadd(acc[9], (double)ts.x); acc[10] = dadd(acc[10], (double)ts.y); acc[11] = dadd(acc[11], (double)ts.z);
        epack += 1u << 20;
        code |= SFH_FACE_SKIN;
    }
    return code;
}

__device__ __forceinline__ void load_body(BodyDev* sb, const BodyDev* __restrict__ src)
{
    const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
    uint32_t* d = reinterpret_cast<uint32_t*>(sb);
    for (uint32_t i = threadIdx.x; i < sizeof(BodyDev) / 4; i += SFH_BLOCK) d[i] = __ldg(s + i);
}

// current parameters and their fp32 pre-test records into shared memory (published by the caller's next block barrier)
__device__ __forceinline__ void stage_currents(sfh_current* s_cur, CurGuard* s_guard, const OceanDev* __restrict__ ocean, uint32_t n_cur)
{
    const uint32_t* s = reinterpret_cast<const uint32_t*>(ocean->currents);
    uint32_t* d = reinterpret_cast<uint32_t*>(s_cur);
    for (uint32_t i = threadIdx.x; i < n_cur * (uint32_t)(sizeof(sfh_current) / 4); i += blockDim.x) d[i] = __ldg(s + i);
    for (uint32_t i = threadIdx.x; i < n_cur; i += blockDim.x) s_guard[i] = make_cur_guard(ocean->currents[i]);
}

// 3 x 10-bit event counters (degenerate | quadratic << 10 | skin << 20) -> shared u32 counters
__device__ __forceinline__ void spill_epack(uint32_t& epack, uint32_t* s_cnt)
{
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        uint32_t v = __reduce_add_sync(SFH_FULL, (epack >> (10 * k)) & 1023u);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_cnt[k], v);
    }
    epack = 0;
}
__device__ __forceinline__ void spill_cpack(uint64_t& cpack, uint32_t* s_cls)
{
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        uint32_t v = __reduce_add_sync(SFH_FULL, (uint32_t)(cpack >> (8 * k)) & 255u);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_cls[k], v);
    }
    cpack = 0;
}

} // namespace sfh

#include "sfh_submerged.cuh" // k_submerged (two faces per thread, packed fp32)

namespace sfh {

// ---------------------------------------------------------------------------------------------
// k_surface — SolidEntity::ComputeHydrodynamicForcesSurface (SolidEntity.cpp:1337-1687)
// acc: [0..11] Fdq Tdq Fdf Tdf | [12..14] Fb | [15..17] Tb | [18] Swet | [19] Vsub | [20..22] CBsub
// ---------------------------------------------------------------------------------------------
#ifndef SFH_SURF_MIN_BLOCKS
#define SFH_SURF_MIN_BLOCKS (4 * 128 / SFH_BLOCK)
#endif
template <bool WAVES, int CUR, bool DEBUG>
__global__ void __launch_bounds__(SFH_BLOCK, SFH_SURF_MIN_BLOCKS) k_surface(StepParams P)
{
    __shared__ ChunkPipe s_pipe;
    __shared__ float4 s_v2[2][SFH_CH_VERTS]; // world x,y,z and depth, one buffer per staging buffer
    __shared__ BodyDev s_body;
    __shared__ double s_red[SFH_BLOCK / 32][SFH_NACC_SURF];
    __shared__ uint32_t s_cnt[12]; // [0..2] degenerate quadratic skin, [4..11] classes
    __shared__ sfh_current s_cur[CUR ? SFH_MAX_CURRENTS : 1];
    __shared__ CurGuard s_guard[CUR ? SFH_MAX_CURRENTS : 1];

    const uint32_t tid = threadIdx.x;
    // effective grid = min(launched blocks, work items): every participating block owns >= 1 item, so the
    // epilogue's [first block, last block] range of a body never contains a block that wrote no record
    __shared__ uint32_t s_seg;
    const uint32_t W = P.item_off[P.n_bodies], SC = P.seg_info[1];
    ItemWalk it;
    it.s = 0;
    it.parity = 0;

    const uint32_t n_cur = CUR ? P.ocean->n_currents : 0;
    if (CUR) stage_currents(s_cur, s_guard, P.ocean, n_cur);
    if (tid < 12) s_cnt[tid] = 0;
    const WavesDev wv = P.waves;
    const float hmax = WAVES ? __fadd_ru(__uint_as_float(__ldg(wv.hmax)), __uint_as_float(__ldg(wv.hmax + 1))) : 0.f;
    const WaveConst wc = WAVES ? wave_const(wv) : WaveConst{};

    double acc[SFH_NACC_SURF];
#pragma unroll
    for (int i = 0; i < SFH_NACC_SURF; ++i) acc[i] = 0.0;
    uint32_t epack = 0, since_spill = 0;
    uint64_t cpack = 0;

    const PipeAddr pa = pipe_init(s_pipe); // includes a block barrier (publishes s_cur / s_guard / s_cnt)
    uint32_t seg = 0;
    BodyRegs B;
    F3 p0;
    bool buoy, damp;
    float zrow;  // >= |row z of T_C|
    float rnorm; // >= the 2-norm of the linear part of T_C (Frobenius norm)

    auto flush = [&](uint32_t body) {
        spill_epack(epack, s_cnt);
        spill_cpack(cpack, s_cnt + 4);
        since_spill = 0;
        PartialSurf* rec = reinterpret_cast<PartialSurf*>(P.partials) + (seg + body); // one record per (segment, body)
        __syncthreads();
        block_reduce_store<SFH_NACC_SURF>(acc, s_red, rec->a);
        if (tid < 3) (&rec->n_degenerate)[tid] = s_cnt[tid];
        if (tid >= 4 && tid < 12) rec->cls[tid - 4] = s_cnt[tid];
        __syncthreads();
        if (tid < 12) s_cnt[tid] = 0;
    };
    auto fetch_regs = [&]() {
        B.p = F3{s_body.p[0], s_body.p[1], s_body.p[2]};
        B.v = F3{s_body.v[0], s_body.v[1], s_body.v[2]};
        B.omega = F3{s_body.omega[0], s_body.omega[1], s_body.omega[2]};
        B.cur = F3{s_body.cur[0], s_body.cur[1], s_body.cur[2]};
        p0 = F3{B.p.x, B.p.y, 0.f}; // SolidEntity.cpp:1332-1334
        buoy = (s_body.flags & BF_BUOYANCY) != 0;
        damp = (s_body.flags & BF_DAMPING) != 0;
        zrow = sqrtf(s_body.R[6] * s_body.R[6] + s_body.R[7] * s_body.R[7] + s_body.R[8] * s_body.R[8]) * 1.000001f;
        float fro = 0.f;
        for (int k = 0; k < 9; ++k) fro += s_body.R[k] * s_body.R[k];
        rnorm = sqrtf(fro) * 1.000001f;
    };

    for (;;) { // segments
    __syncthreads(); // the previous segment is flushed and s_seg has been read by everyone
    if (tid == 0) s_seg = atomicAdd(&P.seg_info[0], 1u);
    __syncthreads();
    seg = s_seg;
    if ((uint64_t)seg * SC >= W) break;
    it.w = seg * SC;
    it.w1 = min(it.w + SC, W);
    it.b = body_of_item(P.item_off, P.n_bodies, it.w);
    it.b_begin = P.item_off[it.b];
    it.b_end = P.item_off[it.b + 1];
    load_body(&s_body, P.bodies + it.b);
    __syncthreads();
    if (tid == 0) walk_issue_first(pa, it, s_body);
    fetch_regs();

    for (;;) { // chunks of the segment
        pipe_wait(pa, it.s, (it.parity >> it.s) & 1u);
        it.parity ^= 1u << it.s;
        const ChunkView cv = chunk_view(s_pipe.buf[it.s]);
        const uint32_t nf = cv.nf, b = it.b, c = it.w - it.b_begin;
        // Entirely above the highest possible water surface? Every vertex v of the chunk satisfies |v - centre| <= r in the mesh
        // frame, so its world z is at most (T_C * centre).z + |row z of T_C| * r (Cauchy-Schwarz; any linear T_C) and its depth
        // at most that + hmax (z points down, the surface is within +-hmax of z = 0); the last term covers the fp32 rounding
        // of both transforms. Then every face is class 0 and adds nothing (SolidEntity.cpp:1353-1360).
        const F3 cw = xform(s_body.R, s_body.t, cv.sphere.x, cv.sphere.y, cv.sphere.z);
        const float reach = fabsf(s_body.t[2]) + fabsf(cv.sphere.x) + fabsf(cv.sphere.y) + fabsf(cv.sphere.z) + cv.sphere.w;
        const bool dry = cw.z + zrow * cv.sphere.w + hmax + (4e-6f * reach + 1e-6f) < 0.f;
        // ... or entirely below the lowest possible surface: every vertex is under water whatever the waves do, every face is
        // class 7, and the per-vertex wave lookups (used for nothing but the classification then) can be skipped
        const bool deep = WAVES && cw.z - zrow * cv.sphere.w - hmax - (4e-6f * reach + 1e-6f) > 0.f;
        // which Jet / Pipe fields can reach this chunk at all (every face centroid lies inside the sphere)
        const uint64_t cur_may =
            CUR ? chunk_current_mask(s_guard, n_cur, cw, rnorm * cv.sphere.w + (4e-6f * (reach + fabsf(s_body.t[0]) + fabsf(s_body.t[1])) + 1e-6f)) : 0ull;

        const uint8_t* nsrc = nullptr;
        uint32_t nbytes = 0;
        // ONE block barrier per chunk. The vertex buffer alternates with the staging buffer, so the vertex phase of chunk c may
        // start while other warps still read chunk c-1 (they use the other pair of buffers); the barrier after the vertex phase
        // publishes s_v AND proves that every thread has left chunk c-1, and only then the copy of chunk c+1 overwrites it
        // (it still has the whole face phase of chunk c to land). With a barrier on either side of the vertex phase the kernel
        // was 2 % slower (0.541 -> 0.531 ms per 16 M faces on cfg 5).
        float4* const s_v = s_v2[it.s];
        bool under = !dry; // every vertex of the chunk is under water: !(depth < 0), the reference's own test (:1353)
        if (!dry) {
#pragma unroll
        for (int k = 0; k < SFH_SURF_FPT; ++k) {
            const uint32_t i = tid + k * SFH_BLOCK;
            if (i < cv.nv) {
                F3 pw = xform(s_body.R, s_body.t, cv.x[i], cv.y[i], cv.z[i]);
                // Ocean::GetDepth per vertex (SolidEntity.cpp:1349-1351): same bits for every incident face
                const float depth = deep ? 1.f : get_depth<WAVES>(wv, wc, pw.x, pw.y, pw.z); // only its sign matters for class 7
                s_v[i] = make_float4(pw.x, pw.y, pw.z, depth);
                under = under && !(depth < 0.f);
            }
        }
        // The pair path pays on a flat sea only: with waves its two centroid lookups (and, with Jet/Pipe, two field evaluations)
        // per thread cost more in registers and instruction fetch than the pairing saves (0.54 -> 0.63 ms on cfg 5).
        }
        const bool all_wet = __syncthreads_and(under) != 0 && !WAVES; // the chunk's one block barrier (dry chunks included)
        if (tid == 0) walk_issue_next(pa, it, P, cv.next_bytes, nsrc, nbytes);
        if (dry) { // block-uniform
#pragma unroll
            for (int k = 0; k < SFH_SURF_FPT; ++k) cpack += tid + k * SFH_BLOCK < nf ? 1ull : 0ull; // class 0, code 0 (the debug buffer is zero-filled)
        } else {

        if (all_wet) {
            // Every face of the chunk is class 7 (no clipping): the triangle branch of the reference (:1605-1687) for faces tid
            // and tid + 128 side by side, in the pair arithmetic of k_submerged (same per-face terms bit for bit).
#pragma unroll 1
            for (uint32_t tA = tid; tA < SFH_CH_FACES; tA += 2u * SFH_BLOCK) {
            const uint32_t tB = tA + SFH_BLOCK;
            const bool vA = tA < nf, vB = tB < nf;
            const uint32_t fA = vA ? cv.faces[tA] : 0u, fB = vB ? cv.faces[tB] : 0u;
            const float4 a1 = s_v[__byte_perm(fA, 0, 0x4440)], a2 = s_v[__byte_perm(fA, 0, 0x4441)], a3 = s_v[__byte_perm(fA, 0, 0x4442)];
            const float4 b1 = s_v[__byte_perm(fB, 0, 0x4440)], b2 = s_v[__byte_perm(fB, 0, 0x4441)], b3 = s_v[__byte_perm(fB, 0, 0x4442)];
            const V3P q1 = {pk(a1.x, b1.x), pk(a1.y, b1.y), pk(a1.z, b1.z)};
            const V3P q2 = {pk(a2.x, b2.x), pk(a2.y, b2.y), pk(a2.z, b2.z)};
            const V3P q3 = {pk(a3.x, b3.x), pk(a3.y, b3.y), pk(a3.z, b3.z)};
            cpack += ((vA ? 1ull : 0ull) + (vB ? 1ull : 0ull)) << 56; // class 7
            if (buoy) { // tetrahedron (p0, p1, p2, p3) of the submerged volume (SolidEntity.cpp:1612-1618)
                const V3P o0 = {pk1(p0.x), pk1(p0.y), pk1(p0.z)};
                const V3P t1 = q1 - o0, t2 = q2 - o0, t3 = q3 - o0;
                const P2 v6 = dot(t1, cross(t2, t3));
                if (!WAVES) {
                    const V3P m = (((t1 + t2) + t3) * pk1(0.25f)) * v6;
                    acc3_if(vA, acc + 20, lo(m.x), lo(m.y), lo(m.z));
                    acc3_if(vB, acc + 20, hi(m.x), hi(m.y), hi(m.z));
                }
                acc1_if(vA, acc[19], lo(v6));
                acc1_if(vB, acc[19], hi(v6));
            }
            const V3P fn = cross(q2 - q1, q3 - q1);
            const P2 len2 = dot(fn, fn);
            const bool dA = lo(len2) < 1e-12f, dB = hi(len2) < 1e-12f;
            const bool mA = vA && !dA, mB = vB && !dB;
            epack += (vA && dA ? 1u : 0u) + (vB && dB ? 1u : 0u);
            uint32_t code = 0;
            if (__any_sync(SFH_FULL, mA | mB)) {
                const P2 len = ieee_sqrt2(len2); // lanes with len2 < 1e-12 compute garbage that is masked out
                const P2 A = len * pk1(0.5f);
                const V3P csum = (q1 + q2) + q3;
                V3P fn1, fc;
                if ((min_key(fn) >= SFH_NUM_KEY) & (min_key(csum) >= SFH_TINY_KEY)) {
                    fn1 = normalize_fast(fn, len);
                    fc = V3P{ieee_div3_2(csum.x), ieee_div3_2(csum.y), ieee_div3_2(csum.z)};
                } else {
                    normalize_centroid_slow(fn, len, csum, fn1, fc);
                }
                const V3P pp = {pk1(B.p.x), pk1(B.p.y), pk1(B.p.z)};
                if (WAVES && buoy) { // pressure buoyancy at the centroid (SolidEntity.cpp:1650-1658)
                    const float dcA = get_depth<WAVES>(wv, wc, lo(fc.x), lo(fc.y), lo(fc.z));
                    const float dcB = get_depth<WAVES>(wv, wc, hi(fc.x), hi(fc.y), hi(fc.z));
                    const P2 nA = neg(A);
                    const V3P Fbi = (fn1 * nA) * pk(dcA, dcB); // ((-fn1) * A) * depthc: negating A instead of fn1 flips the same sign bits
                    const V3P Tbi = cross(fc - pp, Fbi);
                    acc6_if(mA, acc + 12, lo(Fbi.x), lo(Fbi.y), lo(Fbi.z), lo(Tbi.x), lo(Tbi.y), lo(Tbi.z));
                    acc6_if(mB, acc + 12, hi(Fbi.x), hi(Fbi.y), hi(Fbi.z), hi(Tbi.x), hi(Tbi.y), hi(Tbi.z));
                }
                if (damp) {
                    BodyP2 BP;
                    BP.p = pp;
                    BP.v = V3P{pk1(B.v.x), pk1(B.v.y), pk1(B.v.z)};
                    BP.omega = V3P{pk1(B.omega.x), pk1(B.omega.y), pk1(B.omega.z)};
                    BP.cur = V3P{pk1(B.cur.x), pk1(B.cur.y), pk1(B.cur.z)};
                    code = damping_pair<CUR>(BP, s_cur, cur_may, n_cur, fc, fn1, A, mA, mB, acc, epack);
                }
                acc1_if(mA, acc[18], lo(A)); // Swet (SolidEntity.cpp:1686)
                acc1_if(mB, acc[18], hi(A));
            }
            if (DEBUG) {
                const uint32_t* of = s_body.orig_face + (size_t)c * SFH_CH_FACES;
                if (vA) P.codes[P.code_base[b] + of[tA]] = (uint8_t)(7u | (dA ? SFH_FACE_DEGENERATE : 0u) | (code & 255u));
                if (vB) P.codes[P.code_base[b] + of[tB]] = (uint8_t)(7u | (dB ? SFH_FACE_DEGENERATE : 0u) | (code >> 8));
            }
            } // pairs
        } else {

#pragma unroll 1 // interleaving the two faces costs more in instruction fetch than it gains in ILP (0.72 -> 0.87 ms)
        for (int k = 0; k < SFH_SURF_FPT; ++k) {
            const uint32_t slot = tid + k * SFH_BLOCK;
            if (slot >= nf) continue;
            const uint32_t fp = cv.faces[slot];
            const float4 a4 = s_v[fp & 255u], b4 = s_v[(fp >> 8) & 255u], c4 = s_v[(fp >> 16) & 255u];
            F3 p1 = {a4.x, a4.y, a4.z}, p2 = {b4.x, b4.y, b4.z}, p3 = {c4.x, c4.y, c4.z};
            const float d0 = a4.w, d1 = b4.w, d2 = c4.w;
            const bool a0 = d0 < 0.f, a1 = d1 < 0.f, a2 = d2 < 0.f;
            const uint32_t cls = (a0 ? 0u : 4u) | (a1 ? 0u : 2u) | (a2 ? 0u : 1u);
            cpack += 1ull << (8 * cls);
            uint32_t code = cls;
            if (cls != 0) {
                // waterline point on the edge from the submerged vertex pu (du >= 0) to the dry vertex pa (da < 0)
                auto clip = [](F3 pa, F3 pu, float da, float du) { return pu + (pa - pu) * fdiv(du, fadd(fabsf(da), du)); };
                // tetrahedron (p0, a, b, c) of the submerged volume (SolidEntity.cpp:1371-1377 and twins)
                // Vsub / CBsub only reach an output through the buoyancy block (SolidEntity.cpp:1690-1708), and CBsub only on a flat
                // sea (:1703-1706): they are not accumulated where the reference computes them in vain.
                auto tetra = [&](F3 pa, F3 pb, F3 pc) {
                    if (!buoy) return;
                    float v6 = dot(pa, cross(pb, pc));
                    if (!WAVES) {
                        F3 cg = ((pa + pb) + pc) * 0.25f; // "/ 4.f": an exact scaling either way
                        F3 m = cg * v6;
                        acc[20] = dadd(acc[20], (double)m.x); acc[21] = dadd(acc[21], (double)m.y); acc[22] = dadd(acc[22], (double)m.z);
                    }
                    acc[19] = dadd(acc[19], (double)v6);
                };
                F3 p4 = p1, fv1, fv2, fv3, fv4;
                bool is_quad = false;
                switch (cls) {
                case 1: p1 = clip(p1, p3, d0, d2); p2 = clip(p2, p3, d1, d2); break;
                case 2: p1 = clip(p1, p2, d0, d1); p3 = clip(p3, p2, d2, d1); break;
                case 3: p4 = clip(p1, p3, d0, d2); p1 = clip(p1, p2, d0, d1); is_quad = true; break;
                case 4: p2 = clip(p2, p1, d1, d0); p3 = clip(p3, p1, d2, d0); break;
                case 5: p4 = clip(p2, p3, d1, d2); p2 = clip(p2, p1, d1, d0); is_quad = true; break;
                case 6: p4 = clip(p3, p1, d2, d0); p3 = clip(p3, p2, d2, d1); is_quad = true; break;
                default: break;
                }
                const F3 p01 = p1 - p0, p02 = p2 - p0, p03 = p3 - p0;
                tetra(p01, p02, p03);
                F3 fc;
                if (is_quad) {
                    const F3 p04 = p4 - p0;
                    if (cls == 5) { tetra(p02, p04, p03); fv2 = p3 - p1; }
                    else          { tetra(p01, p03, p04); fv2 = p4 - p1; }
                    fv1 = p2 - p1; fv3 = p2 - p3; fv4 = p4 - p3;
                    fc = (((p1 + p2) + p3) + p4) * 0.25f;
                } else {
                    fv1 = p2 - p1; fv2 = p3 - p1;
                    const F3 cs = (p1 + p2) + p3;
                    fc = F3{div3(cs.x), div3(cs.y), div3(cs.z)};
                }
                F3 fn = cross(fv1, fv2);
                float len = length2(fn);
                if (len < 1e-12f) {
                    epack += 1u;
                    code |= SFH_FACE_DEGENERATE;
                } else {
                    len = ieee_sqrt1(len); // len >= 1e-12
                    // fn / len: |fn_i| <= len, len >= 1e-6 > 2^-20; components below 2^-36 (not zero) take the hardware division
                    F3 fn1;
                    if (min(min(tiny_key(fn.x), tiny_key(fn.y)), tiny_key(fn.z)) >= SFH_NUM_KEY) {
                        const float r = refined_rcp1(len);
                        fn1 = F3{ieee_div_r1(fn.x, len, r), ieee_div_r1(fn.y, len, r), ieee_div_r1(fn.z, len, r)};
                    } else {
                        fn1 = fn / len;
                    }
                    float A; // "/ 2.f": exact scaling
                    if (is_quad) {
                        const float l2 = length2(cross(fv3, fv4));
                        A = fmul(fadd(len, l2 >= 0x1p-100f ? ieee_sqrt1(l2) : fsqrt(l2)), 0.5f);
                    } else {
                        A = fmul(len, 0.5f);
                    }
                    if (WAVES && buoy) { // SolidEntity.cpp:1650-1658
                        float depthc = get_depth<WAVES>(wv, wc, fc.x, fc.y, fc.z);
                        F3 Fbi = ((-fn1) * A) * depthc;
                        F3 Tbi = cross(fc - B.p, Fbi);
                        acc[12] = dadd(acc[12], (double)Fbi.x); acc[13] = dadd(acc[13], (double)Fbi.y); acc[14] = dadd(acc[14], (double)Fbi.z);
                        acc[15] = dadd(acc[15], (double)Tbi.x); acc[16] = dadd(acc[16], (double)Tbi.y); acc[17] = dadd(acc[17], (double)Tbi.z);
                    }
                    if (damp) code |= damping_terms<CUR>(B, s_cur, cur_may, n_cur, fc, fn1, A, acc, epack);
                    acc[18] = dadd(acc[18], (double)A); // SolidEntity.cpp:1686
                }
            }
            if (DEBUG) P.codes[P.code_base[b] + s_body.orig_face[(size_t)c * SFH_CH_FACES + slot]] = (uint8_t)code;
        }
        } // !all_wet
        } // !dry
        if (++since_spill == SFH_SURF_SPILL_EVERY) { spill_epack(epack, s_cnt); spill_cpack(cpack, s_cnt + 4); since_spill = 0; }

        if (++it.w >= it.w1) break;
        it.s ^= 1;
        it.src = nsrc;
        it.bytes = nbytes;
        if (it.w >= it.b_end) {
            flush(it.b);
            do { ++it.b; it.b_end = P.item_off[it.b + 1]; } while (it.w >= it.b_end);
            it.b_begin = P.item_off[it.b];
            load_body(&s_body, P.bodies + it.b);
            __syncthreads();
            fetch_regs();
        }
    }
    it.s ^= 1; // the next segment's first copy goes to the other staging buffer
    flush(it.b);
    } // segments
}

// ---------------------------------------------------------------------------------------------
// k_prologue — per-body gates and narrowing, then the two work-list scans. One block of 1024 threads.
// ---------------------------------------------------------------------------------------------
struct PrologueParams {
    const sfh_body_state* states;
    const MeshDev* meshes;
    uint32_t n_meshes;
    uint32_t n_bodies;
    const OceanDev* ocean;
    WavesDev waves;
    BodyDev* bodies;
    uint32_t* off_sub;  // n_bodies + 1
    uint32_t* off_surf; // n_bodies + 1
    uint32_t* status;   // [0] |= 1 if a body referenced an unknown mesh
    uint32_t aero;      // sfh_aero_step: every body runs the (aero variant of the) submerged face loop
};

// SolidEntity::CheckBodyFluidPosition (SolidEntity.cpp:1239-1261); each corner narrowed to fp32 by
// Ocean::GetDepth(const Vector3&) (Ocean.cpp:170-173)
__device__ inline uint32_t check_body_fluid_position(bool waves, const WavesDev& wv, const double* mn, const double* mx)
{
    const double d[3] = {dsub(mx[0], mn[0]), dsub(mx[1], mn[1]), dsub(mx[2], mn[2])};
    const double c[8][3] = {{mn[0], mn[1], mn[2]},
                            {mx[0], mx[1], mx[2]},
                            {dadd(mn[0], d[0]), dadd(mn[1], 0.0), dadd(mn[2], 0.0)},
                            {dadd(mn[0], 0.0), dadd(mn[1], d[1]), dadd(mn[2], 0.0)},
                            {dadd(mn[0], d[0]), dadd(mn[1], d[1]), dadd(mn[2], 0.0)},
                            {dadd(mn[0], 0.0), dadd(mn[1], 0.0), dadd(mn[2], d[2])},
                            {dadd(mn[0], d[0]), dadd(mn[1], 0.0), dadd(mn[2], d[2])},
                            {dadd(mn[0], 0.0), dadd(mn[1], d[1]), dadd(mn[2], d[2])}};
    uint32_t under = 0;
    for (int k = 0; k < 8; ++k) {
        const float x = (float)c[k][0], y = (float)c[k][1], z = (float)c[k][2];
        const float depth = waves ? get_depth<true>(wv, x, y, z) : z;
        if ((double)depth > 0.0) ++under;
    }
    return under == 0 ? SFH_POS_OUTSIDE : (under == 8 ? SFH_POS_INSIDE : SFH_POS_CROSSING);
}

__device__ __forceinline__ void prologue_body(const PrologueParams& P, uint32_t i)
{
    const OceanDev* __restrict__ oc = P.ocean;
    const bool waves = oc->has_waves != 0;

    // position-independent part of Ocean::GetFluidVelocity (Ocean.cpp:183-201)
    float cur[3] = {0.f, 0.f, 0.f};
    if (oc->currents_enabled) {
        D3 fv = {0.0, 0.0, 0.0};
        for (uint32_t k = 0; k < oc->n_currents; ++k)
            if (oc->currents[k].enabled && oc->currents[k].type == SFH_CURRENT_UNIFORM)
                fv = fv + D3{oc->currents[k].a[0], oc->currents[k].a[1], oc->currents[k].a[2]};
        cur[0] = (float)fv.x; cur[1] = (float)fv.y; cur[2] = (float)fv.z;
    }

    const sfh_body_state& st = P.states[i];
    BodyDev bd;
    uint32_t pos = st.mode;
    if (pos == SFH_POS_AUTO) pos = check_body_fluid_position(waves, P.waves, st.aabb_min, st.aabb_max);
    if (pos > SFH_POS_CROSSING) pos = SFH_POS_OUTSIDE;
    const bool dispatch = (st.flags & SFH_F_DISPATCH) != 0;
    bool damping = (st.flags & SFH_F_DAMPING) != 0, buoyancy = (st.flags & SFH_F_BUOYANCY) != 0;
    uint32_t kind = KIND_NONE;
    if (P.aero) { // SolidEntity::ComputeAerodynamicForces has no position test and no settings (SolidEntity.cpp:1852-1923)
        pos = SFH_POS_OUTSIDE;
        kind = KIND_SUBMERGED;
    } else if (pos == SFH_POS_INSIDE) {
        // raw integrator: always runs (it has no settings); dispatcher: only `if(settings.dampingForces)`
        if (!dispatch || damping) kind = KIND_SUBMERGED;
        damping = true;
    } else if (pos == SFH_POS_CROSSING) {
        if (dispatch && !(st.flags & SFH_F_BUOYANT)) buoyancy = false; // SolidEntity.cpp:1844
        kind = KIND_SURFACE;
    }
    MeshDev m = {nullptr, nullptr, nullptr, 0, 0};
    bool known = st.mesh_id < P.n_meshes;
    if (known) {
        m = P.meshes[st.mesh_id];
        if (m.n_faces == SFH_MESH_TOMBSTONE) { known = false; m.n_faces = 0; } // released by sfh_mesh_release
    }
    if (!known && kind != KIND_NONE) atomicOr(P.status, 1u);
    // glMatrixFromTransform / glVectorFromVector (OpenGLDataStructs.cpp:148-171)
    for (int k = 0; k < 9; ++k) bd.R[k] = (float)st.T_C[k];
    for (int k = 0; k < 3; ++k) {
        bd.t[k] = (float)st.T_C[9 + k];
        bd.p[k] = (float)st.T_CG[9 + k];
        bd.v[k] = (float)st.v[k];
        bd.omega[k] = (float)st.omega[k];
        bd.cur[k] = cur[k];
    }
    bd.flags = (damping ? BF_DAMPING : 0u) | (buoyancy ? BF_BUOYANCY : 0u);
    bd.kind = kind | (pos << 8);
    bd.n_chunks = (kind != KIND_NONE) ? m.n_chunks : 0u;
    bd.n_faces = m.n_faces;
    bd.chunks = m.chunks;
    bd.first_bytes = m.n_chunks ? (m.chunk_off[1] - m.chunk_off[0]) * 16u : 0u;
    bd.pad = 0;
    bd.chunk_off = m.chunk_off;
    bd.orig_face = m.orig_face;
    P.bodies[i] = bd;
    // chunk counts; k_scan_items turns them into the two work lists
    P.off_sub[i] = (kind == KIND_SUBMERGED) ? bd.n_chunks : 0u;
    P.off_surf[i] = (kind == KIND_SURFACE) ? bd.n_chunks : 0u;
}

__global__ void __launch_bounds__(128) k_prologue(PrologueParams P)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P.n_bodies) prologue_body(P, i);
}

// In-place exclusive scan of the per-body chunk counts (n_bodies entries -> n_bodies + 1 offsets). One block.
__device__ __forceinline__ void scan_items(uint32_t* off_sub, uint32_t* off_surf, uint32_t n_bodies, uint32_t* seg_info, uint32_t grid_surf)
{
    __shared__ uint32_t s_warp[2][32];
    __shared__ uint32_t s_carry[2];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 2) s_carry[tid] = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_bodies; base += blockDim.x) {
        const uint32_t i = base + tid;
        const uint32_t n_sub = i < n_bodies ? off_sub[i] : 0u, n_surf = i < n_bodies ? off_surf[i] : 0u;
        uint32_t xs = n_sub, xf = n_surf;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            uint32_t ys = __shfl_up_sync(SFH_FULL, xs, off), yf = __shfl_up_sync(SFH_FULL, xf, off);
            if (lane >= (uint32_t)off) { xs += ys; xf += yf; }
        }
        if (lane == 31) { s_warp[0][warp] = xs; s_warp[1][warp] = xf; }
        __syncthreads();
        if (warp == 0) {
            uint32_t ws = s_warp[0][lane], wf = s_warp[1][lane];
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                uint32_t ys = __shfl_up_sync(SFH_FULL, ws, off), yf = __shfl_up_sync(SFH_FULL, wf, off);
                if (lane >= (uint32_t)off) { ws += ys; wf += yf; }
            }
            s_warp[0][lane] = ws; s_warp[1][lane] = wf; // inclusive over warps
        }
        __syncthreads();
        const uint32_t pre_s = s_carry[0] + (warp ? s_warp[0][warp - 1] : 0u) + xs - n_sub;
        const uint32_t pre_f = s_carry[1] + (warp ? s_warp[1][warp - 1] : 0u) + xf - n_surf;
        if (i < n_bodies) { off_sub[i] = pre_s; off_surf[i] = pre_f; }
        __syncthreads();
        if (tid == 0) { s_carry[0] += s_warp[0][31]; s_carry[1] += s_warp[1][31]; }
        __syncthreads();
    }
    if (tid == 0) {
        off_sub[n_bodies] = s_carry[0];
        off_surf[n_bodies] = s_carry[1];
        // k_surface's segment size: at most SFH_SEG_PER_BLOCK segments per resident block
        const uint32_t cap = max(grid_surf, 1u) * SFH_SEG_PER_BLOCK;
        seg_info[0] = 0;
        seg_info[1] = max(SFH_SEG_MIN, (s_carry[1] + cap - 1) / cap);
    }
}
__global__ void __launch_bounds__(1024) k_scan_items(uint32_t* off_sub, uint32_t* off_surf, uint32_t n_bodies, uint32_t* seg_info,
                                                     uint32_t grid_surf)
{
    scan_items(off_sub, off_surf, n_bodies, seg_info, grid_surf);
}
// Steps of at most 1024 bodies: prologue and scan in one launch (single-body steps are launch-latency bound)
__global__ void __launch_bounds__(1024) k_prologue_scan(PrologueParams P, uint32_t* seg_info, uint32_t grid_surf)
{
    if (threadIdx.x < P.n_bodies) prologue_body(P, threadIdx.x);
    __syncthreads();
    scan_items(P.off_sub, P.off_surf, P.n_bodies, seg_info, grid_surf);
}

// ---------------------------------------------------------------------------------------------
// k_epilogue — one warp per body: reduce partial records in block order, finalise, correct.
// ---------------------------------------------------------------------------------------------
struct EpilogueParams {
    const sfh_body_state* states;
    const BodyDev* bodies;
    uint32_t n_bodies;
    const OceanDev* ocean;
    const uint32_t* off_sub;
    const uint32_t* off_surf;
    const PartialSub* part_sub;
    const PartialSurf* part_surf;
    uint32_t grid_sub;
    const uint32_t* seg_info; // [1] = items per k_surface segment
    uint32_t aero;
    sfh_info* out;
    // zero-copy steps (few bodies: `states` and `out` are pinned HOST memory, no copy nodes): the last block to finish
    // publishes the status word and a sequence number to the host, which spins on it instead of synchronising the stream
    uint32_t* zc_host;  // pinned: [0] status, [1] sequence number; nullptr = not a zero-copy step
    uint32_t* zc_dev;   // device: [0] status flags (k_prologue), [1] blocks finished, [4] sequence counter
};

// btMatrix3x3::inverse (cofactors / determinant) as three rows, and btVector3::safeNormalize of M^-1 * X
struct Inv3 { D3 r0, r1, r2; };
__device__ inline Inv3 basis_inverse(const double* B)
{
#define EL(r, c) B[3 * (r) + (c)]
#define COFAC(r1, c1, r2, c2) dsub(dmul(EL(r1, c1), EL(r2, c2)), dmul(EL(r1, c2), EL(r2, c1)))
    const D3 co = {COFAC(1, 1, 2, 2), COFAC(1, 2, 2, 0), COFAC(1, 0, 2, 1)};
    const double det = ddot(D3{EL(0, 0), EL(0, 1), EL(0, 2)}, co);
    const double s = ddiv(1.0, det);
    Inv3 I;
    I.r0 = D3{dmul(co.x, s), dmul(COFAC(0, 2, 2, 1), s), dmul(COFAC(0, 1, 1, 2), s)};
    I.r1 = D3{dmul(co.y, s), dmul(COFAC(0, 0, 2, 2), s), dmul(COFAC(0, 2, 1, 0), s)};
    I.r2 = D3{dmul(co.z, s), dmul(COFAC(0, 1, 2, 0), s), dmul(COFAC(0, 0, 1, 1), s)};
#undef EL
#undef COFAC
    return I;
}
__device__ inline D3 to_frame_normalized(const Inv3& I, const double* X)
{
    const D3 v = {X[0], X[1], X[2]};
    D3 q = {ddot(I.r0, v), ddot(I.r1, v), ddot(I.r2, v)};
    const double l2 = ddot(q, q);
    if (l2 >= 2.2204460492503131e-16 * 2.2204460492503131e-16) q = q * ddiv(1.0, __dsqrt_rn(l2));
    else q = D3{1.0, 0.0, 0.0};
    return q;
}

// SolidEntity::CorrectHydrodynamicForces (SolidEntity.cpp:1263-1287)
__device__ inline void correct_forces(double density, double* Fdq, double* Tdq, double* Fdf, double* Tdf, const double* Cd,
                                      const double* Cf, const double* B)
{
    const Inv3 I = basis_inverse(B);
    auto coeff = [&](const double* X, const double* C) {
        const D3 q = to_frame_normalized(I, X);
        return dadd(dadd(dmul(fabs(q.x), C[0]), dmul(fabs(q.y), C[1])), dmul(fabs(q.z), C[2]));
    };
    double k;
    k = dmul(dmul(0.5, density), coeff(Fdq, Cd)); Fdq[0] = dmul(Fdq[0], k); Fdq[1] = dmul(Fdq[1], k); Fdq[2] = dmul(Fdq[2], k);
    k = dmul(dmul(0.5, density), coeff(Tdq, Cd)); Tdq[0] = dmul(Tdq[0], k); Tdq[1] = dmul(Tdq[1], k); Tdq[2] = dmul(Tdq[2], k);
    k = dmul(density, coeff(Fdf, Cf)); Fdf[0] = dmul(Fdf[0], k); Fdf[1] = dmul(Fdf[1], k); Fdf[2] = dmul(Fdf[2], k);
    k = dmul(density, coeff(Tdf, Cf)); Tdf[0] = dmul(Tdf[0], k); Tdf[1] = dmul(Tdf[1], k); Tdf[2] = dmul(Tdf[2], k);
}

// SolidEntity::CorrectAerodynamicForces (SolidEntity.cpp:1925-1956): B = T_CG2H.getBasis(); Cd = the proxy's vector,
// used only for CYLINDER / ELLIPSOID (use_cd)
__device__ inline void correct_aero(double* Fda, double* Tda, const double* Cd, const double* B, bool use_cd)
{
    const D3 Fdan = to_frame_normalized(basis_inverse(B), Fda);
    double cor = 1.0;
    if (use_cd) cor = dadd(dadd(dmul(Cd[0], Fdan.x), dmul(Cd[1], Fdan.y)), dmul(Cd[2], Fdan.z));
    cor = fabs(cor);
    for (int k = 0; k < 3; ++k) { Fda[k] = dmul(Fda[k], cor); Tda[k] = dmul(Tda[k], cor); }
}

__device__ __forceinline__ void epilogue_body(const EpilogueParams& P, uint32_t b, uint32_t lane);

__global__ void __launch_bounds__(128) k_epilogue(EpilogueParams P)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b < P.n_bodies) epilogue_body(P, b, lane);
    if (P.zc_host) {
        __threadfence_system(); // this thread's record (if it wrote one) is visible to the host before the flag can be
        __syncthreads();
        if (threadIdx.x == 0 && atomicAdd(&P.zc_dev[1], 1u) + 1u == gridDim.x) {
            P.zc_dev[1] = 0;
            const uint32_t seq = P.zc_dev[4] + 1u;
            P.zc_dev[4] = seq;
            P.zc_host[0] = P.zc_dev[0];
            __threadfence_system();
            *reinterpret_cast<volatile uint32_t*>(P.zc_host + 1) = seq;
        }
    }
}

__device__ __forceinline__ void epilogue_body(const EpilogueParams& P, uint32_t b, uint32_t lane)
{
    const BodyDev& bd = P.bodies[b];
    const uint32_t kind = bd.kind & 255u, pos = bd.kind >> 8;

    double acc[SFH_NACC_SURF];
#pragma unroll
    for (int i = 0; i < SFH_NACC_SURF; ++i) acc[i] = 0.0;
    uint32_t cnt[11]; // degenerate quadratic skin | cls[8]
#pragma unroll
    for (int i = 0; i < 11; ++i) cnt[i] = 0;

    if (kind == KIND_SUBMERGED) {
        const uint32_t o0 = P.off_sub[b], o1 = P.off_sub[b + 1];
        if (o1 > o0) {
            const uint64_t W = P.off_sub[P.n_bodies], G = min((uint64_t)P.grid_sub, W);
            const uint32_t k0 = block_of_item(o0, G, W), k1 = block_of_item(o1 - 1, G, W);
            for (uint32_t k = k0 + lane; k <= k1; k += 32) {
                const PartialSub& r = P.part_sub[k + b];
#pragma unroll
                for (int i = 0; i < SFH_NACC_SUB; ++i) acc[i] = dadd(acc[i], r.a[i]);
                cnt[0] += r.n_degenerate; cnt[1] += r.n_quadratic; cnt[2] += r.n_skin;
            }
        }
    } else if (kind == KIND_SURFACE) {
        const uint32_t o0 = P.off_surf[b], o1 = P.off_surf[b + 1];
        if (o1 > o0) {
            const uint32_t SC = P.seg_info[1], k0 = o0 / SC, k1 = (o1 - 1) / SC; // segments that overlap this body's items
            for (uint32_t k = k0 + lane; k <= k1; k += 32) {
                const PartialSurf& r = P.part_surf[k + b];
#pragma unroll
                for (int i = 0; i < SFH_NACC_SURF; ++i) acc[i] = dadd(acc[i], r.a[i]);
                cnt[0] += r.n_degenerate; cnt[1] += r.n_quadratic; cnt[2] += r.n_skin;
#pragma unroll
                for (int i = 0; i < 8; ++i) cnt[3 + i] += r.cls[i];
            }
        }
    }
    if (kind != KIND_NONE) {
#pragma unroll
        for (int i = 0; i < SFH_NACC_SURF; ++i) acc[i] = warp_sum(acc[i]);
#pragma unroll
        for (int i = 0; i < 11; ++i) cnt[i] = __reduce_add_sync(SFH_FULL, cnt[i]);
    }
    if (lane != 0) return;

    const sfh_body_state& st = P.states[b];
    const OceanDev& oc = *P.ocean;
    const bool dispatch = (st.flags & SFH_F_DISPATCH) != 0;
    sfh_info o;
    for (int k = 0; k < 3; ++k) { o.Fb[k] = 0; o.Tb[k] = 0; o.Fdq[k] = 0; o.Tdq[k] = 0; o.Fdf[k] = 0; o.Tdf[k] = 0; }
    o.Swet = 0; o.Vsub = 0; o.wrote_buoyancy = 0; o.wrote_damping = 0; o.position = pos; o.reserved = 0;
    o.n_faces = (kind != KIND_NONE) ? bd.n_faces : 0u;
    o.n_degenerate = cnt[0]; o.n_quadratic = cnt[1]; o.n_skin = cnt[2];
    for (int k = 0; k < 8; ++k) o.class_count[k] = cnt[3 + k];

    if (P.aero) { // SolidEntity.cpp:1920-1922, then CorrectAerodynamicForces
        const double k = dmul(0.5, oc.density);
        for (int c = 0; c < 3; ++c) { o.Fdq[c] = dmul(acc[c], k); o.Tdq[c] = dmul(acc[3 + c], k); }
        o.wrote_damping = 1;
        o.class_count[7] = o.n_faces;
        if (dispatch) correct_aero(o.Fdq, o.Tdq, st.Cd, st.T_O_basis, (st.flags & SFH_F_AERO_CD) != 0);
        P.out[b] = o;
        return;
    }
    if (pos == SFH_POS_OUTSIDE) { // SolidEntity.cpp:1810-1821
        o.wrote_buoyancy = 1; o.wrote_damping = 1;
    } else if (pos == SFH_POS_INSIDE) {
        if (kind == KIND_SUBMERGED) { // SolidEntity.cpp:1793-1796
            for (int k = 0; k < 3; ++k) { o.Fdq[k] = acc[k]; o.Tdq[k] = acc[3 + k]; o.Fdf[k] = acc[6 + k]; o.Tdf[k] = acc[9 + k]; }
            o.wrote_damping = 1;
            o.class_count[7] = o.n_faces;
        }
        if (dispatch) { // SolidEntity.cpp:1828-1841
            if (st.flags & SFH_F_BUOYANT) {
                const double s = dmul(-st.volume, oc.density);
                const D3 Fb = {dmul(0.0, s), dmul(0.0, s), dmul(oc.gravity_z, s)};
                const double* T = st.T_CG;
                const D3 pcb = {st.P_CB[0], st.P_CB[1], st.P_CB[2]}, org = {T[9], T[10], T[11]};
                const D3 wcb = {dadd(ddot(D3{T[0], T[1], T[2]}, pcb), org.x), dadd(ddot(D3{T[3], T[4], T[5]}, pcb), org.y),
                                dadd(ddot(D3{T[6], T[7], T[8]}, pcb), org.z)};
                const D3 Tb = dcross(wcb - org, Fb);
                o.Fb[0] = Fb.x; o.Fb[1] = Fb.y; o.Fb[2] = Fb.z; o.Tb[0] = Tb.x; o.Tb[1] = Tb.y; o.Tb[2] = Tb.z;
                o.wrote_buoyancy = 1;
            }
            o.Swet = st.surface;
        }
    } else { // CROSSING: SolidEntity.cpp:1689-1720 on the fp64 sums (F64_ACC contract)
        const bool buoy = (bd.flags & BF_BUOYANCY) != 0, damp = (bd.flags & BF_DAMPING) != 0;
        const double Vsub = acc[19];
        if (buoy && Vsub > (double)1e-9f) {
            o.wrote_buoyancy = 1;
            o.Vsub = ddiv(Vsub, (double)6.f);
            if (oc.has_waves) {
                const double kk = (double)(float)dmul(oc.density, oc.gravity_z);
                for (int k = 0; k < 3; ++k) { o.Fb[k] = dmul(acc[12 + k], kk); o.Tb[k] = dmul(acc[15 + k], kk); }
            } else {
                const D3 cb = {dadd(ddiv(acc[20], Vsub), (double)bd.p[0]), dadd(ddiv(acc[21], Vsub), (double)bd.p[1]),
                               dadd(ddiv(acc[22], Vsub), (double)0.f)};
                const double s = dmul(-o.Vsub, oc.density);
                const D3 Fb = {dmul(0.0, s), dmul(0.0, s), dmul(oc.gravity_z, s)};
                const D3 Tb = dcross(cb - D3{st.T_CG[9], st.T_CG[10], st.T_CG[11]}, Fb);
                o.Fb[0] = Fb.x; o.Fb[1] = Fb.y; o.Fb[2] = Fb.z; o.Tb[0] = Tb.x; o.Tb[1] = Tb.y; o.Tb[2] = Tb.z;
            }
        }
        if (damp) {
            o.wrote_damping = 1;
            for (int k = 0; k < 3; ++k) { o.Fdq[k] = acc[k]; o.Tdq[k] = acc[3 + k]; o.Fdf[k] = acc[6 + k]; o.Tdf[k] = acc[9 + k]; }
        }
        o.Swet = acc[18];
    }
    if (dispatch && pos != SFH_POS_OUTSIDE && (st.flags & SFH_F_DAMPING)) // SolidEntity.cpp:1848-1849
        correct_forces(oc.density, o.Fdq, o.Tdq, o.Fdf, o.Tdf, st.Cd, st.Cf, st.T_O_basis);
    P.out[b] = o;
}

// point queries: Ocean::GetDepth(glm::vec3) / Ocean::GetFluidVelocity(glm::vec3) for n points
__global__ void k_query_depth(const float* __restrict__ xyz, uint32_t n, bool waves, WavesDev wv, float* __restrict__ out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    out[i] = waves ? get_depth<true>(wv, x, y, z) : z;
}
__global__ void k_query_velocity(const float* __restrict__ xyz, uint32_t n, const OceanDev* __restrict__ ocean, float* __restrict__ out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F3 v = {0.f, 0.f, 0.f};
    if (ocean->currents_enabled) v = fluid_velocity_general(ocean->currents, ocean->n_currents, F3{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]});
    out[3 * i] = v.x; out[3 * i + 1] = v.y; out[3 * i + 2] = v.z;
}

// wave field de-interleave: RGBA32F layer 0 (+ spill-over row) -> planar channel 0 / 1
// Also leaves max|ch0| and max|ch1| in hmax[0..1] (float bits; the caller zeroes them first): their sum bounds the wave
// height anywhere, because the bilinear lookup is a convex combination of texels.
__global__ void k_waves_planar(const float* __restrict__ rgba, uint32_t n_floats_valid, uint32_t n_texels, float* ch0, float* ch1,
                               uint32_t* hmax)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float a = 0.f, b = 0.f;
    if (i < n_texels) {
        a = (4 * i + 0 < n_floats_valid) ? rgba[4 * i + 0] : 0.f;
        b = (4 * i + 1 < n_floats_valid) ? rgba[4 * i + 1] : 0.f;
        ch0[i] = a;
        ch1[i] = b;
    }
    // |x| as an unsigned integer orders like |x| (NaN sorts above infinity: a NaN texel disables the dry-chunk test)
    const uint32_t ma = __reduce_max_sync(SFH_FULL, __float_as_uint(a) & 0x7fffffffu), mb = __reduce_max_sync(SFH_FULL, __float_as_uint(b) & 0x7fffffffu);
    if ((threadIdx.x & 31) == 0) { atomicMax(hmax, ma); atomicMax(hmax + 1, mb); }
}

} // namespace sfh
