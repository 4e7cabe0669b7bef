// sfh_kernels.cuh — the kernels of one hydrodynamics step (sm_100a).
//
//   k_prologue      1 thread/body  per-body: CheckBodyFluidPosition (a9), dispatcher gates (a10), fp64->fp32
//                                  narrowing of transforms/velocities, position-independent current sum,
//   k_scan_items    1 block        exclusive scan of the per-body chunk counts -> two work lists (submerged / surface)
//   k_submerged     persistent     per-face ComputeHydrodynamicForcesSubmerged (a1) over (body, chunk) items, one stream per WARP
//   k_classify      persistent     surface list -> dry (dropped) / wet / clip chunk lists          (sfh_surface.cuh)
//   k_surface_lists persistent     per-face ComputeHydrodynamicForcesSurface (a2) incl. waves (a3/a4), one launch, two roles:
//                                  wet = chunks entirely under water, clip = chunks the water line may cross (eight classes)
//   k_epilogue      1 warp/body    fixed-order reduction of the partial records, finalisation (A.5),
//                                  INSIDE analytic buoyancy and CorrectHydrodynamicForces (a11)
//
// Work decomposition of k_submerged: the work list is the concatenation of every active body's chunks; stream `blk` of G streams
// (a stream = one warp) owns the contiguous item range [blk*W/G, (blk+1)*W/G).
// A stream keeps fp64 partial sums in registers and writes ONE record per (stream, body) it touched at index blk + body — those indices are unique
// because block ranges are contiguous and monotone in the body index — so no atomics, no zero-fill, and the
// epilogue reduces each body's records in a fixed order: results are bitwise reproducible for a given G.
// The surface kernels hand out SEGMENTS of their lists through an atomic counter and write one record per (segment, body).
#pragma once

#include "sfh_device.cuh"

namespace sfh {

#define SFH_FULL 0xffffffffu
#define SFH_SPILL_EVERY 100u     /* chunks between counter spills of k_submerged */
#define SFH_LIST_SPILL_EVERY 31u /* ... of the surface kernels */
// Event counters are packed per lane: 10-bit fields (degenerate | quadratic | skin) and, in k_surface_clip, 8-bit class fields.
// A lane sees SFH_CH_FACES / 32 faces per chunk, so the fields hold that many events times the chunks between two spills.
static_assert((SFH_CH_FACES / 32u) * SFH_SPILL_EVERY < 1024u, "k_submerged: 10-bit event counters would overflow between spills");
static_assert((SFH_CH_FACES / 32u) * SFH_LIST_SPILL_EVERY < 256u, "k_surface_clip: 8-bit class counters would overflow between spills");
static_assert(SFH_CH_FACES == 256u && SFH_CH_VERTS == 256u, "vertex slots are bytes of the packed face word; the warp kernels transform 8 x 32 vertex slots");

struct StepParams {
    const BodyDev* bodies;
    const uint32_t* item_off; // n_bodies + 1, exclusive scan of chunk counts of this kernel's kind
    uint32_t n_bodies;
    const OceanDev* ocean;
    WavesDev waves;
    void* partials;
    uint8_t* codes;            // debug: per-face codes, nullptr otherwise
    const uint64_t* code_base; // debug: first code index of each body
};

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

// Sums of N per-lane values over the warp, stored to out[0..N) and the per-lane values cleared. The values are FOLDED: at
// every level a lane keeps one half of its values and sends the other half to its partner at xor distance 16, 8, 4, 2, 1 —
// 2N double shuffles (N, N/2, ... per level) instead of 5N, and the 40 stores spread over the lanes. The pairing is the tree of
// warp_sum ((l, l^16), then (l, l^8) ... on the partial sums; fp64 addition commutes), so every sum has the bits warp_sum gives.
template <int N, int H>
__device__ __forceinline__ void warp_fold(double* a, uint32_t lane, uint32_t bit)
{
    const bool upper = (lane & bit) != 0u;
#pragma unroll
    for (int k = 0; k < H; ++k) {
        const double lo = a[k], hi = (k + H < N) ? a[k + H] : 0.0;
        const double send = upper ? lo : hi, keep = upper ? hi : lo;
        a[k] = dadd(keep, __shfl_xor_sync(SFH_FULL, send, bit));
    }
}
template <int N>
__device__ __forceinline__ void warp_sum_store(double* acc, double* out, uint32_t lane)
{
    constexpr int H1 = (N + 1) / 2, H2 = (H1 + 1) / 2, H3 = (H2 + 1) / 2, H4 = (H3 + 1) / 2, H5 = (H4 + 1) / 2;
    warp_fold<N, H1>(acc, lane, 16u);
    warp_fold<H1, H2>(acc, lane, 8u);
    warp_fold<H2, H3>(acc, lane, 4u);
    warp_fold<H3, H4>(acc, lane, 2u);
    warp_fold<H4, H5>(acc, lane, 1u);
#pragma unroll
    for (int k = 0; k < H5; ++k) { // value k of this lane is the sum of which index? (an index past a level's size was padding)
        uint32_t i = (uint32_t)k + ((lane & 1u) ? (uint32_t)H5 : 0u);
        bool ok = i < (uint32_t)H4;
        i += (lane & 2u) ? (uint32_t)H4 : 0u;  ok = ok && i < (uint32_t)H3;
        i += (lane & 4u) ? (uint32_t)H3 : 0u;  ok = ok && i < (uint32_t)H2;
        i += (lane & 8u) ? (uint32_t)H2 : 0u;  ok = ok && i < (uint32_t)H1;
        i += (lane & 16u) ? (uint32_t)H1 : 0u; ok = ok && i < (uint32_t)N;
        if (ok) out[i] = acc[k];
    }
#pragma unroll
    for (int i = 0; i < N; ++i) acc[i] = 0.0;
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

// current parameters and their fp32 pre-test records into shared memory (published by the caller's next block barrier)
__device__ __forceinline__ void stage_currents(sfh_current* s_cur, CurGuard* s_guard, const OceanDev* __restrict__ ocean, uint32_t n_cur)
{
    const uint32_t* s = reinterpret_cast<const uint32_t*>(ocean->currents);
    uint32_t* d = reinterpret_cast<uint32_t*>(s_cur);
    for (uint32_t i = threadIdx.x; i < n_cur * (uint32_t)(sizeof(sfh_current) / 4); i += blockDim.x) d[i] = __ldg(s + i);
    for (uint32_t i = threadIdx.x; i < n_cur; i += blockDim.x) s_guard[i] = make_cur_guard(ocean->currents[i]);
}

} // namespace sfh

#include "sfh_submerged.cuh" // k_submerged (two faces per lane side by side)
#include "sfh_surface.cuh"   // k_classify, k_surface_lists

namespace sfh {

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
    uint32_t no_sub, no_surf; // the caller excluded the path (its kernels are not launched): a body that takes it is flagged, status |= 4
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
    if ((kind == KIND_SUBMERGED && P.no_sub) || (kind == KIND_SURFACE && P.no_surf)) {
        atomicOr(P.status, 4u);
        kind = KIND_NONE;
    }
    MeshDev m = {nullptr, nullptr, nullptr, nullptr, 0, 0};
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
    bd.bounds = m.bounds;
    bd.pad2 = 0;
    P.bodies[i] = bd;
    // chunk counts; k_scan_items turns them into the two work lists
    P.off_sub[i] = (kind == KIND_SUBMERGED) ? bd.n_chunks : 0u;
    P.off_surf[i] = (kind == KIND_SURFACE) ? bd.n_chunks : 0u;
}

#ifndef SFH_PROLOGUE_THREADS
#define SFH_PROLOGUE_THREADS 32 /* bodies per block of k_prologue: few, so that the records' round trips spread over the SMs */
#endif
__global__ void __launch_bounds__(128) k_prologue(PrologueParams P)
{
    pdl_trigger();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P.n_bodies) prologue_body(P, i);
}

// In-place exclusive scan of the per-body chunk counts (n_bodies entries -> n_bodies + 1 offsets). One block.
__device__ __forceinline__ void scan_items(uint32_t* off_sub, uint32_t* off_surf, uint32_t n_bodies)
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
    }
}
__global__ void __launch_bounds__(1024) k_scan_items(uint32_t* off_sub, uint32_t* off_surf, uint32_t n_bodies)
{
    pdl_trigger();
    pdl_wait(); // k_prologue's chunk counts
    scan_items(off_sub, off_surf, n_bodies);
}
// Steps of at most 1024 bodies: prologue and scan in one launch (single-body steps are launch-latency bound)
__global__ void __launch_bounds__(1024) k_prologue_scan(PrologueParams P)
{
    pdl_trigger(); // the face kernels may be scheduled now; they wait for this grid before they read its lists
    if (threadIdx.x < P.n_bodies) prologue_body(P, threadIdx.x);
    __syncthreads();
    scan_items(P.off_sub, P.off_surf, P.n_bodies);
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
    const PartialSurf* part_wet;  // one record per (segment, body) of k_surface_wet
    const PartialSurf* part_clip; // ... of k_surface_clip
    uint32_t grid_sub;
    SurfLists L;                  // the surface path's lists: which segments hold a body's wet / clip chunks
    uint32_t direct_streams;      // T of direct_split (sfh_surface.cuh)
    uint32_t team;                // threads that reduce one body's records: 32 (4 bodies per block) or 128 (one body per block)
    // Few bodies with thousands of partial records each (a single hull: one record per warp of the face kernel): `slices` blocks
    // per body add a share of the records each and leave a slice total; the last one to finish adds the totals in slice order.
    uint32_t slices;              // blocks per body (team == 128 only); 1 = no second stage
    PartialSurf* slice_part;      // n_bodies * slices
    uint32_t* slice_ticket;       // n_bodies, zero between steps
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

// The per-body tail of the reference (SolidEntity.cpp:1689-1720, 1793-1850, 1920-1956) on the fp64 sums: one thread per body
__device__ __forceinline__ void epilogue_finish(const EpilogueParams& P, uint32_t b, const BodyDev& bd, uint32_t kind, uint32_t pos,
                                                const double* acc, const uint32_t* cnt, sfh_info& o)
{
    const sfh_body_state& st = P.states[b];
    const OceanDev& oc = *P.ocean;
    const bool dispatch = (st.flags & SFH_F_DISPATCH) != 0;
    for (int k = 0; k < 3; ++k) { o.Fb[k] = 0; o.Tb[k] = 0; o.Fdq[k] = 0; o.Tdq[k] = 0; o.Fdf[k] = 0; o.Tdf[k] = 0; }
    o.Swet = 0; o.Vsub = 0; o.wrote_buoyancy = 0; o.wrote_damping = 0; o.position = pos; o.reserved = 0;
    o.n_faces = (kind != KIND_NONE) ? bd.n_faces : 0u;
    o.n_degenerate = cnt[0]; o.n_quadratic = cnt[1]; o.n_skin = cnt[2];
    for (int k = 0; k < 8; ++k) o.class_count[k] = cnt[3 + k];
    if (kind == KIND_SURFACE) { // every face is in exactly one class; class 0 (all three vertices dry) is the one nobody counts
        uint32_t others = 0;
        for (int k = 1; k < 8; ++k) others += cnt[3 + k];
        o.class_count[0] = bd.n_faces - others;
    }

    if (P.aero) { // SolidEntity.cpp:1920-1922, then CorrectAerodynamicForces
        const double k = dmul(0.5, oc.density);
        for (int c = 0; c < 3; ++c) { o.Fdq[c] = dmul(acc[c], k); o.Tdq[c] = dmul(acc[3 + c], k); }
        o.wrote_damping = 1;
        o.class_count[7] = o.n_faces;
        if (dispatch) correct_aero(o.Fdq, o.Tdq, st.Cd, st.T_O_basis, (st.flags & SFH_F_AERO_CD) != 0);
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
}

__device__ __forceinline__ void epilogue_body(const EpilogueParams& P, uint32_t b, uint32_t t, uint32_t team, double (*s_red)[SFH_NACC_SURF],
                                              uint32_t (*s_cnt)[11]);

// Few bodies (a small scene, a single hull with thousands of partial records): one body per BLOCK, its records spread over 128
// threads; otherwise one body per warp. Either way the order of the sums is fixed.
__global__ void __launch_bounds__(128) k_epilogue(EpilogueParams P)
{
    __shared__ double s_red[4][SFH_NACC_SURF];
    __shared__ uint32_t s_cnt[4][11];
    pdl_wait();
    const uint32_t per_block = 128u / P.team;
    const uint32_t b = P.team == 128u ? blockIdx.x / P.slices : blockIdx.x * per_block + threadIdx.x / P.team;
    if (b < P.n_bodies) epilogue_body(P, b, threadIdx.x % P.team, P.team, s_red, s_cnt);
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

__device__ __forceinline__ void epilogue_body(const EpilogueParams& P, uint32_t b, uint32_t t, uint32_t team, double (*s_red)[SFH_NACC_SURF],
                                              uint32_t (*s_cnt)[11])
{
    const BodyDev& bd = P.bodies[b];
    const uint32_t kind = bd.kind & 255u, pos = bd.kind >> 8;

    double acc[SFH_NACC_SURF];
#pragma unroll
    for (int i = 0; i < SFH_NACC_SURF; ++i) acc[i] = 0.0;
    uint32_t cnt[11]; // degenerate quadratic skin | cls[8]
#pragma unroll
    for (int i = 0; i < 11; ++i) cnt[i] = 0;

    // this block's share of the records: every (slices)-th group of `team` records, starting at group `sl`
    const uint32_t S = team == 128u ? P.slices : 1u, sl = team == 128u ? blockIdx.x % S : 0u, stride = team * S, first = t + sl * team;
    if (kind == KIND_SUBMERGED) {
        const uint32_t o0 = P.off_sub[b], o1 = P.off_sub[b + 1];
        if (o1 > o0) {
            const uint64_t W = P.off_sub[P.n_bodies], G = min((uint64_t)P.grid_sub, W);
            const uint32_t k0 = block_of_item(o0, G, W), k1 = block_of_item(o1 - 1, G, W);
            for (uint32_t k = k0 + first; k <= k1; k += stride) {
                const PartialSub& r = P.part_sub[k + b];
#pragma unroll
                for (int i = 0; i < SFH_NACC_SUB; ++i) acc[i] = dadd(acc[i], r.a[i]);
                cnt[0] += r.n_degenerate; cnt[1] += r.n_quadratic; cnt[2] += r.n_skin;
            }
        }
    } else if (kind == KIND_SURFACE) {
        const uint32_t o0 = P.off_surf[b], o1 = P.off_surf[b + 1], Wall = P.off_surf[P.n_bodies];
        auto add_records = [&](const PartialSurf* part, uint32_t r0, uint32_t r1, uint32_t SC) {
            if (r1 <= r0) return;
            const uint32_t k0 = r0 / SC, k1 = (r1 - 1) / SC; // segments that overlap this body's entries
            for (uint32_t k = k0 + first; k <= k1; k += stride) {
                const PartialSurf& r = part[k + b];
#pragma unroll
                for (int i = 0; i < SFH_NACC_SURF; ++i) acc[i] = dadd(acc[i], r.a[i]);
                cnt[0] += r.n_degenerate; cnt[1] += r.n_quadratic; cnt[2] += r.n_skin;
#pragma unroll
                for (int i = 1; i < 8; ++i) cnt[3 + i] += r.cls[i];
            }
        };
        const uint32_t dsplit = Wall <= P.L.items_cap ? direct_split(Wall, P.direct_streams) : 0u;
        if (dsplit) { // direct mode: one record per sub-item, `dsplit` consecutive ones per chunk
            add_records(P.part_wet, o0 * dsplit, o1 * dsplit, 1u);
        } else if (o1 > o0 && o1 <= P.L.ctrl[SC_ITEMS]) { // SC_ITEMS = 0: the lists did not fit (status flag)
            // position of item w in the wet (shift 0) / clip (shift 16) list = entries of earlier tiles + earlier entries of its tile
            auto rank = [&](const uint32_t* off, uint32_t w, int shift) {
                return (w & (SFH_TILE - 1u)) == 0u ? off[w / SFH_TILE] : off[w / SFH_TILE] + ((P.L.before[w] >> shift) & 0xffffu);
            };
            // segments count sub-items: `split` consecutive ones per list entry
            const uint32_t ws = P.L.ctrl[SC_WET_SPLIT], cs = P.L.ctrl[SC_CLIP_SPLIT];
            add_records(P.part_wet, rank(P.L.wet_off, o0, 0) * ws, rank(P.L.wet_off, o1, 0) * ws, P.L.ctrl[SC_WET_SEG]);
            add_records(P.part_clip, rank(P.L.clip_off, o0, 16) * cs, rank(P.L.clip_off, o1, 16) * cs, P.L.ctrl[SC_CLIP_SEG]);
        }
    }
    // warp sums into shared memory (folded: sfh_kernels.cuh warp_sum_store), read back by the thread that goes on
    const uint32_t lane = t & 31u, wp = threadIdx.x >> 5;
    if (kind != KIND_NONE) {
        warp_sum_store<SFH_NACC_SURF>(acc, s_red[wp], lane);
#pragma unroll
        for (int i = 0; i < 11; ++i) cnt[i] = __reduce_add_sync(SFH_FULL, cnt[i]);
    } else {
        if (lane < (uint32_t)SFH_NACC_SURF) s_red[wp][lane] = 0.0;
    }
    if (team == 32u) {
        __syncwarp();
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < SFH_NACC_SURF; ++i) acc[i] = s_red[wp][i];
        }
    } else { // the four warps' sums, added in warp order
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 11; ++i) s_cnt[wp][i] = cnt[i];
        }
        __syncthreads();
        if (t == 0) {
#pragma unroll
            for (int i = 0; i < SFH_NACC_SURF; ++i) acc[i] = dadd(dadd(dadd(s_red[0][i], s_red[1][i]), s_red[2][i]), s_red[3][i]);
#pragma unroll
            for (int i = 0; i < 11; ++i) cnt[i] = s_cnt[0][i] + s_cnt[1][i] + s_cnt[2][i] + s_cnt[3][i];
        }
        if (S > 1u) { // second stage: slice totals, added in slice order by the block that finishes last
            __shared__ uint32_t s_last;
            __shared__ double s_slice[16][SFH_NACC_SURF];
            __shared__ uint32_t s_slice_cnt[16][11];
            PartialSurf* mine = P.slice_part + (size_t)b * S;
            if (t == 0) {
                PartialSurf& o = mine[sl];
#pragma unroll
                for (int i = 0; i < SFH_NACC_SURF; ++i) o.a[i] = acc[i];
                o.n_degenerate = cnt[0]; o.n_quadratic = cnt[1]; o.n_skin = cnt[2];
#pragma unroll
                for (int i = 0; i < 8; ++i) o.cls[i] = cnt[3 + i];
                __threadfence();
                s_last = atomicAdd(&P.slice_ticket[b], 1u) == S - 1u ? 1u : 0u;
            }
            __syncthreads();
            if (!s_last) return;
            __threadfence();
            // every slice total fetched at once (independent loads over the 128 threads), then added in slice order
            for (uint32_t i = t; i < S * (uint32_t)SFH_NACC_SURF; i += 128u) {
                const uint32_t k = i / (uint32_t)SFH_NACC_SURF, j = i % (uint32_t)SFH_NACC_SURF;
                s_slice[k][j] = __ldcg(&mine[k].a[j]);
            }
            for (uint32_t i = t; i < S * 11u; i += 128u) {
                const uint32_t k = i / 11u, f = i % 11u;
                s_slice_cnt[k][f] = f < 3u ? __ldcg(&mine[k].n_degenerate + f) : __ldcg(&mine[k].cls[f - 3u]);
            }
            __syncthreads();
            if (t < (uint32_t)SFH_NACC_SURF) {
                double v = 0.0;
                for (uint32_t k = 0; k < S; ++k) v = dadd(v, s_slice[k][t]);
                s_red[0][t] = v;
            } else if (t >= 32u && t < 43u) {
                const uint32_t f = t - 32u;
                uint32_t c = 0;
                for (uint32_t k = 0; k < S; ++k) c += s_slice_cnt[k][f];
                s_cnt[0][f] = c;
            }
            __syncthreads();
            if (t == 0) {
#pragma unroll
                for (int i = 0; i < SFH_NACC_SURF; ++i) acc[i] = s_red[0][i];
#pragma unroll
                for (int i = 0; i < 11; ++i) cnt[i] = s_cnt[0][i];
                P.slice_ticket[b] = 0;
            }
        }
    }
    if (t >= 32u) return;
    // lane 0 of the body's first warp finishes the record in shared memory; the warp then writes it as fourteen 16-byte stores
    // (one or two transactions — over PCIe when the record goes straight into the pinned result buffer)
    static_assert(sizeof(sfh_info) % 16 == 0 && sizeof(sfh_info) / 16 <= 32, "record copy by one warp");
    __shared__ __align__(16) sfh_info s_out[4];
    if (lane == 0) epilogue_finish(P, b, bd, kind, pos, acc, cnt, s_out[wp]);
    __syncwarp();
    if (lane < (uint32_t)(sizeof(sfh_info) / 16))
        reinterpret_cast<uint4*>(P.out + b)[lane] = reinterpret_cast<const uint4*>(&s_out[wp])[lane];
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
