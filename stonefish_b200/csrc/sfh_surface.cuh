// sfh_surface.cuh — SolidEntity::ComputeHydrodynamicForcesSurface (SolidEntity.cpp:1289-1721) as three kernels built around
// CHUNK KINDS. A body that crosses the surface mixes three kinds of (body, chunk) work items whose cost differs by 20x:
//
//   dry   every vertex is above the water: every face is class 0 and adds nothing (:1353-1360)          -> dropped
//   wet   every vertex is under the water: every face is class 7, nothing is clipped (:1605-1687)       -> k_surface_wet
//   clip  the water line may pass through the chunk: per-vertex depths, 8-way clipping (:1361-1604)     -> k_surface_clip
//
//   k_classify      one thread per item: the chunk's box (ChunkBounds, mesh frame) is carried to the world with the body's T_C
//                   and compared with a LOCAL bound of the wave height over the box's footprint (the texels the lookups of
//                   OpenGLRealOcean.cpp:280-319 can touch there, tightened by a Lipschitz bound of the bilinear interpolant
//                   around its value at the footprint's centre). Items are compacted per tile of 1024 into two lists of
//                   descriptors (wet from the front of the tile's region, clip from its back); the last block to finish scans
//                   the tile counts and sizes the consumers' segments.
//   k_surface_lists one launch, two roles by block index:
//    wet role       k_submerged's machinery — one stream per WARP, chunks staged by bulk async copies, vertices transformed once
//                   per chunk, two faces per lane side by side in one straight-line block, no block barrier — plus the tetrahedron
//                   sums of the submerged volume, the pressure term (one wave lookup per face centroid) and Swet. Also under waves.
//    clip role      the reference's eight classes, one face per lane at a time; the same warp streams.
//
// The classification is conservative: "dry" and "wet" are only claimed when the bound proves them for every vertex with the
// reference's own fp32 arithmetic (margins below); everything else is "clip", which is exact for any input. The per-face class
// codes of the parity tests would show a wrong claim as a class mismatch.
// Work is handed out in SEGMENTS of consecutive list entries through an atomic counter; a stream writes one partial record per
// (segment, body) at index segment + body, the epilogue adds a body's records in segment order: no atomics on the sums, results
// do not depend on which warp took which segment.
#pragma once

#include "sfh_submerged.cuh"

namespace sfh {

#define SFH_TILE 128u         /* items per classification tile = threads of a k_classify block. Small: the classifier of a single
                                 hull (783 chunks) then runs on 7 SMs instead of one, whose L1 miss path its ~40 texel loads per item saturated */
#ifndef SFH_SEG_PER_STREAM
#define SFH_SEG_PER_STREAM 16u /* at most this many segments per resident stream (warp) */
#endif
enum { CK_DRY = 0, CK_WET = 1, CK_CLIP = 2 };
// control words of the surface path (device, zeroed once): the classifier's last block fills them in for the consumers
// (totals and segment sizes count SUB-ITEMS: entries x split, see finish_lists)
enum { SC_TICKET = 0, SC_WET_NEXT = 2, SC_WET_SEG = 3, SC_CLIP_NEXT = 4, SC_CLIP_SEG = 5, SC_WET_TOTAL = 6, SC_CLIP_TOTAL = 7, SC_ITEMS = 8,
       SC_WET_SPLIT = 9, SC_CLIP_SPLIT = 10, SC_WET_ENTRIES = 11, SC_CLIP_ENTRIES = 12, SC_WORDS = 16 };

// A work item of the consumers: where the chunk is, whose it is. w = chunk index | size in 16-byte units << 23.
// (16 bytes per item against >= 3 KB of chunk data.)
struct SurfLists {
    uint4* items;        // items_cap entries: tile t owns [t * SFH_TILE, (t + 1) * SFH_TILE)
    uint32_t* before;    // items_cap: wet | clip << 16 items of the tile before this one (exclusive ranks)
    uint32_t* wet_off;   // items_cap / SFH_TILE + 1: exclusive scan of the tiles' wet counts (last = total)
    uint32_t* clip_off;  // same for clip
    uint2* tile_cnt;     // items_cap / SFH_TILE: (wet, clip) of each tile
    uint32_t* ctrl;      // SC_WORDS
    uint32_t items_cap;  // multiple of SFH_TILE
};

__device__ __forceinline__ int floor_div(int a, int n) { return a >= 0 ? a / n : -((n - 1 - a) / n); }

// DIRECT MODE: a step with no more surface items than consumer streams (a single hull, a small scene) is bound by latency, and
// every kernel of the chain costs 5-15 us of dependent round trips whatever it computes. Such a step skips the lists: warp u of
// k_surface_lists classifies its own (sub-)item, works on it and writes one record. Every kernel decides alike from W (surface
// items) and T (the consumer's RESIDENT streams; the launch has twice as many warps): 0 = list mode, else the number of sub-items
// per item. Sub-items only while they all fit the resident warps (a second wave pays the per-warp start-up again: 400k faces
// 50.8 us with one sub-item per chunk in one wave, 57.7 us with two in two waves); whole chunks up to two waves (455k-900k-face
// hulls crossing waves: 53-66 us against 61-72 us through the lists).
__device__ __forceinline__ uint32_t direct_split(uint32_t W, uint32_t T)
{
    if (W > 2u * (uint64_t)T) return 0u; // up to two waves of warps (the launch always has 2 T of them): measured against the lists
    if (W == 0u) return 1u; // nothing to do: no lists either (k_classify may not even have been launched)
    return (uint64_t)W * 4u <= T ? 4u : ((uint64_t)W * 2u <= T ? 2u : 1u);
}

// Range [lo, hi] of the reference's bilinear lookup of one channel (OpenGLRealOcean.cpp:280-319) for every point of the world
// rectangle [cx +- ex] x [cy +- ey]. The lookup at texel coordinates X = x / g * N reads columns floor(X - 0.5) and floor(X + 0.5)
// (mod N) and returns a convex combination of four texels, so the minimum and maximum of the texels the rectangle can touch bound
// it (this function: footprints of up to 8 x 8 texels, period boundaries included). Inside one period the interpolant is also
// continuous with slopes bounded by the largest difference of adjacent texels, which bounds it around its value at the centre —
// tighter when the rectangle is a fraction of a texel, as the chunks of a fine hull are (chan_window / chan_reduce below).
// At a period boundary the reference's index can become N (the fraction 1 - |tiny| rounds to 1.0f, :291-296): it then reads
// column 0 of the NEXT row, or the spill-over row N*N..N*N+N; those texels join the candidates.
// Falls back to the global bound +-hmax (which includes the spill-over texels) for large or non-finite rectangles and NaN texels.
__device__ __noinline__ void wave_channel_bounds(const float* __restrict__ ch, int N, float g, float hmax, float cx, float cy, float ex,
                                                 float ey, float& lo, float& hi)
{
    lo = -hmax;
    hi = hmax;
    const float fN = (float)N, s = fN / g;
    const float Xc = cx * s, Yc = cy * s, wx = fabsf(ex * s), wy = fabsf(ey * s);
    // the reference's own X (division, +-0.5/N, fraction, * N: OpenGLRealOcean.cpp:283-300) is within 4 ulp of |X|; so is ours
    const float dx = 0.02f + 8e-6f * fabsf(Xc), dy = 0.02f + 8e-6f * fabsf(Yc);
    const float fa = floorf(Xc - wx - 0.5f - dx), fb = floorf(Xc + wx + 0.5f + dx);
    const float ga = floorf(Yc - wy - 0.5f - dy), gb = floorf(Yc + wy + 0.5f + dy);
    if (!(fb - fa < 8.f && gb - ga < 8.f && fabsf(fa) < 1e8f && fabsf(ga) < 1e8f)) return; // large, or NaN
    const int ia = (int)fa, ib = (int)fb, ja = (int)ga, jb = (int)gb;
    const bool wrap_i = floor_div(ia, N) != floor_div(ib, N), wrap_j = floor_div(ja, N) != floor_div(jb, N);
    float mn = __int_as_float(0x7f800000), mx = -mn;
    bool bad = false;
    auto see = [&](float v) { mn = fminf(mn, v); mx = fmaxf(mx, v); bad |= (v != v); };
    // minimum and maximum of up to 8 x 8 texels: rows one at a time, the eight loads of a row in flight together (a plain double
    // loop waits for an L2 round trip per texel, which made the classifier of a single hull slower than its face kernels)
    int col[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) { const int i = min(ia + c, ib); col[c] = i - N * floor_div(i, N); }
    for (int j = ja; j <= jb; ++j) {
        const float* row = ch + (j - N * floor_div(j, N)) * N;
        float v[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = __ldg(row + col[c]);
        const float q = wrap_i ? __ldg(row + N) : v[0]; // index N of this row
#pragma unroll
        for (int c = 0; c < 8; ++c) see(v[c]);
        see(q);
    }
    if (wrap_j) {
        const float* row = ch + N * N;
#pragma unroll
        for (int c = 0; c < 8; ++c) see(__ldg(row + col[c]));
        if (wrap_i) see(__ldg(row + N));
    }
    if (bad) return;
    lo = mn;
    hi = mx;
}

// The same bound for footprints of at most 4 x 4 texels that do not touch a period boundary — every chunk of a fine hull — as
// straight-line code: the index arithmetic of BOTH channels first, then all 40 texel loads in flight together, then the
// reductions. (The loop version above waits for one L2 round trip per texel: 27 us for the 783 chunks of a single 200k-face
// hull, more than the face kernels take.)
struct ChanWin {
    bool fast;            // footprint fits the 4 x 4 window, inside one period
    int idx[16], cen[4];  // texel indices of the window (clamped duplicates at its edges) and of the centre's four texels
    float a, b, wx, wy;   // centre weights; half widths incl. margins, texels
};
__device__ __forceinline__ ChanWin chan_window(int N, float g, float cx, float cy, float ex, float ey)
{
    ChanWin w;
    const float fN = (float)N, s = fN / g;
    const float Xc = cx * s, Yc = cy * s, wx = fabsf(ex * s), wy = fabsf(ey * s);
    const float dx = 0.02f + 8e-6f * fabsf(Xc), dy = 0.02f + 8e-6f * fabsf(Yc);
    const float fa = floorf(Xc - wx - 0.5f - dx), fb = floorf(Xc + wx + 0.5f + dx);
    const float ga = floorf(Yc - wy - 0.5f - dy), gb = floorf(Yc + wy + 0.5f + dy);
    w.fast = fb - fa < 4.f && gb - ga < 4.f && fabsf(fa) < 1e8f && fabsf(ga) < 1e8f; // false for NaN
    const int ia = w.fast ? (int)fa : 0, ib = w.fast ? (int)fb : 0, ja = w.fast ? (int)ga : 0, jb = w.fast ? (int)gb : 0;
    if (floor_div(ia, N) != floor_div(ib, N) || floor_div(ja, N) != floor_div(jb, N)) w.fast = false;
    const int i0 = ia - N * floor_div(ia, N), j0 = ja - N * floor_div(ja, N), ni = ib - ia, nj = jb - ja; // no wrap: i0 + ni < N
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) w.idx[4 * r + c] = w.fast ? (j0 + min(r, nj)) * N + i0 + min(c, ni) : 0;
    const float x0 = floorf(Xc - 0.5f), y0 = floorf(Yc - 0.5f);
    w.a = Xc - 0.5f - x0;
    w.b = Yc - 0.5f - y0;
    const int xi = w.fast ? (int)x0 : 0, yi = w.fast ? (int)y0 : 0; // columns xi, xi + 1 and rows yi, yi + 1 lie inside the window
    const int ci = xi - N * floor_div(xi, N), cj = yi - N * floor_div(yi, N);
    w.cen[0] = w.fast ? cj * N + ci : 0; w.cen[1] = w.fast ? cj * N + ci + 1 : 0;
    w.cen[2] = w.fast ? (cj + 1) * N + ci : 0; w.cen[3] = w.fast ? (cj + 1) * N + ci + 1 : 0;
    w.wx = wx + dx;
    w.wy = wy + dy;
    return w;
}
__device__ __forceinline__ void chan_reduce(const ChanWin& w, const float (&v)[16], const float (&c)[4], float& lo, float& hi)
{
    float mn = v[0], mx = v[0], Lx = 0.f, Ly = 0.f;
    bool bad = false;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        mn = fminf(mn, v[k]); mx = fmaxf(mx, v[k]); bad |= (v[k] != v[k]);
        if ((k & 3) != 0) Lx = fmaxf(Lx, fabsf(v[k] - v[k - 1]));
        if (k >= 4) Ly = fmaxf(Ly, fabsf(v[k] - v[k - 4]));
    }
    const float bc = (1.f - w.a) * (1.f - w.b) * c[0] + w.a * (1.f - w.b) * c[1] + (1.f - w.a) * w.b * c[2] + w.a * w.b * c[3];
    const float r = Lx * w.wx + Ly * w.wy;
    if (bad) return;
    lo = fmaxf(mn, bc - r);
    hi = fminf(mx, bc + r);
}

// Kind of one chunk of a crossing body. R, t: T_C as fp32 (BodyDev); b0, b1: the chunk's box in the mesh frame (ChunkBounds).
__device__ __forceinline__ uint32_t classify_chunk(const float* __restrict__ R, const float* __restrict__ t, float4 b0, float2 b1, bool waves,
                                                   const WavesDev& wv, float hmax0, float hmax1, int local = 2)
{
    const float cx = b0.x, cy = b0.y, cz = b0.z, hx = b0.w, hy = b1.x, hz = b1.y;
    float cw[3], ext[3], M = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float r0 = R[3 * k], r1 = R[3 * k + 1], r2 = R[3 * k + 2];
        cw[k] = r0 * cx + r1 * cy + r2 * cz + t[k];
        ext[k] = fabsf(r0) * hx + fabsf(r1) * hy + fabsf(r2) * hz;
        M += fabsf(t[k]) + fabsf(r0) * (fabsf(cx) + hx) + fabsf(r1) * (fabsf(cy) + hy) + fabsf(r2) * (fabsf(cz) + hz);
    }
    // Every vertex v of the chunk has |v - c| <= h per axis in the mesh frame, so its world coordinate k lies within ext[k] of
    // cw[k] in exact arithmetic; e covers the fp32 rounding of the reference's transform (a few ulp of the largest intermediate,
    // M bounds them all) and of the three lines above.
    const float e = 4e-6f * M + 1e-6f;
    const float zmin = cw[2] - ext[2] - e, zmax = cw[2] + ext[2] + e;
    if (!waves) return zmax < 0.f ? CK_DRY : (zmin > 0.f ? CK_WET : CK_CLIP); // depth = z (Ocean.cpp:163-166)
    // depth = z - (0 - b0 - b1), the b's within their bounds up to a few ulp of the largest texel (E)
    const float H = __fadd_ru(hmax0, hmax1), E = e + 1e-5f * H;
    if (zmax + H + E < 0.f) return CK_DRY;
    if (zmin - H - E > 0.f) return CK_WET;
    if (!(zmax + H + E >= 0.f) || local == 0) return CK_CLIP; // NaN somewhere: the general path
    float lo0 = -hmax0, hi0 = hmax0, lo1 = -hmax1, hi1 = hmax1;
    const ChanWin w0 = chan_window(wv.N, wv.gx, cw[0], cw[1], ext[0] + e, ext[1] + e);
    const ChanWin w1 = chan_window(wv.N, wv.gy, cw[0], cw[1], ext[0] + e, ext[1] + e);
    float v0[16], v1[16], c0[4], c1[4];
#pragma unroll
    for (int k = 0; k < 16; ++k) { v0[k] = __ldg(wv.ch0 + w0.idx[k]); v1[k] = __ldg(wv.ch1 + w1.idx[k]); }
#pragma unroll
    for (int k = 0; k < 4; ++k) { c0[k] = __ldg(wv.ch0 + w0.cen[k]); c1[k] = __ldg(wv.ch1 + w1.cen[k]); }
    // local = 2 (k_classify, one thread per chunk): the full local bound; local = 1: a channel whose footprint does not fit the
    // 4 x 4 window keeps its global bound instead of walking up to 8 x 8 texels row by row; local = 0 returned above
    if (w0.fast) chan_reduce(w0, v0, c0, lo0, hi0);
    else if (local == 2) wave_channel_bounds(wv.ch0, wv.N, wv.gx, hmax0, cw[0], cw[1], ext[0] + e, ext[1] + e, lo0, hi0);
    if (w1.fast) chan_reduce(w1, v1, c1, lo1, hi1);
    else if (local == 2) wave_channel_bounds(wv.ch1, wv.N, wv.gy, hmax1, cw[0], cw[1], ext[0] + e, ext[1] + e, lo1, hi1);
    if (zmax + (hi0 + hi1) + E < 0.f) return CK_DRY;
    if (zmin + (lo0 + lo1) - E > 0.f) return CK_WET;
    return CK_CLIP;
}

struct ClassifyParams {
    const BodyDev* bodies;
    const uint32_t* off_surf; // n_bodies + 1: exclusive scan of the crossing bodies' chunk counts
    uint32_t n_bodies;
    const OceanDev* ocean;
    WavesDev waves;
    SurfLists L;
    uint32_t* status;         // [0] |= 2 if the work list does not fit items_cap
    uint32_t wet_streams, clip_streams; // resident warps of the two consumers (segment sizing)
    uint32_t direct_streams;            // T of direct_split (0: direct mode off)
};

// Tiles first_tile, first_tile + stride, ... of the surface work list: classify, rank, compact. Block-wide (SFH_TILE threads).
__device__ __forceinline__ void classify_tiles(const ClassifyParams& P, uint32_t W, uint32_t first_tile, uint32_t stride)
{
    __shared__ uint32_t s_warp[32];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t n_tiles = (W + SFH_TILE - 1) / SFH_TILE;
    const bool waves = P.ocean->has_waves != 0;
    float hmax0 = 0.f, hmax1 = 0.f;
    if (waves) { hmax0 = __uint_as_float(__ldg(P.waves.hmax)); hmax1 = __uint_as_float(__ldg(P.waves.hmax + 1)); }

    for (uint32_t tile = first_tile; tile < n_tiles; tile += stride) {
        const uint32_t w = tile * SFH_TILE + tid;
        uint32_t kind = CK_DRY;
        uint4 desc = make_uint4(0, 0, 0, 0);
        if (w < W) {
            const uint32_t b = body_of_item(P.off_surf, P.n_bodies, w), c = w - P.off_surf[b];
            const BodyDev& bd = P.bodies[b];
            const float4* rec = reinterpret_cast<const float4*>(bd.bounds + c);
            const float4 b0 = __ldg(rec), b1 = __ldg(rec + 1);
            kind = classify_chunk(bd.R, bd.t, b0, make_float2(b1.x, b1.y), waves, P.waves, hmax0, hmax1);
            // the pair code of the wet role is compiled for the usual settings (buoyancy and damping, as Ocean::ApplyFluidForces sets
            // them); the wet chunks of a body without one of them (a non-buoyant part, an internal part of a compound) take the general code
            if (kind == CK_WET && (bd.flags & (BF_BUOYANCY | BF_DAMPING)) != (BF_BUOYANCY | BF_DAMPING)) kind = CK_CLIP;
            const unsigned long long addr = (unsigned long long)(bd.chunks + (size_t)__float_as_uint(b1.z) * 16u);
            desc = make_uint4((uint32_t)addr, (uint32_t)(addr >> 32), b, c | ((__float_as_uint(b1.w) >> 4) << 23));
        }
        // block-wide exclusive ranks of the wet (low half) and clip (high half) items: at most SFH_TILE of each, 16 bits suffice
        const uint32_t flag = (kind == CK_WET ? 1u : 0u) | (kind == CK_CLIP ? 0x10000u : 0u);
        uint32_t x = flag;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t y = __shfl_up_sync(SFH_FULL, x, off);
            if (lane >= (uint32_t)off) x += y;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) {
            uint32_t v = lane < SFH_TILE / 32u ? s_warp[lane] : 0u;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t y = __shfl_up_sync(SFH_FULL, v, off);
                if (lane >= (uint32_t)off) v += y;
            }
            s_warp[lane] = v; // inclusive over warps
        }
        __syncthreads();
        const uint32_t excl = (warp ? s_warp[warp - 1] : 0u) + x - flag;
        P.L.before[w] = excl;
        if (kind == CK_WET) P.L.items[tile * SFH_TILE + (excl & 0xffffu)] = desc;
        if (kind == CK_CLIP) P.L.items[tile * SFH_TILE + (SFH_TILE - 1u) - (excl >> 16)] = desc;
        if (tid == SFH_TILE - 1u) P.L.tile_cnt[tile] = make_uint2((excl + flag) & 0xffffu, (excl + flag) >> 16);
        __syncthreads(); // s_warp is reused by the next tile
    }
}

// Tile counts -> offsets, and the consumers' schedule. Block-wide, by ONE block, after every tile has been written.
// SPLIT: with fewer list entries than resident streams (a single hull: 783 chunks for 1776 warps) the step is bound by the time
// ONE warp needs for ONE chunk (4 face pairs per lane in sequence). A list entry is then handed out as 2 or 4 sub-items — each
// stages the chunk and transforms its vertices again (cheap next to the faces) and evaluates a half / a quarter of the faces.
__device__ __forceinline__ void finish_lists(const ClassifyParams& P, uint32_t W)
{
    constexpr uint32_t PER = 8; // tiles per thread and round
    __shared__ uint32_t s_carry[2];
    __shared__ uint32_t s_w2[2][32];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t n_tiles = (W + SFH_TILE - 1) / SFH_TILE;
    if (tid < 2) s_carry[tid] = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_tiles; base += SFH_TILE * PER) {
        const uint32_t i0 = base + tid * PER;
        uint2 c[PER];
        uint32_t ts = 0, tf = 0;
#pragma unroll
        for (uint32_t k = 0; k < PER; ++k) {
            c[k] = make_uint2(0, 0);
            if (i0 + k < n_tiles) c[k] = __ldcg(&P.L.tile_cnt[i0 + k]); // possibly written by other blocks of this launch: read through L2
            ts += c[k].x; tf += c[k].y;
        }
        uint32_t xs = ts, xf = tf;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t ys = __shfl_up_sync(SFH_FULL, xs, off), yf = __shfl_up_sync(SFH_FULL, xf, off);
            if (lane >= (uint32_t)off) { xs += ys; xf += yf; }
        }
        if (lane == 31) { s_w2[0][warp] = xs; s_w2[1][warp] = xf; }
        __syncthreads();
        if (warp == 0) {
            uint32_t ws = lane < SFH_TILE / 32u ? s_w2[0][lane] : 0u, wf = lane < SFH_TILE / 32u ? s_w2[1][lane] : 0u;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t ys = __shfl_up_sync(SFH_FULL, ws, off), yf = __shfl_up_sync(SFH_FULL, wf, off);
                if (lane >= (uint32_t)off) { ws += ys; wf += yf; }
            }
            s_w2[0][lane] = ws; s_w2[1][lane] = wf;
        }
        __syncthreads();
        uint32_t rs = s_carry[0] + (warp ? s_w2[0][warp - 1] : 0u) + xs - ts, rf = s_carry[1] + (warp ? s_w2[1][warp - 1] : 0u) + xf - tf;
#pragma unroll
        for (uint32_t k = 0; k < PER; ++k) {
            if (i0 + k < n_tiles) { P.L.wet_off[i0 + k] = rs; P.L.clip_off[i0 + k] = rf; }
            rs += c[k].x; rf += c[k].y;
        }
        __syncthreads();
        if (tid == 0) { s_carry[0] += s_w2[0][31]; s_carry[1] += s_w2[1][31]; }
        __syncthreads();
    }
    if (tid == 0) {
        const uint32_t wet = s_carry[0], clip = s_carry[1];
        P.L.wet_off[n_tiles] = wet;
        P.L.clip_off[n_tiles] = clip;
        uint32_t wet_split = 1, clip_split = 1;
        while (wet_split < 4u && (uint64_t)wet * wet_split * 2u <= P.wet_streams) wet_split *= 2u;
        if ((uint64_t)clip * 2u <= P.clip_streams) clip_split = 2u; // its per-vertex wave lookups are not cheap: halves only
        const uint32_t wet_sub = wet * wet_split, clip_sub = clip * clip_split;
        const uint32_t cap_w = max(P.wet_streams, 1u) * SFH_SEG_PER_STREAM, cap_c = max(P.clip_streams, 1u) * SFH_SEG_PER_STREAM;
        uint32_t* ctrl = P.L.ctrl;
        ctrl[SC_TICKET] = 0;
        ctrl[SC_WET_NEXT] = P.wet_streams; // every stream starts with the segment of its own index (no atomic, one round trip less)
        ctrl[SC_WET_SEG] = max(1u, (wet_sub + cap_w - 1) / cap_w);
        ctrl[SC_CLIP_NEXT] = P.clip_streams;
        ctrl[SC_CLIP_SEG] = max(1u, (clip_sub + cap_c - 1) / cap_c);
        ctrl[SC_WET_TOTAL] = wet_sub;
        ctrl[SC_CLIP_TOTAL] = clip_sub;
        ctrl[SC_ITEMS] = W;
        ctrl[SC_WET_SPLIT] = wet_split;
        ctrl[SC_CLIP_SPLIT] = clip_split;
        ctrl[SC_WET_ENTRIES] = wet;
        ctrl[SC_CLIP_ENTRIES] = clip;
    }
}

__device__ __forceinline__ uint32_t surface_items(const ClassifyParams& P)
{
    uint32_t W = P.off_surf[P.n_bodies];
    if (W > P.L.items_cap) { // cannot happen with host-side states (the host sizes the lists); device-resident steps are bounded
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(P.status, 2u);
        W = 0;
    }
    return W;
}

__global__ void __launch_bounds__(SFH_TILE, 4) k_classify(ClassifyParams P)
{
    __shared__ uint32_t s_last;
    pdl_trigger();
    pdl_wait();
    const bool fits = P.off_surf[P.n_bodies] <= P.L.items_cap;
    const uint32_t W = surface_items(P); // 0 (and a status flag) when the items do not fit: empty lists are built
    if (fits && direct_split(W, P.direct_streams)) return; // the consumers classify for themselves
    classify_tiles(P, W, blockIdx.x, gridDim.x);
    // the last block to finish turns the tile counts into offsets and sizes the consumers' segments
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&P.L.ctrl[SC_TICKET], 1u) == gridDim.x - 1u) ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    finish_lists(P, W);
}

// ---------------------------------------------------------------------------------------------------------------
// Walk over one of the two compacted lists: a cursor (tile, position in the tile's run), kept by every lane of the warp
// (uniform loads). Entry R of the list lives in the tile t with off[t] <= R < off[t + 1], at position R - off[t] of its run.
// ---------------------------------------------------------------------------------------------------------------
struct ListWalk {
    const uint4* items;
    const uint32_t* off;
    uint32_t n_tiles;
    bool back;        // clip entries are stored from the back of the tile's region
    uint32_t t, k, n; // cursor

    __device__ __forceinline__ void locate(uint32_t R)
    {
        uint32_t lo = 0, hi = n_tiles; // off[lo] <= R < off[hi]
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (__ldg(off + mid) <= R) lo = mid; else hi = mid;
        }
        t = lo;
        const uint32_t o = __ldg(off + t);
        k = R - o;
        n = __ldg(off + t + 1) - o;
    }
    __device__ __forceinline__ uint4 get() const { return __ldg(items + (size_t)t * SFH_TILE + (back ? (SFH_TILE - 1u) - k : k)); }
    __device__ __forceinline__ void next()
    {
        if (++k < n) return;
        k = 0;
        do {
            ++t;
            n = t < n_tiles ? __ldg(off + t + 1) - __ldg(off + t) : 1u; // past the end: never fetched
        } while (n == 0);
    }
};

struct SurfParams {
    const BodyDev* bodies;
    uint32_t n_bodies;
    const OceanDev* ocean;
    WavesDev waves;
    SurfLists L;
    const uint32_t* off_surf;
    PartialSurf* part_wet;     // one record per (segment, body) of the wet role
    PartialSurf* part_clip;    // ... of the clip role
    uint32_t wet_blocks;       // blocks [0, wet_blocks) of the grid work on the wet list, the others on the clip list
    uint32_t direct_streams;   // T of direct_split (0: direct mode off)
    uint8_t* codes;            // debug: per-face codes, nullptr otherwise
    const uint64_t* code_base; // debug: first code index of each body
};

#ifndef SFH_SURF_MIN_BLOCKS
#define SFH_SURF_MIN_BLOCKS 6
#endif

struct SurfBody { // per-body constants of the consumers
    BodyP2 B;
    F3 p, v, omega, cur, p0;
    bool buoy, damp;
    float rnorm;
};

__device__ __forceinline__ void surf_fetch_body(const BodyDev& sb, SurfBody& S)
{
    S.p = F3{sb.p[0], sb.p[1], sb.p[2]};
    S.v = F3{sb.v[0], sb.v[1], sb.v[2]};
    S.omega = F3{sb.omega[0], sb.omega[1], sb.omega[2]};
    S.cur = F3{sb.cur[0], sb.cur[1], sb.cur[2]};
    S.p0 = F3{S.p.x, S.p.y, 0.f}; // SolidEntity.cpp:1332-1334
    S.B.p = V3P{pk1(S.p.x), pk1(S.p.y), pk1(S.p.z)};
    S.B.v = V3P{pk1(S.v.x), pk1(S.v.y), pk1(S.v.z)};
    S.B.omega = V3P{pk1(S.omega.x), pk1(S.omega.y), pk1(S.omega.z)};
    S.B.cur = V3P{pk1(S.cur.x), pk1(S.cur.y), pk1(S.cur.z)};
    S.buoy = (sb.flags & BF_BUOYANCY) != 0;
    S.damp = (sb.flags & BF_DAMPING) != 0;
    float fro = 0.f;
    for (int k = 0; k < 9; ++k) fro += sb.R[k] * sb.R[k];
    S.rnorm = sqrtf(fro) * 1.000001f; // >= the 2-norm of the linear part of T_C
}

// which Jet / Pipe fields can reach this chunk at all (every face centroid lies inside the chunk's bounding sphere)
__device__ __forceinline__ uint64_t surf_current_mask(const BodyDev& sb, const ChunkView& cv, const CurGuard* s_guard, uint32_t n_cur, float rnorm)
{
    const F3 cw = xform(sb.R, sb.t, cv.sphere.x, cv.sphere.y, cv.sphere.z);
    const float reach = fabsf(sb.t[0]) + fabsf(sb.t[1]) + fabsf(sb.t[2]) + fabsf(cv.sphere.x) + fabsf(cv.sphere.y) + fabsf(cv.sphere.z) + cv.sphere.w;
    return chunk_current_mask(s_guard, n_cur, cw, rnorm * cv.sphere.w + (4e-6f * reach + 1e-6f));
}

// One wet chunk: every face is class 7 (SolidEntity.cpp:1605-1687), faces base+lane and base+lane+32 side by side in the pair
// arithmetic of k_submerged. BUOY / DAMP are the body's settings (warp-uniform), compiled in so that each combination is one
// straight-line block per face pair.
// acc: [0..11] Fdq Tdq Fdf Tdf | [12..14] Fb | [15..17] Tb | [18] Swet | [19] Vsub | [20..22] CBsub
template <bool WAVES, int CUR, bool DEBUG, bool BUOY, bool DAMP>
__device__ __forceinline__ void wet_chunk(const SurfParams& P, const ChunkView& cv, const float4* __restrict__ s_v, const SurfBody& S,
                                          const sfh_current* s_cur, uint64_t cur_may, uint32_t n_cur, const WaveConst& wc, double* acc,
                                          uint32_t& epack, uint32_t body, const uint32_t* __restrict__ orig, uint32_t lane, uint32_t f0,
                                          uint32_t nf /* faces [f0, nf) of the chunk */)
{
#pragma unroll 1
    for (uint32_t base = f0; base < nf; base += 64u) {
        const uint32_t iA = base + lane, iB = iA + 32u;
        const bool vA = iA < nf, vB = iB < nf;
        const uint32_t fA = vA ? cv.faces[iA] : 0u, fB = vB ? cv.faces[iB] : 0u;
        const float4 a1 = s_v[__byte_perm(fA, 0, 0x4440)], a2 = s_v[__byte_perm(fA, 0, 0x4441)], a3 = s_v[__byte_perm(fA, 0, 0x4442)];
        const float4 b1 = s_v[__byte_perm(fB, 0, 0x4440)], b2 = s_v[__byte_perm(fB, 0, 0x4441)], b3 = s_v[__byte_perm(fB, 0, 0x4442)];
        const V3P q1 = {pk(a1.x, b1.x), pk(a1.y, b1.y), pk(a1.z, b1.z)};
        const V3P q2 = {pk(a2.x, b2.x), pk(a2.y, b2.y), pk(a2.z, b2.z)};
        const V3P q3 = {pk(a3.x, b3.x), pk(a3.y, b3.y), pk(a3.z, b3.z)};
        if (BUOY) { // tetrahedron (p0, p1, p2, p3) of the submerged volume, before the degenerate test (SolidEntity.cpp:1612-1618)
            const V3P o0 = {pk1(S.p0.x), pk1(S.p0.y), pk1(S.p0.z)};
            const V3P t1 = q1 - o0, t2 = q2 - o0, t3 = q3 - o0;
            const P2 v6 = dot(t1, cross(t2, t3));
            if (!WAVES) { // CBsub only reaches an output on a flat sea (:1703-1706)
                const V3P m = (((t1 + t2) + t3) * pk1(0.25f)) * v6; // "/ 4.f": an exact scaling
                acc3_if(vA, acc + 20, lo(m.x), lo(m.y), lo(m.z));
                acc3_if(vB, acc + 20, hi(m.x), hi(m.y), hi(m.z));
            }
            acc1_if(vA, acc[19], lo(v6));
            acc1_if(vB, acc[19], hi(v6));
        }
        const V3P fn = cross(q2 - q1, q3 - q1);
        const P2 len2 = dot(fn, fn);
        const bool dA = lo(len2) < 1e-12f, dB = hi(len2) < 1e-12f; // :1632
        const bool mA = vA && !dA, mB = vB && !dB;
        epack += (vA && dA ? 1u : 0u) + (vB && dB ? 1u : 0u);
        uint32_t code = 0;
        {
            const P2 len = ieee_sqrt2(len2); // lanes with len2 < 1e-12 compute garbage that is masked out
            const P2 A = len * pk1(0.5f);
            const V3P csum = (q1 + q2) + q3;
            const P2 sx = fn.x * fn.x, sy = fn.y * fn.y, sz = fn.z * fn.z; // range guard of the fast forms: see k_submerged
            const float m_fn = fminf(fminf(fminf(lo(sx), lo(sy)), lo(sz)), fminf(fminf(hi(sx), hi(sy)), hi(sz)));
            const float m_cs = fminf(fminf(fminf(fabsf(lo(csum.x)), fabsf(lo(csum.y))), fabsf(lo(csum.z))),
                                     fminf(fminf(fabsf(hi(csum.x)), fabsf(hi(csum.y))), fabsf(hi(csum.z))));
            const bool cheap_ok = (m_fn >= 0x1p-72f) & (m_cs >= 0x1p-100f);
            V3P fn1 = normalize_fast(fn, len);
            V3P fc = {ieee_div3_2(csum.x), ieee_div3_2(csum.y), ieee_div3_2(csum.z)};
            if (!cheap_ok && !((min_key(fn) >= SFH_NUM_KEY) & (min_key(csum) >= SFH_TINY_KEY)))
                normalize_centroid_slow(fn, len, csum, fn1, fc);
            if (WAVES && BUOY) { // pressure buoyancy at the centroid (SolidEntity.cpp:1650-1658)
                const V3P Fbi = (fn1 * neg(A)) * depth_pair<WAVES>(P.waves, wc, fc); // ((-fn1) * A) * depthc: negating A instead of fn1 flips the same sign bits
                const V3P Tbi = cross(fc - S.B.p, Fbi);
                acc6_if(mA, acc + 12, lo(Fbi.x), lo(Fbi.y), lo(Fbi.z), lo(Tbi.x), lo(Tbi.y), lo(Tbi.z));
                acc6_if(mB, acc + 12, hi(Fbi.x), hi(Fbi.y), hi(Fbi.z), hi(Tbi.x), hi(Tbi.y), hi(Tbi.z));
            }
            if (DAMP) code = damping_pair<CUR>(S.B, s_cur, cur_may, n_cur, fc, fn1, A, mA, mB, acc, epack);
            acc1_if(mA, acc[18], lo(A)); // Swet (:1686)
            acc1_if(mB, acc[18], hi(A));
        }
        if (DEBUG) {
            if (vA) P.codes[P.code_base[body] + orig[iA]] = (uint8_t)(7u | (dA ? SFH_FACE_DEGENERATE : 0u) | (code & 255u));
            if (vB) P.codes[P.code_base[body] + orig[iB]] = (uint8_t)(7u | (dB ? SFH_FACE_DEGENERATE : 0u) | (code >> 8));
        }
    }
}

// Vertex phase of a clip chunk: world position and Ocean::GetDepth per vertex (SolidEntity.cpp:1343-1351; same bits for every incident
// face). (Two vertices per lane and trip, to overlap their texel loads, was measured: 13 % slower at full size — it evaluates up to
// 63 slots past the last vertex — and no faster on a single small hull.)
template <bool WAVES>
__device__ __forceinline__ void clip_vertices(const WavesDev& wv, const WaveConst& wc, const BodyDev& sb, const ChunkView& cv, float4* __restrict__ s_v,
                                              uint32_t lane)
{
#pragma unroll 1
    for (uint32_t i = lane; i < cv.nv; i += 32u) {
        const F3 pw = xform(sb.R, sb.t, cv.x[i], cv.y[i], cv.z[i]);
        s_v[i] = make_float4(pw.x, pw.y, pw.z, get_depth<WAVES>(wv, wc, pw.x, pw.y, pw.z));
    }
}
// The same for a SUB-ITEM (faces [f0, f1) of the chunk): only the vertices those faces refer to. Vertex slots are handed out in the
// order the faces first use them, so a quarter of the faces touches two or three of the eight groups of 32 slots: the other trips —
// two wave lookups per vertex each — are skipped. `need`: 8 words of the warp's shared memory.
template <bool WAVES>
__device__ __forceinline__ void clip_vertices_of(const WavesDev& wv, const WaveConst& wc, const BodyDev& sb, const ChunkView& cv, float4* __restrict__ s_v,
                                                 uint32_t lane, uint32_t f0, uint32_t f1, uint32_t* __restrict__ need)
{
    if (lane < 8u) need[lane] = 0u;
    __syncwarp();
    for (uint32_t slot = f0 + lane; slot < f1; slot += 32u) {
        const uint32_t fp = cv.faces[slot];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const uint32_t id = (fp >> (8 * k)) & 255u;
            atomicOr(&need[id >> 5], 1u << (id & 31u));
        }
    }
    __syncwarp();
#pragma unroll 1
    for (uint32_t g = 0; g < 8u; ++g) {
        const uint32_t m = need[g]; // warp-uniform
        if (m == 0u) continue;
        const uint32_t i = lane + 32u * g;
        if ((m >> lane) & 1u) {
            const F3 pw = xform(sb.R, sb.t, cv.x[i], cv.y[i], cv.z[i]);
            s_v[i] = make_float4(pw.x, pw.y, pw.z, get_depth<WAVES>(wv, wc, pw.x, pw.y, pw.z));
        }
    }
}

// (Measured and rejected: a first pass that sorts the slots of a clip chunk into a class-7 list — sent through the pair code of the wet
// role — and a list of cut faces for the general code: the clip role got 25 % SLOWER (0.102 -> 0.127 ms per 8212 chunks; twice the code,
// 200-360 bytes of spills, list traffic), cfg 5 fell from 47.3 to 42.6 G tri-evals/s.)
// One chunk the water line may cross: per-vertex depths were left in s_v[].w; the reference's eight classes, one face per lane
// at a time (SolidEntity.cpp:1347-1687).
template <bool WAVES, int CUR, bool DEBUG>
__device__ __forceinline__ void clip_chunk(const SurfParams& P, const ChunkView& cv, const float4* __restrict__ s_v, const SurfBody& S,
                                           const sfh_current* s_cur, uint64_t cur_may, uint32_t n_cur, const WaveConst& wc, double* acc,
                                           uint32_t& epack, uint64_t& cpack, uint32_t body, const uint32_t* __restrict__ orig, uint32_t lane,
                                           uint32_t f0, uint32_t nf /* faces [f0, nf) of the chunk */)
{
    const bool buoy = S.buoy, damp = S.damp;
    BodyRegs B;
    B.p = S.p; B.v = S.v; B.omega = S.omega; B.cur = S.cur;
    const F3 p0 = S.p0;
#pragma unroll 1
    for (uint32_t slot = f0 + lane; slot < nf; slot += 32u) {
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
            // tetrahedron (p0, a, b, c) of the submerged volume (SolidEntity.cpp:1371-1377 and twins). Vsub / CBsub only reach an
            // output through the buoyancy block (:1690-1708), and CBsub only on a flat sea (:1703-1706): they are not accumulated
            // where the reference computes them in vain.
            auto tetra = [&](F3 pa, F3 pb, F3 pc) {
                if (!buoy) return;
                const float v6 = dot(pa, cross(pb, pc));
                if (!WAVES) {
                    const F3 cg = ((pa + pb) + pc) * 0.25f; // "/ 4.f": an exact scaling either way
                    const F3 m = cg * v6;
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
            const F3 fn = cross(fv1, fv2);
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
                    const float depthc = get_depth<WAVES>(P.waves, wc, fc.x, fc.y, fc.z);
                    const F3 Fbi = ((-fn1) * A) * depthc;
                    const F3 Tbi = cross(fc - B.p, Fbi);
                    acc[12] = dadd(acc[12], (double)Fbi.x); acc[13] = dadd(acc[13], (double)Fbi.y); acc[14] = dadd(acc[14], (double)Fbi.z);
                    acc[15] = dadd(acc[15], (double)Tbi.x); acc[16] = dadd(acc[16], (double)Tbi.y); acc[17] = dadd(acc[17], (double)Tbi.z);
                }
                if (damp) code |= damping_terms<CUR>(B, s_cur, cur_may, n_cur, fc, fn1, A, acc, epack);
                acc[18] = dadd(acc[18], (double)A); // SolidEntity.cpp:1686
            }
        }
        if (DEBUG) P.codes[P.code_base[body] + orig[slot]] = (uint8_t)code;
    }
}

// One role of the consumer kernel: CLIP = false works through the wet list, CLIP = true through the clip list.
template <bool CLIP, bool WAVES, int CUR, bool DEBUG>
__device__ __forceinline__ void surface_role(const SurfParams& P, WarpStageSurf& St, const sfh_current* s_cur, const CurGuard* s_guard, uint32_t n_cur,
                                             uint32_t lane, uint32_t stream /* this warp's index among the role's warps */)
{
    uint32_t* const ctrl = P.L.ctrl;
    const uint32_t total = ctrl[CLIP ? SC_CLIP_TOTAL : SC_WET_TOTAL], SC = ctrl[CLIP ? SC_CLIP_SEG : SC_WET_SEG]; // in sub-items
    if (total == 0) return;
    const uint32_t split = ctrl[CLIP ? SC_CLIP_SPLIT : SC_WET_SPLIT], sh = 31u - __clz(split), part = SFH_CH_FACES >> sh;
    BodyDev& s_body = St.body;
    float4* const s_v = St.v;
    const PipeAddr pa = warp_pipe_init(St.pipe, lane);
    const WaveConst wc = WAVES ? wave_const(P.waves) : WaveConst{};
    PartialSurf* const partials = CLIP ? P.part_clip : P.part_wet;

    ListWalk lw;
    lw.items = P.L.items;
    lw.off = CLIP ? P.L.clip_off : P.L.wet_off;
    lw.n_tiles = (ctrl[SC_ITEMS] + SFH_TILE - 1) / SFH_TILE;
    lw.back = CLIP;

    double acc[SFH_NACC_SURF];
#pragma unroll
    for (int i = 0; i < SFH_NACC_SURF; ++i) acc[i] = 0.0;
    uint32_t epack = 0, since_spill = 0;
    uint64_t cpack = 0;
    uint32_t cnt[3] = {0, 0, 0}; // degenerate, quadratic, skin of the current body (warp-uniform)
    uint32_t cls[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int s = 0;
    uint32_t parity = 0;
    SurfBody S;

    auto spill = [&]() { // packed per-lane event counters -> warp-uniform totals
#pragma unroll
        for (int k = 0; k < 3; ++k) cnt[k] += __reduce_add_sync(SFH_FULL, (epack >> (10 * k)) & 1023u);
        epack = 0;
        if (CLIP) {
#pragma unroll
            for (int k = 1; k < 8; ++k) cls[k] += __reduce_add_sync(SFH_FULL, (uint32_t)(cpack >> (8 * k)) & 255u);
            cpack = 0;
        }
        since_spill = 0;
    };
    auto flush = [&](uint32_t seg, uint32_t body) { // one partial record per (segment, body)
        spill();
        PartialSurf* rec = partials + (seg + body);
        warp_sum_store<SFH_NACC_SURF>(acc, rec->a, lane);
        if (lane == 0) {
            rec->n_degenerate = cnt[0]; rec->n_quadratic = cnt[1]; rec->n_skin = cnt[2];
#pragma unroll
            for (int k = 0; k < 8; ++k) rec->cls[k] = cls[k];
        }
        cnt[0] = cnt[1] = cnt[2] = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) cls[k] = 0;
    };
    auto issue = [&](int buf, const uint4& d) {
        pipe_issue(pa, buf, reinterpret_cast<const uint8_t*>(((unsigned long long)d.y << 32) | d.x), (d.w >> 23) << 4);
    };

    bool first = true;
    for (;;) { // segments of sub-items: the stream's own index first, then whatever the counter hands out
        uint32_t seg = stream;
        if (!first) {
            if (lane == 0) seg = atomicAdd(&ctrl[CLIP ? SC_CLIP_NEXT : SC_WET_NEXT], 1u);
            seg = __shfl_sync(SFH_FULL, seg, 0);
        }
        first = false;
        const uint64_t R0 = (uint64_t)seg * SC;
        if (R0 >= total) break;
        const uint32_t R1 = (uint32_t)min(R0 + SC, (uint64_t)total);
        uint32_t R = (uint32_t)R0;
        // sub-item R = part (R & (split - 1)) of list entry R >> sh; (dw, qw) is the furthest sub-item fetched so far
        lw.locate(R >> sh);
        uint4 dw = lw.get();
        uint32_t qw = R & (split - 1u);
        auto step = [&]() {
            if (++qw == split) { qw = 0; lw.next(); dw = lw.get(); }
        };
        uint4 d_cur = dw, d_next = dw, d_after = dw;
        uint32_t q_cur = qw, q_next = qw, q_after = qw;
        if (lane == 0) issue(s, d_cur);
        if (R + 1 < R1) { step(); d_next = dw; q_next = qw; }
        uint32_t body = d_cur.z;
        __syncwarp();
        warp_load_body(&s_body, P.bodies + body, lane);
        surf_fetch_body(s_body, S);

        for (;;) { // sub-items of the segment
            pipe_wait(pa, s, (parity >> s) & 1u);
            parity ^= 1u << s;
            const ChunkView cv = chunk_view(St.pipe.buf[s]);
            __syncwarp(); // every lane is done with the previous chunk: s_v and the other staging buffer are free
            if (lane == 0 && R + 1 < R1) issue(s ^ 1, d_next);
            if (R + 2 < R1) { step(); d_after = dw; q_after = qw; } // the descriptor after next has a whole chunk to arrive

            if (!CLIP) { // vertices: slots 0..191 in one straight-line block, see k_submerged
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
            } else { // ... and Ocean::GetDepth per vertex (SolidEntity.cpp:1349-1351): same bits for every incident face
                if (split == 1u) clip_vertices<WAVES>(P.waves, wc, s_body, cv, s_v, lane);
                else clip_vertices_of<WAVES>(P.waves, wc, s_body, cv, s_v, lane, q_cur * part, min(cv.nf, q_cur * part + part), St.need);
            }
            __syncwarp();

            const uint64_t cur_may = CUR ? surf_current_mask(s_body, cv, s_guard, n_cur, S.rnorm) : 0ull;
            const uint32_t* orig = DEBUG ? s_body.orig_face + (size_t)(d_cur.w & 0x7fffffu) * SFH_CH_FACES : nullptr;
            const uint32_t f0 = q_cur * part, f1 = min(cv.nf, f0 + part); // this sub-item's faces
            if (!CLIP) {
                cls[7] += f1 > f0 ? f1 - f0 : 0u;
#if defined(SFH_WET_ONLY_DAMP) /* timing probes of a split wet role (results are NOT the reference's): DESIGN.md 7b */
                wet_chunk<WAVES, CUR, DEBUG, false, true>(P, cv, s_v, S, s_cur, cur_may, n_cur, wc, acc, epack, body, orig, lane, f0, f1);
#elif defined(SFH_WET_ONLY_BUOY)
                wet_chunk<WAVES, CUR, DEBUG, true, false>(P, cv, s_v, S, s_cur, cur_may, n_cur, wc, acc, epack, body, orig, lane, f0, f1);
#else
                wet_chunk<WAVES, CUR, DEBUG, true, true>(P, cv, s_v, S, s_cur, cur_may, n_cur, wc, acc, epack, body, orig, lane, f0, f1); // k_classify only lists such bodies as wet
#endif
            } else {
                clip_chunk<WAVES, CUR, DEBUG>(P, cv, s_v, S, s_cur, cur_may, n_cur, wc, acc, epack, cpack, body, orig, lane, f0, f1);
            }
            if (++since_spill == SFH_LIST_SPILL_EVERY) spill(); // <= 8 faces per lane per sub-item: 10-bit / 8-bit fields stay in range

            const bool last = ++R >= R1;
            if (last || d_next.z != body) flush(seg, body); // the one flush site of the role
            if (last) break;
            s ^= 1;
            if (d_next.z != body) {
                body = d_next.z;
                __syncwarp();
                warp_load_body(&s_body, P.bodies + body, lane);
                surf_fetch_body(s_body, S);
            }
            d_cur = d_next; q_cur = q_next;
            d_next = d_after; q_next = q_after;
        }
        s ^= 1; // the next segment's first copy goes to the other staging buffer
    }
}

// Direct mode: sub-item u of the UNCLASSIFIED surface list, start to finish by this warp; one record at index u + body, zeros
// for a dry chunk. Same per-face code as the roles, so the per-face terms are the same bits whichever mode ran.
template <bool WAVES, int CUR, bool DEBUG>
__device__ __forceinline__ void surface_direct(const SurfParams& P, WarpStageSurf& St, const sfh_current* s_cur, const CurGuard* s_guard, uint32_t n_cur,
                                               uint32_t lane, uint32_t u, uint32_t split)
{
    const uint32_t sh = 31u - __clz(split), part = SFH_CH_FACES >> sh;
    const uint32_t w = u >> sh, q = u & (split - 1u);
    const uint32_t body = body_of_item(P.off_surf, P.n_bodies, w), c = w - P.off_surf[body];
    BodyDev& s_body = St.body;
    float4* const s_v = St.v;
    warp_load_body(&s_body, P.bodies + body, lane);
    const float4* rec = reinterpret_cast<const float4*>(s_body.bounds + c);
    const float4 b0 = __ldg(rec), b1 = __ldg(rec + 1);
    // the chunk is fetched while it is being classified (a dry chunk's copy is wasted: direct mode is bound by latency, not traffic)
    const PipeAddr pa = warp_pipe_init(St.pipe, lane);
    if (lane == 0) pipe_issue(pa, 0, s_body.chunks + (size_t)__float_as_uint(b1.z) * 16u, __float_as_uint(b1.w));
    float hmax0 = 0.f, hmax1 = 0.f;
    if (WAVES) { hmax0 = __uint_as_float(__ldg(P.waves.hmax)); hmax1 = __uint_as_float(__ldg(P.waves.hmax + 1)); }
    // Direct mode keeps the GLOBAL wave bound (local = 0): the step waits for its slowest warp, and that warp is one with a cut chunk
    // whatever the others are spared; the local bound costs every warp its texel round trips. Measured host to host (5k-face hull /
    // 200k-face hull crossing waves, nearly every chunk inside the global band): full local bound 40.7 / 50.3 us, 4 x 4 window only
    // 36.5 / 46.6 us, global bound 34.6 / 45.3 us (r01, no classification at all: 33.5 / 54.5 us).
#ifndef SFH_DIRECT_LOCAL
#define SFH_DIRECT_LOCAL 0
#endif
    uint32_t kind = classify_chunk(s_body.R, s_body.t, b0, make_float2(b1.x, b1.y), WAVES, P.waves, hmax0, hmax1, SFH_DIRECT_LOCAL);
    if (kind == CK_WET && (s_body.flags & (BF_BUOYANCY | BF_DAMPING)) != (BF_BUOYANCY | BF_DAMPING)) kind = CK_CLIP; // as k_classify

    double acc[SFH_NACC_SURF];
#pragma unroll
    for (int i = 0; i < SFH_NACC_SURF; ++i) acc[i] = 0.0;
    uint32_t epack = 0, cls7 = 0;
    uint64_t cpack = 0;
    if (kind == CK_DRY) pipe_wait(pa, 0, 0u); // no warp leaves with a copy into its shared memory in flight
    if (kind != CK_DRY) {
        const WaveConst wc = WAVES ? wave_const(P.waves) : WaveConst{};
        SurfBody S;
        surf_fetch_body(s_body, S);
        pipe_wait(pa, 0, 0u);
        const ChunkView cv = chunk_view(St.pipe.buf[0]);
        const uint64_t cur_may = CUR ? surf_current_mask(s_body, cv, s_guard, n_cur, S.rnorm) : 0ull;
        const uint32_t* orig = DEBUG ? s_body.orig_face + (size_t)c * SFH_CH_FACES : nullptr;
        const uint32_t f0 = q * part, f1 = min(cv.nf, f0 + part);
        if (kind == CK_WET) {
#pragma unroll 1
            for (uint32_t i = lane; i < cv.nv; i += 32u) {
                const F3 pw = xform(s_body.R, s_body.t, cv.x[i], cv.y[i], cv.z[i]);
                s_v[i] = make_float4(pw.x, pw.y, pw.z, 0.f);
            }
            __syncwarp();
            cls7 = f1 > f0 ? f1 - f0 : 0u;
            wet_chunk<WAVES, CUR, DEBUG, true, true>(P, cv, s_v, S, s_cur, cur_may, n_cur, wc, acc, epack, body, orig, lane, f0, f1);
        } else {
            if (split == 1u) clip_vertices<WAVES>(P.waves, wc, s_body, cv, s_v, lane);
            else clip_vertices_of<WAVES>(P.waves, wc, s_body, cv, s_v, lane, f0, f1, St.need);
            __syncwarp();
            clip_chunk<WAVES, CUR, DEBUG>(P, cv, s_v, S, s_cur, cur_may, n_cur, wc, acc, epack, cpack, body, orig, lane, f0, f1);
        }
    }
    PartialSurf* out = P.part_wet + (u + body);
    warp_sum_store<SFH_NACC_SURF>(acc, out->a, lane);
    uint32_t cnt[3], cl[8];
#pragma unroll
    for (int k = 0; k < 3; ++k) cnt[k] = __reduce_add_sync(SFH_FULL, (epack >> (10 * k)) & 1023u);
#pragma unroll
    for (int k = 1; k < 8; ++k) cl[k] = __reduce_add_sync(SFH_FULL, (uint32_t)(cpack >> (8 * k)) & 255u);
    if (lane == 0) {
        out->n_degenerate = cnt[0]; out->n_quadratic = cnt[1]; out->n_skin = cnt[2];
        out->cls[0] = 0;
#pragma unroll
        for (int k = 1; k < 7; ++k) out->cls[k] = cl[k];
        out->cls[7] = cl[7] + cls7;
    }
}

// The consumer of both lists, ONE launch: blocks [0, wet_blocks) take the wet list, the others the clip list. On a full-size
// step the wet blocks fill the machine first and clip blocks move in as they leave — no kernel boundary in between; on a small
// step both roles are resident together and their latencies overlap.
template <bool WAVES, int CUR, bool DEBUG>
#ifdef SFH_SURF_MAXNREG /* experiments: a register cap between the steps __launch_bounds__ offers */
__global__ void __maxnreg__(SFH_SURF_MAXNREG) k_surface_lists(SurfParams P)
#else
__global__ void __launch_bounds__(SFH_SUB_THREADS, SFH_SURF_MIN_BLOCKS) k_surface_lists(SurfParams P)
#endif
{
    __shared__ WarpStageSurf s_stage[SFH_SUB_WPB];
    __shared__ sfh_current s_cur[CUR ? SFH_MAX_CURRENTS : 1];
    __shared__ CurGuard s_guard[CUR ? SFH_MAX_CURRENTS : 1];

    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t n_cur = CUR ? P.ocean->n_currents : 0;
    if constexpr (CUR != 0) {
        stage_currents(s_cur, s_guard, P.ocean, n_cur);
        __syncthreads(); // the only block barrier of the kernel, before any warp can leave
    }
    pdl_trigger();
    pdl_wait(); // the prologue's (and k_classify's) results
    const uint32_t W = P.off_surf[P.n_bodies], dsplit = W <= P.L.items_cap ? direct_split(W, P.direct_streams) : 0u;
    if (dsplit) { // small step: every (sub-)item has a warp of its own, no lists
        const uint32_t u = blockIdx.x * SFH_SUB_WPB + warp;
        if (u < W * dsplit) surface_direct<WAVES, CUR, DEBUG>(P, s_stage[warp], s_cur, s_guard, n_cur, lane, u, dsplit);
        if (u == 0 && lane == 0) { // what sfh_get_timings reports: no lists were built
            P.L.ctrl[SC_ITEMS] = W; P.L.ctrl[SC_WET_ENTRIES] = 0; P.L.ctrl[SC_CLIP_ENTRIES] = 0;
        }
        return;
    }
    if (blockIdx.x < P.wet_blocks) surface_role<false, WAVES, CUR, DEBUG>(P, s_stage[warp], s_cur, s_guard, n_cur, lane, blockIdx.x * SFH_SUB_WPB + warp);
    else surface_role<true, WAVES, CUR, DEBUG>(P, s_stage[warp], s_cur, s_guard, n_cur, lane, (blockIdx.x - P.wet_blocks) * SFH_SUB_WPB + warp);
}

} // namespace sfh
