// mesh_chunker.hpp — upload-time re-layout of an indexed triangle mesh into device "chunks".
//
// Input  : what the reference's Mesh exposes (OpenGLDataStructs.h:120-262): strided vertex positions and
//          3 x u32 vertex ids per face.
// Output : a packed stream of chunks, each holding <= SFH_CH_FACES faces and the <= SFH_CH_VERTS distinct
//          vertices they reference, sized and aligned so that ONE bulk async copy (cp.async.bulk, 16-byte
//          granularity) brings a whole chunk into shared memory:
//              +0                u32 n_verts, u32 n_faces, u32 next_bytes, u32 first_slot
//              +16               float cx, cy, cz, r  bounding sphere of the chunk's vertices in the mesh frame (r rounded up)
//              +32               float x[nv4]        nv4 = n_verts rounded up to a multiple of 4
//              +32 + 4*nv4       float y[nv4]
//              +32 + 8*nv4       float z[nv4]
//              +32 + 12*nv4      u32   face[nf4]     local ids packed i0 | i1<<8 | i2<<16, nf4 = n_faces up to 4
//          The sphere lets k_surface drop a chunk that lies entirely above the highest possible water surface after reading
//          32 bytes of it: all its faces are class 0 and contribute nothing (SolidEntity.cpp:1353-1360).
//          `next_bytes` is the size of the FOLLOWING chunk of the same mesh (0 for the last), so a block can
//          issue the copy of chunk c+1 while it still works on chunk c without a dependent header read.
//          chunk_off[c] (16-byte units, n_chunks+1 entries) gives random access for a block's first chunk.
// A thread block transforms every vertex of a chunk ONCE into shared memory (the reference transforms each vertex
// once per incident face, SolidEntity.cpp:1343-1345 — same bits either way) and faces read 4 bytes of indices
// instead of 12. Faces are Morton-ordered by centroid first (unless SFH_MESH_KEEP_ORDER) so that consecutive
// faces share vertices; the per-body sums are order-independent within the fp64 contract and the class counts
// exactly. `orig_face[c*SFH_CH_FACES + slot]` maps a chunk slot back to the caller's face index (debug codes).
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#ifndef SFH_CH_FACES
#define SFH_CH_FACES 256u
#endif
#ifndef SFH_CH_VERTS
#define SFH_CH_VERTS 256u
#endif
#define SFH_CH_HDR_BYTES 32u
#define SFH_CH_MAX_BYTES (SFH_CH_HDR_BYTES + 12u * SFH_CH_VERTS + 4u * SFH_CH_FACES) /* 4128 */

namespace sfh {

// One record per chunk, read (coalesced, 32 bytes) by the chunk classifier of the surface path: the axis-aligned box of the
// chunk's vertices in the mesh frame — centre and half extents, the extents rounded up so that the box contains every vertex —
// and where the chunk lies in the packed stream. A box follows a flat, elongated patch of a hull far better than a sphere does
// (cfg 3: 13 % of the chunks straddle the water line by their boxes, 61 % by their spheres).
struct ChunkBounds {
    float cx, cy, cz, hx;
    float hy, hz;
    uint32_t off16;  // offset of the chunk in the stream, 16-byte units
    uint32_t bytes;  // its size (multiple of 16)
};
static_assert(sizeof(ChunkBounds) == 32, "ChunkBounds is read as two float4");

struct ChunkedMesh {
    std::vector<uint8_t> chunks;     // packed chunk stream
    std::vector<ChunkBounds> bounds; // n_chunks
    std::vector<uint32_t> chunk_off; // n_chunks + 1 offsets into `chunks`, in 16-byte units
    std::vector<uint32_t> orig_face; // n_chunks * SFH_CH_FACES, 0xFFFFFFFF for unused slots
    uint32_t n_chunks = 0, n_faces = 0, n_verts = 0, n_chunk_verts = 0;
    uint64_t device_bytes = 0; // bytes one evaluation reads (= chunks.size())
};

inline uint64_t morton_spread21(uint64_t v)
{
    v &= 0x1fffffull;
    v = (v | v << 32) & 0x1f00000000ffffull;
    v = (v | v << 16) & 0x1f0000ff0000ffull;
    v = (v | v << 8) & 0x100f00f00f00f00full;
    v = (v | v << 4) & 0x10c30c30c30c30c3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
}

// Returns false when a face references a vertex >= n_verts.
inline bool chunk_mesh(const void* pos, size_t stride, uint32_t n_verts, const uint32_t* faces, uint32_t n_faces,
                       bool keep_order, ChunkedMesh& out)
{
    auto vptr = [&](uint32_t v) { return reinterpret_cast<const float*>(static_cast<const char*>(pos) + (size_t)v * stride); };
    for (size_t i = 0; i < (size_t)n_faces * 3; ++i)
        if (faces[i] >= n_verts) return false;

    std::vector<uint32_t> order(n_faces);
    for (uint32_t i = 0; i < n_faces; ++i) order[i] = i;
    if (!keep_order && n_faces > SFH_CH_FACES) {
        float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
        for (uint32_t v = 0; v < n_verts; ++v) {
            const float* p = vptr(v);
            for (int k = 0; k < 3; ++k) {
                if (p[k] < lo[k]) lo[k] = p[k];
                if (p[k] > hi[k]) hi[k] = p[k];
            }
        }
        double inv[3];
        for (int k = 0; k < 3; ++k) inv[k] = (hi[k] > lo[k]) ? 2097151.0 / ((double)hi[k] - (double)lo[k]) : 0.0;
        std::vector<std::pair<uint64_t, uint32_t>> keyed(n_faces);
        for (uint32_t f = 0; f < n_faces; ++f) {
            const float* a = vptr(faces[3 * f]);
            const float* b = vptr(faces[3 * f + 1]);
            const float* c = vptr(faces[3 * f + 2]);
            uint64_t code = 0;
            for (int k = 0; k < 3; ++k) {
                double cen = ((double)a[k] + (double)b[k] + (double)c[k]) / 3.0;
                double q = (cen - (double)lo[k]) * inv[k];
                if (!(q >= 0.0)) q = 0.0; // also catches NaN
                if (q > 2097151.0) q = 2097151.0;
                code |= morton_spread21((uint64_t)q) << k;
            }
            keyed[f] = {code, f};
        }
        std::sort(keyed.begin(), keyed.end());
        for (uint32_t i = 0; i < n_faces; ++i) order[i] = keyed[i].second;
    }

    out = ChunkedMesh();
    out.n_faces = n_faces;
    out.n_verts = n_verts;
    std::vector<uint32_t> stamp(n_verts, 0xFFFFFFFFu);
    std::vector<uint16_t> local(n_verts, 0);

    // staging of the chunk being filled
    std::vector<float> cx(SFH_CH_VERTS), cy(SFH_CH_VERTS), cz(SFH_CH_VERTS);
    std::vector<uint32_t> cf(SFH_CH_FACES);
    uint32_t chunk_id = 0, nv = 0, nf = 0;
    size_t prev_hdr = (size_t)-1;

    auto end_chunk = [&]() {
        const uint32_t nv4 = (nv + 3u) & ~3u, nf4 = (nf + 3u) & ~3u;
        const uint32_t bytes = SFH_CH_HDR_BYTES + 12u * nv4 + 4u * nf4;
        const size_t base = out.chunks.size();
        out.chunk_off.push_back((uint32_t)(base / 16));
        out.chunks.resize(base + bytes, 0);
        uint8_t* p = out.chunks.data() + base;
        const uint32_t hdr[4] = {nv, nf, 0u, chunk_id * SFH_CH_FACES};
        std::memcpy(p, hdr, sizeof(hdr));
        // bounding sphere: centre of the AABB, radius = largest distance in double, rounded up (NaN coordinates give a NaN
        // sphere, which never passes the "entirely dry" test)
        double lo3[3] = {1e300, 1e300, 1e300}, hi3[3] = {-1e300, -1e300, -1e300};
        for (uint32_t v = 0; v < nv; ++v) {
            const double q[3] = {cx[v], cy[v], cz[v]};
            for (int k = 0; k < 3; ++k) { lo3[k] = std::min(lo3[k], q[k]); hi3[k] = std::max(hi3[k], q[k]); }
        }
        const float cen[3] = {(float)(0.5 * (lo3[0] + hi3[0])), (float)(0.5 * (lo3[1] + hi3[1])), (float)(0.5 * (lo3[2] + hi3[2]))};
        double r2 = 0.0;
        bool finite = true;
        for (uint32_t v = 0; v < nv; ++v) {
            const double dx = (double)cx[v] - cen[0], dy = (double)cy[v] - cen[1], dz = (double)cz[v] - cen[2];
            const double d2 = dx * dx + dy * dy + dz * dz;
            if (!(d2 <= 1e300)) finite = false;
            r2 = std::max(r2, d2);
        }
        float rad = (float)(std::sqrt(r2) * (1.0 + 1e-6));
        rad = std::nextafter(rad, 3.4e38f);
        const float sphere[4] = {cen[0], cen[1], cen[2], finite ? rad : NAN};
        std::memcpy(p + 16, sphere, sizeof(sphere));
        ChunkBounds cb;
        cb.cx = cen[0]; cb.cy = cen[1]; cb.cz = cen[2];
        float half[3];
        for (int k = 0; k < 3; ++k) { // largest distance from the (rounded) centre along the axis, rounded up; NaN if not finite
            const double hk = std::max(hi3[k] - (double)cen[k], (double)cen[k] - lo3[k]) * (1.0 + 1e-6);
            half[k] = finite && nv ? std::nextafter((float)hk, 3.4e38f) : NAN;
        }
        cb.hx = half[0]; cb.hy = half[1]; cb.hz = half[2];
        cb.off16 = (uint32_t)(base / 16);
        cb.bytes = bytes;
        out.bounds.push_back(cb);
        std::memcpy(p + SFH_CH_HDR_BYTES, cx.data(), 4 * nv);
        std::memcpy(p + SFH_CH_HDR_BYTES + 4 * nv4, cy.data(), 4 * nv);
        std::memcpy(p + SFH_CH_HDR_BYTES + 8 * nv4, cz.data(), 4 * nv);
        std::memcpy(p + SFH_CH_HDR_BYTES + 12 * nv4, cf.data(), 4 * nf);
        if (prev_hdr != (size_t)-1) std::memcpy(out.chunks.data() + prev_hdr + 8, &bytes, 4); // previous.next_bytes
        prev_hdr = base;
        out.n_chunk_verts += nv;
        out.orig_face.resize((size_t)(chunk_id + 1) * SFH_CH_FACES, 0xFFFFFFFFu);
        ++chunk_id;
        nv = nf = 0;
    };

    std::vector<uint32_t> slot_face(SFH_CH_FACES);
    for (uint32_t i = 0; i < n_faces; ++i) {
        const uint32_t f = order[i];
        const uint32_t* id = faces + 3 * (size_t)f;
        uint32_t fresh = 0;
        for (int k = 0; k < 3; ++k) {
            bool seen = stamp[id[k]] == chunk_id;
            for (int j = 0; j < k && !seen; ++j) seen = id[j] == id[k];
            if (!seen) ++fresh;
        }
        if (nf == SFH_CH_FACES || nv + fresh > SFH_CH_VERTS) {
            const uint32_t done = nf;
            end_chunk();
            for (uint32_t s = 0; s < done; ++s) out.orig_face[(size_t)(chunk_id - 1) * SFH_CH_FACES + s] = slot_face[s];
        }
        uint32_t packed = 0;
        for (int k = 0; k < 3; ++k) {
            const uint32_t v = id[k];
            if (stamp[v] != chunk_id) {
                stamp[v] = chunk_id;
                local[v] = (uint16_t)nv;
                const float* p = vptr(v);
                cx[nv] = p[0]; cy[nv] = p[1]; cz[nv] = p[2];
                ++nv;
            }
            packed |= (uint32_t)local[v] << (8 * k);
        }
        cf[nf] = packed;
        slot_face[nf] = f;
        ++nf;
    }
    if (nf > 0) {
        const uint32_t done = nf;
        end_chunk();
        for (uint32_t s = 0; s < done; ++s) out.orig_face[(size_t)(chunk_id - 1) * SFH_CH_FACES + s] = slot_face[s];
    }
    out.chunk_off.push_back((uint32_t)(out.chunks.size() / 16));
    out.n_chunks = chunk_id;
    out.device_bytes = out.chunks.size();
    return true;
}

} // namespace sfh
