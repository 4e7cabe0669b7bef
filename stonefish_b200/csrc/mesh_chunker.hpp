// mesh_chunker.hpp — upload-time re-layout of an indexed triangle mesh into device "chunks".
//
// Input  : what the reference's Mesh exposes (OpenGLDataStructs.h:120-262): strided vertex positions and
//          3 x u32 vertex ids per face.
// Output : fixed-stride chunks, each holding <= SFH_CH_FACES faces and the <= SFH_CH_VERTS distinct
//          vertices they reference:
//              +0     u32 n_verts, u32 n_faces, u32 first_slot, u32 pad
//              +16    float x[SFH_CH_VERTS]      (only the first n_verts are ever read)
//              +1040  float y[SFH_CH_VERTS]
//              +2064  float z[SFH_CH_VERTS]
//              +3088  u32   face[SFH_CH_FACES]   local ids packed i0 | i1<<8 | i2<<16
//          so a thread block transforms every vertex of a chunk ONCE into shared memory (the reference
//          transforms each vertex once per incident face, SolidEntity.cpp:1343-1345 — same bits either way)
//          and faces read 4 bytes of indices instead of 12.
// Faces are Morton-ordered by centroid first (unless SFH_MESH_KEEP_ORDER) so that consecutive faces share
// vertices; the per-body sums are order-independent within the fp64 contract and the class counts exactly.
// `orig_face[slot]` maps a chunk slot back to the caller's face index (debug codes only).
#pragma once

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#define SFH_CH_FACES 256u
#define SFH_CH_VERTS 256u
#define SFH_CH_HDR_BYTES 16u
#define SFH_CH_X_OFF SFH_CH_HDR_BYTES
#define SFH_CH_Y_OFF (SFH_CH_X_OFF + 4u * SFH_CH_VERTS)
#define SFH_CH_Z_OFF (SFH_CH_Y_OFF + 4u * SFH_CH_VERTS)
#define SFH_CH_F_OFF (SFH_CH_Z_OFF + 4u * SFH_CH_VERTS)
#define SFH_CH_BYTES (SFH_CH_F_OFF + 4u * SFH_CH_FACES) /* 4112 = 257 * 16 */

namespace sfh {

struct ChunkedMesh {
    std::vector<uint8_t> chunks;     // n_chunks * SFH_CH_BYTES
    std::vector<uint32_t> orig_face; // n_chunks * SFH_CH_FACES, 0xFFFFFFFF for unused slots
    uint32_t n_chunks = 0, n_faces = 0, n_verts = 0, n_chunk_verts = 0;
    uint64_t device_bytes = 0; // bytes one evaluation reads: headers + 12 B per stored vertex + 4 B per face
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

    uint32_t chunk_id = 0, nv = 0, nf = 0;
    auto begin_chunk = [&]() {
        out.chunks.resize((size_t)(chunk_id + 1) * SFH_CH_BYTES, 0);
        out.orig_face.resize((size_t)(chunk_id + 1) * SFH_CH_FACES, 0xFFFFFFFFu);
        nv = 0;
        nf = 0;
    };
    auto end_chunk = [&]() {
        uint8_t* base = out.chunks.data() + (size_t)chunk_id * SFH_CH_BYTES;
        uint32_t hdr[4] = {nv, nf, chunk_id * SFH_CH_FACES, 0};
        std::memcpy(base, hdr, sizeof(hdr));
        out.n_chunk_verts += nv;
        out.device_bytes += SFH_CH_HDR_BYTES + 12ull * nv + 4ull * nf;
        ++chunk_id;
    };

    if (n_faces > 0) begin_chunk();
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
            end_chunk();
            begin_chunk();
        }
        uint8_t* base = out.chunks.data() + (size_t)chunk_id * SFH_CH_BYTES;
        uint32_t packed = 0;
        for (int k = 0; k < 3; ++k) {
            const uint32_t v = id[k];
            if (stamp[v] != chunk_id) {
                stamp[v] = chunk_id;
                local[v] = (uint16_t)nv;
                const float* p = vptr(v);
                std::memcpy(base + SFH_CH_X_OFF + 4 * nv, &p[0], 4);
                std::memcpy(base + SFH_CH_Y_OFF + 4 * nv, &p[1], 4);
                std::memcpy(base + SFH_CH_Z_OFF + 4 * nv, &p[2], 4);
                ++nv;
            }
            packed |= (uint32_t)local[v] << (8 * k);
        }
        std::memcpy(base + SFH_CH_F_OFF + 4 * nf, &packed, 4);
        out.orig_face[(size_t)chunk_id * SFH_CH_FACES + nf] = f;
        ++nf;
    }
    if (n_faces > 0) end_chunk();
    out.n_chunks = chunk_id;
    return true;
}

} // namespace sfh
