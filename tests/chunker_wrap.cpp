// Test shim: exposes the upload-time mesh chunker (stonefish_b200/csrc/mesh_chunker.hpp, host-only code) to ctypes.
#include "../stonefish_b200/csrc/mesh_chunker.hpp"
extern "C" {
void* chk_build(const float* pos, size_t stride, uint32_t nv, const uint32_t* faces, uint32_t nf, int keep_order)
{
    auto* cm = new sfh::ChunkedMesh();
    if (!sfh::chunk_mesh(pos, stride, nv, faces, nf, keep_order != 0, *cm)) { delete cm; return nullptr; }
    return cm;
}
void chk_free(void* h) { delete static_cast<sfh::ChunkedMesh*>(h); }
uint32_t chk_n_chunks(void* h) { return static_cast<sfh::ChunkedMesh*>(h)->n_chunks; }
uint64_t chk_bytes(void* h) { return static_cast<sfh::ChunkedMesh*>(h)->chunks.size(); }
const uint8_t* chk_data(void* h) { return static_cast<sfh::ChunkedMesh*>(h)->chunks.data(); }
const uint32_t* chk_off(void* h) { return static_cast<sfh::ChunkedMesh*>(h)->chunk_off.data(); }
const uint32_t* chk_orig(void* h) { return static_cast<sfh::ChunkedMesh*>(h)->orig_face.data(); }
uint32_t chk_chunk_verts(void* h) { return static_cast<sfh::ChunkedMesh*>(h)->n_chunk_verts; }
}
