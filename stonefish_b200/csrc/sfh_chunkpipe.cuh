// sfh_chunkpipe.cuh — double-buffered staging of mesh chunks in shared memory with bulk async copies
// (cp.async.bulk = the 1-D TMA path, SASS UBLKCP) completing on mbarriers.
//
// One elected thread per stream issues ONE copy per chunk (<= 4128 bytes, 16-byte granular — the chunk format of
// mesh_chunker.hpp is laid out for exactly this) one whole chunk ahead of its use, so no thread ever waits on a
// global load inside the per-face loop and the copy costs no registers and no scoreboard slots.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "mesh_chunker.hpp"

namespace sfh {

struct ChunkPipe { // in shared memory
    alignas(16) uint8_t buf[2][SFH_CH_MAX_BYTES + 16]; // stride SFH_PIPE_BUF_STRIDE
    alignas(8) unsigned long long bar[2];
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define SFH_PIPE_BUF_STRIDE (SFH_CH_MAX_BYTES + 16u)

struct PipeAddr { // shared-window addresses, computed once per kernel and kept in registers
    uint32_t bar0, buf0;
};

// call from all threads once; contains a __syncthreads()
__device__ __forceinline__ PipeAddr pipe_init(ChunkPipe& p)
{
    PipeAddr a;
    a.bar0 = smem_addr(&p.bar[0]);
    a.buf0 = smem_addr(&p.buf[0][0]);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a.bar0));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a.bar0 + 8u));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    return a;
}

// one thread: start the copy of `bytes` (multiple of 16) from global `src` (16-byte aligned) into buffer s
__device__ __forceinline__ void pipe_issue(const PipeAddr& a, int s, const uint8_t* src, uint32_t bytes)
{
    const uint32_t bar = a.bar0 + 8u * (uint32_t)s, dst = a.buf0 + SFH_PIPE_BUF_STRIDE * (uint32_t)s;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}

// all threads: wait until the copy into buffer s with the given phase parity has landed
__device__ __forceinline__ void pipe_wait(const PipeAddr& a, int s, uint32_t parity)
{
    const uint32_t bar = a.bar0 + 8u * (uint32_t)s;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}

struct ChunkView { // where things are inside a staged chunk
    uint32_t nv, nf, next_bytes;
    float4 sphere; // bounding sphere of the vertices, mesh frame
    const float *x, *y, *z;
    const uint32_t* faces;
};

__device__ __forceinline__ ChunkView chunk_view(const uint8_t* buf)
{
    const uint4 h = *reinterpret_cast<const uint4*>(buf);
    const uint32_t nv4 = (h.x + 3u) & ~3u;
    ChunkView v;
    v.nv = h.x; v.nf = h.y; v.next_bytes = h.z;
    v.sphere = *reinterpret_cast<const float4*>(buf + 16);
    v.x = reinterpret_cast<const float*>(buf + SFH_CH_HDR_BYTES);
    v.y = v.x + nv4;
    v.z = v.y + nv4;
    v.faces = reinterpret_cast<const uint32_t*>(v.z + nv4);
    return v;
}

} // namespace sfh
