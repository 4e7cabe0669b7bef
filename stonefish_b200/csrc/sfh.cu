// sfh.cu — C ABI (include/sfh.h) over the sm_100a kernels in sfh_kernels.cuh.
// One sfh_ctx = one GPU: stream, pinned staging, device arenas, mesh table. No CPU fallback anywhere:
// every entry point either enqueues CUDA work or returns an error.
#include "../../include/sfh.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "sfh_kernels.cuh"
#include "sfh_selftest.cuh"
#include "sfh_wavegen.cuh"
#include "sfh_cable.cuh"

using namespace sfh;

namespace {

thread_local std::string g_create_error = "";

struct Slab {
    uint8_t* base = nullptr;
    size_t size = 0, used = 0;
};

struct HostMesh {
    MeshDev dev{};
    sfh_mesh_stats stats{};
    bool live = false;
};

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0; // elements
};

struct PinBuf {
    void* p = nullptr;
    size_t cap = 0; // bytes
};

} // namespace

struct sfh_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    int num_sms = 0;
    std::string err;
    std::mutex err_mutex;  // uploads may fail concurrently
    std::mutex mesh_mutex;

    // static geometry
    std::vector<Slab> slabs;
    std::vector<HostMesh> meshes;
    DevBuf<MeshDev> d_meshes;
    bool mesh_table_dirty = true;
    PinBuf h_mesh_table;

    // waves
    DevBuf<float> d_wave_rgba, d_wave_ch0, d_wave_ch1;
    DevBuf<uint32_t> d_wave_hmax;
    PinBuf h_wave;
    WavesDev waves{nullptr, nullptr, 0, 0.f, 0.f, nullptr};
    bool have_waves = false;

    // cables (row N4)
    DevBuf<sfh_cable> d_cables;
    DevBuf<double> d_node_pos, d_node_vel;
    DevBuf<sfh_node_forces> d_node_out;

    // wave-field producer (row N3)
    DevBuf<float4> d_wg_s12, d_wg_s34, d_wg_tmp, d_wg_fft;
    DevBuf<int2> d_wg_bf_src;
    DevBuf<float2> d_wg_bf_w;
    std::vector<float> wg_s12, wg_s34; // fp32 spectrum as generated (before half storage)
    sfh_wave_params wg_params{};
    bool wg_ready = false;

    // per-step buffers
    DevBuf<sfh_body_state> d_states;
    DevBuf<sfh_info> d_info;
    DevBuf<BodyDev> d_bodies;
    DevBuf<uint32_t> d_off_sub, d_off_surf, d_status;
    DevBuf<PartialSub> d_part_sub;
    DevBuf<PartialSurf> d_part_wet, d_part_clip;
    // the surface path's chunk lists (sfh_surface.cuh)
    DevBuf<uint4> d_items;
    DevBuf<uint32_t> d_before, d_wet_off, d_clip_off, d_surf_ctrl;
    DevBuf<uint2> d_tile_cnt;
    DevBuf<PartialSurf> d_slice_part; // second stage of the epilogue's reduction (few bodies)
    DevBuf<uint32_t> d_slice_ticket;
    uint32_t max_mesh_chunks = 0; // largest n_chunks of any uploaded mesh: bounds the work list of a device-resident step
    DevBuf<OceanDev> d_ocean;
    DevBuf<uint8_t> d_codes;
    DevBuf<uint64_t> d_code_base;
    PinBuf h_ocean, h_states, h_info, h_status, h_codes, h_code_base;
    // ocean staging of the asynchronous (non-graph) paths: a ring of pinned slots guarded by events, so that a queued copy is
    // never overwritten, and the bytes staged last, so that an unchanged ocean costs no copy at all
    PinBuf h_ocean_ring;
    cudaEvent_t ocean_ev[4] = {nullptr, nullptr, nullptr, nullptr};
    int ocean_slot = 0;
    std::vector<uint8_t> ocean_last;

    // pending host-side completion of the last async step
    sfh_info* pending_out = nullptr;
    size_t pending_out_bytes = 0;
    uint8_t* pending_codes = nullptr;
    size_t pending_codes_bytes = 0;
    bool step_in_flight = false;
    bool device_step_pending = false; // sfh_step_device / sfh_aero_step_device since the last sfh_sync: its status word is still on the device

    // occupancy cache: blocks per SM for [kernel variant]
    int occ_sub[2] = {0, 0};
    int occ_aero[2] = {0, 0};
    int occ_surf[2][2] = {{0, 0}, {0, 0}};

    // CUDA-graph replay of small steps (launch-latency bound): see step_host
    struct GraphKey {
        uint32_t n, n_currents, n_meshes, flags; // flags: general | waves<<1 | may_sub<<2 | may_surf<<3
        uint64_t chunks_ub, surf_ub;
        const void *d_states, *d_info, *h_states, *h_info, *h_ocean, *d_ocean, *d_meshes, *d_bodies, *d_part_sub, *d_part_wet, *wave_ch0;
        int32_t wave_n;
        float wave_gx, wave_gy;
    };
    struct GraphEntry {
        GraphKey key;
        cudaGraphExec_t exec;
    };
    std::vector<GraphEntry> graphs;
    // zero-copy results of small steps: the host spins on a sequence number instead of synchronising the stream
    uint32_t zc_seq = 0;
    bool zc_pending = false;
    GraphKey last_key{};
    bool last_key_valid = false;
    uint64_t graph_launches = 0;

    // profiling
    bool profiling = false;
    bool profile_roles = false; // SFH_PROFILE_ROLES=1: while profiling, time the wet and the clip role in launches of their own
    cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}; // [6], [7]: after k_classify, after the wet role of k_surface_lists
    bool ev_sub = false, ev_surf = false, ev_valid = false;
    // small mixed steps: k_submerged runs beside the surface path on a stream of its own (fork / join by events)
    cudaStream_t side_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    uint32_t last_launches = 0;
    uint64_t last_faces = 0;
};

namespace {

int fail(sfh_ctx* ctx, int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) { std::lock_guard<std::mutex> lock(ctx->err_mutex); ctx->err = buf; } else g_create_error = buf;
    return code;
}

#define CU(ctx, call)                                                                                         \
    do {                                                                                                      \
        cudaError_t e__ = (call);                                                                             \
        if (e__ != cudaSuccess)                                                                               \
            return fail(ctx, e__ == cudaErrorMemoryAllocation ? SFH_ERR_OOM : SFH_ERR_CUDA, "%s: %s (%s:%d)", #call, \
                        cudaGetErrorString(e__), __FILE__, __LINE__);                                         \
    } while (0)

// Captured graphs bake in the addresses of every per-step buffer (states, results, work lists, partial records, status words,
// the ocean record, the wave planes): whenever one of them is reallocated every cached graph is dropped, whatever its key says
// (a key that lists "the pointers that matter" is one forgotten pointer away from replaying freed memory).
void invalidate_graphs(sfh_ctx* ctx)
{
    for (auto& g : ctx->graphs) cudaGraphExecDestroy(g.exec); // deferred by the runtime until in-flight launches have finished
    ctx->graphs.clear();
    ctx->last_key_valid = false;
}

template <typename T>
int ensure_dev(sfh_ctx* ctx, DevBuf<T>& b, size_t n)
{
    if (n <= b.cap) return SFH_OK;
    invalidate_graphs(ctx);
    size_t cap = b.cap ? b.cap : 1;
    while (cap < n) cap *= 2;
    if (b.p) CU(ctx, cudaFree(b.p)); // stream-ordered users finished: callers sync before growing (see prepare)
    b.p = nullptr; b.cap = 0;
    CU(ctx, cudaMalloc(&b.p, cap * sizeof(T)));
    b.cap = cap;
    return SFH_OK;
}

int ensure_pin(sfh_ctx* ctx, PinBuf& b, size_t bytes)
{
    if (bytes <= b.cap) return SFH_OK;
    invalidate_graphs(ctx);
    size_t cap = b.cap ? b.cap : 256;
    while (cap < bytes) cap *= 2;
    if (b.p) CU(ctx, cudaFreeHost(b.p));
    b.p = nullptr; b.cap = 0;
    CU(ctx, cudaHostAlloc(&b.p, cap, cudaHostAllocDefault));
    std::memset(b.p, 0, cap); // flags and sequence numbers polled by the host start from a known value
    b.cap = cap;
    return SFH_OK;
}

bool is_pinned_host(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

int arena_alloc(sfh_ctx* ctx, size_t bytes, uint8_t** out)
{
    bytes = (bytes + 255) & ~(size_t)255;
    for (auto& s : ctx->slabs)
        if (s.size - s.used >= bytes) { *out = s.base + s.used; s.used += bytes; return SFH_OK; }
    Slab s;
    s.size = bytes > ((size_t)256 << 20) ? bytes : ((size_t)256 << 20);
    cudaError_t e = cudaMalloc(&s.base, s.size);
    if (e != cudaSuccess && s.size > bytes) { cudaGetLastError(); s.size = bytes; e = cudaMalloc(&s.base, s.size); }
    if (e != cudaSuccess) return fail(ctx, SFH_ERR_OOM, "cudaMalloc(%zu) for mesh arena: %s", s.size, cudaGetErrorString(e));
    s.used = bytes;
    *out = s.base;
    ctx->slabs.push_back(s);
    return SFH_OK;
}

template <typename K>
int occupancy(sfh_ctx* ctx, K kernel, int* cache, int threads = SFH_SUB_THREADS)
{
    if (*cache == 0) {
        int n = 0;
        CU(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, 0));
        *cache = n > 0 ? n : 1;
    }
    return SFH_OK;
}

#define SFH_GRAPH_MAX_BODIES 1024u
#define SFH_ZEROCOPY_MAX_BODIES 64u /* results of steps this small are written straight into pinned host memory */

// Write the ocean record into the pinned staging buffer (no CUDA call after the first) and report whether some
// enabled current needs the per-face fp64 evaluation.
int fill_ocean_into(sfh_ctx* ctx, const sfh_ocean* o, OceanDev* h, bool* cur_general)
{
    if (o->n_currents > SFH_MAX_CURRENTS) return fail(ctx, SFH_ERR_INVALID, "n_currents %u > %u", o->n_currents, SFH_MAX_CURRENTS);
    if (o->has_waves && !ctx->have_waves) return fail(ctx, SFH_ERR_STATE, "ocean.has_waves set but sfh_waves_upload was never called");
    h->density = o->density;
    h->gravity_z = o->gravity_z;
    h->has_waves = o->has_waves ? 1u : 0u;
    h->currents_enabled = o->currents_enabled ? 1u : 0u;
    h->n_currents = o->n_currents;
    bool general = false;
    for (uint32_t i = 0; i < o->n_currents; ++i) {
        h->currents[i] = o->currents[i];
        if (o->currents[i].type > SFH_CURRENT_PIPE) return fail(ctx, SFH_ERR_INVALID, "current %u: unknown type %u", i, o->currents[i].type);
        if (o->currents_enabled && o->currents[i].enabled && o->currents[i].type != SFH_CURRENT_UNIFORM) general = true;
    }
    h->cur_general = general ? 1u : 0u;
    *cur_general = general;
    return SFH_OK;
}
// the graph path's fixed staging buffer (a replayed copy node reads it; steps on that path are synchronised one by one)
int fill_ocean(sfh_ctx* ctx, const sfh_ocean* o, bool* cur_general)
{
    int rc;
    if ((rc = ensure_pin(ctx, ctx->h_ocean, sizeof(OceanDev))) != SFH_OK) return rc;
    if ((rc = ensure_dev(ctx, ctx->d_ocean, 1)) != SFH_OK) return rc;
    ctx->ocean_last.clear(); // the device copy is about to change behind stage_ocean's back
    return fill_ocean_into(ctx, o, static_cast<OceanDev*>(ctx->h_ocean.p), cur_general);
}

// Ocean record to the device, stream-ordered. `fixed`: through the graph path's staging buffer, always copied (the call is
// being captured or precedes a capture). Otherwise: nothing if the bytes equal what the device already holds, else through
// the next slot of a pinned ring whose previous copy has completed.
int stage_ocean(sfh_ctx* ctx, const sfh_ocean* o, bool* cur_general, bool fixed = false)
{
    int rc;
    const size_t bytes = offsetof(OceanDev, currents) + sizeof(sfh_current) * (o->n_currents && o->n_currents <= SFH_MAX_CURRENTS ? o->n_currents : 1);
    if (fixed) {
        if ((rc = fill_ocean(ctx, o, cur_general)) != SFH_OK) return rc;
        CU(ctx, cudaMemcpyAsync(ctx->d_ocean.p, ctx->h_ocean.p, bytes, cudaMemcpyHostToDevice, ctx->stream));
        return SFH_OK;
    }
    OceanDev tmp;
    std::memset(&tmp, 0, bytes);
    if ((rc = fill_ocean_into(ctx, o, &tmp, cur_general)) != SFH_OK) return rc;
    if (ctx->ocean_last.size() == bytes && std::memcmp(ctx->ocean_last.data(), &tmp, bytes) == 0) return SFH_OK;
    const size_t stride = (sizeof(OceanDev) + 255) & ~(size_t)255;
    if ((rc = ensure_pin(ctx, ctx->h_ocean_ring, 4 * stride)) != SFH_OK) return rc;
    if ((rc = ensure_dev(ctx, ctx->d_ocean, 1)) != SFH_OK) return rc;
    const int slot = ctx->ocean_slot = (ctx->ocean_slot + 1) & 3;
    if (!ctx->ocean_ev[slot]) CU(ctx, cudaEventCreateWithFlags(&ctx->ocean_ev[slot], cudaEventDisableTiming));
    else CU(ctx, cudaEventSynchronize(ctx->ocean_ev[slot]));
    uint8_t* h = static_cast<uint8_t*>(ctx->h_ocean_ring.p) + (size_t)slot * stride;
    std::memcpy(h, &tmp, bytes);
    CU(ctx, cudaMemcpyAsync(ctx->d_ocean.p, h, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaEventRecord(ctx->ocean_ev[slot], ctx->stream));
    ctx->ocean_last.assign(reinterpret_cast<const uint8_t*>(&tmp), reinterpret_cast<const uint8_t*>(&tmp) + bytes);
    return SFH_OK;
}

int sync_mesh_table(sfh_ctx* ctx)
{
    std::lock_guard<std::mutex> lock(ctx->mesh_mutex);
    if (!ctx->mesh_table_dirty) return SFH_OK;
    const size_t n = ctx->meshes.size() ? ctx->meshes.size() : 1;
    int rc;
    if (n > ctx->d_meshes.cap) CU(ctx, cudaStreamSynchronize(ctx->stream));
    if ((rc = ensure_dev(ctx, ctx->d_meshes, n)) != SFH_OK) return rc;
    if (n * sizeof(MeshDev) > ctx->h_mesh_table.cap) CU(ctx, cudaStreamSynchronize(ctx->stream));
    if ((rc = ensure_pin(ctx, ctx->h_mesh_table, n * sizeof(MeshDev))) != SFH_OK) return rc;
    MeshDev* h = static_cast<MeshDev*>(ctx->h_mesh_table.p);
    CU(ctx, cudaStreamSynchronize(ctx->stream)); // previous async copy from this staging buffer is done
    // a released id keeps its slot with n_faces = SFH_MESH_TOMBSTONE: k_prologue flags bodies that still refer to it
    for (size_t i = 0; i < ctx->meshes.size(); ++i) h[i] = ctx->meshes[i].live ? ctx->meshes[i].dev : MeshDev{nullptr, nullptr, nullptr, nullptr, 0, SFH_MESH_TOMBSTONE};
    CU(ctx, cudaMemcpyAsync(ctx->d_meshes.p, h, ctx->meshes.size() * sizeof(MeshDev), cudaMemcpyHostToDevice, ctx->stream));
    ctx->mesh_table_dirty = false;
    return SFH_OK;
}

struct StepPlan {
    bool host_states = true;     // states/out are host pointers
    bool debug = false;
    bool aero = false;           // sfh_aero_step: the "ocean" is the atmosphere and every body runs the aero face loop
    bool graph_path = false;     // the call is (or may next time be) captured: fixed staging buffers, every copy issued
    bool zero_copy_out = false;  // d_out is pinned HOST memory; the epilogue publishes status + sequence number to h_status
    bool may_sub = true, may_surf = true; // host knows no body can reach the kernel -> skip the launch
    uint64_t chunks_ub = 0;      // upper bound of work items (host variant), 0 = unknown
    uint64_t surf_ub = 0;        // ... of the surface path's items (bodies that may cross the surface); 0 = unknown
    uint64_t faces = 0;
};

// A kernel launch that may overlap the tail of its predecessor in the stream (programmatic dependent launch, see pdl_wait() in
// sfh_device.cuh); pdl = false: an ordinary launch.
template <typename... KArgs, typename... Args>
void launch_k(void (*kernel)(KArgs...), unsigned grid, unsigned block, cudaStream_t s, bool pdl, Args... params)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid, 1, 1); cfg.blockDim = dim3(block, 1, 1); cfg.dynamicSmemBytes = 0; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1u : 0u;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(params)...); // errors are picked up by cudaGetLastError at the end of run_step
}
template <int CUR, bool DEBUG, bool AERO = false>
void launch_sub(const StepParams& P, int grid, cudaStream_t s, bool pdl) { launch_k(k_submerged<CUR, DEBUG, AERO>, (unsigned)grid, SFH_SUB_THREADS, s, pdl, P); }
template <bool WAVES, int CUR, bool DEBUG>
void launch_lists(const SurfParams& P, int grid, cudaStream_t s, bool pdl) { launch_k(k_surface_lists<WAVES, CUR, DEBUG>, (unsigned)grid, SFH_SUB_THREADS, s, pdl, P); }
void launch_lists_variant(const SurfParams& P, int grid, cudaStream_t s, bool waves, bool general, bool debug, bool pdl)
{
    if (debug) {
        if (waves) { if (general) launch_lists<true, 1, true>(P, grid, s, pdl); else launch_lists<true, 0, true>(P, grid, s, pdl); }
        else       { if (general) launch_lists<false, 1, true>(P, grid, s, pdl); else launch_lists<false, 0, true>(P, grid, s, pdl); }
    } else {
        if (waves) { if (general) launch_lists<true, 1, false>(P, grid, s, pdl); else launch_lists<true, 0, false>(P, grid, s, pdl); }
        else       { if (general) launch_lists<false, 1, false>(P, grid, s, pdl); else launch_lists<false, 0, false>(P, grid, s, pdl); }
    }
}

#ifndef SFH_GRAPH_MAX_FACES
#define SFH_GRAPH_MAX_FACES 2500000ull
#endif
int run_step(sfh_ctx* ctx, const sfh_ocean* ocean, const sfh_body_state* d_states, uint32_t n, sfh_info* d_out, const StepPlan& plan)
{
    int rc;
    bool general = false;
    if ((rc = stage_ocean(ctx, ocean, &general, plan.graph_path)) != SFH_OK) return rc;
    if ((rc = sync_mesh_table(ctx)) != SFH_OK) return rc;
    const bool waves = ocean->has_waves != 0;
    const int gi = general ? 1 : 0, wi = waves ? 1 : 0;

    // persistent grids
    if (plan.aero) {
        if (general) rc = occupancy(ctx, k_submerged<1, false, true>, &ctx->occ_aero[1]);
        else         rc = occupancy(ctx, k_submerged<0, false, true>, &ctx->occ_aero[0]);
    } else {
        if (general) rc = occupancy(ctx, k_submerged<1, false>, &ctx->occ_sub[1]);
        else         rc = occupancy(ctx, k_submerged<0, false>, &ctx->occ_sub[0]);
    }
    if (rc != SFH_OK) return rc;
    int* occs = &ctx->occ_surf[wi][gi];
    if (waves) { if (general) rc = occupancy(ctx, k_surface_lists<true, 1, false>, occs); else rc = occupancy(ctx, k_surface_lists<true, 0, false>, occs); }
    else       { if (general) rc = occupancy(ctx, k_surface_lists<false, 1, false>, occs); else rc = occupancy(ctx, k_surface_lists<false, 0, false>, occs); }
    if (rc != SFH_OK) return rc;
    // grids count BLOCKS; each block runs SFH_SUB_WPB independent streams (warps)
    uint64_t g_sub = (uint64_t)ctx->num_sms * (uint64_t)(plan.aero ? ctx->occ_aero : ctx->occ_sub)[gi];
    uint64_t g_wet = (uint64_t)ctx->num_sms * (uint64_t)*occs, g_clip = g_wet; // the two roles of k_surface_lists
    if (plan.chunks_ub) {
        if (g_sub * SFH_SUB_WPB > plan.chunks_ub) g_sub = (plan.chunks_ub + SFH_SUB_WPB - 1) / SFH_SUB_WPB;
    }
    // the surface path's lists hold every chunk of every body that may cross the surface: known exactly for host-side states,
    // bounded by (bodies x largest mesh) for device-resident ones (k_classify flags a list that does not fit)
    uint64_t items_ub = plan.surf_ub ? plan.surf_ub : (uint64_t)n * (ctx->max_mesh_chunks ? ctx->max_mesh_chunks : 1u);
    if (items_ub > ((uint64_t)1 << 26)) items_ub = (uint64_t)1 << 26;
    const uint64_t items_cap = ((items_ub + SFH_TILE - 1) / SFH_TILE) * SFH_TILE, tiles_cap = items_cap / SFH_TILE;
    if (plan.surf_ub) { // a list entry is handed out as up to 4 (wet) / 2 (clip) sub-items when there are fewer entries than streams
        if (g_wet * SFH_SUB_WPB > 4 * plan.surf_ub) g_wet = (4 * plan.surf_ub + SFH_SUB_WPB - 1) / SFH_SUB_WPB;
        if (g_clip * SFH_SUB_WPB > 2 * plan.surf_ub) g_clip = (2 * plan.surf_ub + SFH_SUB_WPB - 1) / SFH_SUB_WPB;
    }
    uint64_t g_cls = tiles_cap < (uint64_t)ctx->num_sms * 8 ? tiles_cap : (uint64_t)ctx->num_sms * 8;
    if (g_cls == 0) g_cls = 1;
    // one partial record per (segment, body); k_classify sizes the segments: at most SFH_SEG_PER_STREAM per resident stream
    const uint64_t seg_wet_max = g_wet * SFH_SUB_WPB * SFH_SEG_PER_STREAM + 1, seg_clip_max = g_clip * SFH_SUB_WPB * SFH_SEG_PER_STREAM + 1;
    const bool surf = plan.may_surf && !plan.aero;

    // buffers (growing them requires the stream to be idle: older kernels may still read the old ones)
    bool grow = n > ctx->d_bodies.cap || n + 1 > ctx->d_off_sub.cap || g_sub * SFH_SUB_WPB + n > ctx->d_part_sub.cap;
    if (surf) grow = grow || seg_wet_max + n > ctx->d_part_wet.cap || seg_clip_max + n > ctx->d_part_clip.cap || items_cap > ctx->d_items.cap ||
                     tiles_cap + 1 > ctx->d_wet_off.cap;
    if (grow) CU(ctx, cudaStreamSynchronize(ctx->stream));
    if ((rc = ensure_dev(ctx, ctx->d_bodies, n)) != SFH_OK) return rc;
    if ((rc = ensure_dev(ctx, ctx->d_off_sub, n + 1)) != SFH_OK) return rc;
    if ((rc = ensure_dev(ctx, ctx->d_off_surf, n + 1)) != SFH_OK) return rc;
    if ((rc = ensure_dev(ctx, ctx->d_part_sub, g_sub * SFH_SUB_WPB + n)) != SFH_OK) return rc;
    if (surf) {
        if ((rc = ensure_dev(ctx, ctx->d_part_wet, seg_wet_max + n)) != SFH_OK) return rc;
        if ((rc = ensure_dev(ctx, ctx->d_part_clip, seg_clip_max + n)) != SFH_OK) return rc;
        if ((rc = ensure_dev(ctx, ctx->d_items, items_cap)) != SFH_OK) return rc;
        if ((rc = ensure_dev(ctx, ctx->d_before, items_cap)) != SFH_OK) return rc;
        if ((rc = ensure_dev(ctx, ctx->d_wet_off, tiles_cap + 1)) != SFH_OK) return rc;
        if ((rc = ensure_dev(ctx, ctx->d_clip_off, tiles_cap + 1)) != SFH_OK) return rc;
        if ((rc = ensure_dev(ctx, ctx->d_tile_cnt, tiles_cap)) != SFH_OK) return rc;
    }
    if (!ctx->d_status.p) { // [0] error flags, [1] epilogue blocks finished, [4] zero-copy sequence number
        if ((rc = ensure_dev(ctx, ctx->d_status, 8)) != SFH_OK) return rc;
        CU(ctx, cudaMemsetAsync(ctx->d_status.p, 0, 8 * sizeof(uint32_t), ctx->stream));
    }
    if (!ctx->d_surf_ctrl.p) { // control words of the surface path; SC_ITEMS = 0 until a step classifies something
        if ((rc = ensure_dev(ctx, ctx->d_surf_ctrl, SC_WORDS)) != SFH_OK) return rc;
        CU(ctx, cudaMemsetAsync(ctx->d_surf_ctrl.p, 0, SC_WORDS * sizeof(uint32_t), ctx->stream));
    }

    const bool prof = ctx->profiling;
    if (prof) CU(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));

    SurfLists L;
    L.items = ctx->d_items.p; L.before = ctx->d_before.p; L.wet_off = ctx->d_wet_off.p; L.clip_off = ctx->d_clip_off.p;
    L.tile_cnt = ctx->d_tile_cnt.p; L.ctrl = ctx->d_surf_ctrl.p; L.items_cap = (uint32_t)items_cap;
    ClassifyParams cp;
    cp.bodies = ctx->d_bodies.p; cp.off_surf = ctx->d_off_surf.p; cp.n_bodies = n; cp.ocean = ctx->d_ocean.p; cp.waves = ctx->waves;
    cp.L = L; cp.status = ctx->d_status.p; cp.wet_streams = (uint32_t)(g_wet * SFH_SUB_WPB); cp.clip_streams = (uint32_t)(g_clip * SFH_SUB_WPB);
    // Direct mode (sfh_surface.cuh): decided on the device from the number of surface items and T = the consumer's streams. Off
    // when the two roles are launched one by one to be timed (SFH_PROFILE_ROLES, tools/ablate_surface.py).
    const bool time_roles = prof && ctx->profile_roles;
    const char* no_direct = getenv("SFH_NO_DIRECT"); // tests: small scenes through the LIST path too
    // T = the warps that are RESIDENT at once (the grid is twice that: wet blocks, then clip blocks): with more sub-items than that a
    // second wave of blocks would pay the per-warp start-up (classification, descriptor and chunk round trips) a second time
    const uint32_t direct_streams = (time_roles || (no_direct && no_direct[0] == '1')) ? 0u : (uint32_t)((g_wet > g_clip ? g_wet : g_clip) * SFH_SUB_WPB);
    cp.direct_streams = direct_streams;
    // the host often knows that the step is certain to run in direct mode (exactly, with host-side states): k_classify is not even launched
    // (device-resident states: when even bodies x largest mesh fits)
    const bool surely_direct = direct_streams && items_ub <= 2 * (uint64_t)direct_streams; // direct_split(): up to two waves
    PrologueParams pp;
    pp.states = d_states; pp.meshes = ctx->d_meshes.p; pp.n_meshes = (uint32_t)ctx->meshes.size(); pp.n_bodies = n;
    pp.ocean = ctx->d_ocean.p; pp.waves = ctx->waves; pp.bodies = ctx->d_bodies.p;
    pp.off_sub = ctx->d_off_sub.p; pp.off_surf = ctx->d_off_surf.p; pp.status = ctx->d_status.p; pp.aero = plan.aero ? 1u : 0u;
    pp.no_sub = plan.may_sub ? 0u : 1u; pp.no_surf = (plan.may_surf || plan.aero) ? 0u : 1u;
    const char* no_pdl = getenv("SFH_NO_PDL");   // A/B measurements
    // Device-resident steps only (back-to-back launches: cfg 3 34.9 -> 30.8 us per step). The replayed graphs of host-side small steps
    // gain nothing from programmatic edges (host to host 50.3 vs 51.8 us for cfg 3, +-0.5 us elsewhere; measured).
    const bool pdl = !plan.host_states && !(no_pdl && no_pdl[0] == '1') && !prof;
    const bool pdl_chain = pdl; // k_prologue -> k_scan_items of a large step
    uint32_t launches;
    // One block does gates and scans for a small step (one launch less); from a few hundred bodies on, its single SM is the slower
    // way: 512 bodies 218.2 -> 213.1 us per step with k_prologue spread over the SMs, 1024 bodies 420.5 -> 405.4 us, 192 equal.
#ifndef SFH_PROLOGUE_SCAN_MAX
#define SFH_PROLOGUE_SCAN_MAX 192u
#endif
    if (n <= SFH_PROLOGUE_SCAN_MAX) {
        k_prologue_scan<<<1, 1024, 0, ctx->stream>>>(pp);
        launches = 1;
    } else {
        const unsigned pt = SFH_PROLOGUE_THREADS; // bodies per block: few, so that the records' round trips spread over the SMs
        k_prologue<<<(n + pt - 1) / pt, pt, 0, ctx->stream>>>(pp);
        launch_k(k_scan_items, 1u, 1024u, ctx->stream, pdl_chain, ctx->d_off_sub.p, ctx->d_off_surf.p, n);
        launches = 2;
    }
    if (prof) CU(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));

    StepParams sp;
    sp.bodies = ctx->d_bodies.p; sp.n_bodies = n; sp.ocean = ctx->d_ocean.p; sp.waves = ctx->waves;
    sp.codes = plan.debug ? ctx->d_codes.p : nullptr; sp.code_base = plan.debug ? ctx->d_code_base.p : nullptr;

    ctx->ev_sub = ctx->ev_surf = false;
    // A small step that has both kinds of bodies (the FluidDynamicsTest scene): neither face kernel fills the machine and both are
    // bound by latency, so k_submerged goes to a side stream and runs beside the surface path (two branches of the step's graph).
    // Small = the surface path is certain to take direct mode and the submerged chunks do not outnumber the resident warps either.
    const char* no_fork = getenv("SFH_NO_FORK");
    const bool fork = !(no_fork && no_fork[0] == '1') && !prof && plan.may_sub && g_sub > 0 && surf && g_wet > 0 && g_clip > 0 && surely_direct && plan.chunks_ub &&
                      plan.chunks_ub <= (uint64_t)direct_streams;
    cudaStream_t sub_stream = ctx->stream;
    if (fork) {
        CU(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
        CU(ctx, cudaStreamWaitEvent(ctx->side_stream, ctx->ev_fork, 0));
        sub_stream = ctx->side_stream;
    }
    if (plan.may_sub && g_sub > 0) {
        sp.item_off = ctx->d_off_sub.p; sp.partials = ctx->d_part_sub.p;
        if (plan.aero) {
            if (plan.debug) { if (general) launch_sub<1, true, true>(sp, (int)g_sub, sub_stream, pdl && !fork); else launch_sub<0, true, true>(sp, (int)g_sub, sub_stream, pdl && !fork); }
            else            { if (general) launch_sub<1, false, true>(sp, (int)g_sub, sub_stream, pdl && !fork); else launch_sub<0, false, true>(sp, (int)g_sub, sub_stream, pdl && !fork); }
        } else {
            if (plan.debug) { if (general) launch_sub<1, true>(sp, (int)g_sub, sub_stream, pdl && !fork); else launch_sub<0, true>(sp, (int)g_sub, sub_stream, pdl && !fork); }
            else            { if (general) launch_sub<1, false>(sp, (int)g_sub, sub_stream, pdl && !fork); else launch_sub<0, false>(sp, (int)g_sub, sub_stream, pdl && !fork); }
        }
        ++launches;
        ctx->ev_sub = true;
        if (fork) CU(ctx, cudaEventRecord(ctx->ev_join, ctx->side_stream));
    }
    if (prof) CU(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));

    if (surf && g_wet > 0 && g_clip > 0) {
        if (!surely_direct) {
            launch_k(k_classify, (unsigned)g_cls, SFH_TILE, ctx->stream, pdl && !(plan.may_sub && g_sub > 0 && !fork), cp);
            ++launches;
        }
        if (prof) CU(ctx, cudaEventRecord(ctx->ev[6], ctx->stream));
        SurfParams up;
        up.bodies = ctx->d_bodies.p; up.n_bodies = n; up.ocean = ctx->d_ocean.p; up.waves = ctx->waves; up.L = L; up.off_surf = ctx->d_off_surf.p;
        up.codes = sp.codes; up.code_base = sp.code_base;
        up.part_wet = ctx->d_part_wet.p; up.part_clip = ctx->d_part_clip.p; up.direct_streams = direct_streams;
        if (time_roles) { // one launch per role, so that each can be timed
            up.wet_blocks = (uint32_t)g_wet;
            launch_lists_variant(up, (int)g_wet, ctx->stream, waves, general, plan.debug, pdl);
            CU(ctx, cudaEventRecord(ctx->ev[7], ctx->stream));
            up.wet_blocks = 0;
            launch_lists_variant(up, (int)g_clip, ctx->stream, waves, general, plan.debug, pdl);
            launches += 2;
        } else {
            up.wet_blocks = (uint32_t)g_wet;
            launch_lists_variant(up, (int)(g_wet + g_clip), ctx->stream, waves, general, plan.debug, pdl);
            if (prof) CU(ctx, cudaEventRecord(ctx->ev[7], ctx->stream)); // wet_ms = the whole consumer, clip_ms = 0
            ++launches;
        }
        ctx->ev_surf = true;
    }
    if (fork) CU(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
    if (prof) CU(ctx, cudaEventRecord(ctx->ev[4], ctx->stream));

    EpilogueParams ep;
    ep.states = d_states; ep.bodies = ctx->d_bodies.p; ep.n_bodies = n; ep.ocean = ctx->d_ocean.p;
    ep.off_sub = ctx->d_off_sub.p; ep.off_surf = ctx->d_off_surf.p;
    ep.part_sub = ctx->d_part_sub.p; ep.part_wet = ctx->d_part_wet.p; ep.part_clip = ctx->d_part_clip.p;
    ep.grid_sub = (uint32_t)(g_sub * SFH_SUB_WPB); ep.L = L; ep.aero = plan.aero ? 1u : 0u; ep.out = d_out;
    ep.direct_streams = surf ? direct_streams : 0u;
    // slices of >= 256 records: a body has at most one record per stream (or sub-item) that touched it
    uint64_t rec_ub = (uint64_t)ctx->num_sms * 16 * SFH_SUB_WPB;
    if (plan.chunks_ub && 4 * plan.chunks_ub < rec_ub) rec_ub = 4 * plan.chunks_ub;
    // a block of 128 threads per body pays a block barrier and a four-way sum: only where a body can have more records than a warp
    // reads in two or three rounds
    ep.team = (n <= 256 && rec_ub / n > 96) ? 128u : 32u;
    ep.slices = 1;
    if (ep.team == 128u) {
        const uint64_t want = rec_ub / n / 256, room = 512u / n;
        ep.slices = (uint32_t)(want < 1 ? 1 : (want > 16 ? 16 : want));
        if (ep.slices > room) ep.slices = room ? (uint32_t)room : 1u;
    }
    if (ep.slices > 1) {
        const bool fresh = !ctx->d_slice_ticket.p || n > ctx->d_slice_ticket.cap;
        if (fresh || (size_t)n * ep.slices > ctx->d_slice_part.cap) CU(ctx, cudaStreamSynchronize(ctx->stream));
        if ((rc = ensure_dev(ctx, ctx->d_slice_part, (size_t)n * ep.slices)) != SFH_OK) return rc;
        if ((rc = ensure_dev(ctx, ctx->d_slice_ticket, n)) != SFH_OK) return rc;
        if (fresh) CU(ctx, cudaMemsetAsync(ctx->d_slice_ticket.p, 0, ctx->d_slice_ticket.cap * sizeof(uint32_t), ctx->stream)); // the kernel leaves zeros behind
    }
    ep.slice_part = ctx->d_slice_part.p; ep.slice_ticket = ctx->d_slice_ticket.p;
    ep.zc_host = nullptr; ep.zc_dev = ctx->d_status.p;
    if (plan.zero_copy_out) {
        void* dp = nullptr;
        CU(ctx, cudaHostGetDevicePointer(&dp, ctx->h_status.p, 0));
        ep.zc_host = static_cast<uint32_t*>(dp);
    }
    launch_k(k_epilogue, ep.team == 128u ? n * ep.slices : (n + 3) / 4, 128u, ctx->stream, pdl && !fork, ep);
    ++launches;
    if (prof) CU(ctx, cudaEventRecord(ctx->ev[5], ctx->stream));
    CU(ctx, cudaGetLastError());
    ctx->last_launches = launches;
    ctx->last_faces = plan.faces;
    return SFH_OK;
}

int step_host(sfh_ctx* ctx, const sfh_ocean* ocean, const sfh_body_state* states, uint32_t n, sfh_info* out, uint8_t* codes,
              size_t codes_bytes, bool debug, bool aero = false)
{
    if (!ctx) return SFH_ERR_INVALID;
    if (!ocean || (!states && n) || (!out && n)) return fail(ctx, SFH_ERR_INVALID, "null argument");
    if (ctx->step_in_flight) { int rc = sfh_sync(ctx); if (rc != SFH_OK) return rc; }
    CU(ctx, cudaSetDevice(ctx->device));
    if (n == 0) return SFH_OK;
    int rc;

    // host-side validation and work bound (the mesh table is host-visible)
    StepPlan plan;
    plan.debug = debug;
    plan.aero = aero;
    plan.may_sub = plan.may_surf = false;
    std::vector<uint64_t> code_base;
    if (debug) code_base.resize(n);
    uint64_t total_codes = 0;
    // Large steps from pinned buffers: the state copy is started BEFORE the host walks the records (validation and work bounds read
    // every 464-byte record once: tens of microseconds for thousands of bodies), so the two overlap. A record that then fails
    // validation only means the copy was for nothing.
    bool states_copied = false;
    if (n > SFH_GRAPH_MAX_BODIES && n <= ctx->d_states.cap && n <= ctx->d_info.cap && is_pinned_host(states)) {
        if (ctx->profiling) CU(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
        CU(ctx, cudaMemcpyAsync(ctx->d_states.p, states, (size_t)n * sizeof(sfh_body_state), cudaMemcpyHostToDevice, ctx->stream));
        states_copied = true;
    }
    {
        std::lock_guard<std::mutex> lock(ctx->mesh_mutex);
        for (uint32_t i = 0; i < n; ++i) {
            const sfh_body_state& s = states[i];
            if (s.mode > SFH_POS_AUTO) return fail(ctx, SFH_ERR_INVALID, "body %u: bad mode %u", i, s.mode);
            if (s.mesh_id >= ctx->meshes.size() || !ctx->meshes[s.mesh_id].live)
                return fail(ctx, SFH_ERR_INVALID, "body %u: unknown mesh id %u", i, s.mesh_id);
            const HostMesh& m = ctx->meshes[s.mesh_id];
            if (debug) { code_base[i] = total_codes; total_codes += m.stats.n_faces; }
            if (aero) { plan.chunks_ub += m.stats.n_chunks; plan.may_sub = true; plan.faces += m.stats.n_faces; continue; }
            if (s.mode == SFH_POS_OUTSIDE) continue;
            plan.chunks_ub += m.stats.n_chunks;
            if (s.mode != SFH_POS_CROSSING) plan.may_sub = true;
            if (s.mode != SFH_POS_INSIDE) { plan.may_surf = true; plan.surf_ub += m.stats.n_chunks; }
            plan.faces += m.stats.n_faces; // AUTO bodies that turn out OUTSIDE are subtracted by callers via info.n_faces
        }
    }
    if (plan.chunks_ub == 0) plan.chunks_ub = 1;
    if (debug && codes_bytes < total_codes) return fail(ctx, SFH_ERR_INVALID, "face_codes buffer too small: %zu < %llu", codes_bytes, (unsigned long long)total_codes);

    const bool grow = n > ctx->d_states.cap || n > ctx->d_info.cap;
    if (grow) CU(ctx, cudaStreamSynchronize(ctx->stream));
    if ((rc = ensure_dev(ctx, ctx->d_states, n)) != SFH_OK) return rc;
    if ((rc = ensure_dev(ctx, ctx->d_info, n)) != SFH_OK) return rc;

    const size_t sbytes = (size_t)n * sizeof(sfh_body_state);
    const size_t obytes = (size_t)n * sizeof(sfh_info);

    // Small steps are launch-latency bound (2 copies in, 5 kernels, 2 copies out): once the same step shape has been
    // seen twice in a row it is captured into a CUDA graph and replayed with one launch. Everything a graph node
    // reads is at a fixed address: states/results/ocean go through the context's pinned staging buffers.
    // From PINNED caller buffers a graph costs the two copies through the staging buffers (0.03 us per body), and a step with
    // millions of faces hides its launches behind its own kernels anyway: such steps are launched one by one, the copies straight
    // from / to the caller's buffers. Measured end to end, 50k-face hulls: 32 bodies 54.5 (graph) vs 57.5 us, 64 bodies 75.1 vs
    // 73.6 us, 128 bodies 104 vs 98.3 us, 512 bodies 266 vs 251.5 us, 1024 bodies 482 vs 450.5 us.
    const char* gmf = getenv("SFH_GRAPH_MAX_FACES"); // A/B measurements
    const uint64_t graph_max_faces = gmf ? strtoull(gmf, nullptr, 10) : SFH_GRAPH_MAX_FACES;
    bool graphable = !debug && !ctx->profiling && n <= SFH_GRAPH_MAX_BODIES;
    if (graphable && (plan.faces > graph_max_faces || n > 256u) && is_pinned_host(states) && is_pinned_host(out)) graphable = false;
    if (graphable) {
        if ((rc = ensure_pin(ctx, ctx->h_states, sbytes)) != SFH_OK) return rc;
        if ((rc = ensure_pin(ctx, ctx->h_info, obytes)) != SFH_OK) return rc;
        if ((rc = ensure_pin(ctx, ctx->h_status, 16)) != SFH_OK) return rc;
        if ((rc = sync_mesh_table(ctx)) != SFH_OK) return rc; // may synchronise: never inside a capture
        bool general = false;
        for (uint32_t i = 0; i < ocean->n_currents && i < SFH_MAX_CURRENTS; ++i)
            if (ocean->currents_enabled && ocean->currents[i].enabled && ocean->currents[i].type != SFH_CURRENT_UNIFORM) general = true;
        sfh_ctx::GraphKey key;
        std::memset(&key, 0, sizeof(key));
        key.n = n; key.n_currents = ocean->n_currents; key.n_meshes = (uint32_t)ctx->meshes.size();
        // Few bodies: the epilogue writes the records straight into the pinned result buffer and then a sequence number the host
        // spins on — no device-to-host copy nodes and no stream synchronisation on the critical path of a small step.
        const bool zc = n <= SFH_ZEROCOPY_MAX_BODIES;
        plan.zero_copy_out = zc;
        plan.graph_path = true;
        sfh_info* out_target = ctx->d_info.p;
        if (zc) {
            void* dp = nullptr;
            CU(ctx, cudaHostGetDevicePointer(&dp, ctx->h_info.p, 0));
            out_target = static_cast<sfh_info*>(dp);
        }
        key.flags = (general ? 1u : 0u) | (ocean->has_waves ? 2u : 0u) | (plan.may_sub ? 4u : 0u) | (plan.may_surf ? 8u : 0u) | (aero ? 16u : 0u) |
                    (zc ? 32u : 0u) | ((getenv("SFH_NO_DIRECT") && getenv("SFH_NO_DIRECT")[0] == '1') ? 64u : 0u);
        key.chunks_ub = plan.chunks_ub; key.surf_ub = plan.surf_ub;
        key.d_states = ctx->d_states.p; key.d_info = ctx->d_info.p; key.h_states = ctx->h_states.p; key.h_info = ctx->h_info.p;
        key.h_ocean = ctx->h_ocean.p; key.d_ocean = ctx->d_ocean.p; key.d_meshes = ctx->d_meshes.p; key.d_bodies = ctx->d_bodies.p;
        key.d_part_sub = ctx->d_part_sub.p; key.d_part_wet = ctx->d_part_wet.p;
        key.wave_ch0 = ctx->waves.ch0; key.wave_n = ctx->waves.N; key.wave_gx = ctx->waves.gx; key.wave_gy = ctx->waves.gy;

        std::memcpy(ctx->h_states.p, states, sbytes);
        cudaGraphExec_t exec = nullptr;
        for (auto& g : ctx->graphs)
            if (std::memcmp(&g.key, &key, sizeof(key)) == 0) exec = g.exec;
        const bool capture = !exec && ctx->last_key_valid && std::memcmp(&ctx->last_key, &key, sizeof(key)) == 0 && ctx->d_status.p;
        ctx->last_key = key;
        ctx->last_key_valid = true;
        if (exec) {
            if ((rc = fill_ocean(ctx, ocean, &general)) != SFH_OK) return rc;
            CU(ctx, cudaGraphLaunch(exec, ctx->stream));
            ++ctx->graph_launches;
        } else {
            if (capture) CU(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
            cudaError_t e = cudaMemcpyAsync(ctx->d_states.p, ctx->h_states.p, sbytes, cudaMemcpyHostToDevice, ctx->stream);
            rc = (e == cudaSuccess) ? run_step(ctx, ocean, ctx->d_states.p, n, out_target, plan) : fail(ctx, SFH_ERR_CUDA, "H2D: %s", cudaGetErrorString(e));
            if (rc == SFH_OK && !zc) {
                e = cudaMemcpyAsync(ctx->h_info.p, ctx->d_info.p, obytes, cudaMemcpyDeviceToHost, ctx->stream);
                if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->h_status.p, ctx->d_status.p, 4, cudaMemcpyDeviceToHost, ctx->stream);
                if (e != cudaSuccess) rc = fail(ctx, SFH_ERR_CUDA, "D2H: %s", cudaGetErrorString(e));
            }
            if (capture) {
                cudaGraph_t graph = nullptr;
                e = cudaStreamEndCapture(ctx->stream, &graph);
                if (rc != SFH_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
                if (e != cudaSuccess) return fail(ctx, SFH_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
                e = cudaGraphInstantiate(&exec, graph, 0);
                cudaGraphDestroy(graph);
                if (e != cudaSuccess) return fail(ctx, SFH_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
                if (ctx->graphs.size() >= 8) { cudaGraphExecDestroy(ctx->graphs.front().exec); ctx->graphs.erase(ctx->graphs.begin()); }
                ctx->graphs.push_back({key, exec});
                CU(ctx, cudaGraphLaunch(exec, ctx->stream));
                ++ctx->graph_launches;
            } else if (rc != SFH_OK) {
                return rc;
            }
        }
        ctx->pending_out = out;
        ctx->pending_out_bytes = obytes;
        ctx->pending_codes = nullptr;
        ctx->step_in_flight = true;
        ctx->ev_valid = false;
        if (zc) { ++ctx->zc_seq; ctx->zc_pending = true; }
        return SFH_OK;
    }

    if (!states_copied) {
        if (ctx->profiling) CU(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
        const void* src = states;
        if (!is_pinned_host(states)) {
            if ((rc = ensure_pin(ctx, ctx->h_states, sbytes)) != SFH_OK) return rc;
            std::memcpy(ctx->h_states.p, states, sbytes);
            src = ctx->h_states.p;
        }
        CU(ctx, cudaMemcpyAsync(ctx->d_states.p, src, sbytes, cudaMemcpyHostToDevice, ctx->stream));
    }

    if (debug) {
        if ((rc = ensure_dev(ctx, ctx->d_codes, total_codes ? total_codes : 1)) != SFH_OK) return rc;
        if ((rc = ensure_dev(ctx, ctx->d_code_base, n)) != SFH_OK) return rc;
        if ((rc = ensure_pin(ctx, ctx->h_code_base, n * sizeof(uint64_t))) != SFH_OK) return rc;
        std::memcpy(ctx->h_code_base.p, code_base.data(), n * sizeof(uint64_t));
        CU(ctx, cudaMemcpyAsync(ctx->d_code_base.p, ctx->h_code_base.p, n * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
        CU(ctx, cudaMemsetAsync(ctx->d_codes.p, 0, total_codes ? total_codes : 1, ctx->stream));
    }

    if ((rc = run_step(ctx, ocean, ctx->d_states.p, n, ctx->d_info.p, plan)) != SFH_OK) return rc;

    ctx->pending_out = nullptr;
    if (is_pinned_host(out)) {
        CU(ctx, cudaMemcpyAsync(out, ctx->d_info.p, obytes, cudaMemcpyDeviceToHost, ctx->stream));
    } else {
        if ((rc = ensure_pin(ctx, ctx->h_info, obytes)) != SFH_OK) return rc;
        CU(ctx, cudaMemcpyAsync(ctx->h_info.p, ctx->d_info.p, obytes, cudaMemcpyDeviceToHost, ctx->stream));
        ctx->pending_out = out;
        ctx->pending_out_bytes = obytes;
    }
    ctx->pending_codes = nullptr;
    if (debug && total_codes) {
        if ((rc = ensure_pin(ctx, ctx->h_codes, total_codes)) != SFH_OK) return rc;
        CU(ctx, cudaMemcpyAsync(ctx->h_codes.p, ctx->d_codes.p, total_codes, cudaMemcpyDeviceToHost, ctx->stream));
        ctx->pending_codes = codes;
        ctx->pending_codes_bytes = total_codes;
    }
    if ((rc = ensure_pin(ctx, ctx->h_status, 16)) != SFH_OK) return rc;
    CU(ctx, cudaMemcpyAsync(ctx->h_status.p, ctx->d_status.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->step_in_flight = true;
    ctx->ev_valid = ctx->profiling;
    return SFH_OK;
}

} // namespace

extern "C" {

int sfh_abi_version(void) { return SFH_ABI_VERSION; }

int sfh_create_on_stream(int cuda_device, void* cuda_stream, sfh_ctx** out)
{
    if (!out) return fail(nullptr, SFH_ERR_INVALID, "out == NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(nullptr, SFH_ERR_NO_DEVICE, "no CUDA device available (%s); this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (cuda_device < 0 || cuda_device >= count) return fail(nullptr, SFH_ERR_INVALID, "cuda_device %d out of range [0,%d)", cuda_device, count);
    sfh_ctx* ctx = new sfh_ctx();
    ctx->device = cuda_device;
    auto bail = [&](const char* what, cudaError_t err) {
        fail(nullptr, SFH_ERR_CUDA, "%s: %s", what, cudaGetErrorString(err));
        delete ctx;
        return SFH_ERR_CUDA;
    };
    if ((e = cudaSetDevice(cuda_device)) != cudaSuccess) return bail("cudaSetDevice", e);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, cuda_device)) != cudaSuccess) return bail("cudaGetDeviceProperties", e);
    if (prop.major < 10) {
        fail(nullptr, SFH_ERR_NO_DEVICE, "device %d is sm_%d%d; this library ships sm_100a code only", cuda_device, prop.major, prop.minor);
        delete ctx;
        return SFH_ERR_NO_DEVICE;
    }
    ctx->num_sms = prop.multiProcessorCount;
    if (cuda_stream) {
        ctx->stream = static_cast<cudaStream_t>(cuda_stream);
        ctx->own_stream = false;
    } else {
        if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    }
    for (auto& ev : ctx->ev)
        if ((e = cudaEventCreate(&ev)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    if ((e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    *out = ctx;
    return SFH_OK;
}

int sfh_create(int cuda_device, sfh_ctx** out) { return sfh_create_on_stream(cuda_device, nullptr, out); }

void sfh_destroy(sfh_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto& s : ctx->slabs) cudaFree(s.base);
    cudaFree(ctx->d_cables.p); cudaFree(ctx->d_node_pos.p); cudaFree(ctx->d_node_vel.p); cudaFree(ctx->d_node_out.p);
    cudaFree(ctx->d_wg_s12.p); cudaFree(ctx->d_wg_s34.p); cudaFree(ctx->d_wg_tmp.p); cudaFree(ctx->d_wg_fft.p);
    cudaFree(ctx->d_wg_bf_src.p); cudaFree(ctx->d_wg_bf_w.p);
    cudaFree(ctx->d_meshes.p); cudaFree(ctx->d_wave_rgba.p); cudaFree(ctx->d_wave_ch0.p); cudaFree(ctx->d_wave_ch1.p); cudaFree(ctx->d_wave_hmax.p);
    cudaFree(ctx->d_states.p); cudaFree(ctx->d_info.p); cudaFree(ctx->d_bodies.p); cudaFree(ctx->d_off_sub.p);
    cudaFree(ctx->d_off_surf.p); cudaFree(ctx->d_status.p); cudaFree(ctx->d_part_sub.p); cudaFree(ctx->d_part_wet.p); cudaFree(ctx->d_part_clip.p);
    cudaFree(ctx->d_items.p); cudaFree(ctx->d_before.p); cudaFree(ctx->d_wet_off.p); cudaFree(ctx->d_clip_off.p); cudaFree(ctx->d_surf_ctrl.p); cudaFree(ctx->d_tile_cnt.p); cudaFree(ctx->d_slice_part.p); cudaFree(ctx->d_slice_ticket.p);
    cudaFree(ctx->d_ocean.p); cudaFree(ctx->d_codes.p); cudaFree(ctx->d_code_base.p);
    for (cudaEvent_t e : ctx->ocean_ev)
        if (e) cudaEventDestroy(e);
    for (PinBuf* b : {&ctx->h_ocean_ring, &ctx->h_mesh_table, &ctx->h_wave, &ctx->h_ocean, &ctx->h_states, &ctx->h_info, &ctx->h_status, &ctx->h_codes, &ctx->h_code_base})
        if (b->p) cudaFreeHost(b->p);
    for (auto& g : ctx->graphs) cudaGraphExecDestroy(g.exec);
    for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->side_stream) { cudaStreamSynchronize(ctx->side_stream); cudaStreamDestroy(ctx->side_stream); }
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* sfh_last_error(const sfh_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int sfh_mesh_upload(sfh_ctx* ctx, const void* pos, size_t stride_bytes, uint32_t n_verts, const uint32_t* faces, uint32_t n_faces,
                    uint32_t flags, uint32_t* mesh_id_out)
{
    if (!ctx) return SFH_ERR_INVALID;
    if (!mesh_id_out || (n_verts && !pos) || (n_faces && !faces)) return fail(ctx, SFH_ERR_INVALID, "null argument");
    if (stride_bytes < 12 || (stride_bytes & 3)) return fail(ctx, SFH_ERR_INVALID, "vertex stride %zu must be >= 12 and a multiple of 4", stride_bytes);
    ChunkedMesh cm; // heavy part outside the lock: callers may upload from several threads
    if (!chunk_mesh(pos, stride_bytes, n_verts, faces, n_faces, (flags & SFH_MESH_KEEP_ORDER) != 0, cm))
        return fail(ctx, SFH_ERR_INVALID, "a face references a vertex id >= n_verts (%u)", n_verts);

    std::lock_guard<std::mutex> lock(ctx->mesh_mutex);
    CU(ctx, cudaSetDevice(ctx->device));
    HostMesh hm;
    hm.stats.n_faces = n_faces; hm.stats.n_verts = n_verts; hm.stats.n_chunks = cm.n_chunks; hm.stats.n_chunk_verts = cm.n_chunk_verts;
    hm.stats.device_bytes = cm.device_bytes;
    hm.stats.algorithmic_bytes = 12ull * n_faces + 12ull * n_verts;
    hm.dev.n_chunks = cm.n_chunks; hm.dev.n_faces = n_faces;
    if (cm.n_chunks) {
        uint8_t *d_chunks = nullptr, *d_off = nullptr, *d_orig = nullptr, *d_bounds = nullptr;
        int rc;
        if ((rc = arena_alloc(ctx, cm.chunks.size(), &d_chunks)) != SFH_OK) return rc;
        if ((rc = arena_alloc(ctx, cm.bounds.size() * sizeof(ChunkBounds), &d_bounds)) != SFH_OK) return rc;
        CU(ctx, cudaMemcpy(d_bounds, cm.bounds.data(), cm.bounds.size() * sizeof(ChunkBounds), cudaMemcpyHostToDevice));
        if ((rc = arena_alloc(ctx, cm.chunk_off.size() * sizeof(uint32_t), &d_off)) != SFH_OK) return rc;
        if ((rc = arena_alloc(ctx, cm.orig_face.size() * sizeof(uint32_t), &d_orig)) != SFH_OK) return rc;
        CU(ctx, cudaMemcpy(d_chunks, cm.chunks.data(), cm.chunks.size(), cudaMemcpyHostToDevice));
        CU(ctx, cudaMemcpy(d_off, cm.chunk_off.data(), cm.chunk_off.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        CU(ctx, cudaMemcpy(d_orig, cm.orig_face.data(), cm.orig_face.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        // A pageable cudaMemcpy may return once the data is staged, with the DMA still in flight on the legacy stream, and the
        // context's non-blocking stream is not ordered after it: wait for the legacy stream, so that a step enqueued right after
        // the upload reads the whole mesh.
        CU(ctx, cudaStreamSynchronize(cudaStreamLegacy));
        hm.dev.chunks = d_chunks;
        hm.dev.chunk_off = reinterpret_cast<const uint32_t*>(d_off);
        hm.dev.orig_face = reinterpret_cast<const uint32_t*>(d_orig);
        hm.dev.bounds = reinterpret_cast<const ChunkBounds*>(d_bounds);
        if (cm.n_chunks > ctx->max_mesh_chunks) ctx->max_mesh_chunks = cm.n_chunks;
    }
    hm.live = true;
    ctx->meshes.push_back(hm);
    ctx->mesh_table_dirty = true;
    *mesh_id_out = (uint32_t)(ctx->meshes.size() - 1);
    return SFH_OK;
}

int sfh_mesh_release(sfh_ctx* ctx, uint32_t mesh_id)
{
    if (!ctx) return SFH_ERR_INVALID;
    std::lock_guard<std::mutex> lock(ctx->mesh_mutex);
    if (mesh_id >= ctx->meshes.size() || !ctx->meshes[mesh_id].live) return fail(ctx, SFH_ERR_INVALID, "unknown mesh id %u", mesh_id);
    ctx->meshes[mesh_id].live = false; // arena memory is reclaimed when the context is destroyed
    ctx->mesh_table_dirty = true;
    return SFH_OK;
}

int sfh_mesh_get_stats(sfh_ctx* ctx, uint32_t mesh_id, sfh_mesh_stats* out)
{
    if (!ctx || !out) return SFH_ERR_INVALID;
    std::lock_guard<std::mutex> lock(ctx->mesh_mutex);
    if (mesh_id >= ctx->meshes.size() || !ctx->meshes[mesh_id].live) return fail(ctx, SFH_ERR_INVALID, "unknown mesh id %u", mesh_id);
    *out = ctx->meshes[mesh_id].stats;
    return SFH_OK;
}

int sfh_waves_upload(sfh_ctx* ctx, const sfh_waves* w)
{
    if (!ctx) return SFH_ERR_INVALID;
    if (!w || !w->data) return fail(ctx, SFH_ERR_INVALID, "null waves");
    if (w->fft_size < 2 || w->fft_size > 4096) return fail(ctx, SFH_ERR_INVALID, "fft_size %d out of range", w->fft_size);
    CU(ctx, cudaSetDevice(ctx->device));
    const size_t N = (size_t)w->fft_size;
    const size_t texels = N * N + N + 1; // what OpenGLRealOcean.cpp:289-311 can index with i,j in [0,N]
    const size_t want = texels * 4;
    const size_t have = w->n_floats < want ? w->n_floats : want;
    int rc;
    const bool grow = want > ctx->d_wave_rgba.cap || texels > ctx->d_wave_ch0.cap || want * 4 > ctx->h_wave.cap;
    if (grow || ctx->have_waves) CU(ctx, cudaStreamSynchronize(ctx->stream)); // staging buffer reuse
    if ((rc = ensure_dev(ctx, ctx->d_wave_rgba, want)) != SFH_OK) return rc;
    if ((rc = ensure_dev(ctx, ctx->d_wave_ch0, texels)) != SFH_OK) return rc;
    if ((rc = ensure_dev(ctx, ctx->d_wave_ch1, texels)) != SFH_OK) return rc;
    const void* src = w->data;
    if (!is_pinned_host(w->data)) {
        if ((rc = ensure_pin(ctx, ctx->h_wave, want * sizeof(float))) != SFH_OK) return rc;
        std::memcpy(ctx->h_wave.p, w->data, have * sizeof(float));
        src = ctx->h_wave.p;
    }
    CU(ctx, cudaMemcpyAsync(ctx->d_wave_rgba.p, src, have * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = ensure_dev(ctx, ctx->d_wave_hmax, 2)) != SFH_OK) return rc;
    CU(ctx, cudaMemsetAsync(ctx->d_wave_hmax.p, 0, 2 * sizeof(uint32_t), ctx->stream));
    k_waves_planar<<<(unsigned)((texels + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_wave_rgba.p, (uint32_t)have, (uint32_t)texels,
                                                                               ctx->d_wave_ch0.p, ctx->d_wave_ch1.p, ctx->d_wave_hmax.p);
    CU(ctx, cudaGetLastError());
    ctx->waves.hmax = ctx->d_wave_hmax.p;
    ctx->waves.ch0 = ctx->d_wave_ch0.p;
    ctx->waves.ch1 = ctx->d_wave_ch1.p;
    ctx->waves.N = w->fft_size;
    ctx->waves.gx = w->grid_sizes[0];
    ctx->waves.gy = w->grid_sizes[1];
    ctx->have_waves = true;
    return SFH_OK;
}

int sfh_wavegen_init(sfh_ctx* ctx, const sfh_wave_params* p)
{
    if (!ctx) return SFH_ERR_INVALID;
    if (!p) return fail(ctx, SFH_ERR_INVALID, "null argument");
    if (p->passes < 2 || p->passes > 10) return fail(ctx, SFH_ERR_INVALID, "passes %u outside [2, 10]", p->passes);
    for (int g = 0; g < 4; ++g)
        if (!(p->grid_sizes[g] > 0.f)) return fail(ctx, SFH_ERR_INVALID, "grid_sizes[%d] must be positive", g);
    if (!(p->wind > 0.f) || !(p->omega > 0.f)) return fail(ctx, SFH_ERR_INVALID, "wind and omega must be positive");
    CU(ctx, cudaSetDevice(ctx->device));
    const size_t N = (size_t)1 << p->passes, NN = N * N;
    wavegen_spectrum(*p, ctx->wg_s12, ctx->wg_s34);
    std::vector<int2> bsrc;
    std::vector<float2> bw;
    wavegen_butterfly(p->passes, bsrc, bw);
    std::vector<float> h12(ctx->wg_s12.size()), h34(ctx->wg_s34.size()); // as the GL_RGBA16F textures hold them
    for (size_t i = 0; i < h12.size(); ++i) { h12[i] = half_storage(ctx->wg_s12[i]); h34[i] = half_storage(ctx->wg_s34[i]); }
    int rc;
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if ((rc = ensure_dev(ctx, ctx->d_wg_s12, NN)) != SFH_OK) return rc;
    if ((rc = ensure_dev(ctx, ctx->d_wg_s34, NN)) != SFH_OK) return rc;
    if ((rc = ensure_dev(ctx, ctx->d_wg_tmp, 3 * NN)) != SFH_OK) return rc;
    if ((rc = ensure_dev(ctx, ctx->d_wg_fft, 4 * NN)) != SFH_OK) return rc;
    if ((rc = ensure_dev(ctx, ctx->d_wg_bf_src, N * p->passes)) != SFH_OK) return rc;
    if ((rc = ensure_dev(ctx, ctx->d_wg_bf_w, N * p->passes)) != SFH_OK) return rc;
    CU(ctx, cudaMemcpyAsync(ctx->d_wg_s12.p, h12.data(), NN * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(ctx->d_wg_s34.p, h34.data(), NN * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(ctx->d_wg_bf_src.p, bsrc.data(), bsrc.size() * sizeof(int2), cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(ctx->d_wg_bf_w.p, bw.data(), bw.size() * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream)); // the staging vectors die with this scope
    ctx->wg_params = *p;
    ctx->wg_ready = true;
    return SFH_OK;
}

int sfh_wavegen_simulate(sfh_ctx* ctx, float t)
{
    if (!ctx) return SFH_ERR_INVALID;
    if (!ctx->wg_ready) return fail(ctx, SFH_ERR_STATE, "sfh_wavegen_simulate before sfh_wavegen_init");
    CU(ctx, cudaSetDevice(ctx->device));
    const sfh_wave_params& p = ctx->wg_params;
    const int N = 1 << p.passes;
    const size_t texels = (size_t)N * N + N + 1;
    int rc;
    if (texels > ctx->d_wave_ch0.cap) CU(ctx, cudaStreamSynchronize(ctx->stream));
    if ((rc = ensure_dev(ctx, ctx->d_wave_ch0, texels)) != SFH_OK) return rc;
    if ((rc = ensure_dev(ctx, ctx->d_wave_ch1, texels)) != SFH_OK) return rc;
    WaveGenDev G;
    G.s12 = ctx->d_wg_s12.p; G.s34 = ctx->d_wg_s34.p; G.bf_src = ctx->d_wg_bf_src.p; G.bf_w = ctx->d_wg_bf_w.p;
    G.tmp = ctx->d_wg_tmp.p; G.fft = ctx->d_wg_fft.p; G.N = N; G.passes = (int32_t)p.passes;
    for (int g = 0; g < 4; ++g) G.inv_grid[g] = (float)(2.f * M_PI * (float)N / p.grid_sizes[g]); // OpenGLOcean.cpp:472-475
    const dim3 grid((unsigned)N, 3);
    const size_t smem = (size_t)2 * N * sizeof(float4);
    k_wave_rows<<<grid, N, smem, ctx->stream>>>(G, t);
    k_wave_cols<<<grid, N, smem, ctx->stream>>>(G);
    // the per-face kernels read planar copies of channels 0 and 1 of layer 0 (+ one row and one texel of layer 1)
    if ((rc = ensure_dev(ctx, ctx->d_wave_hmax, 2)) != SFH_OK) return rc;
    CU(ctx, cudaMemsetAsync(ctx->d_wave_hmax.p, 0, 2 * sizeof(uint32_t), ctx->stream));
    k_waves_planar<<<(unsigned)((texels + 255) / 256), 256, 0, ctx->stream>>>(reinterpret_cast<const float*>(ctx->d_wg_fft.p),
                                                                               (uint32_t)(texels * 4), (uint32_t)texels, ctx->d_wave_ch0.p,
                                                                               ctx->d_wave_ch1.p, ctx->d_wave_hmax.p);
    CU(ctx, cudaGetLastError());
    ctx->waves.hmax = ctx->d_wave_hmax.p;
    ctx->waves.ch0 = ctx->d_wave_ch0.p;
    ctx->waves.ch1 = ctx->d_wave_ch1.p;
    ctx->waves.N = N;
    ctx->waves.gx = p.grid_sizes[0];
    ctx->waves.gy = p.grid_sizes[1];
    ctx->have_waves = true;
    return SFH_OK;
}

int sfh_wavegen_download(sfh_ctx* ctx, float* fft_data, size_t n_floats)
{
    if (!ctx) return SFH_ERR_INVALID;
    if (!ctx->wg_ready) return fail(ctx, SFH_ERR_STATE, "sfh_wavegen_download before sfh_wavegen_init");
    if (!fft_data) return fail(ctx, SFH_ERR_INVALID, "null argument");
    CU(ctx, cudaSetDevice(ctx->device));
    const size_t N = (size_t)1 << ctx->wg_params.passes, all = 4 * N * N * 4;
    const size_t n = n_floats < all ? n_floats : all;
    CU(ctx, cudaMemcpyAsync(fft_data, ctx->d_wg_fft.p, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return SFH_OK;
}

int sfh_wavegen_get_spectrum(sfh_ctx* ctx, float* spectrum12, float* spectrum34, size_t n_floats_each)
{
    if (!ctx) return SFH_ERR_INVALID;
    if (!ctx->wg_ready) return fail(ctx, SFH_ERR_STATE, "sfh_wavegen_get_spectrum before sfh_wavegen_init");
    if (!spectrum12 || !spectrum34) return fail(ctx, SFH_ERR_INVALID, "null argument");
    const size_t n = n_floats_each < ctx->wg_s12.size() ? n_floats_each : ctx->wg_s12.size();
    std::memcpy(spectrum12, ctx->wg_s12.data(), n * sizeof(float));
    std::memcpy(spectrum34, ctx->wg_s34.data(), n * sizeof(float));
    return SFH_OK;
}

int sfh_step_async(sfh_ctx* ctx, const sfh_ocean* ocean, const sfh_body_state* states, uint32_t n_bodies, sfh_info* out)
{
    return step_host(ctx, ocean, states, n_bodies, out, nullptr, 0, false);
}

int sfh_sync(sfh_ctx* ctx)
{
    if (!ctx) return SFH_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    bool published = false;
    const bool polled = ctx->zc_pending;
    const volatile uint32_t* seq = ctx->h_status.p ? static_cast<const volatile uint32_t*>(ctx->h_status.p) + 1 : nullptr;
    if (polled) { // a zero-copy step: its epilogue writes the sequence number last (after a system-wide fence)
        ctx->zc_pending = false;
        for (uint32_t spin = 0; spin < (1u << 22) && !published; ++spin) published = *seq == ctx->zc_seq;
    }
    if (!published) {
        CU(ctx, cudaStreamSynchronize(ctx->stream)); // also the path that reports a failed kernel
        if (polled) ctx->zc_seq = *seq;               // stay in step with the device's counter whatever happened
    }
    if (ctx->step_in_flight) {
        ctx->step_in_flight = false;
        if (ctx->pending_out) std::memcpy(ctx->pending_out, ctx->h_info.p, ctx->pending_out_bytes);
        if (ctx->pending_codes) std::memcpy(ctx->pending_codes, ctx->h_codes.p, ctx->pending_codes_bytes);
        ctx->pending_out = nullptr;
        ctx->pending_codes = nullptr;
        if (ctx->h_status.p && *static_cast<uint32_t*>(ctx->h_status.p) != 0) {
            cudaMemsetAsync(ctx->d_status.p, 0, sizeof(uint32_t), ctx->stream); // the error flags only
            *static_cast<uint32_t*>(ctx->h_status.p) = 0;
            ctx->device_step_pending = false;
            return fail(ctx, SFH_ERR_INVALID, "a body referenced an unknown or released mesh id");
        }
    }
    if (ctx->device_step_pending) { // device-resident steps cannot be validated on the host: fetch the status word they left behind
        ctx->device_step_pending = false;
        uint32_t st = 0;
        CU(ctx, cudaMemcpyAsync(&st, ctx->d_status.p, sizeof(st), cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        if (st != 0) {
            CU(ctx, cudaMemsetAsync(ctx->d_status.p, 0, sizeof(uint32_t), ctx->stream));
            return fail(ctx, SFH_ERR_INVALID, (st & 4u) ? "a body of a device-resident step took a path its hints excluded" :
                        ((st & 2u) ? "the surface work list of a device-resident step did not fit" :
                                     "a body of a device-resident step referenced an unknown or released mesh id"));
        }
    }
    return SFH_OK;
}

int sfh_step(sfh_ctx* ctx, const sfh_ocean* ocean, const sfh_body_state* states, uint32_t n_bodies, sfh_info* out)
{
    int rc = sfh_step_async(ctx, ocean, states, n_bodies, out);
    if (rc != SFH_OK) return rc;
    return sfh_sync(ctx);
}

int sfh_step_debug(sfh_ctx* ctx, const sfh_ocean* ocean, const sfh_body_state* states, uint32_t n_bodies, sfh_info* out,
                   uint8_t* face_codes, size_t face_codes_bytes)
{
    if (!face_codes) return fail(ctx, SFH_ERR_INVALID, "face_codes == NULL");
    int rc = step_host(ctx, ocean, states, n_bodies, out, face_codes, face_codes_bytes, true);
    if (rc != SFH_OK) return rc;
    return sfh_sync(ctx);
}

// sf::Atmosphere as the kernels' fluid record: wind = currents that are always enabled (Atmosphere.cpp:146-152)
static int atmosphere_as_fluid(sfh_ctx* ctx, const sfh_atmosphere* atm, sfh_ocean* fluid)
{
    if (!ctx) return SFH_ERR_INVALID;
    if (!atm) return fail(ctx, SFH_ERR_INVALID, "null argument");
    if (atm->n_winds > SFH_MAX_CURRENTS) return fail(ctx, SFH_ERR_INVALID, "n_winds %u > %u", atm->n_winds, SFH_MAX_CURRENTS);
    fluid->density = atm->density;
    fluid->gravity_z = 0.0;
    fluid->has_waves = 0;
    fluid->currents_enabled = 1;
    fluid->n_currents = atm->n_winds;
    fluid->reserved = 0;
    for (uint32_t i = 0; i < atm->n_winds; ++i) { fluid->currents[i] = atm->winds[i]; fluid->currents[i].enabled = 1; }
    return SFH_OK;
}

int sfh_aero_step(sfh_ctx* ctx, const sfh_atmosphere* atm, const sfh_body_state* states, uint32_t n_bodies, sfh_info* out)
{
    sfh_ocean fluid;
    int rc = atmosphere_as_fluid(ctx, atm, &fluid);
    if (rc != SFH_OK) return rc;
    rc = step_host(ctx, &fluid, states, n_bodies, out, nullptr, 0, false, true);
    if (rc != SFH_OK) return rc;
    return sfh_sync(ctx);
}

int sfh_aero_step_debug(sfh_ctx* ctx, const sfh_atmosphere* atm, const sfh_body_state* states, uint32_t n_bodies, sfh_info* out,
                        uint8_t* face_codes, size_t face_codes_bytes)
{
    if (ctx && !face_codes && n_bodies) return fail(ctx, SFH_ERR_INVALID, "null argument");
    sfh_ocean fluid;
    int rc = atmosphere_as_fluid(ctx, atm, &fluid);
    if (rc != SFH_OK) return rc;
    rc = step_host(ctx, &fluid, states, n_bodies, out, face_codes, face_codes_bytes, true, true);
    if (rc != SFH_OK) return rc;
    return sfh_sync(ctx);
}

int sfh_step_device(sfh_ctx* ctx, const sfh_ocean* ocean, const sfh_body_state* d_states, uint32_t n_bodies, sfh_info* d_out)
{
    return sfh_step_device_hinted(ctx, ocean, d_states, n_bodies, d_out, 0u);
}

int sfh_step_device_hinted(sfh_ctx* ctx, const sfh_ocean* ocean, const sfh_body_state* d_states, uint32_t n_bodies, sfh_info* d_out, uint32_t hints)
{
    if (!ctx) return SFH_ERR_INVALID;
    if (!ocean || (!d_states && n_bodies) || (!d_out && n_bodies)) return fail(ctx, SFH_ERR_INVALID, "null argument");
    CU(ctx, cudaSetDevice(ctx->device));
    if (n_bodies == 0) return SFH_OK;
    StepPlan plan;
    plan.host_states = false;
    plan.may_surf = !(hints & SFH_HINT_NO_SURFACE);
    plan.may_sub = !(hints & SFH_HINT_NO_SUBMERGED);
    if (ctx->profiling) CU(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    int rc = run_step(ctx, ocean, d_states, n_bodies, d_out, plan);
    if (rc != SFH_OK) return rc;
    ctx->ev_valid = ctx->profiling;
    ctx->device_step_pending = true;
    return SFH_OK;
}

int sfh_aero_step_device(sfh_ctx* ctx, const sfh_atmosphere* atm, const sfh_body_state* d_states, uint32_t n_bodies, sfh_info* d_out)
{
    sfh_ocean fluid;
    int rc = atmosphere_as_fluid(ctx, atm, &fluid);
    if (rc != SFH_OK) return rc;
    if ((!d_states && n_bodies) || (!d_out && n_bodies)) return fail(ctx, SFH_ERR_INVALID, "null argument");
    CU(ctx, cudaSetDevice(ctx->device));
    if (n_bodies == 0) return SFH_OK;
    StepPlan plan;
    plan.host_states = false;
    plan.aero = true;
    if (ctx->profiling) CU(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    rc = run_step(ctx, &fluid, d_states, n_bodies, d_out, plan);
    if (rc != SFH_OK) return rc;
    ctx->ev_valid = ctx->profiling;
    ctx->device_step_pending = true;
    return SFH_OK;
}

int sfh_cable_step(sfh_ctx* ctx, const sfh_ocean* ocean, const sfh_cable* cables, uint32_t n_cables, const double* node_pos,
                   const double* node_vel, uint32_t n_nodes_total, sfh_node_forces* out)
{
    if (!ctx) return SFH_ERR_INVALID;
    if (!ocean || (n_cables && !cables) || (n_nodes_total && (!node_pos || !node_vel || !out))) return fail(ctx, SFH_ERR_INVALID, "null argument");
    if (ctx->step_in_flight) { int rc0 = sfh_sync(ctx); if (rc0 != SFH_OK) return rc0; }
    CU(ctx, cudaSetDevice(ctx->device));
    uint32_t max_nodes = 0;
    for (uint32_t i = 0; i < n_cables; ++i) {
        if (cables[i].n_nodes < 2) return fail(ctx, SFH_ERR_INVALID, "cable %u: needs at least 2 nodes", i);
        if ((uint64_t)cables[i].first_node + cables[i].n_nodes > n_nodes_total) return fail(ctx, SFH_ERR_INVALID, "cable %u: nodes out of range", i);
        if (!(cables[i].radius > 0.0)) return fail(ctx, SFH_ERR_INVALID, "cable %u: radius must be positive", i);
        if (cables[i].n_nodes > max_nodes) max_nodes = cables[i].n_nodes;
    }
    if (n_nodes_total == 0) return SFH_OK;
    int rc;
    bool general = false;
    CU(ctx, cudaStreamSynchronize(ctx->stream)); // buffers below may grow; this entry is synchronous anyway
    if ((rc = stage_ocean(ctx, ocean, &general)) != SFH_OK) return rc;
    if ((rc = ensure_dev(ctx, ctx->d_cables, n_cables ? n_cables : 1)) != SFH_OK) return rc;
    if ((rc = ensure_dev(ctx, ctx->d_node_pos, (size_t)3 * n_nodes_total)) != SFH_OK) return rc;
    if ((rc = ensure_dev(ctx, ctx->d_node_vel, (size_t)3 * n_nodes_total)) != SFH_OK) return rc;
    if ((rc = ensure_dev(ctx, ctx->d_node_out, n_nodes_total)) != SFH_OK) return rc;
    CU(ctx, cudaMemcpyAsync(ctx->d_cables.p, cables, sizeof(sfh_cable) * n_cables, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(ctx->d_node_pos.p, node_pos, sizeof(double) * 3 * n_nodes_total, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(ctx->d_node_vel.p, node_vel, sizeof(double) * 3 * n_nodes_total, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemsetAsync(ctx->d_node_out.p, 0, sizeof(sfh_node_forces) * n_nodes_total, ctx->stream)); // nodes outside every cable
    if (n_cables) {
        CableParams cp;
        cp.cables = ctx->d_cables.p; cp.n_cables = n_cables; cp.pos = ctx->d_node_pos.p; cp.vel = ctx->d_node_vel.p;
        cp.ocean = ctx->d_ocean.p; cp.waves = ctx->waves; cp.out = ctx->d_node_out.p;
        const dim3 grid((max_nodes + 127) / 128, n_cables);
        k_cable_nodes<<<grid, 128, 0, ctx->stream>>>(cp);
        CU(ctx, cudaGetLastError());
    }
    CU(ctx, cudaMemcpyAsync(out, ctx->d_node_out.p, sizeof(sfh_node_forces) * n_nodes_total, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return SFH_OK;
}

static int point_query(sfh_ctx* ctx, const sfh_ocean* ocean, const float* xyz, uint32_t n, float* out, bool velocity)
{
    if (!ctx) return SFH_ERR_INVALID;
    if (!ocean || (n && (!xyz || !out))) return fail(ctx, SFH_ERR_INVALID, "null argument");
    CU(ctx, cudaSetDevice(ctx->device));
    if (ctx->step_in_flight) { int rc = sfh_sync(ctx); if (rc != SFH_OK) return rc; }
    if (n == 0) return SFH_OK;
    bool general = false;
    int rc = stage_ocean(ctx, ocean, &general);
    if (rc != SFH_OK) return rc;
    float *d_in = nullptr, *d_out = nullptr;
    const size_t out_floats = velocity ? 3 * (size_t)n : (size_t)n;
    CU(ctx, cudaMalloc(&d_in, 3 * (size_t)n * sizeof(float)));
    cudaError_t e = cudaMalloc(&d_out, out_floats * sizeof(float));
    if (e != cudaSuccess) { cudaFree(d_in); return fail(ctx, SFH_ERR_OOM, "cudaMalloc: %s", cudaGetErrorString(e)); }
    cudaMemcpyAsync(d_in, xyz, 3 * (size_t)n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
    if (velocity) k_query_velocity<<<(n + 127) / 128, 128, 0, ctx->stream>>>(d_in, n, ctx->d_ocean.p, d_out);
    else k_query_depth<<<(n + 127) / 128, 128, 0, ctx->stream>>>(d_in, n, ocean->has_waves != 0, ctx->waves, d_out);
    cudaMemcpyAsync(out, d_out, out_floats * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
    e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_in); cudaFree(d_out);
    if (e != cudaSuccess) return fail(ctx, SFH_ERR_CUDA, "point query: %s", cudaGetErrorString(e));
    return SFH_OK;
}

int sfh_query_depth(sfh_ctx* ctx, const sfh_ocean* ocean, const float* xyz, uint32_t n, float* depth_out)
{
    return point_query(ctx, ocean, xyz, n, depth_out, false);
}

int sfh_query_fluid_velocity(sfh_ctx* ctx, const sfh_ocean* ocean, const float* xyz, uint32_t n, float* vel_out)
{
    return point_query(ctx, ocean, xyz, n, vel_out, true);
}

int sfh_selftest(sfh_ctx* ctx, uint32_t which, uint64_t n, uint64_t* tested_out, uint64_t* mismatches_out)
{
    if (!ctx) return SFH_ERR_INVALID;
    if (which > 4 || !tested_out || !mismatches_out) return fail(ctx, SFH_ERR_INVALID, "bad selftest arguments");
    CU(ctx, cudaSetDevice(ctx->device));
    unsigned long long* d = nullptr;
    CU(ctx, cudaMalloc(&d, 3 * sizeof(unsigned long long)));
    cudaMemsetAsync(d, 0, 3 * sizeof(unsigned long long), ctx->stream);
    k_selftest<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(which, n, d, d + 1);
    unsigned long long h[3] = {0, 0, 0};
    cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail(ctx, SFH_ERR_CUDA, "selftest: %s", cudaGetErrorString(e));
    *mismatches_out = h[0];
    *tested_out = h[1];
    if (h[0]) fail(ctx, SFH_OK, "selftest %u: %llu mismatches, e.g. operand/index 0x%llx", which, h[0], h[2] & ~(1ull << 63));
    return SFH_OK;
}

int sfh_host_alloc(sfh_ctx* ctx, size_t bytes, void** out)
{
    if (!ctx || !out) return SFH_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return SFH_OK;
}

int sfh_host_free(sfh_ctx* ctx, void* p)
{
    if (!ctx) return SFH_ERR_INVALID;
    if (p) CU(ctx, cudaFreeHost(p));
    return SFH_OK;
}

int sfh_set_profiling(sfh_ctx* ctx, int enabled)
{
    if (!ctx) return SFH_ERR_INVALID;
    ctx->profiling = enabled != 0;
    const char* roles = getenv("SFH_PROFILE_ROLES");
    ctx->profile_roles = roles && roles[0] == '1';
    ctx->ev_valid = false;
    return SFH_OK;
}

int sfh_get_timings(sfh_ctx* ctx, sfh_timings* out)
{
    if (!ctx || !out) return SFH_ERR_INVALID;
    std::memset(out, 0, sizeof(*out));
    if (!ctx->ev_valid) return fail(ctx, SFH_ERR_STATE, "no profiled step: call sfh_set_profiling(ctx,1) before the step");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaEventSynchronize(ctx->ev[5]));
    CU(ctx, cudaEventElapsedTime(&out->prologue_ms, ctx->ev[1], ctx->ev[2]));
    if (ctx->ev_sub) CU(ctx, cudaEventElapsedTime(&out->submerged_ms, ctx->ev[2], ctx->ev[3]));
    if (ctx->ev_surf) {
        CU(ctx, cudaEventElapsedTime(&out->surface_ms, ctx->ev[3], ctx->ev[4]));
        CU(ctx, cudaEventElapsedTime(&out->classify_ms, ctx->ev[3], ctx->ev[6]));
        CU(ctx, cudaEventElapsedTime(&out->wet_ms, ctx->ev[6], ctx->ev[7]));
        CU(ctx, cudaEventElapsedTime(&out->clip_ms, ctx->ev[7], ctx->ev[4]));
        uint32_t ctrl[SC_WORDS];
        CU(ctx, cudaMemcpyAsync(ctrl, ctx->d_surf_ctrl.p, sizeof(ctrl), cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        out->chunks_surface = ctrl[SC_ITEMS]; out->chunks_wet = ctrl[SC_WET_ENTRIES]; out->chunks_clip = ctrl[SC_CLIP_ENTRIES];
    }
    CU(ctx, cudaEventElapsedTime(&out->epilogue_ms, ctx->ev[4], ctx->ev[5]));
    CU(ctx, cudaEventElapsedTime(&out->total_ms, ctx->ev[0], ctx->ev[5]));
    out->launches = ctx->last_launches;
    out->faces = ctx->last_faces;
    return SFH_OK;
}

} // extern "C"
