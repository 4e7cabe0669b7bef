/*
 * sfh.h — C ABI of the B200 (sm_100a) implementation of Stonefish's per-triangle hydrodynamic force
 * integration. Plain pointers and sizes only; the shared library is stonefish_b200/libsfh.so.
 *
 * The reference has no FFI layer for this path: the boundary is the C++ call site
 *   Ocean::ApplyFluidForces -> SolidEntity::ComputeHydrodynamicForces
 *   (Library/src/entities/forcefields/Ocean.cpp:219-268, Library/src/entities/SolidEntity.cpp:1799-1850)
 * driven once per ocean-ghost overlap pair by SimulationManager::SimulationTickCallback
 * (Library/src/core/SimulationManager.cpp:1619-1655). Each entry point below names the reference
 * interface it replaces; INTEGRATION.md shows the call-site patch.
 *
 * Error convention: the reference's integrators are `void` and cannot fail. Here every call returns
 * SFH_OK (0) or a negative SFH_ERR_*; sfh_last_error() gives the text. There is NO CPU fallback: without a
 * usable CUDA device sfh_create fails and nothing else can be called.
 *
 * Threading: one sfh_ctx per GPU (per process/rank). A context is not re-entrant; calls on different
 * contexts are independent. sfh_mesh_upload may be called concurrently from several host threads.
 */
#ifndef SFH_H
#define SFH_H

#include "sfh_types.h"

#ifdef __cplusplus
extern "C" {
#endif

#define SFH_ABI_VERSION 2

enum {
    SFH_OK = 0,
    SFH_ERR_INVALID = -1, /* bad argument (null pointer, index out of range, too many currents, ...) */
    SFH_ERR_CUDA = -2,    /* CUDA runtime error; text in sfh_last_error                               */
    SFH_ERR_NO_DEVICE = -3,
    SFH_ERR_OOM = -4,
    SFH_ERR_STATE = -5    /* e.g. has_waves without sfh_waves_upload, sync without step              */
};

/* sfh_mesh_upload flags */
enum {
    SFH_MESH_KEEP_ORDER = 1u << 0 /* do not Morton-sort faces before chunking (default: sort; the sums are
                                     order-independent within the 1e-9 contract, class counts exactly)  */
};

typedef struct sfh_ctx sfh_ctx;

typedef struct sfh_mesh_stats {
    uint32_t n_faces, n_verts;
    uint32_t n_chunks;          /* device chunks of <=256 faces / <=256 local vertices              */
    uint32_t n_chunk_verts;     /* vertices stored after per-chunk duplication                     */
    uint64_t device_bytes;      /* bytes actually read per evaluation of this mesh                 */
    uint64_t algorithmic_bytes; /* 12*n_faces + 12*n_verts                           */
} sfh_mesh_stats;

typedef struct sfh_timings {
    float prologue_ms;  /* per-body prologue + work-list scan                                       */
    float submerged_ms; /* per-face kernel, INSIDE bodies     (0 when not launched)                  */
    float surface_ms;   /* per-face kernel, CROSSING bodies   (0 when not launched)                  */
    float epilogue_ms;  /* per-body reduction + finalisation                                        */
    float total_ms;     /* first enqueue to last enqueue of the step, incl. copies                   */
    uint32_t launches;  /* kernels launched by the step                                             */
    uint32_t reserved;
    uint64_t faces;     /* triangle-evals of the step                                               */
    /* ABI 2: surface_ms = classify_ms + wet_ms + clip_ms, and how the chunks of the crossing bodies were classified
     * (dry chunks = chunks_surface - chunks_wet - chunks_clip are dropped unread)                              */
    float classify_ms;  /* chunk classifier of the surface path                                     */
    float wet_ms;       /* chunks entirely under water (no clipping)                                */
    float clip_ms;      /* chunks the water line may cross (eight clipping classes)                 */
    uint32_t chunks_surface, chunks_wet, chunks_clip;
} sfh_timings;

int sfh_abi_version(void);

/* Context = one GPU: stream, pinned staging, device arenas.
 * Replaces: the ThreadPool the reference creates for hydrodynamics (SimulationApp.cpp:132-142). */
int sfh_create(int cuda_device, sfh_ctx** out);
/* Same, but enqueue on a caller-owned cudaStream_t (e.g. torch's current stream) instead of a private one. Everything a step does
 * is ordered with that stream: a small step with both submerged and surface-crossing bodies runs one of its kernels on a private side
 * stream, forked after the step's first kernel and joined before its last by events (two branches inside the step, invisible outside). */
int sfh_create_on_stream(int cuda_device, void* cuda_stream, sfh_ctx** out);
void sfh_destroy(sfh_ctx* ctx);
/* Text of the last error on ctx (ctx == NULL: last error of sfh_create on this thread). Never NULL. */
const char* sfh_last_error(const sfh_ctx* ctx);

/* Static geometry, uploaded once per mesh (SolidEntity::phyMesh is immutable after BuildScenario,
 * SolidEntity.cpp:108-112,1153-1181). `pos` points at the first vertex position and may be a strided AoS
 * exactly as Mesh::getVertexDataPointer()/getVertexSize() give it (stride 24 or 44 bytes,
 * OpenGLDataStructs.h:190-262); `faces` is Mesh::getFaceDataPointer() (3 u32 per face). */
int sfh_mesh_upload(sfh_ctx* ctx, const void* pos, size_t stride_bytes, uint32_t n_verts, const uint32_t* faces,
                    uint32_t n_faces, uint32_t flags, uint32_t* mesh_id_out);
int sfh_mesh_release(sfh_ctx* ctx, uint32_t mesh_id);
int sfh_mesh_get_stats(sfh_ctx* ctx, uint32_t mesh_id, sfh_mesh_stats* out);

/* Wave height field = OpenGLRealOcean::fftData (OpenGLRealOcean.cpp:191-198, refreshed :333-356).
 * Copies what the lookup can touch (layer 0 plus one row and one texel) to the device; call again whenever
 * the field changes. The copy is ordered before the next step on the context's stream. */
int sfh_waves_upload(sfh_ctx* ctx, const sfh_waves* waves);

/* Wave-field PRODUCER on the device: what OpenGLOcean::InitializeSimulation / Simulate do with fragment
 * shaders and OpenGLRealOcean::Simulate reads back through a PBO (OpenGLOcean.cpp:414-534,857-1047, OpenGLRealOcean.cpp:333-356;
 * shaders oceanInit.frag, oceanFFTX.frag, oceanFFTY.frag), so headless / batched runs have waves without OpenGL.
 *   sfh_wavegen_init      GenerateWavesSpectrum + ComputeButterflyLookupTable for `params` (host, once), uploaded the way
 *                         the reference's GL_RGBA16F textures store them (half precision, round to nearest even)
 *   sfh_wavegen_simulate  the h(k,t) pass and the 2 x passes FFT passes at time t (the reference advances t by dt per rendered
 *                         frame, OpenGLOcean.cpp:476-477; here the caller owns the clock). The result REPLACES the context's
 *                         wave field exactly as sfh_waves_upload of that fftData would; asynchronous on the context's stream.
 *   sfh_wavegen_download  fftData [4][N][N][4] floats (layer 3 zero) to the host = the PBO read-back; synchronises
 *   sfh_wavegen_get_spectrum  spectrum12 / spectrum34 (N*N*4 floats each) as generated, before half storage */
int sfh_wavegen_init(sfh_ctx* ctx, const sfh_wave_params* params);
int sfh_wavegen_simulate(sfh_ctx* ctx, float t);
int sfh_wavegen_download(sfh_ctx* ctx, float* fft_data, size_t n_floats);
int sfh_wavegen_get_spectrum(sfh_ctx* ctx, float* spectrum12, float* spectrum34, size_t n_floats_each);

/* One recompute of the hydrodynamic forces of n_bodies bodies:
 * replaces the `for each overlapping pair: ocean->ApplyFluidForces(...)` loop
 * (SimulationManager.cpp:1625-1651) up to, not including, ApplyHydrodynamicForces.
 * states/out are HOST buffers (pageable or pinned via sfh_host_alloc); the call enqueues H2D copies,
 * kernels and the D2H copy of `out` and returns; `out` is valid after sfh_sync. */
int sfh_step_async(sfh_ctx* ctx, const sfh_ocean* ocean, const sfh_body_state* states, uint32_t n_bodies,
                   sfh_info* out);
int sfh_sync(sfh_ctx* ctx);
/* sfh_step_async + sfh_sync */
int sfh_step(sfh_ctx* ctx, const sfh_ocean* ocean, const sfh_body_state* states, uint32_t n_bodies, sfh_info* out);
/* As sfh_step, plus one SFH_FACE_* code per face in ORIGINAL face order, bodies concatenated in input
 * order (sum of their meshes' n_faces bytes). Test/diagnostic entry; mirrors what DEBUG_HYDRO exposes. */
int sfh_step_debug(sfh_ctx* ctx, const sfh_ocean* ocean, const sfh_body_state* states, uint32_t n_bodies,
                   sfh_info* out, uint8_t* face_codes, size_t face_codes_bytes);
/* Device-resident variant: d_states / d_out are DEVICE pointers (state already in HBM, results stay in HBM).
 * Only the ocean parameters (a few KB) are copied. Asynchronous on the context's stream. The states cannot be validated on the
 * host: a body that names an unknown or released mesh id yields zero forces and makes the NEXT sfh_sync return SFH_ERR_INVALID
 * (sfh_sync after a device-resident step waits for the stream and fetches the step's status word). */
int sfh_step_device(sfh_ctx* ctx, const sfh_ocean* ocean, const sfh_body_state* d_states, uint32_t n_bodies,
                    sfh_info* d_out);
/* The same with what the caller knows about states the host cannot see (an environment batch knows its bodies' modes): the kernels of
 * a path no body can take are not launched (sfh_step derives this itself from its host-side states). A body that takes an excluded
 * path after all yields zero forces and makes the next sfh_sync return SFH_ERR_INVALID. */
enum {
    SFH_HINT_NO_SURFACE   = 1u << 0, /* no body crosses the surface: mode is SFH_POS_INSIDE or SFH_POS_OUTSIDE for all of them */
    SFH_HINT_NO_SUBMERGED = 1u << 1  /* no body is entirely under water: mode is SFH_POS_CROSSING or SFH_POS_OUTSIDE           */
};
int sfh_step_device_hinted(sfh_ctx* ctx, const sfh_ocean* ocean, const sfh_body_state* d_states, uint32_t n_bodies,
                           sfh_info* d_out, uint32_t hints);

/* Aerodynamic twin: one recompute of the air drag of n_bodies bodies — replaces the
 * `for each atmosphere-ghost pair: atmosphere->ApplyFluidForces(...)` recompute, i.e.
 * SolidEntity::ComputeAerodynamicForces (SolidEntity.cpp:1852-1923, Atmosphere.cpp:160-193), up to, not including,
 * ApplyAerodynamicForces. Same buffers and rules as sfh_step. Of each sfh_body_state it reads mesh_id, T_CG, T_C, v,
 * omega (mode is ignored: every face is evaluated) and, with SFH_F_DISPATCH, CorrectAerodynamicForces' inputs:
 * T_O_basis = T_CG2H.getBasis(), Cd = the proxy's coefficient vector ((0.5,0.5,1) for a cylinder, (1/a,1/b,1/c)/max for
 * an ellipsoid) and SFH_F_AERO_CD. Results: Fda -> out.Fdq, Tda -> out.Tdq, wrote_damping = 1, n_quadratic = faces that
 * took the drag branch; everything else is zero. Without SFH_F_DISPATCH the result is the static integrator's
 * (0.5 * density * sum, uncorrected). */
int sfh_aero_step(sfh_ctx* ctx, const sfh_atmosphere* atm, const sfh_body_state* states, uint32_t n_bodies, sfh_info* out);
int sfh_aero_step_debug(sfh_ctx* ctx, const sfh_atmosphere* atm, const sfh_body_state* states, uint32_t n_bodies,
                        sfh_info* out, uint8_t* face_codes, size_t face_codes_bytes);
/* Device-resident variant (see sfh_step_device). */
int sfh_aero_step_device(sfh_ctx* ctx, const sfh_atmosphere* atm, const sfh_body_state* d_states, uint32_t n_bodies,
                         sfh_info* d_out);

/* Cables (row N4): CableEntity::ComputeHydrodynamicForces (CableEntity.cpp:239-372, called from the same
 * Ocean::ApplyFluidForces loop, Ocean.cpp:257-266) for n_cables cables whose nodes lie in node_pos / node_vel (3 doubles per
 * node, n_nodes_total nodes; host buffers). Writes one sfh_node_forces per node (nodes not covered by a cable: zeros);
 * ApplyHydrodynamicForces (:374-381) stays on the host. fp64 throughout like the reference; the circular-segment geometry uses
 * atan/cos/sin/acos, whose last bit differs between CUDA and glibc: parity is 1e-12 relative, not bit-exact. Synchronous. */
int sfh_cable_step(sfh_ctx* ctx, const sfh_ocean* ocean, const sfh_cable* cables, uint32_t n_cables, const double* node_pos,
                   const double* node_vel, uint32_t n_nodes_total, sfh_node_forces* out);

/* Batched point queries with the same arithmetic as the per-face kernels (host buffers, synchronous):
 *   sfh_query_depth          = Ocean::GetDepth(const glm::vec3&)          (Ocean.cpp:149-168), n points -> n floats
 *   sfh_query_fluid_velocity = Ocean::GetFluidVelocity(const glm::vec3&)  (Ocean.cpp:198-201), n points -> 3n floats
 * The reference calls these one point at a time from actuators and sensors (Thruster.cpp:130-139,
 * Pressure.cpp:44); here they also let tests check rows a3-a8 of the scope table in isolation. */
int sfh_query_depth(sfh_ctx* ctx, const sfh_ocean* ocean, const float* xyz, uint32_t n, float* depth_out);
int sfh_query_fluid_velocity(sfh_ctx* ctx, const sfh_ocean* ocean, const float* xyz, uint32_t n, float* vel_out);

/* On-device checks of the hand-written IEEE sequences the kernels use (stonefish_b200/csrc/sfh_packed.cuh):
 *   which 0: square root, exhaustive over [2^-101, FLT_MAX]      which 2: shared-reciprocal division, n random pairs
 *   which 1: division by 3, exhaustive over its domain            which 3: packed f32x2 operators vs scalar, n samples
 *   which 4: the same division with unrelated magnitudes (coordinate / grid size), n random pairs
 * Writes the number of operands tested and of results that differ from the hardware's div.rn/sqrt.rn/add.rn/... */
int sfh_selftest(sfh_ctx* ctx, uint32_t which, uint64_t n, uint64_t* tested_out, uint64_t* mismatches_out);

/* Pinned host memory for states/out so the per-step copies are truly asynchronous. */
int sfh_host_alloc(sfh_ctx* ctx, size_t bytes, void** out);
int sfh_host_free(sfh_ctx* ctx, void* p);

/* CUDA-event timing of the kernels of each step (off by default; adds event records to the stream). */
int sfh_set_profiling(sfh_ctx* ctx, int enabled);
int sfh_get_timings(sfh_ctx* ctx, sfh_timings* out); /* synchronises the stream */

#ifdef __cplusplus
}
#endif
#endif /* SFH_H */
