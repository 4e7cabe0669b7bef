/*
 * hydro_oracle.h — C interface of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY. This library is a CPU restatement of the reference's algorithm
 * (Stonefish v1.6.0) used as the checker by tests/, by __graft_entry__.smoke() and as the
 * `cpu_baseline` / `--impl reference` leg of bench.py. Nothing in the product path
 * (stonefish_b200/, include/sfh.h) links, imports or calls it.
 *
 * PARITY: PINNED against the reference's own source text, with one stated residue. The reference's build cannot run here (GLM,
 * SDL2, Freetype and GL headers are absent, no network) and its repository holds no golden vectors or enabled known-answer tests
 * for this path (the only numeric checks are commented out, Tests/FluidDynamicsTest/FluidDynamicsTestManager.cpp:168-249). But
 * the DEFINITIONS of the path compile on their own: oracle/ref_extract.py pulls them out of the reference's files where they lie
 * (SolidEntity's integrators, dispatcher and corrections, Compound, CableEntity, Ocean, the wave lookup, Uniform/Jet/Pipe,
 * Atmosphere, glMatrixFromTransform — 31 definitions), oracle/ref_harness.cpp compiles that text with the reference's flags
 * against the reference's vendored Bullet and oracle/ref_shim/shim.hpp into oracle/_ref/librefsrc.so, and this oracle, in the
 * reference's own accumulation mode, reproduces its outputs BIT FOR BIT (tests/golden/ref_golden.json written from it;
 * tests/test_oracle_refsource.py). THE RESIDUE: GLM itself is not in the image, so shim.hpp stands in for the dozen GLM
 * primitives those definitions use (vec3/vec4/mat4 operators, dot, cross, length2, sqrt, mat4*vec4, gtx/scalar_multiplication),
 * following GLM 0.9.9's scalar code path as listed in SURVEY.md Appendix A.1 — that layer is UNPINNED. The fp64 pieces are
 * additionally validated against Bullet directly (oracle/bt_check.cpp -> oracle/_ref/libbtcheck.so, tests/golden/bt_golden.json).
 * The wave-field PRODUCER (wave_oracle.cpp) restates shader code whose arithmetic is implementation-defined: parity unpinned there.
 */
#ifndef HYDRO_ORACLE_H
#define HYDRO_ORACLE_H

#include "../include/sfh_types.h"

#ifdef __cplusplus
extern "C" {
#endif

/* accumulation modes */
enum {
    ORC_REF_F32_SEQ = 0, /* what the reference does: fp32 accumulators, face-index order (SolidEntity.cpp:1317-1324) */
    ORC_F64_ACC     = 1  /* identical fp32 per-face terms, summed in fp64: the parity contract of the CUDA path  */
};

typedef struct orc_mesh {
    const void* pos;      /* first vertex position; 3 floats at pos + i*stride_bytes (Mesh::getVertexDataPointer) */
    size_t stride_bytes;  /* Mesh::getVertexSize(): 24 (Vertex) or 44 (TexturableVertex)                          */
    uint32_t n_verts;
    const uint32_t* faces; /* 3*n_faces vertex ids (Mesh::getFaceDataPointer)                                      */
    uint32_t n_faces;
} orc_mesh;

/* a4: OpenGLRealOcean::ComputeInterpolatedWaveData / ComputeWaveHeight (OpenGLRealOcean.cpp:280-331) */
float orc_wave_bilinear(const sfh_waves* w, float x, float y, uint32_t channel);
float orc_wave_height(const sfh_waves* w, float x, float y);
/* a3: Ocean::GetDepth(glm::vec3) (Ocean.cpp:149-168) and the double overload (:170-173) */
float orc_get_depth(const sfh_ocean* o, const sfh_waves* w, float x, float y, float z);
double orc_get_depth_d(const sfh_ocean* o, const sfh_waves* w, const double p[3]);
/* a6-a8: VelocityField::GetVelocityAtPoint */
void orc_current_velocity(const sfh_current* c, const double p[3], double out[3]);
/* a5: Ocean::GetFluidVelocity, Vector3 overload (Ocean.cpp:183-196) and glm overload (:198-201) */
void orc_fluid_velocity_d(const sfh_ocean* o, const double p[3], double out[3]);
void orc_fluid_velocity_f(const sfh_ocean* o, const float p[3], float out[3]);
/* a9: SolidEntity::CheckBodyFluidPosition (SolidEntity.cpp:1239-1261) -> SFH_POS_* */
uint32_t orc_check_body_fluid_position(const sfh_ocean* o, const sfh_waves* w, const double aabb_min[3], const double aabb_max[3]);
/* a11: SolidEntity::CorrectHydrodynamicForces (SolidEntity.cpp:1263-1287); vectors corrected in place */
void orc_correct(double density, double Fdq[3], double Tdq[3], double Fdf[3], double Tdf[3],
                 const double Cd[3], const double Cf[3], const double T_O_basis[9]);
/* glMatrixFromTransform (OpenGLDataStructs.cpp:148-166): column-major float[16] */
void orc_gl_matrix(const double T[12], float out[16]);

/* a1: SolidEntity::ComputeHydrodynamicForcesSubmerged (SolidEntity.cpp:1723-1797).
 * a2: SolidEntity::ComputeHydrodynamicForcesSurface   (SolidEntity.cpp:1289-1721).
 * Raw static integrators; settings come from st->flags (SFH_F_DAMPING / SFH_F_BUOYANCY).
 * face_codes (nullable): n_faces bytes, SFH_FACE_* per face. mesh may be NULL (reference zeroes outputs). */
void orc_submerged(const orc_mesh* mesh, const sfh_ocean* o, const sfh_body_state* st, int acc_mode,
                   sfh_info* out, uint8_t* face_codes);
void orc_surface(const orc_mesh* mesh, const sfh_ocean* o, const sfh_waves* w, const sfh_body_state* st,
                 int acc_mode, sfh_info* out, uint8_t* face_codes);

/* One body through the same contract as sfh_step (include/sfh.h): honours st->mode (incl. AUTO) and
 * SFH_F_DISPATCH (= SolidEntity::ComputeHydrodynamicForces, SolidEntity.cpp:1799-1850). */
void orc_body(const orc_mesh* mesh, const sfh_ocean* o, const sfh_waves* w, const sfh_body_state* st,
              int acc_mode, sfh_info* out, uint8_t* face_codes);

/* Compound::ComputeHydrodynamicForces (Library/src/entities/solids/Compound.cpp:291-437).
 * `compound` carries what belongs to the whole body: mode (incl. AUTO with the compound's AABB), flags
 * (SFH_F_DAMPING/BUOYANCY = settings, SFH_F_BUOYANT = Compound::isBuoyant()), T_CG, v, omega, volume, surface, P_CB.
 * parts[i] carries what belongs to part i: T_C = T_O * origin_i * O2C_i, T_O_basis of T_O * origin_i, Cd, Cf, and in
 * flags SFH_F_BUOYANT = parts[i].solid->isBuoyant(); external[i] = parts[i].isExternal. Parts whose physics mode is not
 * SUBMERGED/FLOATING are simply not passed. The temporaries Fbp..Vsubp persist across parts as in the reference. */
void orc_compound(const orc_mesh* const* part_meshes, const sfh_body_state* parts, const uint32_t* external, uint32_t n_parts,
                  const sfh_ocean* o, const sfh_waves* w, const sfh_body_state* compound, int acc_mode, sfh_info* out);

/* N2, the aerodynamic twin.
 * orc_wind_velocity_f: Atmosphere::GetFluidVelocity(glm::vec3) (Atmosphere.cpp:146-157) — every wind, no isEnabled() test.
 * orc_correct_aero:    SolidEntity::CorrectAerodynamicForces (SolidEntity.cpp:1925-1956); B = T_CG2H.getBasis(), Cd = the
 *                      proxy vector, use_cd = CYLINDER/ELLIPSOID.
 * orc_aero:            SolidEntity::ComputeAerodynamicForces, static (SolidEntity.cpp:1865-1923) and, with SFH_F_DISPATCH,
 *                      the member (SolidEntity.cpp:1852-1863) = static + CorrectAerodynamicForces. Same contract as
 *                      sfh_aero_step (include/sfh.h): Fda -> out->Fdq, Tda -> out->Tdq. */
void orc_wind_velocity_f(const sfh_atmosphere* atm, const float p[3], float out[3]);
void orc_correct_aero(double Fda[3], double Tda[3], const double Cd[3], const double B[9], int use_cd);
void orc_aero(const orc_mesh* mesh, const sfh_atmosphere* atm, const sfh_body_state* st, int acc_mode, sfh_info* out,
              uint8_t* face_codes);
/* Compound::ComputeAerodynamicForces (Compound.cpp:439-464): external parts only, each corrected with its own proxy.
 * parts[i]: T_C = T_O * origin_i * O2C_i, and (SFH_F_DISPATCH set) the part's T_O_basis = T_CG2H basis, Cd, SFH_F_AERO_CD;
 * compound: T_CG, v, omega. */
void orc_compound_aero(const orc_mesh* const* part_meshes, const sfh_body_state* parts, const uint32_t* external, uint32_t n_parts,
                       const sfh_atmosphere* atm, const sfh_body_state* compound, int acc_mode, sfh_info* out);

/* N3, the wave-field producer (oracle/wave_oracle.cpp): OpenGLOcean::GenerateWavesSpectrum, ComputeButterflyLookupTable and
 * Simulate (oceanInit.frag + 16 FFT passes). See the header of wave_oracle.cpp for what "parity" can mean for shader code. */
float orc_half_round(float x);
void orc_wave_spectrum(const sfh_wave_params* p, float* spectrum12, float* spectrum34);
void orc_wave_butterfly(uint32_t size, uint32_t passes, float* table);
void orc_wave_simulate(const sfh_wave_params* p, const float* spectrum12, const float* spectrum34, const float* butterfly, float t,
                       int half_storage, float* fft_data);

/* N4: CableEntity::ComputeHydrodynamicForces (CableEntity.cpp:239-372) for one cable; pos/vel = that cable's nodes
 * (3 doubles each), out = n_nodes records. circularSegmentArea: CableEntity.cpp:217-237. */
void orc_cable(const sfh_ocean* o, const sfh_waves* w, const sfh_cable* cable, const double* pos, const double* vel,
               sfh_node_forces* out);
double orc_circular_segment_area(double radius, double h);

/* Whole step, scheduled like the reference: one task per body on a pool of n_threads
 * (SimulationManager.cpp:1641-1651, utils/ThreadPool.hpp). meshes[i] is the mesh of body i.
 * Returns wall seconds spent between the first enqueue and waitAll (the PerformanceMonitor bracket,
 * SimulationManager.cpp:1623,1653). */
double orc_step(const orc_mesh* const* meshes, const sfh_ocean* o, const sfh_waves* w,
                const sfh_body_state* states, uint32_t n_bodies, int acc_mode, int n_threads, sfh_info* out);

/* Not the reference's schedule (labelled so in reports): one body split into n_threads face ranges.
 * Only ORC_F64_ACC-style partial sums make sense here; used to give the CPU "all the host threads it can
 * use" on single-body workloads. Returns wall seconds. */
double orc_body_split(const orc_mesh* mesh, const sfh_ocean* o, const sfh_waves* w, const sfh_body_state* st,
                      int n_threads, sfh_info* out);

#ifdef __cplusplus
}
#endif
#endif
