/*
 * sfh_types.h — plain-C data carriers of the Stonefish hydrodynamics hot path.
 *
 * These PODs are the exact argument lists of the reference call sites, flattened so they can
 * cross a C ABI (no glm, no Bullet, no torch types):
 *   - sfh_current      <- sf::Uniform / sf::Jet / sf::Pipe members
 *                         (Library/src/entities/forcefields/Uniform.cpp:31-57, Jet.cpp:31-75, Pipe.cpp:31-78)
 *   - sfh_ocean        <- what the integrators read from sf::Ocean
 *                         (Library/src/entities/forcefields/Ocean.cpp:149-201; liquid.density; gravity,
 *                          Library/src/core/SimulationManager.cpp:657-665)
 *   - sfh_waves        <- OpenGLRealOcean::fftData + OceanParams{fftSize,gridSizes}
 *                         (Library/src/graphics/OpenGLRealOcean.cpp:191-198,280-331; OpenGLOcean.cpp:63-70)
 *   - sfh_body_state   <- the arguments of SolidEntity::ComputeHydrodynamicForcesSurface/Submerged
 *                         (Library/include/entities/SolidEntity.h:135-137,152-153) plus the members
 *                         SolidEntity::ComputeHydrodynamicForces reads (SolidEntity.cpp:1799-1850)
 *   - sfh_info         <- "HydrodynamicsInfo": the outputs SolidEntity stores
 *                         (Library/include/entities/SolidEntity.h:430-437) + validation counters
 *
 * Conventions: NED world frame (z down, depth > 0 is under water). Transforms are Bullet
 * btTransform flattened as basis rows m_el[0..2] (9 doubles, row-major) followed by the origin (3 doubles).
 */
#ifndef SFH_TYPES_H
#define SFH_TYPES_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFH_MAX_CURRENTS 64u /* physics list is unbounded in the reference (Ocean.h:167); 64 is its UBO cap
                                (OpenGLDataStructs.h:92). The C ABI rejects more with SFH_ERR_INVALID. */

/* sf::VelocityFieldType */
enum { SFH_CURRENT_UNIFORM = 0, SFH_CURRENT_JET = 1, SFH_CURRENT_PIPE = 2 };

/* sf::BodyFluidPosition (SolidEntity.h:41) + AUTO = run CheckBodyFluidPosition from the AABB on the device */
enum { SFH_POS_OUTSIDE = 0, SFH_POS_INSIDE = 1, SFH_POS_CROSSING = 2, SFH_POS_AUTO = 3 };

/* sfh_body_state.flags */
enum {
    SFH_F_DAMPING   = 1u << 0, /* HydrodynamicsSettings::dampingForces     (Ocean.h:36-40) */
    SFH_F_BUOYANCY  = 1u << 1, /* HydrodynamicsSettings::reallisticBuoyancy                */
    SFH_F_BUOYANT   = 1u << 2, /* SolidEntity::isBuoyant()                                 */
    SFH_F_DISPATCH  = 1u << 3, /* 0: raw static integrator selected by `mode`
                                     (INSIDE -> ...Submerged, CROSSING -> ...Surface), outputs uncorrected.
                                  1: whole SolidEntity::ComputeHydrodynamicForces (SolidEntity.cpp:1799-1850):
                                     analytic INSIDE buoyancy, Swet=surface, reallisticBuoyancy&=isBuoyant,
                                     CorrectHydrodynamicForces with Cd/Cf/T_O. */
    /* --- read only by sfh_aero_step ------------------------------------------------------------------- */
    SFH_F_AERO_CD   = 1u << 4  /* CorrectAerodynamicForces (SolidEntity.cpp:1925-1956) scales by |Cd . Fdan|
                                  (GeometryApproxType CYLINDER / ELLIPSOID); clear: factor 1 (AUTO / SPHERE) */
};

typedef struct sfh_current {
    uint32_t type;    /* SFH_CURRENT_* */
    uint32_t enabled; /* VelocityField::isEnabled() */
    double a[3];      /* Uniform: v        | Jet: c (outlet centre) | Pipe: p1           */
    double n[3];      /* unused            | Jet: n (unit)          | Pipe: n (unit)     */
    double r1;        /*                   | Jet: r                 | Pipe: r1           */
    double r2;        /*                   |                        | Pipe: r2           */
    double l;         /*                   |                        | Pipe: l            */
    double vel;       /*                   | Jet: vout              | Pipe: vin          */
    double gamma;     /*                   |                        | Pipe: gamma        */
} sfh_current;

typedef struct sfh_ocean {
    double density;            /* Ocean::getLiquid().density                                  */
    double gravity_z;          /* SimulationManager::getGravity().getZ() (gravity = (0,0,g))  */
    uint32_t has_waves;        /* Ocean::hasWaves(); needs a prior sfh_waves_upload           */
    uint32_t currents_enabled; /* Ocean::currentsEnabled                                      */
    uint32_t n_currents;       /* <= SFH_MAX_CURRENTS, evaluated in list order                */
    uint32_t reserved;
    sfh_current currents[SFH_MAX_CURRENTS];
} sfh_ocean;

/* What the aerodynamic integrator reads from sf::Atmosphere: getGas().density and the wind list, which
 * Atmosphere::GetFluidVelocity sums WITHOUT looking at VelocityField::isEnabled() (Atmosphere.cpp:146-157):
 * winds[i].enabled is ignored. */
typedef struct sfh_atmosphere {
    double density;
    uint32_t n_winds; /* <= SFH_MAX_CURRENTS */
    uint32_t reserved;
    sfh_current winds[SFH_MAX_CURRENTS];
} sfh_atmosphere;

typedef struct sfh_waves {
    int32_t fft_size;     /* OceanParams::fftSize (256) */
    float grid_sizes[4];  /* OceanParams::gridSizes (893,101,21,11); only x,y are used */
    const float* data;    /* fftData: [layer][j][i][4] floats; the lookup touches indices
                             < (fft_size*fft_size + fft_size + 1)*4 (one row + one texel past layer 0,
                             see OpenGLRealOcean.cpp:289-311 with i0f==1.0f) */
    size_t n_floats;      /* floats readable at `data`; texels beyond it read as 0 */
} sfh_waves;

/* Parameters of the wave-field producer: OceanParams as OpenGLOcean / OpenGLRealOcean set them
 * (Library/src/graphics/OpenGLOcean.cpp:62-72, OpenGLRealOcean.cpp:41-45). */
typedef struct sfh_wave_params {
    float wind;          /* U10 [m/s]: OpenGLRealOcean: state * 5.f + 2.f                       */
    float omega;         /* inverse wave age: 5.f * expf(-state) + 0.2f                         */
    float A;             /* spectrum amplitude factor (1.f)                                     */
    float km, cm;        /* 370.f, 0.23f                                                        */
    uint32_t propagate;  /* 1: one-sided spectrum (kx < 0 removed, rest doubled)                 */
    uint32_t passes;     /* log2(fftSize): 8                                                    */
    float grid_sizes[4]; /* 893, 101, 21, 11 [m]                                                */
    uint32_t reserved;
    int64_t seed;        /* `static long seed = 1234` of OpenGLOcean::GetSpectrumSample (:1008) */
} sfh_wave_params;

typedef struct sfh_body_state {
    uint32_t mesh_id;
    uint32_t mode;  /* SFH_POS_* */
    uint32_t flags; /* SFH_F_*   */
    uint32_t reserved;
    double T_CG[12]; /* getCGTransform()                         */
    double T_C[12];  /* getCTransform(): mesh frame -> world     */
    double v[3];     /* getLinearVelocity()                      */
    double omega[3]; /* getAngularVelocity()                     */
    /* --- read only with SFH_POS_AUTO ---------------------------------------------------- */
    double aabb_min[3], aabb_max[3]; /* SolidEntity::getAABB (world)                        */
    /* --- read only with SFH_F_DISPATCH -------------------------------------------------- */
    double T_O_basis[9]; /* getOTransform().getBasis(), row-major                           */
    double Cd[3], Cf[3]; /* fdCd, fdCf                                                      */
    double volume, surface; /* SolidEntity::volume, ::surface                               */
    double P_CB[3];      /* centre of buoyancy in the CG frame                              */
} sfh_body_state;

/* Cables: what CableEntity::ComputeHydrodynamicForces reads (Library/src/entities/CableEntity.cpp:239-372).
 * A cable is a run of consecutive nodes in the step's node arrays (btSoftBody::m_nodes[i].m_x / m_v). */
typedef struct sfh_cable {
    uint32_t first_node; /* index of node 0 of this cable in the node arrays                                  */
    uint32_t n_nodes;    /* >= 2                                                                               */
    uint32_t flags;      /* SFH_F_DAMPING / SFH_F_BUOYANCY = settings, SFH_F_BUOYANT = phy_.buoyancy           */
    uint32_t reserved;
    double radius;       /* radius_                                                                            */
} sfh_cable;

/* CableEntity::NodalForces (CableEntity.h:42-55), one per node */
typedef struct sfh_node_forces {
    double Fb[3], Fdq[3], Fdf[3];
} sfh_node_forces;

typedef struct sfh_info {
    double Fb[3], Tb[3], Fdq[3], Tdq[3], Fdf[3], Tdf[3];
    double Swet, Vsub;
    uint32_t wrote_buoyancy; /* 0: the reference leaves _Fb/_Tb/_Vsub untouched (SolidEntity.cpp:1690-1708) */
    uint32_t wrote_damping;  /* 0: _Fdq.._Tdf untouched                                                      */
    uint32_t position;       /* resolved SFH_POS_*                                                           */
    uint32_t n_faces;        /* triangle-evals spent on this body this step                                   */
    uint32_t class_count[8]; /* faces per clipping class, bit2/1/0 = vertex 0/1/2 under water (depth>=0)      */
    uint32_t n_degenerate;   /* faces dropped by |fn|^2 < 1e-12f                                              */
    uint32_t n_quadratic;    /* faces that took the form-drag branch (vc.n < -1e-12f)                         */
    uint32_t n_skin;         /* faces that took the skin-friction branch (|vt|^2 > 1e-9f)                     */
    uint32_t reserved;
} sfh_info;

/* per-face debug code written by the *_debug entry points */
enum {
    SFH_FACE_CLASS_MASK = 7u,
    SFH_FACE_DEGENERATE = 1u << 3,
    SFH_FACE_QUADRATIC  = 1u << 4,
    SFH_FACE_SKIN       = 1u << 5
};

#ifdef __cplusplus
}
#endif
#endif /* SFH_TYPES_H */
