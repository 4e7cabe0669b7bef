/*
 * ref_harness.cpp — C entry points around the reference's OWN definitions of the hot path, compiled from the text that
 * oracle/ref_extract.py pulls out of /root/reference at build time (oracle/_ref/ref_extract.inc, a build intermediate that is
 * deleted after compilation and never committed) against oracle/ref_shim/shim.hpp.
 *
 * What this pins: every expression, float/double promotion, branch and accumulation order of
 *   glMatrixFromTransform, glVectorFromVector, Uniform/Jet/Pipe (constructors and GetVelocityAtPoint),
 *   OpenGLRealOcean::ComputeInterpolatedWaveData / ComputeWaveHeight, Ocean::GetDepth / GetFluidVelocity (both overloads),
 *   Atmosphere::GetFluidVelocity, SolidEntity::CheckBodyFluidPosition / CorrectHydrodynamicForces /
 *   ComputeHydrodynamicForcesSurface / ...Submerged / ComputeHydrodynamicForces / ComputeAerodynamicForces (both) /
 *   CorrectAerodynamicForces, Compound::ComputeHydrodynamicForces / ComputeAerodynamicForces,
 *   CableEntity::circularSegmentArea / ComputeHydrodynamicForces,
 *   OpenGLOcean::GenerateWavesSpectrum / spectrum / GetSpectrumSample / ComputeButterflyLookupTable (+ lrandom / frandom)
 * as g++ compiles the reference's text with the reference's flags. What it cannot pin: GLM's primitives, for which shim.hpp
 * stands in (GLM is not in this image).
 *
 * TEST INFRASTRUCTURE ONLY: built into oracle/_ref/librefsrc.so by `make -C oracle refsrc` where the reference checkout is
 * mounted; used by tests/golden/make_ref_golden.py to write tests/golden/ref_golden.json and by tests/test_oracle_refsource.py.
 */
#include "ref_shim/shim.hpp"

#include <atomic>
#include <chrono>
#include <thread>

#include "../include/sfh_types.h"
#include "hydro_oracle.h"

namespace sf {
#include "_ref/ref_extract.inc"
} // namespace sf

extern "C" {

/* a velocity field by its CONSTRUCTOR arguments (the constructors normalise: Jet.cpp:31-37, Pipe.cpp:31-40)
 * Uniform: a = velocity | Jet: a = point, b = direction, r1 = radius, vel = outletVelocity
 * Pipe: a = point1, b = point2, r1, r2, vel = inletVelocity, gamma = exponent */
typedef struct refx_field {
    uint32_t type, enabled;
    double a[3], b[3];
    double r1, r2, vel, gamma;
} refx_field;

typedef struct refx_fluid {
    double density, gravity_z;
    uint32_t has_waves, currents_enabled, n_fields, reserved;
    refx_field fields[SFH_MAX_CURRENTS];
    const float* fft_data;
    int32_t fft_size;
    float grid_sizes[4];
} refx_fluid;

} // extern "C"

namespace {

sf::Vector3 V(const double* p) { return sf::Vector3(p[0], p[1], p[2]); }
sf::Transform T(const double* m)
{
    return sf::Transform(sf::Matrix3(m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8]), sf::Vector3(m[9], m[10], m[11]));
}
void put(double* d, const sf::Vector3& v) { d[0] = v.x(); d[1] = v.y(); d[2] = v.z(); }

struct World { // an Ocean (or the wind list of an Atmosphere) built from a refx_fluid; owns its fields
    sf::Ocean ocean;
    sf::Atmosphere atmosphere;
    sf::OpenGLRealOcean gl;
    std::vector<std::unique_ptr<sf::VelocityField>> fields;
    explicit World(const refx_fluid* f)
    {
        for (uint32_t i = 0; i < f->n_fields; ++i) {
            const refx_field& d = f->fields[i];
            sf::VelocityField* vf;
            if (d.type == SFH_CURRENT_UNIFORM) vf = new sf::Uniform(V(d.a));
            else if (d.type == SFH_CURRENT_JET) vf = new sf::Jet(V(d.a), V(d.b), d.r1, d.vel);
            else vf = new sf::Pipe(V(d.a), V(d.b), d.r1, d.r2, d.vel, d.gamma);
            vf->setEnabled(d.enabled != 0);
            fields.emplace_back(vf);
            ocean.currents.push_back(vf);
            atmosphere.wind.push_back(vf);
        }
        ocean.currentsEnabled = f->currents_enabled != 0;
        ocean.liquid.density = f->density;
        atmosphere.gas.density = f->density;
        ocean.waves = f->has_waves != 0;
        gl.fftData = const_cast<float*>(f->fft_data);
        gl.params.fftSize = f->fft_size;
        gl.params.gridSizes = glm::vec4(f->grid_sizes[0], f->grid_sizes[1], f->grid_sizes[2], f->grid_sizes[3]);
        ocean.glOcean = &gl;
        sf::SimulationApp::getApp()->getSimulationManager()->g = sf::Vector3(0, 0, f->gravity_z);
    }
};

void fill_mesh(sf::Mesh& m, const orc_mesh* src)
{
    m.pos.resize(src->n_verts);
    for (uint32_t i = 0; i < src->n_verts; ++i) {
        const float* p = reinterpret_cast<const float*>(static_cast<const char*>(src->pos) + (size_t)i * src->stride_bytes);
        m.pos[i] = glm::vec3(p[0], p[1], p[2]);
    }
    m.faces.resize(src->n_faces);
    for (uint32_t i = 0; i < src->n_faces; ++i)
        for (int k = 0; k < 3; ++k) m.faces[i].vertexID[k] = src->faces[3 * (size_t)i + k];
}

void fill_solid(sf::SolidEntity& s, const sfh_body_state* st, const sf::Mesh* mesh)
{
    s.T_CG = T(st->T_CG);
    s.T_C = T(st->T_C);
    s.T_O = sf::Transform(sf::Matrix3(st->T_O_basis[0], st->T_O_basis[1], st->T_O_basis[2], st->T_O_basis[3], st->T_O_basis[4], st->T_O_basis[5],
                                      st->T_O_basis[6], st->T_O_basis[7], st->T_O_basis[8]),
                          sf::Vector3(0, 0, 0));
    s.T_CG2H = s.T_O; // the aerodynamic contract passes T_CG2H's basis in T_O_basis
    s.linVel = V(st->v);
    s.angVel = V(st->omega);
    s.aabbMin = V(st->aabb_min);
    s.aabbMax = V(st->aabb_max);
    s.fdCd = V(st->Cd);
    s.fdCf = V(st->Cf);
    s.volume = st->volume;
    s.surface = st->surface;
    s.P_CB = V(st->P_CB);
    s.phy.buoyancy = (st->flags & SFH_F_BUOYANT) != 0;
    s.phy.mode = sf::PhysicsMode::SUBMERGED;
    s.phyMesh = mesh;
}

} // namespace

extern "C" {

/* Ocean::GetDepth(glm::vec3) / Ocean::GetFluidVelocity(glm::vec3) / (Vector3) and the current fields on their own */
float refx_depth_f(const refx_fluid* f, const float p[3])
{
    World w(f);
    return w.ocean.GetDepth(glm::vec3(p[0], p[1], p[2]));
}
void refx_velocity_f(const refx_fluid* f, const float p[3], float out[3])
{
    World w(f);
    const glm::vec3 v = w.ocean.GetFluidVelocity(glm::vec3(p[0], p[1], p[2]));
    out[0] = v.x; out[1] = v.y; out[2] = v.z;
}
void refx_velocity_d(const refx_fluid* f, const double p[3], double out[3])
{
    World w(f);
    put(out, w.ocean.GetFluidVelocity(V(p)));
}
void refx_wind_f(const refx_fluid* f, const float p[3], float out[3])
{
    World w(f);
    const glm::vec3 v = w.atmosphere.GetFluidVelocity(glm::vec3(p[0], p[1], p[2]));
    out[0] = v.x; out[1] = v.y; out[2] = v.z;
}
/* normalised members after construction, for the host mirrors: Jet -> n; Pipe -> n, l */
void refx_field_axis(const refx_field* d, double n_out[3], double* l_out)
{
    // GetVelocityAtPoint is the only window on the private members: probe it is not needed — rebuild what the constructors do
    if (d->type == SFH_CURRENT_JET) { put(n_out, V(d->b).normalized()); *l_out = 0; }
    else { const sf::Scalar l = (V(d->b) - V(d->a)).length(); put(n_out, (V(d->b) - V(d->a)) / l); *l_out = l; }
}

/* One body through the sfh_step contract (include/sfh.h) with the reference's code:
 *   SFH_F_DISPATCH (mode must be SFH_POS_AUTO): SolidEntity::ComputeHydrodynamicForces, outputs that it leaves untouched are
 *       reported through wrote_buoyancy / wrote_damping = 0 (the members start as NaN so "untouched" is observable);
 *   otherwise the static integrator selected by mode. position: 0 OUTSIDE, 1 INSIDE, 2 CROSSING as in sfh_types.h. */
static void run_body(World& w, const sf::Mesh* mesh, const sfh_body_state* st, sfh_info* out)
{
    sf::SolidEntity s;
    fill_solid(s, st, mesh);
    const double nan = std::nan("");
    s.Fb = s.Tb = s.Fdq = s.Tdq = s.Fdf = s.Tdf = sf::Vector3(nan, nan, nan);
    s.Swet = s.Vsub = nan;
    sf::HydrodynamicsSettings settings;
    settings.dampingForces = (st->flags & SFH_F_DAMPING) != 0;
    settings.reallisticBuoyancy = (st->flags & SFH_F_BUOYANCY) != 0;
    std::memset(out, 0, sizeof(*out));
    // A position given by the caller (sfh contract) is imposed through the AABB, the only thing CheckBodyFluidPosition reads
    if (st->mode == SFH_POS_INSIDE) { s.aabbMin = sf::Vector3(0, 0, 1e6); s.aabbMax = sf::Vector3(1, 1, 2e6); }
    else if (st->mode == SFH_POS_OUTSIDE) { s.aabbMin = sf::Vector3(0, 0, -2e6); s.aabbMax = sf::Vector3(1, 1, -1e6); }
    else if (st->mode == SFH_POS_CROSSING) { s.aabbMin = sf::Vector3(0, 0, -1e6); s.aabbMax = sf::Vector3(1, 1, 1e6); }
    const sf::BodyFluidPosition bf = s.CheckBodyFluidPosition(&w.ocean);
    const uint32_t pos = bf == sf::BodyFluidPosition::OUTSIDE ? SFH_POS_OUTSIDE : (bf == sf::BodyFluidPosition::INSIDE ? SFH_POS_INSIDE : SFH_POS_CROSSING);
    if (st->flags & SFH_F_DISPATCH) {
        out->position = pos;
        s.ComputeHydrodynamicForces(settings, &w.ocean);
    } else if (pos == SFH_POS_INSIDE) {
        out->position = SFH_POS_INSIDE;
        sf::SolidEntity::ComputeHydrodynamicForcesSubmerged(s.phyMesh, &w.ocean, s.T_CG, s.T_C, s.linVel, s.angVel, s.Fdq, s.Tdq, s.Fdf, s.Tdf);
    } else if (pos == SFH_POS_CROSSING) {
        out->position = SFH_POS_CROSSING;
        sf::SolidEntity::ComputeHydrodynamicForcesSurface(settings, s.phyMesh, &w.ocean, s.T_CG, s.T_C, s.linVel, s.angVel, s.Fb, s.Tb, s.Fdq, s.Tdq,
                                                          s.Fdf, s.Tdf, s.Swet, s.Vsub, s.submerged);
    } else {
        out->position = SFH_POS_OUTSIDE;
        out->wrote_buoyancy = out->wrote_damping = 1;
        return;
    }
    out->wrote_buoyancy = !std::isnan(s.Fb.x());
    out->wrote_damping = !std::isnan(s.Fdq.x());
    if (out->wrote_buoyancy) { put(out->Fb, s.Fb); put(out->Tb, s.Tb); }
    if (out->wrote_damping) { put(out->Fdq, s.Fdq); put(out->Tdq, s.Tdq); put(out->Fdf, s.Fdf); put(out->Tdf, s.Tdf); }
    out->Swet = std::isnan(s.Swet) ? 0.0 : s.Swet;
    out->Vsub = std::isnan(s.Vsub) ? 0.0 : s.Vsub;
}

void refx_body(const orc_mesh* mesh_in, const refx_fluid* f, const sfh_body_state* st, sfh_info* out)
{
    World w(f);
    sf::Mesh mesh;
    if (mesh_in) fill_mesh(mesh, mesh_in);
    run_body(w, mesh_in ? &mesh : nullptr, st, out);
}

/* Whole step scheduled like the reference: one task per body on a pool of n_threads (SimulationManager.cpp:1641-1651,
 * utils/ThreadPool.hpp). Meshes are converted before the clock starts (the reference holds them resident). Returns the wall
 * seconds between the first task and the join. This is the CPU arm of bench.py (`--impl reference`, cpu_baseline.kind
 * "reference"): the reference's own code on the host cores. */
double refx_step(const orc_mesh* const* meshes, const refx_fluid* f, const sfh_body_state* states, uint32_t n, int n_threads, sfh_info* out)
{
    World w(f);
    std::vector<const orc_mesh*> uniq;
    std::vector<std::unique_ptr<sf::Mesh>> built;
    std::vector<const sf::Mesh*> of_body(n, nullptr);
    for (uint32_t i = 0; i < n; ++i) {
        if (!meshes[i]) continue;
        size_t k = 0;
        while (k < uniq.size() && uniq[k] != meshes[i]) ++k;
        if (k == uniq.size()) {
            uniq.push_back(meshes[i]);
            built.emplace_back(new sf::Mesh());
            fill_mesh(*built.back(), meshes[i]);
        }
        of_body[i] = built[k].get();
    }
    if (n_threads < 1) n_threads = 1;
    std::atomic<uint32_t> next(0);
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; ++t)
        pool.emplace_back([&]() {
            for (uint32_t i = next.fetch_add(1); i < n; i = next.fetch_add(1)) run_body(w, of_body[i], &states[i], &out[i]);
        });
    for (auto& th : pool) th.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

/* SolidEntity::ComputeAerodynamicForces: the static integrator, or (dispatch != 0) the member = static + CorrectAerodynamicForces
 * with the proxy given as in the reference: approx_type (GeometryApproxType) and fdApproxParams; T_CG2H basis = st->T_O_basis */
void refx_aero(const orc_mesh* mesh_in, const refx_fluid* atm, const sfh_body_state* st, int dispatch, int approx_type, const double params[3],
               double Fda[3], double Tda[3])
{
    World w(atm);
    sf::Mesh mesh;
    if (mesh_in) fill_mesh(mesh, mesh_in);
    sf::SolidEntity s;
    fill_solid(s, st, mesh_in ? &mesh : nullptr);
    s.phy.mode = sf::PhysicsMode::AERODYNAMIC;
    s.fdApproxType = (sf::GeometryApproxType)approx_type;
    s.fdApproxParams = {params[0], params[1], params[2]};
    if (dispatch) s.ComputeAerodynamicForces(&w.atmosphere);
    else sf::SolidEntity::ComputeAerodynamicForces(s.phyMesh, &w.atmosphere, s.T_CG, s.T_C, s.linVel, s.angVel, s.Fda, s.Tda);
    put(Fda, s.Fda);
    put(Tda, s.Tda);
}

/* Compound::ComputeHydrodynamicForces / ComputeAerodynamicForces. The compound: `comp` (T_CG, v, omega, aabb, volume, surface,
 * P_CB, flags: settings + SFH_F_BUOYANT) and its own T_O (12 doubles). Part i: mesh, origin_i and O2C_i (12 doubles each),
 * external_i, and in parts[i]: Cd, Cf, flags & SFH_F_BUOYANT; for the aerodynamic variant approx_type[i], params[3*i..] and the
 * T_CG2H basis in parts[i].T_O_basis. Outputs as refx_body (hydro) or Fdq/Tdq = Fda/Tda (aero). */
void refx_compound(const orc_mesh* const* meshes, const sfh_body_state* parts, const uint32_t* external, const double* origins,
                   const double* o2c, uint32_t n_parts, const refx_fluid* f, const sfh_body_state* comp, const double T_O[12], int aero,
                   const int* approx_type, const double* params, sfh_info* out)
{
    World w(f);
    std::vector<sf::Mesh> ms(n_parts);
    std::vector<sf::SolidEntity> solids(n_parts);
    sf::Compound c;
    fill_solid(c, comp, nullptr);
    c.T_O = T(T_O);
    for (uint32_t i = 0; i < n_parts; ++i) {
        fill_mesh(ms[i], meshes[i]);
        fill_solid(solids[i], &parts[i], &ms[i]);
        solids[i].T_O2C = T(o2c + 12 * i);
        if (aero) {
            solids[i].fdApproxType = (sf::GeometryApproxType)approx_type[i];
            solids[i].fdApproxParams = {params[3 * i], params[3 * i + 1], params[3 * i + 2]};
        }
        sf::CompoundPart p;
        p.solid = &solids[i];
        p.origin = T(origins + 12 * i);
        p.isExternal = external[i] != 0;
        p.alwaysVisible = false;
        c.parts.push_back(p);
    }
    const double nan = std::nan("");
    c.Fb = c.Tb = c.Fdq = c.Tdq = c.Fdf = c.Tdf = sf::Vector3(nan, nan, nan);
    c.Swet = c.Vsub = nan;
    std::memset(out, 0, sizeof(*out));
    if (aero) {
        c.phy.mode = sf::PhysicsMode::AERODYNAMIC;
        c.ComputeAerodynamicForces(&w.atmosphere);
        put(out->Fdq, c.Fda);
        put(out->Tdq, c.Tda);
        out->wrote_damping = 1;
        return;
    }
    sf::HydrodynamicsSettings settings;
    settings.dampingForces = (comp->flags & SFH_F_DAMPING) != 0;
    settings.reallisticBuoyancy = (comp->flags & SFH_F_BUOYANCY) != 0;
    const sf::BodyFluidPosition bf = c.CheckBodyFluidPosition(&w.ocean);
    out->position = bf == sf::BodyFluidPosition::OUTSIDE ? SFH_POS_OUTSIDE : (bf == sf::BodyFluidPosition::INSIDE ? SFH_POS_INSIDE : SFH_POS_CROSSING);
    c.ComputeHydrodynamicForces(settings, &w.ocean);
    out->wrote_buoyancy = !std::isnan(c.Fb.x());
    out->wrote_damping = !std::isnan(c.Fdq.x());
    if (out->wrote_buoyancy) { put(out->Fb, c.Fb); put(out->Tb, c.Tb); }
    if (out->wrote_damping) { put(out->Fdq, c.Fdq); put(out->Tdq, c.Tdq); put(out->Fdf, c.Fdf); put(out->Tdf, c.Tdf); }
    out->Swet = std::isnan(c.Swet) ? 0.0 : c.Swet;
    out->Vsub = std::isnan(c.Vsub) ? 0.0 : c.Vsub;
}

/* CableEntity::ComputeHydrodynamicForces: out9 = Fb, Fdq, Fdf per node */
void refx_cable(const refx_fluid* f, uint32_t n_nodes, double radius, uint32_t flags, const double* pos, const double* vel, double* out9)
{
    World w(f);
    sf::SoftBody body;
    body.m_nodes.resize(n_nodes);
    for (uint32_t i = 0; i < n_nodes; ++i) { body.m_nodes[i].m_x = V(pos + 3 * i); body.m_nodes[i].m_v = V(vel + 3 * i); }
    sf::CableEntity c;
    c.cableBody_ = &body;
    c.radius_ = radius;
    c.phy_.mode = sf::PhysicsMode::SUBMERGED;
    c.phy_.buoyancy = (flags & SFH_F_BUOYANT) != 0;
    c.nodalForces_.resize(n_nodes);
    sf::HydrodynamicsSettings settings;
    settings.dampingForces = (flags & SFH_F_DAMPING) != 0;
    settings.reallisticBuoyancy = (flags & SFH_F_BUOYANCY) != 0;
    c.ComputeHydrodynamicForces(settings, &w.ocean);
    for (uint32_t i = 0; i < n_nodes; ++i) { put(out9 + 9 * i, c.nodalForces_[i].Fb); put(out9 + 9 * i + 3, c.nodalForces_[i].Fdq); put(out9 + 9 * i + 6, c.nodalForces_[i].Fdf); }
}

/* OpenGLOcean::GenerateWavesSpectrum / ComputeButterflyLookupTable with the parameters of sfh_wave_params. The reference keeps its
 * random seed in a function-local static (`static long seed = 1234`, OpenGLOcean.cpp:1008) that cannot be reset from outside:
 * call this ONCE per process for the spectrum (the butterfly table has no state). */
void refx_wave_tables(const sfh_wave_params* p, float* spectrum12, float* spectrum34, float* butterfly)
{
    sf::OpenGLRealOcean o;
    o.params.passes = p->passes;
    o.params.fftSize = 1 << p->passes;
    o.params.gridSizes = glm::vec4(p->grid_sizes[0], p->grid_sizes[1], p->grid_sizes[2], p->grid_sizes[3]);
    o.params.propagate = p->propagate != 0;
    o.params.wind = p->wind; o.params.omega = p->omega; o.params.A = p->A; o.params.km = p->km; o.params.cm = p->cm; o.params.t = 0.f;
    const size_t n = (size_t)o.params.fftSize * o.params.fftSize * 4;
    if (spectrum12 && spectrum34) {
        o.GenerateWavesSpectrum();
        std::memcpy(spectrum12, o.params.spectrum12, n * sizeof(float));
        std::memcpy(spectrum34, o.params.spectrum34, n * sizeof(float));
        delete[] o.params.spectrum12;
        delete[] o.params.spectrum34;
    }
    if (butterfly) {
        GLfloat* t = o.ComputeButterflyLookupTable((unsigned)o.params.fftSize, p->passes);
        std::memcpy(butterfly, t, (size_t)o.params.fftSize * p->passes * 4 * sizeof(float));
        delete[] t;
    }
}

void refx_gl_matrix(const double Tm[12], float out[16])
{
    const glm::mat4 M = sf::glMatrixFromTransform(T(Tm));
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) out[4 * c + r] = M[c][r];
}

} // extern "C"
