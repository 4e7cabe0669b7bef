/*
 * adapter_harness.cpp — the call-site adapter of integration/stonefish_v1.6.0.patch (sf::HydroGPU, the file the patch ADDS to
 * Stonefish: integration/files/Library/src/core/HydroGPU.cpp) compiled here, unchanged, against the same stand-ins for the
 * Stonefish headers that the reference's own definitions are compiled against (ref_harness.cpp, ref_shim/shim.hpp), next to
 * those definitions. The tests run a scene of sf::SolidEntity / sf::Compound objects two ways from identical states:
 *     arm "gpu": HydroGPU::ComputeHydrodynamicForces(ocean, solids)        — the patched call site, libsfh.so underneath
 *     arm "ref": ent->ComputeHydrodynamicForces({true, true}, ocean)        — what Ocean::ApplyFluidForces does, the reference's text
 * and compare the members ApplyHydrodynamicForces reads. Members persist from step to step in both arms, so the reference's
 * "leave untouched" semantics are part of the comparison.
 *
 * TEST INFRASTRUCTURE ONLY: built into oracle/_ref/libadapter.so by `make -C oracle adapter` where the reference checkout is
 * mounted (it needs the extracted reference text); used by tests/test_integration_gpu.py.
 */
#include "ref_harness.cpp"

#define STONEFISH_B200_HYDRO
#define SFH_ADAPTER_HARNESS
#include "../integration/files/Library/include/core/HydroGPU.h"
#include "../integration/files/Library/src/core/HydroGPU.cpp"

namespace {

struct Entity {
    bool compound = false;
    std::unique_ptr<sf::SolidEntity> gpu, ref;       // plain solids
    std::unique_ptr<sf::Compound> cgpu, cref;        // compounds
    std::vector<std::unique_ptr<sf::SolidEntity>> parts_gpu, parts_ref;
    sf::SolidEntity* G() { return compound ? cgpu.get() : gpu.get(); }
    sf::SolidEntity* R() { return compound ? cref.get() : ref.get(); }
};

struct Scene {
    World world;
    std::vector<std::unique_ptr<sf::Mesh>> meshes;
    std::vector<Entity> entities;
    sf::HydroGPU hydro;
    explicit Scene(const refx_fluid* f) : world(f), hydro(0) {}
};

const double kSentinel = 123.0; // initial value of every output member: "never assigned" is observable in both arms

void init_outputs(sf::SolidEntity& s)
{
    s.Fb = s.Tb = s.Fdq = s.Tdq = s.Fdf = s.Tdf = sf::Vector3(kSentinel, kSentinel, kSentinel);
    s.Swet = s.Vsub = kSentinel;
}

void read_outputs(const sf::SolidEntity& s, sfh_info* out)
{
    std::memset(out, 0, sizeof(*out));
    put(out->Fb, s.Fb); put(out->Tb, s.Tb); put(out->Fdq, s.Fdq); put(out->Tdq, s.Tdq); put(out->Fdf, s.Fdf); put(out->Tdf, s.Tdf);
    out->Swet = s.Swet;
    out->Vsub = s.Vsub;
}

void set_motion(sf::SolidEntity& s, const sfh_body_state* st)
{
    s.T_CG = T(st->T_CG);
    s.T_C = T(st->T_C);
    s.linVel = V(st->v);
    s.angVel = V(st->omega);
    s.aabbMin = V(st->aabb_min);
    s.aabbMax = V(st->aabb_max);
}

} // namespace

extern "C" {

void* adp_new(const refx_fluid* f)
{
    try { return new Scene(f); }
    catch (const std::exception& e) { std::fprintf(stderr, "adp_new: %s\n", e.what()); return nullptr; }
}
void adp_free(void* h) { delete static_cast<Scene*>(h); }

static const sf::Mesh* add_mesh(Scene& sc, const orc_mesh* m)
{
    sc.meshes.emplace_back(new sf::Mesh());
    fill_mesh(*sc.meshes.back(), m);
    return sc.meshes.back().get();
}

/* a plain solid: mesh and everything of `st` that SolidEntity::ComputeHydrodynamicForces reads (T_O = st->T_O_basis, no origin) */
int adp_add_solid(void* h, const orc_mesh* mesh, const sfh_body_state* st)
{
    Scene& sc = *static_cast<Scene*>(h);
    const sf::Mesh* m = add_mesh(sc, mesh);
    Entity e;
    e.gpu.reset(new sf::SolidEntity());
    e.ref.reset(new sf::SolidEntity());
    for (sf::SolidEntity* s : {e.gpu.get(), e.ref.get()}) { fill_solid(*s, st, m); init_outputs(*s); }
    sc.entities.push_back(std::move(e));
    return (int)sc.entities.size() - 1;
}

/* a compound: as refx_compound (ref_harness.cpp) */
int adp_add_compound(void* h, const orc_mesh* const* meshes, const sfh_body_state* parts, const uint32_t* external, const double* origins,
                     const double* o2c, uint32_t n_parts, const sfh_body_state* comp, const double T_O[12])
{
    Scene& sc = *static_cast<Scene*>(h);
    Entity e;
    e.compound = true;
    e.cgpu.reset(new sf::Compound());
    e.cref.reset(new sf::Compound());
    for (int arm = 0; arm < 2; ++arm) {
        sf::Compound& c = arm ? *e.cref : *e.cgpu;
        fill_solid(c, comp, nullptr);
        c.T_O = T(T_O);
        init_outputs(c);
        for (uint32_t i = 0; i < n_parts; ++i) {
            const sf::Mesh* m = add_mesh(sc, meshes[i]);
            auto& store = arm ? e.parts_ref : e.parts_gpu;
            store.emplace_back(new sf::SolidEntity());
            fill_solid(*store.back(), &parts[i], m);
            store.back()->T_O2C = T(o2c + 12 * i);
            sf::CompoundPart p;
            p.solid = store.back().get();
            p.origin = T(origins + 12 * i);
            p.isExternal = external[i] != 0;
            p.alwaysVisible = false;
            c.parts.push_back(p);
        }
    }
    sc.entities.push_back(std::move(e));
    return (int)sc.entities.size() - 1;
}

/* new pose / velocities / AABB of entity `k` (both arms); compounds also take their T_O */
void adp_set_state(void* h, int k, const sfh_body_state* st, const double* T_O /* compounds, else NULL */)
{
    Scene& sc = *static_cast<Scene*>(h);
    Entity& e = sc.entities[(size_t)k];
    for (sf::SolidEntity* s : {e.G(), e.R()}) {
        set_motion(*s, st);
        if (T_O) s->T_O = T(T_O);
        else s->T_O = sf::Transform(sf::Matrix3(st->T_O_basis[0], st->T_O_basis[1], st->T_O_basis[2], st->T_O_basis[3], st->T_O_basis[4], st->T_O_basis[5],
                                                st->T_O_basis[6], st->T_O_basis[7], st->T_O_basis[8]), sf::Vector3(0, 0, 0));
    }
}

/* one recompute of every entity, both ways; out_gpu / out_ref: one record per entity (members after the step; position of the
 * ref arm's CheckBodyFluidPosition in both). Returns 0, or -1 with the message on stderr if the GPU arm threw. */
int adp_step(void* h, sfh_info* out_gpu, sfh_info* out_ref)
{
    Scene& sc = *static_cast<Scene*>(h);
    std::vector<sf::SolidEntity*> solids;
    for (Entity& e : sc.entities) solids.push_back(e.G());
    try { sc.hydro.ComputeHydrodynamicForces(&sc.world.ocean, solids); }
    catch (const std::exception& ex) { std::fprintf(stderr, "adp_step: %s\n", ex.what()); return -1; }
    sf::HydrodynamicsSettings settings;
    settings.dampingForces = true;      // Ocean.cpp:250-251
    settings.reallisticBuoyancy = true;
    for (size_t k = 0; k < sc.entities.size(); ++k) {
        Entity& e = sc.entities[k];
        e.R()->ComputeHydrodynamicForces(settings, &sc.world.ocean);
        read_outputs(*e.G(), &out_gpu[k]);
        read_outputs(*e.R(), &out_ref[k]);
        const sf::BodyFluidPosition bf = e.R()->CheckBodyFluidPosition(&sc.world.ocean);
        out_gpu[k].position = out_ref[k].position =
            bf == sf::BodyFluidPosition::OUTSIDE ? SFH_POS_OUTSIDE : (bf == sf::BodyFluidPosition::INSIDE ? SFH_POS_INSIDE : SFH_POS_CROSSING);
    }
    return 0;
}

} // extern "C"
