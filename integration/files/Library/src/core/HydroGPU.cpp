/*
    This file is an ADDITION to Stonefish (integration/stonefish_v1.6.0.patch of the stonefish_b200 repository).
    Compiled only with -DSTONEFISH_B200_HYDRO (CMake option WITH_B200_HYDRO).
*/
#ifdef STONEFISH_B200_HYDRO

#ifndef SFH_ADAPTER_HARNESS // (the repository's test harness compiles this file against its stand-ins for the Stonefish headers)
#include "core/HydroGPU.h"
#include "core/SimulationApp.h"
#include "core/SimulationManager.h"
#include "entities/SolidEntity.h"
#include "entities/solids/Compound.h"
#include "entities/forcefields/Ocean.h"
#include "entities/forcefields/Uniform.h"
#include "entities/forcefields/Jet.h"
#include "entities/forcefields/Pipe.h"
#include "graphics/OpenGLRealOcean.h"
#endif
#include <map>
#include <StonefishHydro.hpp>

namespace sf
{

struct HydroGPU::Impl
{
    sfb::HydroEngine engine;
    std::map<const Mesh*, uint32_t> meshIds; // physics meshes are immutable after construction: uploaded once
    std::vector<sfb::SolidBody> bodies;      // one per solid, one per evaluated part of a compound
    struct Target { SolidEntity* ent; size_t first, count; };
    std::vector<Target> targets;
    explicit Impl(int dev) : engine(dev) {}

    uint32_t MeshId(const Mesh* m)
    {
        auto it = meshIds.find(m);
        if(it != meshIds.end()) return it->second;
        const uint32_t id = engine.UploadMesh(m->getVertexDataPointer(), m->getVertexSize(), m->getNumOfVertices(), m->getFaceDataPointer(), m->faces.size());
        meshIds[m] = id;
        return id;
    }
};

static sfb::HydrodynamicsInfo CurrentOutputs(const Vector3& Fb, const Vector3& Tb, const Vector3& Fdq, const Vector3& Tdq, const Vector3& Fdf, const Vector3& Tdf, Scalar Swet, Scalar Vsub)
{
    sfb::HydrodynamicsInfo i;
    i.Fb = sfb::Vector3::FromBullet(Fb); i.Tb = sfb::Vector3::FromBullet(Tb);
    i.Fdq = sfb::Vector3::FromBullet(Fdq); i.Tdq = sfb::Vector3::FromBullet(Tdq);
    i.Fdf = sfb::Vector3::FromBullet(Fdf); i.Tdf = sfb::Vector3::FromBullet(Tdf);
    i.Swet = Swet; i.Vsub = Vsub;
    return i;
}

static Vector3 ToBullet(const sfb::Vector3& v) { return Vector3(v.x, v.y, v.z); }

HydroGPU::HydroGPU(int cudaDevice) : impl(new Impl(cudaDevice)) {}

HydroGPU::~HydroGPU() { delete impl; }

void HydroGPU::Clear()
{
    for(auto it = impl->meshIds.begin(); it != impl->meshIds.end(); ++it) sfh_mesh_release(impl->engine.raw(), it->second);
    impl->meshIds.clear();
}

#ifndef SFH_ADAPTER_HARNESS
//The casts of Ocean::ApplyFluidForces (Ocean.cpp:223-244), for solids only
SolidEntity* HydroGPU::SolidFromCollisionObject(btCollisionObject* co)
{
    Entity* ent;
    if(btRigidBody* rb = dynamic_cast<btRigidBody*>(co))
    {
        if(rb->isStaticOrKinematicObject()) return nullptr;
        ent = static_cast<Entity*>(rb->getUserPointer());
    }
    else if(btMultiBodyLinkCollider* mbl = dynamic_cast<btMultiBodyLinkCollider*>(co))
    {
        if(mbl->isStaticOrKinematicObject()) return nullptr;
        ent = static_cast<Entity*>(mbl->getUserPointer());
    }
    else
        return nullptr;
    return ent->getType() == EntityType::SOLID ? (SolidEntity*)ent : nullptr;
}
#endif

void HydroGPU::ComputeHydrodynamicForces(Ocean* ocn, const std::vector<SolidEntity*>& solids)
{
    //Mirror of the ocean: density, gravity, currents (members as constructed: no second normalisation), waves
    sfb::Ocean hydroOcean(ocn->getLiquid().density, SimulationApp::getApp()->getSimulationManager()->getGravity().getZ());
    for(size_t i=0; i<ocn->currents.size(); ++i)
    {
        sfb::VelocityField* vf;
        switch(ocn->currents[i]->getType())
        {
            case VelocityFieldType::UNIFORM:
                vf = new sfb::Uniform(sfb::Vector3::FromBullet(((Uniform*)ocn->currents[i])->getVelocity()));
                break;
            case VelocityFieldType::JET:
            {
                Jet* j = (Jet*)ocn->currents[i];
                vf = sfb::Jet::FromMembers(sfb::Vector3::FromBullet(j->c), sfb::Vector3::FromBullet(j->n), j->r, j->vout);
            }
                break;
            default:
            {
                Pipe* p = (Pipe*)ocn->currents[i];
                vf = sfb::Pipe::FromMembers(sfb::Vector3::FromBullet(p->p1), sfb::Vector3::FromBullet(p->n), p->l, p->r1, p->r2, p->vin, p->gamma);
            }
                break;
        }
        vf->setEnabled(ocn->currents[i]->isEnabled());
        hydroOcean.AddVelocityField(vf);
    }
    if(ocn->currentsEnabled) hydroOcean.EnableCurrents();
    if(ocn->hasWaves())
    {
        //fftData is refreshed by the render thread under the hydrodynamics mutex, which the caller holds (SimulationManager.cpp:1621)
        OpenGLRealOcean* rocn = (OpenGLRealOcean*)ocn->getOpenGLOcean();
        const OceanParams& prm = rocn->getOceanParams();
        const float grid[4] = {prm.gridSizes.x, prm.gridSizes.y, prm.gridSizes.z, prm.gridSizes.w};
        hydroOcean.setWaveData(rocn->getFFTData(), (size_t)prm.fftSize * prm.fftSize * 4 * 4, prm.fftSize, grid);
    }

    //Gather: SolidEntity::ComputeHydrodynamicForces reads exactly these (SolidEntity.cpp:1799-1850, Compound.cpp:291-437)
    impl->bodies.clear();
    impl->targets.clear();
    for(size_t i=0; i<solids.size(); ++i)
    {
        SolidEntity* ent = solids[i];
        if(ent->phy.mode != PhysicsMode::FLOATING && ent->phy.mode != PhysicsMode::SUBMERGED) continue;
        Impl::Target t;
        t.ent = ent;
        t.first = impl->bodies.size();
        Vector3 aabbMin, aabbMax;
        ent->getAABB(aabbMin, aabbMax);
        sfb::SolidBody b;
        b.T_CG = sfb::Transform::FromBullet(ent->getCGTransform());
        b.v = sfb::Vector3::FromBullet(ent->getLinearVelocity());
        b.omega = sfb::Vector3::FromBullet(ent->getAngularVelocity());
        b.aabbMin = sfb::Vector3::FromBullet(aabbMin);
        b.aabbMax = sfb::Vector3::FromBullet(aabbMax);
        Compound* comp = dynamic_cast<Compound*>(ent);
        if(comp == nullptr)
        {
            b.meshId = impl->MeshId(ent->getPhysicsMesh());
            b.buoyant = ent->isBuoyant();
            b.fdCd = sfb::Vector3::FromBullet(ent->fdCd);
            b.fdCf = sfb::Vector3::FromBullet(ent->fdCf);
            b.volume = ent->volume;
            b.surface = ent->surface;
            b.P_CB = sfb::Vector3::FromBullet(ent->P_CB);
            b.T_C = sfb::Transform::FromBullet(ent->getCTransform());
            b.T_O = sfb::Transform::FromBullet(ent->getOTransform());
            b.info = CurrentOutputs(ent->Fb, ent->Tb, ent->Fdq, ent->Tdq, ent->Fdf, ent->Tdf, ent->Swet, ent->Vsub);
            impl->bodies.push_back(b);
        }
        else //One body per part that the compound would evaluate, with the compound's CG frame, velocities and AABB
        {
            for(size_t h=0; h<comp->parts.size(); ++h)
            {
                SolidEntity* part = comp->parts[h].solid;
                if(part->getPhysicsMode() != PhysicsMode::SUBMERGED && part->getPhysicsMode() != PhysicsMode::FLOATING) continue;
                sfb::SolidBody pb = b;
                Vector3 Cd, Cf;
                part->getHydrodynamicCoefficients(Cd, Cf);
                pb.meshId = impl->MeshId(part->getPhysicsMesh());
                pb.buoyant = part->isBuoyant();
                pb.dampingForces = comp->parts[h].isExternal;
                pb.fdCd = sfb::Vector3::FromBullet(Cd);
                pb.fdCf = sfb::Vector3::FromBullet(Cf);
                pb.T_C = sfb::Transform::FromBullet(comp->getOTransform() * comp->parts[h].origin * part->getO2CTransform());
                pb.T_O = sfb::Transform::FromBullet(comp->getOTransform() * comp->parts[h].origin);
                impl->bodies.push_back(pb);
            }
        }
        t.count = impl->bodies.size() - t.first;
        impl->targets.push_back(t);
    }
    if(impl->bodies.empty()) return;

    //One step on the device for the whole batch
    std::vector<sfb::SolidBody*> batch(impl->bodies.size());
    for(size_t i=0; i<batch.size(); ++i) batch[i] = &impl->bodies[i];
    sfb::HydrodynamicsSettings settings; //dampingForces = reallisticBuoyancy = true, as set by Ocean::ApplyFluidForces
    impl->engine.ComputeHydrodynamicForces(settings, hydroOcean, batch);

    //Scatter into the members ApplyHydrodynamicForces reads
    for(size_t i=0; i<impl->targets.size(); ++i)
    {
        const Impl::Target& t = impl->targets[i];
        SolidEntity* ent = t.ent;
        sfb::HydrodynamicsInfo out;
        Compound* comp = dynamic_cast<Compound*>(ent);
        if(comp == nullptr)
            out = impl->bodies[t.first].info;
        else
        {
            out = CurrentOutputs(ent->Fb, ent->Tb, ent->Fdq, ent->Tdq, ent->Fdf, ent->Tdf, ent->Swet, ent->Vsub);
            std::vector<sfb::CompoundPart> parts;
            for(size_t h=0; h<t.count; ++h)
            {
                sfb::CompoundPart cp;
                cp.body = &impl->bodies[t.first + h];
                cp.isExternal = cp.body->dampingForces;
                parts.push_back(cp);
            }
            if(parts.empty()) //No part takes part in the hydrodynamics: the compound still knows where it is (Compound.cpp:299-313)
                continue;
            sfb::SumCompoundParts(parts, settings, ent->isBuoyant(), ent->volume, ent->surface, sfb::Vector3::FromBullet(ent->P_CB),
                                  sfb::Transform::FromBullet(ent->getCGTransform()), hydroOcean, out);
        }
        ent->Fb = ToBullet(out.Fb); ent->Tb = ToBullet(out.Tb);
        ent->Fdq = ToBullet(out.Fdq); ent->Tdq = ToBullet(out.Tdq);
        ent->Fdf = ToBullet(out.Fdf); ent->Tdf = ToBullet(out.Tdf);
        ent->Swet = out.Swet; ent->Vsub = out.Vsub;
    }
}

}

#endif
