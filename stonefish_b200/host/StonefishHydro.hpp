// StonefishHydro.hpp — C++ host-side mirror of the reference's interface for the hydrodynamics path, over the C ABI
// (include/sfh.h, libsfh.so). Header-only; this is what a Stonefish build links instead of running
// SolidEntity::ComputeHydrodynamicForces on its thread pool (see INTEGRATION.md for the call-site patch).
//
// Names, argument meaning and behaviour follow the reference:
//   sf::HydrodynamicsSettings                      Library/include/entities/forcefields/Ocean.h:36-40
//   sf::Uniform / sf::Jet / sf::Pipe               Library/src/entities/forcefields/{Uniform,Jet,Pipe}.cpp
//   sf::Ocean::{AddVelocityField,EnableCurrents,DisableCurrents,hasWaves}   Ocean.cpp:139-211
//   sf::SolidEntity::ComputeHydrodynamicForces     SolidEntity.cpp:1799-1850  (per body; here: one batch per step)
//   sf::SolidEntity outputs Fb,Tb,Fdq,Tdq,Fdf,Tdf,Swet,Vsub               SolidEntity.h:430-437 ("HydrodynamicsInfo")
// Error behaviour: the reference's methods are void and cannot fail; here a failing CUDA call throws
// std::runtime_error with sfh_last_error() — there is no CPU fallback to fall back to.
//
// The math types are plain PODs so this header needs neither Bullet nor glm; FromBullet() adapters take anything
// with the btVector3 / btTransform accessor names (x(),y(),z(), getBasis()[r][c], getOrigin()).
#pragma once

#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "../../include/sfh.h"

namespace sfb {

typedef double Scalar;

struct Vector3 {
    Scalar x = 0, y = 0, z = 0;
    Vector3() {}
    Vector3(Scalar x_, Scalar y_, Scalar z_) : x(x_), y(y_), z(z_) {}
    template <class BtVec>
    static Vector3 FromBullet(const BtVec& v) { return Vector3(v.x(), v.y(), v.z()); }
    Vector3 operator+(const Vector3& o) const { return Vector3(x + o.x, y + o.y, z + o.z); }
};

struct Transform { // btTransform flattened: basis rows then origin
    Scalar m[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
    template <class BtTransform>
    static Transform FromBullet(const BtTransform& T)
    {
        Transform r;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) r.m[3 * i + j] = T.getBasis()[i][j];
        r.m[9] = T.getOrigin().x(); r.m[10] = T.getOrigin().y(); r.m[11] = T.getOrigin().z();
        return r;
    }
};

//! A structure holding the settings of the hydrodynamics computation (Ocean.h:36-40, spelling as in the reference).
struct HydrodynamicsSettings {
    bool dampingForces = true;
    bool reallisticBuoyancy = true;
};

enum class BodyFluidPosition { OUTSIDE = SFH_POS_OUTSIDE, INSIDE = SFH_POS_INSIDE, CROSSING_SURFACE = SFH_POS_CROSSING };

//! What SolidEntity stores after ComputeHydrodynamicForces (SolidEntity.h:430-437) plus the validation counters.
struct HydrodynamicsInfo {
    Vector3 Fb, Tb, Fdq, Tdq, Fdf, Tdf;
    Scalar Swet = 0, Vsub = 0;
    BodyFluidPosition position = BodyFluidPosition::OUTSIDE;
    uint32_t faces = 0, classCount[8] = {0, 0, 0, 0, 0, 0, 0, 0}, degenerate = 0, quadratic = 0, skin = 0;
    bool buoyancyAssigned = false, dampingAssigned = false; // did the last step assign Fb/Tb/Vsub resp. Fdq..Tdf (Appendix B.1)
    //! SolidEntity::ApplyHydrodynamicForces (SolidEntity.cpp:1958-1962): what goes to ApplyCentralForce / ApplyTorque
    Vector3 Force() const { return Fb + Fdq + Fdf; }
    Vector3 Torque() const { return Tb + Tdq + Tdf; }
};

// ------------------------------------------------------------------------------------------------ velocity fields
enum class VelocityFieldType { UNIFORM, JET, PIPE };

class VelocityField {
public:
    virtual ~VelocityField() {}
    void setEnabled(bool en) { enabled = en; }
    bool isEnabled() const { return enabled; }
    virtual VelocityFieldType getType() const = 0;
    virtual void Fill(sfh_current& c) const = 0;

private:
    bool enabled = true; // VelocityField.cpp:31
};

class Uniform : public VelocityField { // Uniform.cpp:31-57
public:
    explicit Uniform(const Vector3& velocity) : v(velocity) {}
    void setVelocity(const Vector3& x) { v = x; }
    Vector3 getVelocity() const { return v; }
    VelocityFieldType getType() const override { return VelocityFieldType::UNIFORM; }
    void Fill(sfh_current& c) const override { c.type = SFH_CURRENT_UNIFORM; c.a[0] = v.x; c.a[1] = v.y; c.a[2] = v.z; }

private:
    Vector3 v;
};

class Jet : public VelocityField { // Jet.cpp:31-75
public:
    Jet(const Vector3& point, const Vector3& direction, Scalar radius, Scalar outletVelocity) : c(point), r(radius), vout(outletVelocity)
    {
        // btVector3::normalized(): *this / length() == *this * (1 / sqrt(x*x + y*y + z*z))
        const Scalar s = Scalar(1.0) / std::sqrt(direction.x * direction.x + direction.y * direction.y + direction.z * direction.z);
        n = Vector3(direction.x * s, direction.y * s, direction.z * s);
    }
    //! From the members sf::Jet holds after construction (c, n already normalised, r, vout): no second normalisation, same bits
    static Jet* FromMembers(const Vector3& c, const Vector3& n, Scalar r, Scalar vout)
    {
        Jet* j = new Jet(c, Vector3(1, 0, 0), r, vout);
        j->n = n;
        return j;
    }
    void setOutletVelocity(Scalar x) { vout = x; }
    Scalar getOutletVelocity() const { return vout; }
    VelocityFieldType getType() const override { return VelocityFieldType::JET; }
    void Fill(sfh_current& k) const override
    {
        k.type = SFH_CURRENT_JET;
        k.a[0] = c.x; k.a[1] = c.y; k.a[2] = c.z;
        k.n[0] = n.x; k.n[1] = n.y; k.n[2] = n.z;
        k.r1 = r; k.vel = vout;
    }

private:
    Vector3 c, n;
    Scalar r, vout;
};

class Pipe : public VelocityField { // Pipe.cpp:31-78
public:
    Pipe(const Vector3& point1, const Vector3& point2, Scalar radius1, Scalar radius2, Scalar inletVelocity, Scalar exponent)
        : p1(point1), r1(radius1), r2(radius2), vin(inletVelocity), gamma(exponent)
    {
        const Vector3 d(point2.x - point1.x, point2.y - point1.y, point2.z - point1.z);
        l = std::sqrt(d.x * d.x + d.y * d.y + d.z * d.z);
        const Scalar s = Scalar(1.0) / l; // btVector3 operator/(v, s) == v * (1/s)
        n = Vector3(d.x * s, d.y * s, d.z * s);
    }
    //! From the members sf::Pipe holds after construction (p1, n, l, r1, r2, vin, gamma)
    static Pipe* FromMembers(const Vector3& p1, const Vector3& n, Scalar l, Scalar r1, Scalar r2, Scalar vin, Scalar gamma)
    {
        Pipe* p = new Pipe(p1, Vector3(p1.x + 1, p1.y, p1.z), r1, r2, vin, gamma);
        p->n = n;
        p->l = l;
        return p;
    }
    void setInletVelocity(Scalar x) { vin = x; }
    Scalar getInletVelocity() const { return vin; }
    VelocityFieldType getType() const override { return VelocityFieldType::PIPE; }
    void Fill(sfh_current& k) const override
    {
        k.type = SFH_CURRENT_PIPE;
        k.a[0] = p1.x; k.a[1] = p1.y; k.a[2] = p1.z;
        k.n[0] = n.x; k.n[1] = n.y; k.n[2] = n.z;
        k.r1 = r1; k.r2 = r2; k.l = l; k.vel = vin; k.gamma = gamma;
    }

private:
    Vector3 p1, n;
    Scalar l, r1, r2, vin, gamma;
};

// ------------------------------------------------------------------------------------------------ ocean
//! What the integrators read from sf::Ocean. Owns its velocity fields like the reference (Ocean.cpp:64-75).
class Ocean {
public:
    Ocean(Scalar liquidDensity, Scalar gravity) : density(liquidDensity), g(gravity) {}
    void AddVelocityField(VelocityField* field) { currents.emplace_back(field); }
    void EnableCurrents() { currentsEnabled = true; }
    void DisableCurrents() { currentsEnabled = false; }
    bool hasWaves() const { return fftData != nullptr || generatedWaves; }
    //! The wave field is the one HydroEngine::SimulateWaves left on the device: no fftData, no read-back, no upload.
    void useGeneratedWaves(bool on) { generatedWaves = on; }
    //! OpenGLRealOcean::fftData (float[fftSize*fftSize*4*layers]); nullptr = flat sea. Not owned. Call again when the
    //! render thread refreshes it (OpenGLRealOcean.cpp:333-356).
    void setWaveData(const float* data, size_t nFloats, int fftSize = 256, const float gridSizes[4] = nullptr)
    {
        fftData = data; fftFloats = nFloats; this->fftSize = fftSize;
        if (gridSizes) std::memcpy(grid, gridSizes, sizeof(grid));
        wavesDirty = true;
    }
    Scalar getLiquidDensity() const { return density; }
    Scalar getGravity() const { return g; }

    void Fill(sfh_ocean& o) const
    {
        if (currents.size() > SFH_MAX_CURRENTS) throw std::runtime_error("more than 64 velocity fields");
        std::memset(&o, 0, offsetof(sfh_ocean, currents));
        o.density = density; o.gravity_z = g;
        o.has_waves = hasWaves(); o.currents_enabled = currentsEnabled; o.n_currents = (uint32_t)currents.size();
        for (size_t i = 0; i < currents.size(); ++i) {
            std::memset(&o.currents[i], 0, sizeof(sfh_current));
            o.currents[i].enabled = currents[i]->isEnabled();
            currents[i]->Fill(o.currents[i]);
        }
    }
    bool TakeWavesDirty(sfh_waves& w)
    {
        if (!wavesDirty || !fftData) return false;
        w.fft_size = fftSize; std::memcpy(w.grid_sizes, grid, sizeof(grid)); w.data = fftData; w.n_floats = fftFloats;
        wavesDirty = false;
        return true;
    }

private:
    Scalar density, g;
    std::vector<std::unique_ptr<VelocityField>> currents;
    bool currentsEnabled = false;
    const float* fftData = nullptr;
    size_t fftFloats = 0;
    int fftSize = 256;
    float grid[4] = {893.f, 101.f, 21.f, 11.f}; // OpenGLOcean.cpp:70
    bool wavesDirty = false;
    bool generatedWaves = false;
};

//! OceanParams as OpenGLOcean / OpenGLRealOcean set them for a sea state (OpenGLOcean.cpp:62-72, OpenGLRealOcean.cpp:41-45).
inline sfh_wave_params WaveParamsFromSeaState(float state, int64_t seed = 1234)
{
    sfh_wave_params p;
    std::memset(&p, 0, sizeof(p));
    p.wind = state * 5.f + 2.f;
    p.A = 1.f;
    p.omega = 5.f * expf(-state) + 0.2f;
    p.km = 370.f; p.cm = 0.23f;
    p.propagate = 1; p.passes = 8;
    p.grid_sizes[0] = 893.f; p.grid_sizes[1] = 101.f; p.grid_sizes[2] = 21.f; p.grid_sizes[3] = 11.f;
    p.seed = seed;
    return p;
}

// ------------------------------------------------------------------------------------------------ atmosphere
//! What the aerodynamic integrator reads from sf::Atmosphere (Atmosphere.cpp:72-75,117-157): the gas density and the
//! wind fields, which the reference sums without consulting isEnabled(). Owns its fields like the reference.
class Atmosphere {
public:
    explicit Atmosphere(Scalar gasDensity) : density(gasDensity) {}
    void AddVelocityField(VelocityField* field) { wind.emplace_back(field); }
    Scalar getGasDensity() const { return density; }
    void Fill(sfh_atmosphere& a) const
    {
        if (wind.size() > SFH_MAX_CURRENTS) throw std::runtime_error("more than 64 velocity fields");
        a.density = density; a.n_winds = (uint32_t)wind.size(); a.reserved = 0;
        for (size_t i = 0; i < wind.size(); ++i) {
            std::memset(&a.winds[i], 0, sizeof(sfh_current));
            a.winds[i].enabled = wind[i]->isEnabled();
            wind[i]->Fill(a.winds[i]);
        }
    }

private:
    Scalar density;
    std::vector<std::unique_ptr<VelocityField>> wind;
};

//! sf::GeometryApproxType (SolidEntity.h:39)
enum class GeometryApproxType { AUTO, SPHERE, CYLINDER, ELLIPSOID };

// ------------------------------------------------------------------------------------------------ bodies
//! The members SolidEntity::ComputeHydrodynamicForces reads from a body (SolidEntity.cpp:1799-1850).
struct SolidBody {
    uint32_t meshId = 0;           // from HydroEngine::UploadMesh(getPhysicsMesh())
    bool buoyant = true;           // isBuoyant()
    bool dampingForces = true;     // false: evaluated like an internal part of a Compound (pSettings.dampingForces = false, Compound.cpp:420)
    Vector3 fdCd{1, 1, 1}, fdCf{0, 0, 0};
    Scalar volume = 0, surface = 0;
    Vector3 P_CB;                  // centre of buoyancy in the CG frame
    // per step:
    Transform T_CG, T_C, T_O;      // getCGTransform(), getCTransform(), getOTransform()
    Vector3 v, omega;              // getLinearVelocity(), getAngularVelocity()
    Vector3 aabbMin, aabbMax;      // getAABB()
    HydrodynamicsInfo info;        // outputs, with the reference's "leave untouched" semantics
    // aerodynamics (SolidEntity.h:423-427,436): proxy type / parameters, T_CG2H, and the outputs Fda, Tda
    GeometryApproxType fdApproxType = GeometryApproxType::AUTO;
    Scalar fdApproxParams[3] = {1, 1, 1};
    Transform T_CG2H;
    Vector3 Fda, Tda;
};

//! The members CableEntity::ComputeHydrodynamicForces reads (CableEntity.cpp:239-372) and the NodalForces it fills
//! (CableEntity.h:42-55). nodes / velocities = btSoftBody::m_nodes[i].m_x / m_v.
struct CableBody {
    Scalar radius = 0.01;
    bool buoyancy = true;            // phy_.buoyancy
    std::vector<Vector3> nodes, velocities;
    struct NodalForces { Vector3 Fb, Fdq, Fdf; };
    std::vector<NodalForces> nodalForces;   // output; ApplyHydrodynamicForces adds Fb + Fdf + Fdq to node i (:374-381)
};

//! One GPU worth of hydrodynamics: replaces the ThreadPool-driven loop of SimulationManager::SimulationTickCallback
//! (SimulationManager.cpp:1619-1655). Not copyable; one per GPU.
//! Array of trivially copyable records in page-locked host memory (sfh_host_alloc): sfh_step copies straight from / to it — no
//! staging copy inside the library, and the state copy of a large step overlaps the library's host-side validation. Contents are
//! NOT kept when it grows (the engine rewrites every record each step).
template <typename T>
class PinnedArray {
public:
    PinnedArray() = default;
    PinnedArray(const PinnedArray&) = delete;
    PinnedArray& operator=(const PinnedArray&) = delete;
    ~PinnedArray() { release(); }
    void bind(sfh_ctx* c) { ctx = c; }
    void release()
    {
        if (p) sfh_host_free(ctx, p);
        p = nullptr; n = cap = 0;
    }
    void resize(size_t count)
    {
        if (count > cap) {
            const size_t want = count > cap + cap / 2 ? count : cap + cap / 2;
            void* q = nullptr;
            if (sfh_host_alloc(ctx, want * sizeof(T), &q) != SFH_OK) throw std::runtime_error(std::string("sfh_host_alloc: ") + sfh_last_error(ctx));
            if (p) sfh_host_free(ctx, p);
            p = static_cast<T*>(q);
            cap = want;
        }
        n = count;
    }
    void clear() { n = 0; }
    size_t size() const { return n; }
    T* data() { return p; }
    const T* data() const { return p; }
    T& operator[](size_t i) { return p[i]; }
    const T& operator[](size_t i) const { return p[i]; }
    const T* begin() const { return p; }
    const T* end() const { return p + n; }

private:
    sfh_ctx* ctx = nullptr;
    T* p = nullptr;
    size_t n = 0, cap = 0;
};

class HydroEngine {
public:
    explicit HydroEngine(int cudaDevice = 0)
    {
        if (sfh_create(cudaDevice, &ctx) != SFH_OK) throw std::runtime_error(std::string("sfh_create: ") + sfh_last_error(nullptr));
        states.bind(ctx);
        infos.bind(ctx);
    }
    ~HydroEngine()
    {
        states.release(); // before the context they were allocated through goes away
        infos.release();
        sfh_destroy(ctx);
    }
    HydroEngine(const HydroEngine&) = delete;
    HydroEngine& operator=(const HydroEngine&) = delete;

    //! Upload a physics mesh once: pass Mesh::getVertexDataPointer(), getVertexSize(), getNumOfVertices(),
    //! getFaceDataPointer(), faces.size() (OpenGLDataStructs.h:120-262).
    uint32_t UploadMesh(const void* vertexData, size_t vertexSize, size_t numVertices, const void* faceData, size_t numFaces)
    {
        uint32_t id = 0;
        Check(sfh_mesh_upload(ctx, vertexData, vertexSize, (uint32_t)numVertices, static_cast<const uint32_t*>(faceData), (uint32_t)numFaces, 0, &id),
              "sfh_mesh_upload");
        return id;
    }

    //! settings = {true, true} and recompute = true reproduce Ocean::ApplyFluidForces (Ocean.cpp:246-254) for every
    //! body of the batch: CheckBodyFluidPosition, the Surface/Submerged integrators, CorrectHydrodynamicForces.
    //! Results land in bodies[i]->info; buoyancy/damping members the reference would leave untouched stay untouched.
    void ComputeHydrodynamicForces(const HydrodynamicsSettings& settings, Ocean& ocn, const std::vector<SolidBody*>& bodies)
    {
        sfh_waves w;
        if (ocn.TakeWavesDirty(w)) Check(sfh_waves_upload(ctx, &w), "sfh_waves_upload");
        ocn.Fill(ocean);
        states.resize(bodies.size());
        infos.resize(bodies.size());
        for (size_t i = 0; i < bodies.size(); ++i) {
            const SolidBody& b = *bodies[i];
            sfh_body_state& s = states[i];
            std::memset(&s, 0, sizeof(s));
            s.mesh_id = b.meshId;
            s.mode = SFH_POS_AUTO;
            s.flags = (uint32_t)SFH_F_DISPATCH | (settings.dampingForces && b.dampingForces ? (uint32_t)SFH_F_DAMPING : 0u) |
                      (settings.reallisticBuoyancy ? (uint32_t)SFH_F_BUOYANCY : 0u) | (b.buoyant ? (uint32_t)SFH_F_BUOYANT : 0u);
            std::memcpy(s.T_CG, b.T_CG.m, sizeof(s.T_CG));
            std::memcpy(s.T_C, b.T_C.m, sizeof(s.T_C));
            std::memcpy(s.T_O_basis, b.T_O.m, sizeof(s.T_O_basis));
            Put(s.v, b.v); Put(s.omega, b.omega); Put(s.aabb_min, b.aabbMin); Put(s.aabb_max, b.aabbMax);
            Put(s.Cd, b.fdCd); Put(s.Cf, b.fdCf); Put(s.P_CB, b.P_CB);
            s.volume = b.volume; s.surface = b.surface;
        }
        Check(sfh_step(ctx, &ocean, states.data(), (uint32_t)states.size(), infos.data()), "sfh_step");
        for (size_t i = 0; i < bodies.size(); ++i) Scatter(infos[i], bodies[i]->info);
    }

    //! Atmosphere::ApplyFluidForces' recompute for every body of the batch (Atmosphere.cpp:187-192):
    //! SolidEntity::ComputeAerodynamicForces = static integrator + CorrectAerodynamicForces (SolidEntity.cpp:1852-1956).
    //! Results land in bodies[i]->Fda / Tda. The caller filters on PhysicsMode::AERODYNAMIC like the reference (:1854).
    void ComputeAerodynamicForces(const Atmosphere& atm, const std::vector<SolidBody*>& bodies)
    {
        atm.Fill(atmosphere);
        states.resize(bodies.size());
        infos.resize(bodies.size());
        for (size_t i = 0; i < bodies.size(); ++i) {
            const SolidBody& b = *bodies[i];
            sfh_body_state& s = states[i];
            std::memset(&s, 0, sizeof(s));
            s.mesh_id = b.meshId;
            s.flags = (uint32_t)SFH_F_DISPATCH;
            std::memcpy(s.T_CG, b.T_CG.m, sizeof(s.T_CG));
            std::memcpy(s.T_C, b.T_C.m, sizeof(s.T_C));
            std::memcpy(s.T_O_basis, b.T_CG2H.m, sizeof(s.T_O_basis));
            Put(s.v, b.v); Put(s.omega, b.omega);
            if (b.fdApproxType == GeometryApproxType::CYLINDER) { // SolidEntity.cpp:1937-1941
                s.Cd[0] = 0.5; s.Cd[1] = 0.5; s.Cd[2] = 1.0;
                s.flags |= (uint32_t)SFH_F_AERO_CD;
            } else if (b.fdApproxType == GeometryApproxType::ELLIPSOID) { // :1943-1950, Cd /= maxCd multiplies by 1/maxCd
                const Scalar cx = Scalar(1) / b.fdApproxParams[0], cy = Scalar(1) / b.fdApproxParams[1], cz = Scalar(1) / b.fdApproxParams[2];
                const Scalar mx = cx > cy ? cx : cy, maxCd = mx > cz ? mx : cz, inv = Scalar(1) / maxCd;
                s.Cd[0] = cx * inv; s.Cd[1] = cy * inv; s.Cd[2] = cz * inv;
                s.flags |= (uint32_t)SFH_F_AERO_CD;
            }
        }
        Check(sfh_aero_step(ctx, &atmosphere, states.data(), (uint32_t)states.size(), infos.data()), "sfh_aero_step");
        for (size_t i = 0; i < bodies.size(); ++i) { bodies[i]->Fda = Get(infos[i].Fdq); bodies[i]->Tda = Get(infos[i].Tdq); }
    }

    //! CableEntity::ComputeHydrodynamicForces for every cable of the batch (the CABLE branch of Ocean::ApplyFluidForces,
    //! Ocean.cpp:257-266).
    void ComputeCableForces(const HydrodynamicsSettings& settings, Ocean& ocn, const std::vector<CableBody*>& cables)
    {
        sfh_waves w;
        if (ocn.TakeWavesDirty(w)) Check(sfh_waves_upload(ctx, &w), "sfh_waves_upload");
        ocn.Fill(ocean);
        std::vector<sfh_cable> cs(cables.size());
        std::vector<double> pos, vel;
        for (size_t i = 0; i < cables.size(); ++i) {
            const CableBody& c = *cables[i];
            if (c.nodes.size() != c.velocities.size()) throw std::runtime_error("cable: nodes and velocities differ in size");
            cs[i].first_node = (uint32_t)(pos.size() / 3); cs[i].n_nodes = (uint32_t)c.nodes.size(); cs[i].reserved = 0;
            cs[i].flags = (settings.dampingForces ? (uint32_t)SFH_F_DAMPING : 0u) | (settings.reallisticBuoyancy ? (uint32_t)SFH_F_BUOYANCY : 0u) |
                          (c.buoyancy ? (uint32_t)SFH_F_BUOYANT : 0u);
            cs[i].radius = c.radius;
            for (size_t k = 0; k < c.nodes.size(); ++k) {
                pos.insert(pos.end(), {c.nodes[k].x, c.nodes[k].y, c.nodes[k].z});
                vel.insert(vel.end(), {c.velocities[k].x, c.velocities[k].y, c.velocities[k].z});
            }
        }
        std::vector<sfh_node_forces> out(pos.size() / 3);
        Check(sfh_cable_step(ctx, &ocean, cs.data(), (uint32_t)cs.size(), pos.data(), vel.data(), (uint32_t)out.size(), out.data()), "sfh_cable_step");
        for (size_t i = 0; i < cables.size(); ++i) {
            CableBody& c = *cables[i];
            c.nodalForces.resize(c.nodes.size());
            for (size_t k = 0; k < c.nodes.size(); ++k) {
                const sfh_node_forces& f = out[cs[i].first_node + k];
                c.nodalForces[k].Fb = Get(f.Fb); c.nodalForces[k].Fdq = Get(f.Fdq); c.nodalForces[k].Fdf = Get(f.Fdf);
            }
        }
    }

    //! Wave-field producer on the device (OpenGLOcean::InitializeSimulation / Simulate without OpenGL): InitWaves once,
    //! SimulateWaves(t) whenever the sea should advance (the reference does it once per rendered frame with t += dt,
    //! OpenGLOcean.cpp:476-477). Pair with Ocean::useGeneratedWaves(true).
    void InitWaves(const sfh_wave_params& p) { Check(sfh_wavegen_init(ctx, &p), "sfh_wavegen_init"); }
    void SimulateWaves(float t) { Check(sfh_wavegen_simulate(ctx, t), "sfh_wavegen_simulate"); }
    //! = the PBO read-back of OpenGLRealOcean::Simulate (OpenGLRealOcean.cpp:333-356): fftData, 4*N*N*4 floats
    void DownloadWaves(float* fftData, size_t nFloats) { Check(sfh_wavegen_download(ctx, fftData, nFloats), "sfh_wavegen_download"); }

    //! Batched Ocean::GetDepth(glm::vec3) / Ocean::GetFluidVelocity(glm::vec3) (Ocean.cpp:149-168, 198-201).
    void GetDepth(Ocean& ocn, const float* xyz, size_t n, float* depth)
    {
        sfh_waves w;
        if (ocn.TakeWavesDirty(w)) Check(sfh_waves_upload(ctx, &w), "sfh_waves_upload");
        ocn.Fill(ocean);
        Check(sfh_query_depth(ctx, &ocean, xyz, (uint32_t)n, depth), "sfh_query_depth");
    }
    void GetFluidVelocity(Ocean& ocn, const float* xyz, size_t n, float* vel)
    {
        ocn.Fill(ocean);
        Check(sfh_query_fluid_velocity(ctx, &ocean, xyz, (uint32_t)n, vel), "sfh_query_fluid_velocity");
    }

    sfh_ctx* raw() { return ctx; }
    //! The raw records of the last ComputeHydrodynamicForces / ComputeAerodynamicForces (validation counters included)
    std::vector<sfh_info> LastInfos() const { return std::vector<sfh_info>(infos.begin(), infos.end()); }
    void ClearInfos() { infos.clear(); }
    //! Feed a wave field this engine's Ocean mirror did not hand over itself (HydroEngineMulti shares one Ocean between engines)
    void UploadWaves(const sfh_waves& w) { Check(sfh_waves_upload(ctx, &w), "sfh_waves_upload"); }

private:
    static void Put(double* d, const Vector3& v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }
    static Vector3 Get(const double* d) { return Vector3(d[0], d[1], d[2]); }
    void Check(int rc, const char* what)
    {
        if (rc != SFH_OK) throw std::runtime_error(std::string(what) + ": " + sfh_last_error(ctx));
    }
    static void Scatter(const sfh_info& r, HydrodynamicsInfo& o)
    {
        if (r.wrote_buoyancy) { o.Fb = Get(r.Fb); o.Tb = Get(r.Tb); }
        // Vsub follows the reference: written with the buoyancy by the Surface integrator, zeroed when OUTSIDE,
        // untouched when INSIDE (SolidEntity.cpp:1692,1819,1840)
        if (r.wrote_buoyancy && r.position != SFH_POS_INSIDE) o.Vsub = r.Vsub;
        if (r.wrote_damping) { o.Fdq = Get(r.Fdq); o.Tdq = Get(r.Tdq); o.Fdf = Get(r.Fdf); o.Tdf = Get(r.Tdf); }
        o.Swet = r.Swet;
        o.buoyancyAssigned = r.wrote_buoyancy != 0;
        o.dampingAssigned = r.wrote_damping != 0;
        o.position = (BodyFluidPosition)r.position;
        o.faces = r.n_faces; o.degenerate = r.n_degenerate; o.quadratic = r.n_quadratic; o.skin = r.n_skin;
        for (int k = 0; k < 8; ++k) o.classCount[k] = r.class_count[k];
    }

    sfh_ctx* ctx = nullptr;
    sfh_ocean ocean;
    sfh_atmosphere atmosphere;
    PinnedArray<sfh_body_state> states; // gathered / scattered in place: page-locked, see PinnedArray
    PinnedArray<sfh_info> infos;
};

// ------------------------------------------------------------------------------------------------ several GPUs
//! Order-independent 64-bit checksum of a set of result records (= stonefish_b200/sharding.py:checksum): the wrapping sum of a
//! per-record FNV-style hash. Equal sets of records give equal sums whatever their partition over devices.
inline uint64_t Checksum(const std::vector<sfh_info>& infos)
{
    uint64_t sum = 0;
    for (const sfh_info& r : infos) {
        uint64_t w[sizeof(sfh_info) / 8], h = 0xCBF29CE484222325ull;
        std::memcpy(w, &r, sizeof(w));
        for (uint64_t x : w) { h = (h ^ x) * 0x100000001B3ull; h ^= h >> 29; }
        sum += h;
    }
    return sum;
}

//! The bodies of one simulation spread over several GPUs of one node: ONE context and ONE host thread per device, bodies
//! block-partitioned (sizes differ by at most one), every shard stepped concurrently, the results back in the caller's bodies.
//! Bodies are independent: there is no traffic between the devices. Pass the same device twice ({0, 0}) to get
//! two contexts on one GPU (tests on single-GPU machines).
class HydroEngineMulti {
public:
    explicit HydroEngineMulti(const std::vector<int>& cudaDevices)
    {
        if (cudaDevices.empty()) throw std::runtime_error("HydroEngineMulti: no device given");
        for (int d : cudaDevices) engines.emplace_back(new HydroEngine(d));
        for (size_t r = 0; r < engines.size(); ++r) workers.emplace_back(new Worker());
    }
    size_t size() const { return engines.size(); }
    HydroEngine& engine(size_t r) { return *engines.at(r); }

    //! Contiguous block [first, last) of n units for shard `rank` of `world` (= sharding.py:shard_range)
    static std::pair<size_t, size_t> ShardRange(size_t n, size_t rank, size_t world)
    {
        const size_t base = n / world, rem = n % world, lo = rank * base + (rank < rem ? rank : rem);
        return {lo, lo + base + (rank < rem ? 1 : 0)};
    }

    //! Upload a mesh to EVERY device (a body may land on any of them); the id is valid on all.
    uint32_t UploadMesh(const void* vertexData, size_t vertexSize, size_t numVertices, const void* faceData, size_t numFaces)
    {
        uint32_t id = 0;
        for (size_t r = 0; r < engines.size(); ++r) {
            const uint32_t got = engines[r]->UploadMesh(vertexData, vertexSize, numVertices, faceData, numFaces);
            if (r && got != id) throw std::runtime_error("HydroEngineMulti: mesh ids diverged between devices (mix of UploadMesh and UploadMeshOn?)");
            id = got;
        }
        return id;
    }
    //! Upload a mesh to ONE device only — batches of independent environments whose bodies never change shard (instance i of n
    //! lives on ShardOf(i, n)). Ids are per device then: do not mix with UploadMesh.
    uint32_t UploadMeshOn(size_t shard, const void* vertexData, size_t vertexSize, size_t numVertices, const void* faceData, size_t numFaces)
    {
        return engines.at(shard)->UploadMesh(vertexData, vertexSize, numVertices, faceData, numFaces);
    }
    size_t ShardOf(size_t body, size_t nBodies) const
    {
        for (size_t r = 0; r < engines.size(); ++r)
            if (body < ShardRange(nBodies, r, engines.size()).second) return r;
        return engines.size() - 1;
    }

    //! HydroEngine::ComputeHydrodynamicForces for bodies[ShardRange(r)] on device r, all r at once.
    void ComputeHydrodynamicForces(const HydrodynamicsSettings& settings, Ocean& ocn, const std::vector<SolidBody*>& bodies)
    {
        sfh_waves w;
        if (ocn.TakeWavesDirty(w))
            for (auto& e : engines) e->UploadWaves(w);
        const size_t G = engines.size();
        for (size_t r = 0; r < G; ++r) {
            const auto range = ShardRange(bodies.size(), r, G);
            workers[r]->Run([this, r, range, &settings, &ocn, &bodies]() {
                const std::vector<SolidBody*> mine(bodies.begin() + (ptrdiff_t)range.first, bodies.begin() + (ptrdiff_t)range.second);
                if (!mine.empty()) engines[r]->ComputeHydrodynamicForces(settings, ocn, mine);
                else engines[r]->ClearInfos();
            });
        }
        std::string err;
        for (size_t r = 0; r < G; ++r) {
            const std::string e = workers[r]->Wait();
            if (!e.empty() && err.empty()) err = "device shard " + std::to_string(r) + ": " + e;
        }
        if (!err.empty()) throw std::runtime_error(err);
    }

    //! Checksum of the last step's records over all shards (what bench.py gathers over NCCL, here inside one process)
    uint64_t LastChecksum() const
    {
        uint64_t s = 0;
        for (const auto& e : engines) s += Checksum(e->LastInfos());
        return s;
    }

private:
    struct Worker { // one host thread per device, alive as long as the engine
        std::thread th;
        std::mutex m;
        std::condition_variable cv;
        std::function<void()> job;
        std::string error;
        bool busy = false, quit = false;
        Worker()
        {
            th = std::thread([this]() {
                std::unique_lock<std::mutex> lk(m);
                for (;;) {
                    cv.wait(lk, [this]() { return busy || quit; });
                    if (quit) return;
                    std::function<void()> j = std::move(job);
                    lk.unlock();
                    std::string e;
                    try { j(); } catch (const std::exception& ex) { e = ex.what(); }
                    lk.lock();
                    error = e;
                    busy = false;
                    cv.notify_all();
                }
            });
        }
        ~Worker()
        {
            { std::lock_guard<std::mutex> lk(m); quit = true; }
            cv.notify_all();
            th.join();
        }
        void Run(std::function<void()> j)
        {
            std::lock_guard<std::mutex> lk(m);
            job = std::move(j);
            busy = true;
            cv.notify_all();
        }
        std::string Wait()
        {
            std::unique_lock<std::mutex> lk(m);
            cv.wait(lk, [this]() { return !busy; });
            return error;
        }
    };
    std::vector<std::unique_ptr<HydroEngine>> engines;
    std::vector<std::unique_ptr<Worker>> workers;
};

// ------------------------------------------------------------------------------------------------ compounds
//! One part of a sf::Compound as submitted to HydroEngine: `body` carries the part's mesh, T_C = T_O * origin_i * O2C_i,
//! T_O = T_O * origin_i, fdCd/fdCf and isBuoyant(), plus the COMPOUND's T_CG, v, omega and AABB (so that every part
//! resolves to the compound's BodyFluidPosition). Internal parts are submitted with dampingForces = false.
struct CompoundPart {
    SolidBody* body;
    bool isExternal;
};

//! Host half of Compound::ComputeHydrodynamicForces (Compound.cpp:315-435): combines the parts' results in part order,
//! with the temporaries that persist across parts (a part whose buoyancy is not assigned re-adds the previous part's
//! Fbp/Tbp/Vsubp, as the reference does) and the compound-level analytic buoyancy when INSIDE. `out` keeps the members
//! the reference would leave untouched.
inline void SumCompoundParts(const std::vector<CompoundPart>& parts, const HydrodynamicsSettings& settings, bool compoundBuoyant,
                             Scalar volume, Scalar surface, const Vector3& P_CB, const Transform& T_CG, const Ocean& ocn,
                             HydrodynamicsInfo& out)
{
    if (parts.empty()) return;
    const BodyFluidPosition pos = parts.front().body->info.position;
    out.position = pos;
    auto add = [](Vector3& a, const Vector3& b) { a.x += b.x; a.y += b.y; a.z += b.z; };
    if (pos == BodyFluidPosition::OUTSIDE) { // :302-313
        out.Fb = out.Tb = out.Fdq = out.Tdq = out.Fdf = out.Tdf = Vector3();
        out.Swet = out.Vsub = 0;
        return;
    }
    if (pos == BodyFluidPosition::INSIDE) { // :315-363
        if (compoundBuoyant) {
            const Scalar s = -volume * ocn.getLiquidDensity();
            const Vector3 Fb(Scalar(0) * s, Scalar(0) * s, ocn.getGravity() * s);
            const Scalar* T = T_CG.m;
            const Vector3 d(T[0] * P_CB.x + T[1] * P_CB.y + T[2] * P_CB.z + T[9] - T[9], T[3] * P_CB.x + T[4] * P_CB.y + T[5] * P_CB.z + T[10] - T[10],
                            T[6] * P_CB.x + T[7] * P_CB.y + T[8] * P_CB.z + T[11] - T[11]);
            out.Fb = Fb;
            out.Tb = Vector3(d.y * Fb.z - d.z * Fb.y, d.z * Fb.x - d.x * Fb.z, d.x * Fb.y - d.y * Fb.x);
        }
        if (settings.dampingForces) {
            out.Fdq = out.Tdq = out.Fdf = out.Tdf = Vector3();
            for (const CompoundPart& p : parts)
                if (p.isExternal) { add(out.Fdq, p.body->info.Fdq); add(out.Tdq, p.body->info.Tdq); add(out.Fdf, p.body->info.Fdf); add(out.Tdf, p.body->info.Tdf); }
        }
        out.Swet = surface;
        out.Vsub = volume;
        return;
    }
    if (!(settings.reallisticBuoyancy || settings.dampingForces)) return; // :366
    if (settings.reallisticBuoyancy) out.Fb = out.Tb = Vector3();
    if (settings.dampingForces) out.Fdq = out.Tdq = out.Fdf = out.Tdf = Vector3();
    out.Swet = out.Vsub = 0;
    Vector3 Fbp, Tbp;
    Scalar Vsubp = 0;
    for (const CompoundPart& p : parts) {
        const HydrodynamicsInfo& r = p.body->info;
        const bool pRb = settings.reallisticBuoyancy && p.body->buoyant;
        if (!p.isExternal && !pRb) continue;
        if (r.buoyancyAssigned) { Fbp = r.Fb; Tbp = r.Tb; Vsubp = r.Vsub; }
        add(out.Fb, Fbp); add(out.Tb, Tbp);
        out.Vsub += Vsubp;
        if (p.isExternal) {
            // with dampingForces off the reference adds its zero-initialised temporaries (Compound.cpp:395-417): nothing changes.
            // r.Fdq.. would be stale here (Scatter leaves them untouched when the part's damping was not assigned).
            if (r.dampingAssigned) { add(out.Fdq, r.Fdq); add(out.Tdq, r.Tdq); add(out.Fdf, r.Fdf); add(out.Tdf, r.Tdf); }
            out.Swet += r.Swet;
        }
    }
}

//! Host half of Compound::ComputeAerodynamicForces (Compound.cpp:439-464): Fda/Tda of the external parts, each already
//! corrected with its own proxy by HydroEngine::ComputeAerodynamicForces, summed in part order. Parts are submitted with
//! the compound's T_CG, v, omega and T_C = T_O * origin_i * O2C_i.
inline void SumCompoundPartsAero(const std::vector<CompoundPart>& parts, Vector3& Fda, Vector3& Tda)
{
    Fda = Tda = Vector3();
    for (const CompoundPart& p : parts)
        if (p.isExternal) {
            Fda.x += p.body->Fda.x; Fda.y += p.body->Fda.y; Fda.z += p.body->Fda.z;
            Tda.x += p.body->Tda.x; Tda.y += p.body->Tda.y; Tda.z += p.body->Tda.z;
        }
}

} // namespace sfb
