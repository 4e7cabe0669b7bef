/*
 * shim.hpp — the minimum of declarations that lets the reference's OWN function definitions (pulled out of its source files by
 * oracle/ref_extract.py at build time, never stored here) compile in this image, which has neither GLM nor SDL2/OpenGL:
 *
 *   - Bullet: the reference's vendored LinearMath headers, used as they are (btScalar = double).
 *   - namespace glm: a stand-in for the handful of GLM 0.9.9 primitives those definitions use (vec3/vec4/mat4 arithmetic,
 *     dot, cross, length2, sqrt, mat4 * vec4, gtx/scalar_multiplication). THIS is the only restated arithmetic left under the
 *     reference's integrators: every expression, promotion, branch and accumulation order above it is the reference's own text
 *     compiled by g++. The stand-in follows GLM's scalar code path (no GLM_FORCE_INTRINSICS in the reference build):
 *     component-wise operators, `/` as a true division, dot(a,b) = (a.x*b.x + a.y*b.y) + a.z*b.z on the product vector,
 *     cross per detail/func_geometric.inl, mat4*vec4 per detail/type_mat4x4.inl ((m0*v0 + m1*v1) + (m2*v2 + m3*v3)).
 *   - namespace sf: class declarations with exactly the members the extracted definitions touch. Their DATA is filled by
 *     oracle/ref_harness.cpp from the C structs of include/sfh_types.h.
 *
 * TEST INFRASTRUCTURE ONLY (oracle/hydro_oracle.h).
 */
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <type_traits>
#include <vector>

#include "LinearMath/btMatrix3x3.h"
#include "LinearMath/btQuaternion.h"
#include "LinearMath/btTransform.h"
#include "LinearMath/btVector3.h"

typedef float GLfloat;
typedef unsigned int GLuint;
typedef int GLint;

// ---------------------------------------------------------------------------------------------------------------- glm
namespace glm {

struct vec4;
struct vec3 {
    float x, y, z;
    vec3() : x(0), y(0), z(0) {}
    explicit vec3(float s) : x(s), y(s), z(s) {}
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    explicit vec3(const vec4& v);
    static constexpr int length() { return 3; } // vec<3,T,Q>::length(): the component count
    float& operator[](int i) { return (&x)[i]; }
    const float& operator[](int i) const { return (&x)[i]; }
    vec3& operator+=(const vec3& v) { x += v.x; y += v.y; z += v.z; return *this; }
    vec3& operator-=(const vec3& v) { x -= v.x; y -= v.y; z -= v.z; return *this; }
    template <typename U>
    vec3& operator*=(U s) { x *= static_cast<float>(s); y *= static_cast<float>(s); z *= static_cast<float>(s); return *this; }
    template <typename U>
    vec3& operator/=(U s) { x /= static_cast<float>(s); y /= static_cast<float>(s); z /= static_cast<float>(s); return *this; }
};
struct vec4 {
    float x, y, z, w;
    vec4() : x(0), y(0), z(0), w(0) {}
    explicit vec4(float s) : x(s), y(s), z(s), w(s) {}
    vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    vec4(const vec3& v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
    float& operator[](int i) { return (&x)[i]; }
    const float& operator[](int i) const { return (&x)[i]; }
};
inline vec3::vec3(const vec4& v) : x(v.x), y(v.y), z(v.z) {}

inline vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
inline vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, const vec3& a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
// glm/gtx/scalar_multiplication.hpp (reached through glm/ext.hpp with GLM_ENABLE_EXPERIMENTAL, OpenGLDataStructs.h:30-32):
// arithmetic scalars that are NOT float multiply through `vec *= static_cast<float>(s)`
template <typename T>
typename std::enable_if<!std::is_same<T, float>::value && std::is_arithmetic<T>::value, vec3>::type operator*(vec3 lh, T const& s)
{
    return lh *= static_cast<float>(s);
}
template <typename T>
typename std::enable_if<!std::is_same<T, float>::value && std::is_arithmetic<T>::value, vec3>::type operator*(T const& s, vec3 rh)
{
    return rh *= static_cast<float>(s);
}

inline vec4 operator*(const vec4& a, const vec4& b) { return vec4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
inline vec4 operator+(const vec4& a, const vec4& b) { return vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

struct mat4 { // column-major
    vec4 c[4];
    mat4() {}
    mat4(float x0, float y0, float z0, float w0, float x1, float y1, float z1, float w1, float x2, float y2, float z2, float w2, float x3,
         float y3, float z3, float w3)
    {
        c[0] = vec4(x0, y0, z0, w0); c[1] = vec4(x1, y1, z1, w1); c[2] = vec4(x2, y2, z2, w2); c[3] = vec4(x3, y3, z3, w3);
    }
    vec4& operator[](int i) { return c[i]; }
    const vec4& operator[](int i) const { return c[i]; }
};
// detail/type_mat4x4.inl, operator*(mat<4,4>, vec<4>), scalar path
inline vec4 operator*(const mat4& m, const vec4& v)
{
    const vec4 Mov0(v[0]), Mov1(v[1]);
    const vec4 Mul0 = m[0] * Mov0, Mul1 = m[1] * Mov1;
    const vec4 Add0 = Mul0 + Mul1;
    const vec4 Mov2(v[2]), Mov3(v[3]);
    const vec4 Mul2 = m[2] * Mov2, Mul3 = m[3] * Mov3;
    const vec4 Add1 = Mul2 + Mul3;
    return Add0 + Add1;
}

inline float dot(const vec3& a, const vec3& b)
{
    const vec3 tmp(a * b);
    return tmp.x + tmp.y + tmp.z;
}
inline vec3 cross(const vec3& x, const vec3& y) { return vec3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); }
inline float length2(const vec3& v) { return dot(v, v); }
inline float sqrt(float x) { return std::sqrt(x); }
inline float length(const vec3& v) { return sqrt(dot(v, v)); }

} // namespace glm

// ----------------------------------------------------------------------------------------------------------------- sf
namespace sf {

typedef btScalar Scalar;
typedef btVector3 Vector3;
typedef btTransform Transform;
typedef btMatrix3x3 Matrix3;
typedef btQuaternion Quaternion;
inline Vector3 V0() { return Vector3(0, 0, 0); }

glm::mat4 glMatrixFromTransform(const Transform& T);
glm::vec3 glVectorFromVector(const Vector3& v);

struct Fluid {
    std::string name;
    Scalar density = 0, viscosity = 0, IOR = 0;
};

// debug sinks the integrators poke when DEBUG_HYDRO / DEBUG_WAVES are defined (they are not)
struct Renderable {
    std::shared_ptr<std::vector<glm::vec3>> points;
    std::shared_ptr<std::vector<glm::vec3>> getDataAsPoints() { return points; }
};

struct HydrodynamicsSettings {
    bool dampingForces;
    bool reallisticBuoyancy;
};
enum class PhysicsMode { DISABLED, SURFACE, FLOATING, SUBMERGED, AERODYNAMIC };
enum class BodyFluidPosition { INSIDE, OUTSIDE, CROSSING_SURFACE };
enum class GeometryApproxType { AUTO, SPHERE, CYLINDER, ELLIPSOID };
enum class VelocityFieldType { UNIFORM, JET, PIPE };
struct PhysicsSettings {
    PhysicsMode mode = PhysicsMode::SUBMERGED;
    bool collisions = true;
    bool buoyancy = true;
};

// graphics/OpenGLDataStructs.h: what the integrators read of a mesh
struct Face { GLuint vertexID[3]; };
struct Mesh {
    std::vector<Face> faces;
    std::vector<glm::vec3> pos;
    glm::vec3 getVertexPos(size_t faceID, unsigned short index) const { return pos[faces[faceID].vertexID[index]]; }
    // what HydroGPU (integration/files/Library/src/core/HydroGPU.cpp) uploads: OpenGLDataStructs.h:174-200
    void* getFaceDataPointer() const { return (void*)&faces[0].vertexID[0]; }
    size_t getNumOfVertices() const { return pos.size(); }
    void* getVertexDataPointer() const { return (void*)&pos[0].x; }
    size_t getVertexSize() const { return sizeof(glm::vec3); }
};

// SimulationApp::getApp()->getSimulationManager()->getGravity()
struct SimulationManager {
    Vector3 g;
    Vector3 getGravity() const { return g; }
};
struct SimulationApp {
    SimulationManager sm;
    SimulationManager* getSimulationManager() { return &sm; }
    static SimulationApp* getApp() { static SimulationApp app; return &app; }
};

struct VelocityFieldUBO;
class VelocityField {
public:
    virtual ~VelocityField() {}
    virtual Vector3 GetVelocityAtPoint(const Vector3& p) const = 0;
    void setEnabled(bool en) { enabled = en; }
    bool isEnabled() const { return enabled; }
    virtual VelocityFieldType getType() const = 0;
protected:
    bool enabled = true;
};
class Uniform : public VelocityField {
public:
    Uniform(const Vector3& velocity);
    void setVelocity(const Vector3& x);
    Vector3 getVelocity() const { return v; } // Uniform.cpp:59-62
    Vector3 GetVelocityAtPoint(const Vector3& p) const;
    VelocityFieldType getType() const { return VelocityFieldType::UNIFORM; }
private:
    Vector3 v;
};
class Jet : public VelocityField {
public:
    Jet(const Vector3& point, const Vector3& direction, Scalar radius, Scalar outletVelocity);
    void setOutletVelocity(Scalar x);
    Vector3 GetVelocityAtPoint(const Vector3& p) const;
    VelocityFieldType getType() const { return VelocityFieldType::JET; }
private:
    friend class HydroGPU; // as integration/stonefish_v1.6.0.patch adds to Jet.h
    Vector3 c, n;
    Scalar r, vout;
};
class Pipe : public VelocityField {
public:
    Pipe(const Vector3& point1, const Vector3& point2, Scalar radius1, Scalar radius2, Scalar inletVelocity, Scalar exponent);
    void setInletVelocity(Scalar v);
    Vector3 GetVelocityAtPoint(const Vector3& p) const;
    VelocityFieldType getType() const { return VelocityFieldType::PIPE; }
private:
    friend class HydroGPU; // as the patch adds to Pipe.h
    Vector3 p1, n;
    Scalar l, r1, r2, vin, gamma;
};

struct OceanParams { // graphics/OpenGLOcean.h:36-51
    unsigned int passes;
    unsigned int slopeVarianceSize;
    int fftSize;
    glm::vec4 gridSizes;
    float* spectrum12 = nullptr;
    float* spectrum34 = nullptr;
    bool propagate;
    float wind, omega, A, km, cm, t;
};
class OpenGLOcean {
public:
    virtual ~OpenGLOcean() {}
    virtual GLfloat ComputeWaveHeight(GLfloat x, GLfloat y) = 0;
    // the wave-field generator's host half (graphics/OpenGLOcean.cpp:857-1047)
    float sqr(float x);
    GLfloat* ComputeButterflyLookupTable(unsigned int size, unsigned int passes);
    int bitReverse(int i, int N);
    void computeWeight(int N, int k, float& Wr, float& Wi);
    void GenerateWavesSpectrum();
    void GetSpectrumSample(int i, int j, float lengthScale, float kMin, float* result);
    float spectrum(float kx, float ky, bool omnispectrum = false);
    float omega(float k);
    const OceanParams& getOceanParams() const { return params; } // accessor added by the patch (OpenGLOcean.h)
    OceanParams params;
};
class OpenGLRealOcean : public OpenGLOcean {
public:
    GLfloat ComputeInterpolatedWaveData(GLfloat x, GLfloat y, GLuint channel);
    GLfloat ComputeWaveHeight(GLfloat x, GLfloat y);
    const GLfloat* getFFTData() const { return fftData; }        // accessor added by the patch (OpenGLRealOcean.h)
    GLfloat* fftData = nullptr;
};

class Ocean {
public:
    float GetDepth(const glm::vec3& point);
    Scalar GetDepth(const Vector3& point);
    Vector3 GetFluidVelocity(const Vector3& point) const;
    glm::vec3 GetFluidVelocity(const glm::vec3& point) const;
    bool hasWaves() const { return waves; }
    Fluid getLiquid() const { return liquid; }
    OpenGLOcean* getOpenGLOcean() { return glOcean; } // Ocean.h:141
    // data (filled by the harness)
    bool waves = false;
    Fluid liquid;
    OpenGLOcean* glOcean = nullptr;
    std::vector<VelocityField*> currents;
    bool currentsEnabled = false;
};

class Atmosphere {
public:
    Vector3 GetFluidVelocity(const Vector3& point) const;
    glm::vec3 GetFluidVelocity(const glm::vec3& point) const;
    Fluid getGas() const { return gas; }
    Fluid gas;
    std::vector<VelocityField*> wind;
};

class SolidEntity {
public:
    virtual ~SolidEntity() {}
    // the definitions under test
    BodyFluidPosition CheckBodyFluidPosition(Ocean* ocn);
    static void CorrectHydrodynamicForces(Ocean* ocn, Vector3& _Fdq, Vector3& _Tdq, Vector3& _Fdf, Vector3& _Tdf, const Vector3& fdCd,
                                          const Vector3& fdCf, const Transform& T_O);
    static void ComputeHydrodynamicForcesSurface(const HydrodynamicsSettings& settings, const Mesh* mesh, Ocean* ocn, const Transform& T_CG,
                                                 const Transform& T_C, const Vector3& _v, const Vector3& _omega, Vector3& _Fb, Vector3& _Tb,
                                                 Vector3& _Fdq, Vector3& _Tdq, Vector3& _Fdf, Vector3& _Tdf, Scalar& _Swet, Scalar& _Vsub,
                                                 Renderable& debug);
    static void ComputeHydrodynamicForcesSubmerged(const Mesh* mesh, Ocean* ocn, const Transform& T_CG, const Transform& T_C, const Vector3& _v,
                                                   const Vector3& _omega, Vector3& _Fdq, Vector3& _Tdq, Vector3& _Fdf, Vector3& _Tdf);
    virtual void ComputeHydrodynamicForces(HydrodynamicsSettings settings, Ocean* ocn);
    virtual void ComputeAerodynamicForces(Atmosphere* atm);
    static void ComputeAerodynamicForces(const Mesh* mesh, Atmosphere* atm, const Transform& T_CG, const Transform& T_C, const Vector3& _v,
                                         const Vector3& _omega, Vector3& _Fda, Vector3& _Tda);
    void CorrectAerodynamicForces(Atmosphere* atm, Vector3& _Fda, Vector3& _Tda);

    // what those definitions call on `this` (state set by the harness)
    void getAABB(Vector3& min, Vector3& max) { min = aabbMin; max = aabbMax; }
    Vector3 getLinearVelocity() const { return linVel; }
    Vector3 getAngularVelocity() const { return angVel; }
    Transform getCGTransform() const { return T_CG; }
    Transform getCTransform() const { return T_C; }
    Transform getOTransform() const { return T_O; }
    Transform getO2CTransform() const { return T_O2C; }
    const Mesh* getPhysicsMesh() { return phyMesh; }
    bool isBuoyant() const { return phy.buoyancy; }
    PhysicsMode getPhysicsMode() const { return phy.mode; }
    void getHydrodynamicCoefficients(Vector3& Cd, Vector3& Cf) const { Cd = fdCd; Cf = fdCf; }

    Vector3 aabbMin, aabbMax, linVel, angVel;
    Transform T_CG, T_C, T_O, T_O2C, T_CG2H;
    const Mesh* phyMesh = nullptr;
    PhysicsSettings phy;
    Renderable submerged;
    Vector3 Fb, Tb, Fdq, Tdq, Fdf, Tdf, Fda, Tda;
    Scalar Swet = 0, Vsub = 0, volume = 0, surface = 0;
    Vector3 P_CB, fdCd, fdCf;
    GeometryApproxType fdApproxType = GeometryApproxType::AUTO;
    std::vector<Scalar> fdApproxParams;
};

struct CompoundPart {
    SolidEntity* solid;
    Transform origin;
    bool isExternal;
    bool alwaysVisible;
};
class Compound : public SolidEntity {
public:
    void ComputeHydrodynamicForces(HydrodynamicsSettings settings, Ocean* ocn);
    void ComputeAerodynamicForces(Atmosphere* atm);
    std::vector<CompoundPart> parts;
};

// btSoftBody as far as CableEntity reads it
struct SoftNode { Vector3 m_x, m_v; };
struct SoftBody { std::vector<SoftNode> m_nodes; };
class CableEntity {
public:
    struct NodalForces {
        Vector3 Fb, Fdq, Fdf;
        void clearForces() { Fb.setZero(); Fdq.setZero(); Fdf.setZero(); }
    };
    Scalar circularSegmentArea(Scalar h) const;
    void ComputeHydrodynamicForces(HydrodynamicsSettings settings, Ocean* ocn);
    PhysicsSettings phy_;
    Scalar radius_ = 0;
    SoftBody* cableBody_ = nullptr;
    std::vector<NodalForces> nodalForces_;
};

} // namespace sf
