/*
 * hydro_oracle.cpp — CPU restatement of Stonefish's per-triangle hydrodynamic force integration.
 *
 * TEST INFRASTRUCTURE ONLY (see hydro_oracle.h). Parity: pinned bit for bit against the reference's own definitions compiled from
 * its source text (oracle/ref_harness.cpp, tests/test_oracle_refsource.py); only the GLM primitives under them are a stand-in.
 *
 * Every function cites the reference lines it follows (paths relative to the Stonefish v1.6.0 tree).
 * Arithmetic rules:
 *   - fp32 chain: GLM 0.9.9 scalar semantics, one IEEE RN rounding per operator, nothing fused, true
 *     divisions. Build with -O3 -ffp-contract=off and without -march=native / -ffast-math
 *     (the reference builds with -O3 -DNDEBUG for baseline x86-64: CMakeLists.txt:15,18).
 *   - fp64 pieces: Bullet LinearMath (BT_USE_DOUBLE_PRECISION, scalar code path) operator order.
 */
#include "hydro_oracle.h"

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

namespace {

/* ---------------------------------------------------------------- glm::vec3 restated (Appendix A.1) */
struct V3 {
    float x, y, z;
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(float s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
/* glm::dot(vec3): tmp = a*b; tmp.x + tmp.y + tmp.z */
inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float length2(V3 a) { return dot(a, a); }
inline float length(V3 a) { return std::sqrt(dot(a, a)); }
/* glm::cross */
inline V3 cross(V3 x, V3 y) { return {x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y}; }

/* glm::mat4, column-major as built by glMatrixFromTransform (OpenGLDataStructs.cpp:148-156) */
struct M4 {
    float m[16];
};

/* btTransform::getOpenGLMatrix (3rdparty/LinearMath/btTransform.h:136-143, btMatrix3x3.h:404-416), each
 * element narrowed with (GLfloat) */
M4 gl_matrix(const double T[12])
{
    M4 M;
    M.m[0] = (float)T[0];  M.m[1] = (float)T[3];  M.m[2] = (float)T[6];   M.m[3] = (float)0.0;
    M.m[4] = (float)T[1];  M.m[5] = (float)T[4];  M.m[6] = (float)T[7];   M.m[7] = (float)0.0;
    M.m[8] = (float)T[2];  M.m[9] = (float)T[5];  M.m[10] = (float)T[8];  M.m[11] = (float)0.0;
    M.m[12] = (float)T[9]; M.m[13] = (float)T[10]; M.m[14] = (float)T[11]; M.m[15] = (float)1.0;
    return M;
}

/* glm::vec3(TC * glm::vec4(p, 1.f)): (m[0]*v.x + m[1]*v.y) + (m[2]*v.z + m[3]*v.w), w dropped */
inline V3 xform(const M4& M, V3 p)
{
    const float* m = M.m;
    V3 r;
    r.x = (m[0] * p.x + m[4] * p.y) + (m[8] * p.z + m[12] * 1.f);
    r.y = (m[1] * p.x + m[5] * p.y) + (m[9] * p.z + m[13] * 1.f);
    r.z = (m[2] * p.x + m[6] * p.y) + (m[10] * p.z + m[14] * 1.f);
    return r;
}

/* ---------------------------------------------------------------- btVector3 (double) restated */
struct D3 {
    double x, y, z;
};
inline D3 operator-(D3 a, D3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline D3 operator+(D3 a, D3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline D3 operator*(D3 v, double s) { return {v.x * s, v.y * s, v.z * s}; } /* btVector3.h:820-832 */
inline D3 operator*(double s, D3 v) { return v * s; }                        /* btVector3.h:836-840 */
inline double ddot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; } /* btVector3.h:229-247 */
inline D3 dcross(D3 a, D3 v)                                                 /* btVector3.h:380-416 */
{
    return {a.y * v.z - a.z * v.y, a.z * v.x - a.x * v.z, a.x * v.y - a.y * v.x};
}
inline double dnorm(D3 a) { return std::sqrt(ddot(a, a)); }

inline const float* vertex_ptr(const orc_mesh* mesh, uint32_t id)
{
    return reinterpret_cast<const float*>(static_cast<const char*>(mesh->pos) + (size_t)id * mesh->stride_bytes);
}
/* Mesh::getVertexPos(faceID, index) (OpenGLDataStructs.h:207-210, 259-262) */
inline V3 vertex_pos(const orc_mesh* mesh, size_t face, int k)
{
    const float* p = vertex_ptr(mesh, mesh->faces[3 * face + k]);
    return {p[0], p[1], p[2]};
}

inline float wave_texel(const sfh_waves* w, long idx)
{
    if (idx < 0 || (size_t)idx >= w->n_floats) return 0.f; /* outside what the caller handed over */
    return w->data[idx];
}

} // namespace

extern "C" {

/* OpenGLRealOcean::ComputeInterpolatedWaveData (OpenGLRealOcean.cpp:280-319) */
float orc_wave_bilinear(const sfh_waves* w, float x, float y, uint32_t channel)
{
    const int N = w->fft_size;
    float tmp;
    float i0f = modff(x - 0.5f / (float)N, &tmp);
    float j0f = modff(y - 0.5f / (float)N, &tmp);
    if (i0f < 0.f) i0f = 1.f - fabsf(i0f);
    if (j0f < 0.f) j0f = 1.f - fabsf(j0f);
    int i0 = (int)truncf(i0f * (float)N);
    int j0 = (int)truncf(j0f * (float)N);

    float i1f = modff(x + 0.5f / (float)N, &tmp);
    float j1f = modff(y + 0.5f / (float)N, &tmp);
    if (i1f < 0.f) i1f = 1.f - fabsf(i1f);
    if (j1f < 0.f) j1f = 1.f - fabsf(j1f);
    int i1 = (int)truncf(i1f * (float)N);
    int j1 = (int)truncf(j1f * (float)N);

    float alpha = modff(i0f * (float)N, &tmp);
    float beta = modff(j0f * (float)N, &tmp);

    float t[4];
    t[0] = wave_texel(w, ((long)j0 * N + i0) * 4 + channel);
    t[1] = wave_texel(w, ((long)j0 * N + i1) * 4 + channel);
    t[2] = wave_texel(w, ((long)j1 * N + i0) * 4 + channel);
    t[3] = wave_texel(w, ((long)j1 * N + i1) * 4 + channel);

    float h = (1.f - alpha) * (1.f - beta) * t[0] + alpha * (1.f - beta) * t[1] + (1.f - alpha) * beta * t[2] +
              alpha * beta * t[3];
    return h;
}

/* OpenGLRealOcean::ComputeWaveHeight (OpenGLRealOcean.cpp:321-331) */
float orc_wave_height(const sfh_waves* w, float x, float y)
{
    float z = 0.f;
    z -= orc_wave_bilinear(w, x / w->grid_sizes[0], y / w->grid_sizes[0], 0);
    z -= orc_wave_bilinear(w, x / w->grid_sizes[1], y / w->grid_sizes[1], 1);
    return z;
}

/* Ocean::GetDepth(const glm::vec3&) (Ocean.cpp:149-168) */
float orc_get_depth(const sfh_ocean* o, const sfh_waves* w, float x, float y, float z)
{
    if (o->has_waves) {
        float waveHeight = orc_wave_height(w, x, y);
        return z - waveHeight;
    }
    return z;
}

/* Ocean::GetDepth(const Vector3&) (Ocean.cpp:170-173) */
double orc_get_depth_d(const sfh_ocean* o, const sfh_waves* w, const double p[3])
{
    return (double)orc_get_depth(o, w, (float)p[0], (float)p[1], (float)p[2]);
}

/* Uniform::GetVelocityAtPoint (Uniform.cpp:54-57), Jet:: (Jet.cpp:54-75), Pipe:: (Pipe.cpp:57-78) */
void orc_current_velocity(const sfh_current* c, const double pp[3], double out[3])
{
    D3 p{pp[0], pp[1], pp[2]};
    D3 r{0.0, 0.0, 0.0};
    if (c->type == SFH_CURRENT_UNIFORM) {
        r = D3{c->a[0], c->a[1], c->a[2]};
    } else if (c->type == SFH_CURRENT_JET) {
        D3 cc{c->a[0], c->a[1], c->a[2]}, n{c->n[0], c->n[1], c->n[2]};
        double rr = c->r1, vout = c->vel;
        D3 cp = p - cc;
        double d = dnorm(dcross(cp, n));
        double t = ddot(cp, n);
        if (!(t < 0.0)) {
            double r_ = 1.0 / 5.0 * (t + 5.0 * rr);
            if (!(d >= r_)) {
                D3 vmax = 10.0 * rr / (t + 5.0 * rr) * vout * n;
                double f = std::exp(-50.0 * d * d / (t * t));
                r = f * vmax;
            }
        }
    } else if (c->type == SFH_CURRENT_PIPE) {
        D3 p1{c->a[0], c->a[1], c->a[2]}, n{c->n[0], c->n[1], c->n[2]};
        double l = c->l, r1 = c->r1, r2 = c->r2, vin = c->vel, gamma = c->gamma;
        D3 p1p = p - p1;
        double d = dnorm(dcross(p1p, n));
        double t = ddot(p1p, n);
        if (!(t < 0.0 || t > l)) {
            double rad = r1 + (r2 - r1) * t / l;
            if (!(d >= rad)) {
                D3 v = r1 / rad * vin * n;
                double f = std::pow(1.0 - d / rad, gamma);
                r = f * v;
            }
        }
    }
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}

/* Ocean::GetFluidVelocity(const Vector3&) (Ocean.cpp:183-196) */
void orc_fluid_velocity_d(const sfh_ocean* o, const double p[3], double out[3])
{
    D3 fv{0.0, 0.0, 0.0};
    if (o->currents_enabled) {
        for (uint32_t i = 0; i < o->n_currents; ++i) {
            if (o->currents[i].enabled) {
                double v[3];
                orc_current_velocity(&o->currents[i], p, v);
                fv = fv + D3{v[0], v[1], v[2]};
            }
        }
    }
    out[0] = fv.x; out[1] = fv.y; out[2] = fv.z;
}

/* Ocean::GetFluidVelocity(const glm::vec3&) (Ocean.cpp:198-201) + glVectorFromVector (OpenGLDataStructs.cpp:168-171) */
void orc_fluid_velocity_f(const sfh_ocean* o, const float p[3], float out[3])
{
    double pd[3] = {(double)p[0], (double)p[1], (double)p[2]}, v[3];
    orc_fluid_velocity_d(o, pd, v);
    out[0] = (float)v[0]; out[1] = (float)v[1]; out[2] = (float)v[2];
}

/* SolidEntity::CheckBodyFluidPosition (SolidEntity.cpp:1239-1261) */
uint32_t orc_check_body_fluid_position(const sfh_ocean* o, const sfh_waves* w, const double mn[3], const double mx[3])
{
    D3 aabbMin{mn[0], mn[1], mn[2]}, aabbMax{mx[0], mx[1], mx[2]};
    D3 d = aabbMax - aabbMin;
    unsigned underwater = 0;
    auto depth = [&](D3 p) {
        double q[3] = {p.x, p.y, p.z};
        return orc_get_depth_d(o, w, q);
    };
    if (depth(aabbMin) > 0.0) ++underwater;
    if (depth(aabbMax) > 0.0) ++underwater;
    if (depth(aabbMin + D3{d.x, 0, 0}) > 0.0) ++underwater;
    if (depth(aabbMin + D3{0, d.y, 0}) > 0.0) ++underwater;
    if (depth(aabbMin + D3{d.x, d.y, 0}) > 0.0) ++underwater;
    if (depth(aabbMin + D3{0, 0, d.z}) > 0.0) ++underwater;
    if (depth(aabbMin + D3{d.x, 0, d.z}) > 0.0) ++underwater;
    if (depth(aabbMin + D3{0, d.y, d.z}) > 0.0) ++underwater;
    if (underwater == 0) return SFH_POS_OUTSIDE;
    if (underwater == 8) return SFH_POS_INSIDE;
    return SFH_POS_CROSSING;
}

/* SolidEntity::CorrectHydrodynamicForces (SolidEntity.cpp:1263-1287);
 * btMatrix3x3::inverse (btMatrix3x3.h:1092-1103), btVector3::safeNormalize (btVector3.h:286-299) */
void orc_correct(double density, double Fdq[3], double Tdq[3], double Fdf[3], double Tdf[3], const double Cd[3],
                 const double Cf[3], const double B[9])
{
    auto el = [&](int r, int c) { return B[3 * r + c]; };
    auto cofac = [&](int r1, int c1, int r2, int c2) { return el(r1, c1) * el(r2, c2) - el(r1, c2) * el(r2, c1); };
    D3 co{cofac(1, 1, 2, 2), cofac(1, 2, 2, 0), cofac(1, 0, 2, 1)};
    double det = ddot(D3{el(0, 0), el(0, 1), el(0, 2)}, co);
    double s = 1.0 / det;
    D3 r0{co.x * s, cofac(0, 2, 2, 1) * s, cofac(0, 1, 1, 2) * s};
    D3 r1{co.y * s, cofac(0, 0, 2, 2) * s, cofac(0, 2, 1, 0) * s};
    D3 r2{co.z * s, cofac(0, 1, 2, 0) * s, cofac(0, 0, 1, 1) * s};

    auto coeff = [&](const double X[3], const double C[3]) {
        D3 v{X[0], X[1], X[2]};
        D3 q{ddot(r0, v), ddot(r1, v), ddot(r2, v)}; /* toOrigin * X (btMatrix3x3.h:1215-1222) */
        double l2 = ddot(q, q);
        if (l2 >= 2.2204460492503131e-16 * 2.2204460492503131e-16) /* SIMD_EPSILON = DBL_EPSILON */
            q = q * (1.0 / std::sqrt(l2));                         /* operator/= : *this *= 1/s */
        else
            q = D3{1, 0, 0};
        return std::fabs(q.x) * C[0] + std::fabs(q.y) * C[1] + std::fabs(q.z) * C[2];
    };
    auto scale = [&](double X[3], double k) {
        X[0] = X[0] * k; X[1] = X[1] * k; X[2] = X[2] * k; /* scalar * Vector3 -> v * s */
    };
    double Fdqc = coeff(Fdq, Cd);
    scale(Fdq, 0.5 * density * Fdqc);
    double Tdqc = coeff(Tdq, Cd);
    scale(Tdq, 0.5 * density * Tdqc);
    double Fdfc = coeff(Fdf, Cf);
    scale(Fdf, density * Fdfc);
    double Tdfc = coeff(Tdf, Cf);
    scale(Tdf, density * Tdfc);
}

void orc_gl_matrix(const double T[12], float out[16])
{
    M4 M = gl_matrix(T);
    std::memcpy(out, M.m, sizeof(M.m));
}

/* Atmosphere::GetFluidVelocity, glm overload (Atmosphere.cpp:154-157) over the Vector3 one (:146-152) */
void orc_wind_velocity_f(const sfh_atmosphere* atm, const float p[3], float out[3])
{
    const double pd[3] = {(double)p[0], (double)p[1], (double)p[2]};
    D3 fv{0, 0, 0};
    for (uint32_t i = 0; i < atm->n_winds; ++i) {
        double v[3];
        orc_current_velocity(&atm->winds[i], pd, v);
        fv = fv + D3{v[0], v[1], v[2]};
    }
    out[0] = (float)fv.x; out[1] = (float)fv.y; out[2] = (float)fv.z;
}

/* SolidEntity::CorrectAerodynamicForces (SolidEntity.cpp:1925-1956) */
void orc_correct_aero(double Fda[3], double Tda[3], const double Cd[3], const double B[9], int use_cd)
{
    auto el = [&](int r, int c) { return B[3 * r + c]; };
    auto cofac = [&](int r1, int c1, int r2, int c2) { return el(r1, c1) * el(r2, c2) - el(r1, c2) * el(r2, c1); };
    D3 co{cofac(1, 1, 2, 2), cofac(1, 2, 2, 0), cofac(1, 0, 2, 1)};
    double det = ddot(D3{el(0, 0), el(0, 1), el(0, 2)}, co);
    double s = 1.0 / det;
    D3 r0{co.x * s, cofac(0, 2, 2, 1) * s, cofac(0, 1, 1, 2) * s};
    D3 r1{co.y * s, cofac(0, 0, 2, 2) * s, cofac(0, 2, 1, 0) * s};
    D3 r2{co.z * s, cofac(0, 1, 2, 0) * s, cofac(0, 0, 1, 1) * s};
    D3 v{Fda[0], Fda[1], Fda[2]};
    D3 Fdan{ddot(r0, v), ddot(r1, v), ddot(r2, v)};
    double l2 = ddot(Fdan, Fdan);
    if (l2 >= 2.2204460492503131e-16 * 2.2204460492503131e-16)
        Fdan = Fdan * (1.0 / std::sqrt(l2));
    else
        Fdan = D3{1, 0, 0};
    double corFactor = 1.0;
    if (use_cd) corFactor = ddot(D3{Cd[0], Cd[1], Cd[2]}, Fdan);
    const double k = std::fabs(corFactor);
    for (int c = 0; c < 3; ++c) { Fda[c] *= k; Tda[c] *= k; }
}

} // extern "C"

namespace {

/* ------------------------------------------------------------------------------------------------
 * Accumulators of one integrator call. T = float reproduces the reference's glm::vec3 accumulators
 * (SolidEntity.cpp:1317-1329, 1736-1739); T = double sums the identical fp32 terms in fp64.
 */
template <typename T>
struct Acc {
    T Fb[3] = {0, 0, 0}, Tb[3] = {0, 0, 0}, Fdq[3] = {0, 0, 0}, Tdq[3] = {0, 0, 0}, Fdf[3] = {0, 0, 0}, Tdf[3] = {0, 0, 0};
    T Swet = 0, Vsub = 0, CBsub[3] = {0, 0, 0};
    uint32_t cls[8] = {0, 0, 0, 0, 0, 0, 0, 0}, n_degenerate = 0, n_quadratic = 0, n_skin = 0, n_faces = 0;
    inline void add3(T a[3], V3 v) { a[0] += (T)v.x; a[1] += (T)v.y; a[2] += (T)v.z; }
    void merge(const Acc<T>& o)
    {
        for (int k = 0; k < 3; ++k) {
            Fb[k] += o.Fb[k]; Tb[k] += o.Tb[k]; Fdq[k] += o.Fdq[k]; Tdq[k] += o.Tdq[k];
            Fdf[k] += o.Fdf[k]; Tdf[k] += o.Tdf[k]; CBsub[k] += o.CBsub[k];
        }
        Swet += o.Swet; Vsub += o.Vsub;
        for (int k = 0; k < 8; ++k) cls[k] += o.cls[k];
        n_degenerate += o.n_degenerate; n_quadratic += o.n_quadratic; n_skin += o.n_skin; n_faces += o.n_faces;
    }
};

struct Prologue {
    M4 TC;
    V3 v, omega, p, p0;
};

/* SolidEntity.cpp:1325-1334 / 1740-1746 */
Prologue make_prologue(const sfh_body_state* st)
{
    Prologue P;
    M4 TCG = gl_matrix(st->T_CG);
    P.TC = gl_matrix(st->T_C);
    P.v = V3{(float)st->v[0], (float)st->v[1], (float)st->v[2]};
    P.omega = V3{(float)st->omega[0], (float)st->omega[1], (float)st->omega[2]};
    P.p = V3{TCG.m[12], TCG.m[13], TCG.m[14]};
    P.p0 = P.p;
    P.p0.z = 0.f;
    return P;
}

inline V3 fluid_velocity(const sfh_ocean* o, V3 fc)
{
    float p[3] = {fc.x, fc.y, fc.z}, v[3];
    orc_fluid_velocity_f(o, p, v);
    return V3{v[0], v[1], v[2]};
}

/* Damping block shared by both integrators (SolidEntity.cpp:1661-1683, 1771-1790) */
template <typename T>
inline uint8_t damping(Acc<T>& a, const sfh_ocean* o, const Prologue& P, V3 fc, V3 fn1, float A)
{
    uint8_t code = 0;
    V3 vc = fluid_velocity(o, fc) - (P.v + cross(P.omega, fc - P.p));
    float vc_n = dot(vc, fn1);
    V3 vn = vc_n * fn1;
    V3 vt = vc - vn;
    if (vc_n < -1e-12f) {
        float vmag2 = length2(vc);
        V3 quadratic = vc * sqrtf(vmag2) * -vc_n * A;
        a.add3(a.Fdq, quadratic);
        a.add3(a.Tdq, cross(fc - P.p, quadratic));
        ++a.n_quadratic;
        code |= SFH_FACE_QUADRATIC;
    }
    float vmag2 = length2(vt);
    if (vmag2 > 1e-9f) {
        V3 skin = vt * A;
        a.add3(a.Fdf, skin);
        a.add3(a.Tdf, cross(fc - P.p, skin));
        ++a.n_skin;
        code |= SFH_FACE_SKIN;
    }
    return code;
}

/* Face loop of ComputeHydrodynamicForcesSubmerged (SolidEntity.cpp:1749-1791) over faces [f0,f1) */
template <typename T>
void submerged_faces(Acc<T>& a, const orc_mesh* mesh, const sfh_ocean* o, const Prologue& P, size_t f0, size_t f1,
                     uint8_t* codes)
{
    for (size_t i = f0; i < f1; ++i) {
        V3 p1 = xform(P.TC, vertex_pos(mesh, i, 0));
        V3 p2 = xform(P.TC, vertex_pos(mesh, i, 1));
        V3 p3 = xform(P.TC, vertex_pos(mesh, i, 2));
        ++a.n_faces;
        ++a.cls[7];
        V3 fv1 = p2 - p1;
        V3 fv2 = p3 - p1;
        V3 fn = cross(fv1, fv2);
        float len = length2(fn);
        if (len < 1e-12f) {
            ++a.n_degenerate;
            if (codes) codes[i] = 7 | SFH_FACE_DEGENERATE;
            continue;
        }
        len = std::sqrt(len);
        V3 fn1 = fn / len;
        float A = len / 2.f;
        V3 fc = (p1 + p2 + p3) / 3.f;
        uint8_t code = 7 | damping(a, o, P, fc, fn1, A);
        if (codes) codes[i] = code;
    }
}

/* Face loop of ComputeHydrodynamicForcesSurface (SolidEntity.cpp:1337-1687) over faces [f0,f1) */
template <typename T>
void surface_faces(Acc<T>& a, const orc_mesh* mesh, const sfh_ocean* o, const sfh_waves* w, const Prologue& P,
                   bool reallisticBuoyancy, bool dampingForces, size_t f0, size_t f1, uint8_t* codes)
{
    const V3 p0 = P.p0;
    /* one tetrahedron (p0, a, b, c) of the submerged volume, SolidEntity.cpp:1371-1377 and twins */
    auto tetra = [&](V3 pa, V3 pb, V3 pc) {
        V3 tetraCG = (pa + pb + pc) / 4.f;
        float tetraV6 = dot(pa, cross(pb, pc));
        a.add3(a.CBsub, tetraCG * tetraV6);
        a.Vsub += (T)tetraV6;
    };
    /* waterline point on the edge from the submerged vertex pu (depth du >= 0) to the dry vertex pa (da < 0) */
    auto clip = [](V3 pa, V3 pu, float da, float du) { return pu + (pa - pu) * (du / (fabsf(da) + du)); };

    for (size_t i = f0; i < f1; ++i) {
        V3 p1 = xform(P.TC, vertex_pos(mesh, i, 0));
        V3 p2 = xform(P.TC, vertex_pos(mesh, i, 1));
        V3 p3 = xform(P.TC, vertex_pos(mesh, i, 2));
        ++a.n_faces;

        float depth[3];
        depth[0] = orc_get_depth(o, w, p1.x, p1.y, p1.z);
        depth[1] = orc_get_depth(o, w, p2.x, p2.y, p2.z);
        depth[2] = orc_get_depth(o, w, p3.x, p3.y, p3.z);

        const bool a0 = depth[0] < 0.f, a1 = depth[1] < 0.f, a2 = depth[2] < 0.f;
        const uint8_t cls = (uint8_t)((a0 ? 0 : 4) | (a1 ? 0 : 2) | (a2 ? 0 : 1));
        ++a.cls[cls];
        if (codes) codes[i] = cls;
        if (a0 && a1 && a2) continue; /* :1353-1354 */

        V3 fc, fn, fn1;
        float A;
        bool degenerate = false;
        /* triangle face properties, e.g. :1380-1389 */
        auto tri = [&]() {
            V3 fv1 = p2 - p1;
            V3 fv2 = p3 - p1;
            fc = (p1 + p2 + p3) / 3.f;
            fn = cross(fv1, fv2);
            float len = length2(fn);
            if (len < 1e-12f) { degenerate = true; return; }
            len = std::sqrt(len);
            fn1 = fn / len;
            A = len / 2.f;
        };
        /* quad face properties, e.g. :1459-1471 */
        auto quad = [&](V3 fv1, V3 fv2, V3 fv3, V3 fv4, V3 p4) {
            fc = (p1 + p2 + p3 + p4) / 4.f;
            fn = cross(fv1, fv2);
            float len = length2(fn);
            if (len < 1e-12f) { degenerate = true; return; }
            len = std::sqrt(len);
            fn1 = fn / len;
            A = (len + length(cross(fv3, fv4))) / 2.f;
        };

        if (a0) {
            if (a1) { /* class 1, :1364-1398 */
                p1 = clip(p1, p3, depth[0], depth[2]);
                p2 = clip(p2, p3, depth[1], depth[2]);
                tetra(p1 - p0, p2 - p0, p3 - p0);
                tri();
            } else if (a2) { /* class 2, :1399-1433 */
                p1 = clip(p1, p2, depth[0], depth[1]);
                p3 = clip(p3, p2, depth[2], depth[1]);
                tetra(p1 - p0, p2 - p0, p3 - p0);
                tri();
            } else { /* class 3, :1434-1482 */
                V3 p4 = clip(p1, p3, depth[0], depth[2]);
                p1 = clip(p1, p2, depth[0], depth[1]);
                V3 p01 = p1 - p0, p02 = p2 - p0, p03 = p3 - p0;
                tetra(p01, p02, p03);
                V3 p04 = p4 - p0;
                tetra(p01, p03, p04);
                quad(p2 - p1, p4 - p1, p2 - p3, p4 - p3, p4);
            }
        } else if (a1) {
            if (a2) { /* class 4, :1486-1520 */
                p2 = clip(p2, p1, depth[1], depth[0]);
                p3 = clip(p3, p1, depth[2], depth[0]);
                tetra(p1 - p0, p2 - p0, p3 - p0);
                tri();
            } else { /* class 5, :1521-1568 */
                V3 p4 = clip(p2, p3, depth[1], depth[2]);
                p2 = clip(p2, p1, depth[1], depth[0]);
                V3 p01 = p1 - p0, p02 = p2 - p0, p03 = p3 - p0;
                tetra(p01, p02, p03);
                V3 p04 = p4 - p0;
                tetra(p02, p04, p03);
                quad(p2 - p1, p3 - p1, p2 - p3, p4 - p3, p4);
            }
        } else if (a2) { /* class 6, :1570-1617 */
            V3 p4 = clip(p3, p1, depth[2], depth[0]);
            p3 = clip(p3, p2, depth[2], depth[1]);
            V3 p01 = p1 - p0, p02 = p2 - p0, p03 = p3 - p0;
            tetra(p01, p02, p03);
            V3 p04 = p4 - p0;
            tetra(p01, p03, p04);
            quad(p2 - p1, p4 - p1, p2 - p3, p4 - p3, p4);
        } else { /* class 7, :1618-1647 */
            tetra(p1 - p0, p2 - p0, p3 - p0);
            tri();
        }
        if (degenerate) {
            ++a.n_degenerate;
            if (codes) codes[i] |= SFH_FACE_DEGENERATE;
            continue;
        }

        /* :1649-1658 */
        if (reallisticBuoyancy && o->has_waves) {
            float depthc = orc_get_depth(o, w, fc.x, fc.y, fc.z);
            V3 Fbi = -fn1 * A * depthc;
            a.add3(a.Fb, Fbi);
            a.add3(a.Tb, cross(fc - P.p, Fbi));
        }
        /* :1660-1683 */
        if (dampingForces) {
            uint8_t c = damping(a, o, P, fc, fn1, A);
            if (codes) codes[i] |= c;
        }
        /* :1685-1686 */
        a.Swet += (T)A;
    }
}

void zero_info(sfh_info* out)
{
    std::memset(out, 0, sizeof(*out));
}

template <typename T>
void copy_counters(const Acc<T>& a, sfh_info* out)
{
    for (int k = 0; k < 8; ++k) out->class_count[k] = a.cls[k];
    out->n_degenerate = a.n_degenerate;
    out->n_quadratic = a.n_quadratic;
    out->n_skin = a.n_skin;
    out->n_faces = a.n_faces;
}

/* SolidEntity.cpp:1793-1796 */
template <typename T>
void finalize_submerged(const Acc<T>& a, sfh_info* out)
{
    for (int k = 0; k < 3; ++k) {
        out->Fdq[k] = (double)a.Fdq[k]; out->Tdq[k] = (double)a.Tdq[k];
        out->Fdf[k] = (double)a.Fdf[k]; out->Tdf[k] = (double)a.Tdf[k];
    }
    out->wrote_damping = 1;
    copy_counters(a, out);
}

/* SolidEntity.cpp:1689-1720. With T=double the same expressions are evaluated on the fp64 sums
 * (thresholds and constants keep their fp32 values): this is the F64_ACC contract of DESIGN.md. */
template <typename T>
void finalize_surface(Acc<T>& a, const sfh_ocean* o, const sfh_body_state* st, const Prologue& P,
                      bool reallisticBuoyancy, bool dampingForces, sfh_info* out)
{
    if (reallisticBuoyancy && a.Vsub > (T)1e-9f) {
        out->wrote_buoyancy = 1;
        out->Vsub = (double)(a.Vsub / (T)6.f);
        if (o->has_waves) {
            /* glm: vec3 *= double  ->  component *= static_cast<float>(scalar) */
            const float k = (float)(o->density * o->gravity_z);
            for (int c = 0; c < 3; ++c) {
                out->Fb[c] = (double)(a.Fb[c] * (T)k);
                out->Tb[c] = (double)(a.Tb[c] * (T)k);
            }
        } else {
            T CB[3];
            CB[0] = a.CBsub[0] / a.Vsub + (T)P.p0.x;
            CB[1] = a.CBsub[1] / a.Vsub + (T)P.p0.y;
            CB[2] = a.CBsub[2] / a.Vsub + (T)P.p0.z;
            D3 _CBsub{(double)CB[0], (double)CB[1], (double)CB[2]};
            D3 gravity{0.0, 0.0, o->gravity_z};
            D3 Fb = -out->Vsub * o->density * gravity;
            D3 Tb = dcross(_CBsub - D3{st->T_CG[9], st->T_CG[10], st->T_CG[11]}, Fb);
            out->Fb[0] = Fb.x; out->Fb[1] = Fb.y; out->Fb[2] = Fb.z;
            out->Tb[0] = Tb.x; out->Tb[1] = Tb.y; out->Tb[2] = Tb.z;
        }
    }
    if (dampingForces) {
        out->wrote_damping = 1;
        for (int k = 0; k < 3; ++k) {
            out->Fdq[k] = (double)a.Fdq[k]; out->Tdq[k] = (double)a.Tdq[k];
            out->Fdf[k] = (double)a.Fdf[k]; out->Tdf[k] = (double)a.Tdf[k];
        }
    }
    out->Swet = (double)a.Swet;
    copy_counters(a, out);
}

/* Face loop of the static ComputeAerodynamicForces (SolidEntity.cpp:1886-1919) over faces [f0,f1); Fda/Tda live in a.Fdq/a.Tdq */
template <typename T>
void aero_faces(Acc<T>& a, const orc_mesh* mesh, const sfh_atmosphere* atm, const Prologue& P, size_t f0, size_t f1, uint8_t* codes)
{
    for (size_t i = f0; i < f1; ++i) {
        V3 p1 = xform(P.TC, vertex_pos(mesh, i, 0));
        V3 p2 = xform(P.TC, vertex_pos(mesh, i, 1));
        V3 p3 = xform(P.TC, vertex_pos(mesh, i, 2));
        ++a.n_faces;
        ++a.cls[7];
        V3 fv1 = p2 - p1;
        V3 fv2 = p3 - p1;
        V3 fn = cross(fv1, fv2);
        float len = length2(fn);
        if (len < 1e-12f) {
            ++a.n_degenerate;
            if (codes) codes[i] = 7 | SFH_FACE_DEGENERATE;
            continue;
        }
        len = std::sqrt(len);
        V3 fn1 = fn / len;
        float A = len / 2.f;
        V3 fc = (p1 + p2 + p3) / 3.f;
        float pf[3] = {fc.x, fc.y, fc.z}, wf[3];
        orc_wind_velocity_f(atm, pf, wf);
        V3 vc = V3{wf[0], wf[1], wf[2]} - (P.v + cross(P.omega, fc - P.p));
        V3 vn = dot(vc, fn1) * fn1;
        uint8_t code = 7;
        if (dot(fn1, vn) < -1e-12f) {
            /* `vn * vn.length() * A`: glm::vec3::length() is the component count (int 3); vec3 * int resolves to
             * glm/gtx/scalar_multiplication.hpp (pulled in by glm/ext.hpp with GLM_ENABLE_EXPERIMENTAL,
             * OpenGLDataStructs.h:30-32): lh *= static_cast<float>(3) */
            V3 quadratic = vn * 3.f * A;
            a.add3(a.Fdq, quadratic);
            a.add3(a.Tdq, cross(fc - P.p, quadratic));
            ++a.n_quadratic;
            code |= SFH_FACE_QUADRATIC;
        }
        if (codes) codes[i] = code;
    }
}

template <typename T>
void run_aero(const orc_mesh* mesh, const sfh_atmosphere* atm, const sfh_body_state* st, sfh_info* out, uint8_t* codes)
{
    Prologue P = make_prologue(st);
    Acc<T> a;
    aero_faces(a, mesh, atm, P, 0, mesh->n_faces, codes);
    /* SolidEntity.cpp:1921-1923: Scalar(0.5) * density * Vector3(Fda) */
    const double k = 0.5 * atm->density;
    for (int c = 0; c < 3; ++c) {
        out->Fdq[c] = (double)a.Fdq[c] * k;
        out->Tdq[c] = (double)a.Tdq[c] * k;
    }
    copy_counters(a, out);
}

template <typename T>
void run_submerged(const orc_mesh* mesh, const sfh_ocean* o, const sfh_body_state* st, sfh_info* out, uint8_t* codes)
{
    Prologue P = make_prologue(st);
    Acc<T> a;
    submerged_faces(a, mesh, o, P, 0, mesh->n_faces, codes);
    finalize_submerged(a, out);
}

template <typename T>
void run_surface(const orc_mesh* mesh, const sfh_ocean* o, const sfh_waves* w, const sfh_body_state* st, bool rb,
                 bool df, sfh_info* out, uint8_t* codes)
{
    Prologue P = make_prologue(st);
    Acc<T> a;
    surface_faces(a, mesh, o, w, P, rb, df, 0, mesh->n_faces, codes);
    finalize_surface(a, o, st, P, rb, df, out);
}

} // namespace

extern "C" {

void orc_submerged(const orc_mesh* mesh, const sfh_ocean* o, const sfh_body_state* st, int acc_mode, sfh_info* out,
                   uint8_t* codes)
{
    zero_info(out);
    out->position = SFH_POS_INSIDE;
    if (mesh == nullptr) { /* :1726-1733 */
        out->wrote_damping = 1;
        return;
    }
    if (acc_mode == ORC_REF_F32_SEQ)
        run_submerged<float>(mesh, o, st, out, codes);
    else
        run_submerged<double>(mesh, o, st, out, codes);
}

void orc_surface(const orc_mesh* mesh, const sfh_ocean* o, const sfh_waves* w, const sfh_body_state* st, int acc_mode,
                 sfh_info* out, uint8_t* codes)
{
    const bool rb = (st->flags & SFH_F_BUOYANCY) != 0, df = (st->flags & SFH_F_DAMPING) != 0;
    zero_info(out);
    out->position = SFH_POS_CROSSING;
    if (mesh == nullptr) { /* :1293-1312 */
        out->wrote_buoyancy = rb;
        out->wrote_damping = df;
        return;
    }
    if (acc_mode == ORC_REF_F32_SEQ)
        run_surface<float>(mesh, o, w, st, rb, df, out, codes);
    else
        run_surface<double>(mesh, o, w, st, rb, df, out, codes);
}

void orc_aero(const orc_mesh* mesh, const sfh_atmosphere* atm, const sfh_body_state* st, int acc_mode, sfh_info* out, uint8_t* codes)
{
    zero_info(out);
    out->position = SFH_POS_OUTSIDE;
    out->wrote_damping = 1;
    if (mesh != nullptr) { /* :1868-1873: no mesh -> zero */
        if (acc_mode == ORC_REF_F32_SEQ)
            run_aero<float>(mesh, atm, st, out, codes);
        else
            run_aero<double>(mesh, atm, st, out, codes);
    }
    if (st->flags & SFH_F_DISPATCH) /* :1862 */
        orc_correct_aero(out->Fdq, out->Tdq, st->Cd, st->T_O_basis, (st->flags & SFH_F_AERO_CD) != 0);
}

void orc_compound_aero(const orc_mesh* const* part_meshes, const sfh_body_state* parts, const uint32_t* external, uint32_t n_parts,
                       const sfh_atmosphere* atm, const sfh_body_state* compound, int acc_mode, sfh_info* out)
{
    zero_info(out);
    out->position = SFH_POS_OUTSIDE;
    out->wrote_damping = 1; /* Compound.cpp:444-445 */
    for (uint32_t i = 0; i < n_parts; ++i) {
        if (!external[i]) continue; /* :456 */
        sfh_body_state st = parts[i];
        std::memcpy(st.T_CG, compound->T_CG, sizeof(st.T_CG));
        std::memcpy(st.v, compound->v, sizeof(st.v));
        std::memcpy(st.omega, compound->omega, sizeof(st.omega));
        st.flags |= SFH_F_DISPATCH; /* :459-460: static integrator, then the part's CorrectAerodynamicForces */
        sfh_info pi;
        orc_aero(part_meshes[i], atm, &st, acc_mode, &pi, nullptr);
        for (int c = 0; c < 3; ++c) { out->Fdq[c] += pi.Fdq[c]; out->Tdq[c] += pi.Tdq[c]; }
        out->n_faces += pi.n_faces; out->n_degenerate += pi.n_degenerate; out->n_quadratic += pi.n_quadratic;
        out->class_count[7] += pi.class_count[7];
    }
}

/* The sfh_step contract for one body. With SFH_F_DISPATCH: SolidEntity::ComputeHydrodynamicForces
 * (SolidEntity.cpp:1799-1850). */
void orc_body(const orc_mesh* mesh, const sfh_ocean* o, const sfh_waves* w, const sfh_body_state* st, int acc_mode,
              sfh_info* out, uint8_t* codes)
{
    uint32_t pos = st->mode;
    if (pos == SFH_POS_AUTO) pos = orc_check_body_fluid_position(o, w, st->aabb_min, st->aabb_max);
    const bool dispatch = (st->flags & SFH_F_DISPATCH) != 0;

    if (pos == SFH_POS_OUTSIDE) { /* :1810-1821 — all eight outputs set to zero */
        zero_info(out);
        out->position = SFH_POS_OUTSIDE;
        out->wrote_buoyancy = 1;
        out->wrote_damping = 1;
        if (codes && mesh) std::memset(codes, 0, mesh->n_faces);
        return;
    }
    if (!dispatch) {
        if (pos == SFH_POS_INSIDE)
            orc_submerged(mesh, o, st, acc_mode, out, codes);
        else
            orc_surface(mesh, o, w, st, acc_mode, out, codes);
        return;
    }

    const bool damping = (st->flags & SFH_F_DAMPING) != 0;
    const bool buoyant = (st->flags & SFH_F_BUOYANT) != 0;
    if (pos == SFH_POS_INSIDE) { /* :1828-1841 */
        zero_info(out);
        if (damping) {
            orc_submerged(mesh, o, st, acc_mode, out, codes);
        } else if (codes && mesh) {
            std::memset(codes, 0, mesh->n_faces);
        }
        out->position = SFH_POS_INSIDE;
        if (buoyant) {
            D3 gravity{0.0, 0.0, o->gravity_z};
            D3 Fb = -st->volume * o->density * gravity;
            const double* B = st->T_CG;
            D3 pcb{st->P_CB[0], st->P_CB[1], st->P_CB[2]};
            D3 org{B[9], B[10], B[11]};
            /* btTransform::operator()(x) = x.dot3(basis rows) + origin (btTransform.h:96-99) */
            D3 wcb{ddot(D3{B[0], B[1], B[2]}, pcb) + org.x, ddot(D3{B[3], B[4], B[5]}, pcb) + org.y,
                   ddot(D3{B[6], B[7], B[8]}, pcb) + org.z};
            D3 Tb = dcross(wcb - org, Fb);
            out->Fb[0] = Fb.x; out->Fb[1] = Fb.y; out->Fb[2] = Fb.z;
            out->Tb[0] = Tb.x; out->Tb[1] = Tb.y; out->Tb[2] = Tb.z;
            out->wrote_buoyancy = 1;
        }
        out->Swet = st->surface; /* :1840; Vsub is not touched for INSIDE bodies (Appendix B.2) */
    } else {                     /* :1842-1846 */
        sfh_body_state s2 = *st;
        if (!buoyant) s2.flags &= ~(uint32_t)SFH_F_BUOYANCY;
        orc_surface(mesh, o, w, &s2, acc_mode, out, codes);
    }
    if (damping) /* :1848-1849 */
        orc_correct(o->density, out->Fdq, out->Tdq, out->Fdf, out->Tdf, st->Cd, st->Cf, st->T_O_basis);
}

/* Compound::ComputeHydrodynamicForces (Compound.cpp:291-437) */
void orc_compound(const orc_mesh* const* part_meshes, const sfh_body_state* parts, const uint32_t* external, uint32_t n_parts,
                  const sfh_ocean* o, const sfh_waves* w, const sfh_body_state* compound, int acc_mode, sfh_info* out)
{
    zero_info(out);
    uint32_t pos = compound->mode;
    if (pos == SFH_POS_AUTO) pos = orc_check_body_fluid_position(o, w, compound->aabb_min, compound->aabb_max); /* :299 */
    out->position = pos;
    if (pos == SFH_POS_OUTSIDE) { /* :302-313 */
        out->wrote_buoyancy = out->wrote_damping = 1;
        return;
    }
    const bool damping = (compound->flags & SFH_F_DAMPING) != 0, rb = (compound->flags & SFH_F_BUOYANCY) != 0;
    auto add3 = [](double* a, const double* b) { a[0] += b[0]; a[1] += b[1]; a[2] += b[2]; }; /* btVector3 += */
    auto part_state = [&](uint32_t i, bool p_rb, bool p_df) {
        sfh_body_state s = parts[i];
        std::memcpy(s.T_CG, compound->T_CG, sizeof(s.T_CG)); /* getCGTransform() of the compound for every part */
        std::memcpy(s.v, compound->v, sizeof(s.v));
        std::memcpy(s.omega, compound->omega, sizeof(s.omega));
        s.flags = (p_rb ? (uint32_t)SFH_F_BUOYANCY : 0u) | (p_df ? (uint32_t)SFH_F_DAMPING : 0u);
        return s;
    };
    auto add_counters = [&](const sfh_info& r) {
        for (int k = 0; k < 8; ++k) out->class_count[k] += r.class_count[k];
        out->n_degenerate += r.n_degenerate; out->n_quadratic += r.n_quadratic; out->n_skin += r.n_skin; out->n_faces += r.n_faces;
    };

    if (pos == SFH_POS_INSIDE) { /* :315-363 */
        if (compound->flags & SFH_F_BUOYANT) {
            D3 gravity{0.0, 0.0, o->gravity_z};
            D3 Fb = -compound->volume * o->density * gravity;
            const double* B = compound->T_CG;
            D3 pcb{compound->P_CB[0], compound->P_CB[1], compound->P_CB[2]}, org{B[9], B[10], B[11]};
            D3 wcb{ddot(D3{B[0], B[1], B[2]}, pcb) + org.x, ddot(D3{B[3], B[4], B[5]}, pcb) + org.y, ddot(D3{B[6], B[7], B[8]}, pcb) + org.z};
            D3 Tb = dcross(wcb - org, Fb);
            out->Fb[0] = Fb.x; out->Fb[1] = Fb.y; out->Fb[2] = Fb.z;
            out->Tb[0] = Tb.x; out->Tb[1] = Tb.y; out->Tb[2] = Tb.z;
            out->wrote_buoyancy = 1;
        }
        if (damping) {
            out->wrote_damping = 1;
            for (uint32_t i = 0; i < n_parts; ++i) {
                if (!external[i]) continue; /* drag only for external parts, :342-345 */
                sfh_body_state s = part_state(i, false, true);
                sfh_info r;
                orc_submerged(part_meshes[i], o, &s, acc_mode, &r, nullptr);
                orc_correct(o->density, r.Fdq, r.Tdq, r.Fdf, r.Tdf, parts[i].Cd, parts[i].Cf, parts[i].T_O_basis);
                add3(out->Fdq, r.Fdq); add3(out->Tdq, r.Tdq); add3(out->Fdf, r.Fdf); add3(out->Tdf, r.Tdf);
                add_counters(r);
            }
        }
        out->Swet = compound->surface;
        out->Vsub = compound->volume; /* :361-362 */
        return;
    }

    /* CROSSING (compound body but not necessarily all parts!), :364-435 */
    if (!(rb || damping)) return;
    if (rb) out->wrote_buoyancy = 1;       /* Fb, Tb set to zero and accumulated below */
    if (damping) out->wrote_damping = 1;
    double Fbp[3] = {0, 0, 0}, Tbp[3] = {0, 0, 0}, Fdqp[3] = {0, 0, 0}, Tdqp[3] = {0, 0, 0}, Fdfp[3] = {0, 0, 0}, Tdfp[3] = {0, 0, 0};
    double Swetp = 0, Vsubp = 0; /* temporaries that persist across parts (:391-398) */
    for (uint32_t i = 0; i < n_parts; ++i) {
        const bool p_rb = rb && (parts[i].flags & SFH_F_BUOYANT) != 0; /* pSettings.reallisticBuoyancy &= isBuoyant(), :409 */
        auto run_surface = [&](bool p_df) {
            sfh_body_state s = part_state(i, p_rb, p_df);
            sfh_info r;
            orc_surface(part_meshes[i], o, w, &s, acc_mode, &r, nullptr);
            if (r.wrote_buoyancy) { std::memcpy(Fbp, r.Fb, sizeof(Fbp)); std::memcpy(Tbp, r.Tb, sizeof(Tbp)); Vsubp = r.Vsub; }
            if (r.wrote_damping) {
                std::memcpy(Fdqp, r.Fdq, sizeof(Fdqp)); std::memcpy(Tdqp, r.Tdq, sizeof(Tdqp));
                std::memcpy(Fdfp, r.Fdf, sizeof(Fdfp)); std::memcpy(Tdfp, r.Tdf, sizeof(Tdfp));
            }
            Swetp = r.Swet;
            add_counters(r);
        };
        if (external[i]) { /* buoyancy and drag, :411-425 */
            run_surface(damping);
            orc_correct(o->density, Fdqp, Tdqp, Fdfp, Tdfp, parts[i].Cd, parts[i].Cf, parts[i].T_O_basis);
            add3(out->Fb, Fbp); add3(out->Tb, Tbp);
            add3(out->Fdq, Fdqp); add3(out->Tdq, Tdqp); add3(out->Fdf, Fdfp); add3(out->Tdf, Tdfp);
            out->Swet += Swetp;
            out->Vsub += Vsubp;
        } else if (p_rb) { /* only buoyancy, :426-433 */
            run_surface(false);
            add3(out->Fb, Fbp); add3(out->Tb, Tbp);
            out->Vsub += Vsubp;
        }
    }
}

double orc_step(const orc_mesh* const* meshes, const sfh_ocean* o, const sfh_waves* w, const sfh_body_state* states,
                uint32_t n_bodies, int acc_mode, int n_threads, sfh_info* out)
{
    auto t0 = std::chrono::steady_clock::now();
    if (n_threads <= 1) {
        for (uint32_t i = 0; i < n_bodies; ++i) orc_body(meshes[i], o, w, &states[i], acc_mode, &out[i], nullptr);
    } else {
        std::atomic<uint32_t> next{0};
        std::vector<std::thread> pool;
        for (int t = 0; t < n_threads; ++t)
            pool.emplace_back([&]() {
                for (;;) {
                    uint32_t i = next.fetch_add(1);
                    if (i >= n_bodies) break;
                    orc_body(meshes[i], o, w, &states[i], acc_mode, &out[i], nullptr);
                }
            });
        for (auto& th : pool) th.join();
    }
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

double orc_body_split(const orc_mesh* mesh, const sfh_ocean* o, const sfh_waves* w, const sfh_body_state* st,
                      int n_threads, sfh_info* out)
{
    auto t0 = std::chrono::steady_clock::now();
    uint32_t pos = st->mode;
    if (pos == SFH_POS_AUTO) pos = orc_check_body_fluid_position(o, w, st->aabb_min, st->aabb_max);
    zero_info(out);
    out->position = pos;
    if (pos != SFH_POS_OUTSIDE && mesh != nullptr) {
        if (n_threads < 1) n_threads = 1;
        const bool rb = (st->flags & SFH_F_BUOYANCY) != 0, df = (st->flags & SFH_F_DAMPING) != 0;
        Prologue P = make_prologue(st);
        std::vector<Acc<double>> part((size_t)n_threads);
        std::vector<std::thread> pool;
        const size_t F = mesh->n_faces;
        for (int t = 0; t < n_threads; ++t)
            pool.emplace_back([&, t]() {
                size_t f0 = F * (size_t)t / (size_t)n_threads, f1 = F * (size_t)(t + 1) / (size_t)n_threads;
                if (pos == SFH_POS_INSIDE)
                    submerged_faces(part[(size_t)t], mesh, o, P, f0, f1, nullptr);
                else
                    surface_faces(part[(size_t)t], mesh, o, w, P, rb, df, f0, f1, nullptr);
            });
        for (auto& th : pool) th.join();
        Acc<double> a;
        for (auto& q : part) a.merge(q);
        if (pos == SFH_POS_INSIDE)
            finalize_submerged(a, out);
        else
            finalize_surface(a, o, st, P, rb, df, out);
    }
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

/* 1 if a*b+c is evaluated unfused in this build (guards against a -march/-ffp-contract slip) */
int orc_selftest_unfused(void)
{
    volatile float a = 1.0f + 0x1p-12f, b = 1.0f + 0x1p-12f, c = -(1.0f + 0x1p-11f);
    float r = a * b + c; /* exact product 1 + 2^-11 + 2^-24 is a tie -> rounds to 1 + 2^-11: 0 unfused, 2^-24 fused */
    return r == 0.0f;
}

/* ------------------------------------------------------------------------------------------------
 * N4: cables. fp64 with Bullet's operator order (btVector3.h), libm transcendentals as btAtan/btCos/btSin/btAcos are.
 */
double orc_circular_segment_area(double radius_, double h) /* CableEntity.cpp:217-237 */
{
    if (h <= 0.0)
        return 0.0;
    else if (h >= 2.0 * radius_)
        return M_PI * radius_ * radius_;
    else {
        if (h > radius_) {
            h = 2.0 * radius_ - h;
            double theta = 2.0 * std::acos((radius_ - h) / radius_);
            return radius_ * radius_ * (M_PI - (theta - std::sin(theta)) / 2.0);
        } else {
            double theta = 2.0 * std::acos((radius_ - h) / radius_);
            return (radius_ * radius_ / 2.0) * (theta - std::sin(theta));
        }
    }
}

void orc_cable(const sfh_ocean* o, const sfh_waves* w, const sfh_cable* cable, const double* pos, const double* vel,
               sfh_node_forces* out)
{
    const uint32_t n = cable->n_nodes;
    const double radius_ = cable->radius;
    const bool buoyancy = (cable->flags & SFH_F_BUOYANT) && (cable->flags & SFH_F_BUOYANCY);
    const bool damping = (cable->flags & SFH_F_DAMPING) != 0;
    const double EPS = 2.2204460492503131e-16; /* SIMD_EPSILON = DBL_EPSILON (btScalar.h:536) */
    auto fuzzy_zero = [&](double x) { return std::fabs(x) < EPS; };
    auto add_half = [](double* acc, D3 F) { /* nodalForces_[k].X += F / Scalar(2): operator/ multiplies by 1/s (btVector3.h:853-858) */
        D3 hF = F * (1.0 / 2.0);
        acc[0] += hF.x; acc[1] += hF.y; acc[2] += hF.z;
    };
    std::memset(out, 0, sizeof(sfh_node_forces) * n); /* :244-245 clearForces */

    for (uint32_t i = 0; i + 1 < n; ++i) {
        D3 p1{pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]};
        D3 p2{pos[3 * i + 3], pos[3 * i + 4], pos[3 * i + 5]};
        double pp1[3] = {p1.x, p1.y, p1.z}, pp2[3] = {p2.x, p2.y, p2.z};
        double d1 = orc_get_depth_d(o, w, pp1);
        double d2 = orc_get_depth_d(o, w, pp2);
        if (d1 > d2) { /* the shallower point first */
            std::swap(p1, p2);
            std::swap(d1, d2);
        }
        double dz = d2 - d1;
        D3 hv{p2.x - p1.x, p2.y - p1.y, 0.0};
        double hl2 = ddot(hv, hv);
        double dxy = hl2 > EPS ? std::sqrt(hl2) : 0.0; /* btVector3::safeNorm (btVector3.h:269-276) */
        D3 segmentVec = p2 - p1;
        double segmentLength = dnorm(segmentVec);
        double alpha = std::atan(dz / dxy);
        double margin = fuzzy_zero(dxy) ? 0.0 : radius_ * std::cos(alpha);

        double submergedV = 0.0;
        if (d2 < -margin) {
            continue;
        } else if (d1 > margin) {
            submergedV = M_PI * radius_ * radius_ * segmentLength;
        } else {
            if (fuzzy_zero(dxy)) {
                double submergedL = std::min(d2, dz);
                submergedV = M_PI * radius_ * radius_ * submergedL;
            } else if (fuzzy_zero(dz)) {
                double submergedA = orc_circular_segment_area(radius_, d1 + radius_);
                submergedV = submergedA * segmentLength;
            } else if (std::fabs(d1) <= margin && std::fabs(d2) <= margin) {
                double A1 = orc_circular_segment_area(radius_, d1 / std::cos(alpha) + radius_);
                double A2 = orc_circular_segment_area(radius_, d2 / std::cos(alpha) + radius_);
                submergedV = (A1 + A2) / 2.0 * segmentLength;
            } else if (d1 < -margin && d2 > margin) {
                double l1 = (d2 - radius_ * std::cos(alpha)) / std::sin(alpha);
                double l2 = 2.0 * radius_ * std::cos(alpha);
                submergedV = M_PI * radius_ * radius_ * (l1 + l2 / 2.0);
            } else if (d2 <= margin) {
                double a = radius_ + d2 / std::cos(alpha);
                double l = a / std::sin(alpha);
                submergedV = orc_circular_segment_area(radius_, a) * l / 2.0;
            } else if (d1 >= -margin) {
                double a = radius_ - d1 / std::cos(alpha);
                double l = a / std::sin(alpha);
                submergedV = M_PI * radius_ * radius_ * segmentLength - orc_circular_segment_area(radius_, a) * l / 2.0;
            }
        }

        if (buoyancy) { /* :319-326 */
            D3 gravity{0.0, 0.0, o->gravity_z};
            D3 b = -submergedV * o->density * gravity;
            add_half(out[i].Fb, b);
            add_half(out[i + 1].Fb, b);
        }
        if (damping && submergedV > 0.0) { /* :329-359 */
            D3 p = (p1 + p2) * (1.0 / 2.0);
            D3 v1{vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]}, v2{vel[3 * i + 3], vel[3 * i + 4], vel[3 * i + 5]};
            D3 v = (v1 + v2) * (1.0 / 2.0);
            double pp[3] = {p.x, p.y, p.z}, wv[3];
            orc_fluid_velocity_d(o, pp, wv);
            D3 relV = v - D3{wv[0], wv[1], wv[2]};
            const double Cd = 0.5;
            double f1 = std::fabs(ddot(segmentVec, relV));
            double f2 = submergedV / (M_PI * radius_ * radius_ * segmentLength);
            {
                double S = 2.0 * radius_ * segmentLength;
                D3 Fdq = -(1.0 - f1) * f2 * 0.5 * o->density * Cd * S * dnorm(relV) * relV;
                add_half(out[i].Fdq, Fdq);
                add_half(out[i + 1].Fdq, Fdq);
            }
            {
                double S = f2 * ((2 * M_PI) * radius_ * segmentLength);
                D3 Fdf = -S * relV;
                add_half(out[i].Fdf, Fdf);
                add_half(out[i + 1].Fdf, Fdf);
            }
        }
    }
    /* :363-371: end nodes carry one segment only */
    for (int c = 0; c < 3; ++c) {
        out[0].Fb[c] *= 2.0; out[n - 1].Fb[c] *= 2.0;
        out[0].Fdq[c] *= 2.0; out[n - 1].Fdq[c] *= 2.0;
        out[0].Fdf[c] *= 2.0; out[n - 1].Fdf[c] *= 2.0;
    }
}

} // extern "C"
