/*
 * bt_check.cpp — the fp64 pieces of the hot path evaluated with the reference's OWN vendored Bullet
 * LinearMath headers (3rdparty/LinearMath, BT_USE_DOUBLE_PRECISION as in CMakeLists.txt:53).
 *
 * TEST INFRASTRUCTURE ONLY. Built into oracle/_ref/libbtcheck.so by oracle/Makefile from the headers where
 * they lie under /root/reference (never copied). It pins the oracle's hand-written double-precision
 * operator order (btVector3 dot/cross/norm/safeNormalize, btMatrix3x3::inverse, btTransform::operator(),
 * getOpenGLMatrix) against the real classes. The statement bodies mirror the reference lines cited below;
 * only the carrier (sfh_current instead of sf::Jet members) differs, because the reference's
 * Jet.cpp/Pipe.cpp/Ocean.cpp/SolidEntity.cpp need glm + SDL2 + GL + a cmake-generated version.h and do
 * not compile in this image.
 */
#include "../include/sfh_types.h"

#include "LinearMath/btMatrix3x3.h"
#include "LinearMath/btQuaternion.h"
#include "LinearMath/btTransform.h"
#include "LinearMath/btVector3.h"

#include <algorithm>
#include <vector>

typedef btScalar Scalar;
typedef btVector3 Vector3;
typedef btMatrix3x3 Matrix3;
typedef btTransform Transform;

static Transform make_transform(const double T[12])
{
    return Transform(Matrix3(T[0], T[1], T[2], T[3], T[4], T[5], T[6], T[7], T[8]), Vector3(T[9], T[10], T[11]));
}

extern "C" {

int btc_scalar_size(void) { return (int)sizeof(btScalar); }

/* Uniform.cpp:54-57, Jet.cpp:54-75, Pipe.cpp:57-78 */
void btc_current_velocity(const sfh_current* cur, const double pp[3], double out[3])
{
    Vector3 p(pp[0], pp[1], pp[2]);
    Vector3 res(0, 0, 0);
    if (cur->type == SFH_CURRENT_UNIFORM) {
        res = Vector3(cur->a[0], cur->a[1], cur->a[2]);
    } else if (cur->type == SFH_CURRENT_JET) {
        Vector3 c(cur->a[0], cur->a[1], cur->a[2]), n(cur->n[0], cur->n[1], cur->n[2]);
        Scalar r = cur->r1, vout = cur->vel;
        Vector3 cp = p - c;
        Scalar d = cp.cross(n).norm();
        Scalar t = cp.dot(n);
        if (t < 0.0) { res = Vector3(0, 0, 0); }
        else {
            Scalar r_ = Scalar(1) / Scalar(5) * (t + Scalar(5) * r);
            if (d >= r_) { res = Vector3(0, 0, 0); }
            else {
                Vector3 vmax = Scalar(10) * r / (t + Scalar(5) * r) * vout * n;
                Scalar f = btExp(-Scalar(50) * d * d / (t * t));
                res = f * vmax;
            }
        }
    } else {
        Vector3 p1(cur->a[0], cur->a[1], cur->a[2]), n(cur->n[0], cur->n[1], cur->n[2]);
        Scalar l = cur->l, r1 = cur->r1, r2 = cur->r2, vin = cur->vel, gamma = cur->gamma;
        Vector3 p1p = p - p1;
        Scalar d = p1p.cross(n).norm();
        Scalar t = p1p.dot(n);
        if (t < 0.0 || t > l) { res = Vector3(0, 0, 0); }
        else {
            Scalar r = r1 + (r2 - r1) * t / l;
            if (d >= r) { res = Vector3(0, 0, 0); }
            else {
                Vector3 v = r1 / r * vin * n;
                Scalar f = btPow(Scalar(1) - d / r, gamma);
                res = f * v;
            }
        }
    }
    out[0] = res.x(); out[1] = res.y(); out[2] = res.z();
}

/* Jet.cpp:31-37 and Pipe.cpp:31-40: constructor-time normalisation of the axis */
void btc_jet_axis(const double dir[3], double n_out[3])
{
    Vector3 n = Vector3(dir[0], dir[1], dir[2]).normalized();
    n_out[0] = n.x(); n_out[1] = n.y(); n_out[2] = n.z();
}
void btc_pipe_axis(const double p1[3], const double p2[3], double n_out[3], double* l_out)
{
    Vector3 point1(p1[0], p1[1], p1[2]), point2(p2[0], p2[1], p2[2]);
    Scalar l = (point2 - point1).length();
    Vector3 n = (point2 - point1) / l;
    n_out[0] = n.x(); n_out[1] = n.y(); n_out[2] = n.z();
    *l_out = l;
}

/* Ocean.cpp:183-196 */
void btc_fluid_velocity(const sfh_ocean* o, const double pp[3], double out[3])
{
    Vector3 fv(0, 0, 0);
    if (o->currents_enabled) {
        for (size_t i = 0; i < o->n_currents; ++i) {
            if (o->currents[i].enabled) {
                double v[3];
                btc_current_velocity(&o->currents[i], pp, v);
                fv += Vector3(v[0], v[1], v[2]);
            }
        }
    }
    out[0] = fv.x(); out[1] = fv.y(); out[2] = fv.z();
}

/* SolidEntity.cpp:1263-1287 */
void btc_correct(double density, double Fdq_[3], double Tdq_[3], double Fdf_[3], double Tdf_[3], const double Cd[3],
                 const double Cf[3], const double B[9])
{
    Transform T_O(Matrix3(B[0], B[1], B[2], B[3], B[4], B[5], B[6], B[7], B[8]), Vector3(0, 0, 0));
    Vector3 _Fdq(Fdq_[0], Fdq_[1], Fdq_[2]), _Tdq(Tdq_[0], Tdq_[1], Tdq_[2]);
    Vector3 _Fdf(Fdf_[0], Fdf_[1], Fdf_[2]), _Tdf(Tdf_[0], Tdf_[1], Tdf_[2]);
    Vector3 fdCd(Cd[0], Cd[1], Cd[2]), fdCf(Cf[0], Cf[1], Cf[2]);

    Matrix3 toOrigin = T_O.getBasis().inverse();

    Vector3 Fdq = toOrigin * _Fdq;
    Fdq = Fdq.safeNormalize();
    Scalar Fdqc = btFabs(Fdq.getX()) * fdCd.getX() + btFabs(Fdq.getY()) * fdCd.getY() + btFabs(Fdq.getZ()) * fdCd.getZ();
    _Fdq = Scalar(0.5) * density * Fdqc * _Fdq;

    Vector3 Tdq = toOrigin * _Tdq;
    Tdq = Tdq.safeNormalize();
    Scalar Tdqc = btFabs(Tdq.getX()) * fdCd.getX() + btFabs(Tdq.getY()) * fdCd.getY() + btFabs(Tdq.getZ()) * fdCd.getZ();
    _Tdq = Scalar(0.5) * density * Tdqc * _Tdq;

    Vector3 Fdf = toOrigin * _Fdf;
    Fdf = Fdf.safeNormalize();
    Scalar Fdfc = btFabs(Fdf.getX()) * fdCf.getX() + btFabs(Fdf.getY()) * fdCf.getY() + btFabs(Fdf.getZ()) * fdCf.getZ();
    _Fdf = density * Fdfc * _Fdf;

    Vector3 Tdf = toOrigin * _Tdf;
    Tdf = Tdf.safeNormalize();
    Scalar Tdfc = btFabs(Tdf.getX()) * fdCf.getX() + btFabs(Tdf.getY()) * fdCf.getY() + btFabs(Tdf.getZ()) * fdCf.getZ();
    _Tdf = density * Tdfc * _Tdf;

    for (int k = 0; k < 3; ++k) { Fdq_[k] = _Fdq[k]; Tdq_[k] = _Tdq[k]; Fdf_[k] = _Fdf[k]; Tdf_[k] = _Tdf[k]; }
}

/* SolidEntity.cpp:1925-1956. approx_type: 0 AUTO, 1 SPHERE, 2 CYLINDER, 3 ELLIPSOID (SolidEntity.h:39) */
void btc_correct_aero(double Fda_[3], double Tda_[3], int approx_type, const double params[3], const double B[9])
{
    Transform T_CG2H(Matrix3(B[0], B[1], B[2], B[3], B[4], B[5], B[6], B[7], B[8]), Vector3(0, 0, 0));
    Vector3 _Fda(Fda_[0], Fda_[1], Fda_[2]), _Tda(Tda_[0], Tda_[1], Tda_[2]);
    const double* fdApproxParams = params;

    Vector3 Fdan = (T_CG2H.getBasis().inverse() * _Fda).safeNormalize();
    Scalar corFactor(1.0);
    switch (approx_type) {
    case 0:
    case 1:
        break;
    case 2: {
        Vector3 Cd(0.5, 0.5, 1.0);
        corFactor = Cd.dot(Fdan);
    } break;
    case 3: {
        Vector3 Cd(Scalar(1) / fdApproxParams[0], Scalar(1) / fdApproxParams[1], Scalar(1) / fdApproxParams[2]);
        Scalar maxCd = btMax(btMax(Cd.x(), Cd.y()), Cd.z());
        Cd /= maxCd;
        corFactor = Cd.dot(Fdan);
    } break;
    }
    _Fda *= btFabs(corFactor);
    _Tda *= btFabs(corFactor);
    for (int k = 0; k < 3; ++k) { Fda_[k] = _Fda[k]; Tda_[k] = _Tda[k]; }
}
/* the ELLIPSOID coefficient vector alone (SolidEntity.cpp:1945-1947), for the host mirrors */
void btc_ellipsoid_cd(const double params[3], double cd_out[3])
{
    Vector3 Cd(Scalar(1) / params[0], Scalar(1) / params[1], Scalar(1) / params[2]);
    Scalar maxCd = btMax(btMax(Cd.x(), Cd.y()), Cd.z());
    Cd /= maxCd;
    for (int k = 0; k < 3; ++k) cd_out[k] = Cd[k];
}

/* CableEntity.cpp:217-372 with the real btVector3 / btScalar helpers, flat sea (Ocean::GetDepth(Vector3) narrows to float,
 * Ocean.cpp:149-173). flags: SFH_F_DAMPING, SFH_F_BUOYANCY (settings), SFH_F_BUOYANT (phy_.buoyancy). */
static Scalar btc_circularSegmentArea(Scalar radius_, Scalar h)
{
    if (h <= Scalar(0))
        return Scalar(0);
    else if (h >= Scalar(2) * radius_)
        return Scalar(M_PI) * radius_ * radius_;
    else {
        if (h > radius_) {
            h = Scalar(2) * radius_ - h;
            Scalar theta = Scalar(2) * btAcos((radius_ - h) / radius_);
            return radius_ * radius_ * (Scalar(M_PI) - (theta - btSin(theta)) / Scalar(2));
        } else {
            Scalar theta = Scalar(2) * btAcos((radius_ - h) / radius_);
            return (radius_ * radius_ / Scalar(2)) * (theta - btSin(theta));
        }
    }
}

void btc_cable(const sfh_ocean* ocn, uint32_t n_nodes, double radius_, uint32_t flags, const double* pos, const double* vel, double* out9)
{
    struct NodalForces { Vector3 Fb, Fdq, Fdf; };
    std::vector<NodalForces> nodalForces_(n_nodes);
    for (size_t i = 0; i < nodalForces_.size(); ++i) { nodalForces_[i].Fb.setZero(); nodalForces_[i].Fdq.setZero(); nodalForces_[i].Fdf.setZero(); }
    auto GetDepth = [](const Vector3& point) { return Scalar((float)point.getZ()); };
    const Vector3 gravity(0, 0, ocn->gravity_z);
    const bool buoyancy = (flags & SFH_F_BUOYANT) != 0, reallisticBuoyancy = (flags & SFH_F_BUOYANCY) != 0, dampingForces = (flags & SFH_F_DAMPING) != 0;

    for (size_t i = 0; i < n_nodes - 1; ++i) {
        Vector3 p1(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
        Vector3 p2(pos[3 * i + 3], pos[3 * i + 4], pos[3 * i + 5]);
        Scalar d1 = GetDepth(p1);
        Scalar d2 = GetDepth(p2);
        if (d1 > d2) {
            std::swap(p1, p2);
            std::swap(d1, d2);
        }
        Scalar dz = d2 - d1;
        Scalar dxy = Vector3(p2.getX() - p1.getX(), p2.getY() - p1.getY(), 0.0).safeNorm();
        Vector3 segmentVec = p2 - p1;
        Scalar segmentLength = segmentVec.length();
        Scalar alpha = btAtan(dz / dxy);
        Scalar margin = btFuzzyZero(dxy) ? Scalar(0) : radius_ * btCos(alpha);
        Scalar submergedV{0};
        if (d2 < -margin) {
            continue;
        } else if (d1 > margin) {
            submergedV = Scalar(M_PI) * radius_ * radius_ * segmentLength;
        } else {
            if (btFuzzyZero(dxy)) {
                Scalar submergedL = btMin(d2, dz);
                submergedV = Scalar(M_PI) * radius_ * radius_ * submergedL;
            } else if (btFuzzyZero(dz)) {
                Scalar submergedA = btc_circularSegmentArea(radius_, d1 + radius_);
                submergedV = submergedA * segmentLength;
            } else if (btFabs(d1) <= margin && btFabs(d2) <= margin) {
                Scalar A1 = btc_circularSegmentArea(radius_, d1 / btCos(alpha) + radius_);
                Scalar A2 = btc_circularSegmentArea(radius_, d2 / btCos(alpha) + radius_);
                submergedV = (A1 + A2) / Scalar(2) * segmentLength;
            } else if (d1 < -margin && d2 > margin) {
                Scalar l1 = (d2 - radius_ * btCos(alpha)) / btSin(alpha);
                Scalar l2 = Scalar(2) * radius_ * btCos(alpha);
                submergedV = Scalar(M_PI) * radius_ * radius_ * (l1 + l2 / Scalar(2));
            } else if (d2 <= margin) {
                Scalar a = radius_ + d2 / btCos(alpha);
                Scalar l = a / btSin(alpha);
                submergedV = btc_circularSegmentArea(radius_, a) * l / Scalar(2);
            } else if (d1 >= -margin) {
                Scalar a = radius_ - d1 / btCos(alpha);
                Scalar l = a / btSin(alpha);
                submergedV = Scalar(M_PI) * radius_ * radius_ * segmentLength - btc_circularSegmentArea(radius_, a) * l / Scalar(2);
            }
        }
        if (buoyancy && reallisticBuoyancy) {
            Vector3 b = -submergedV * ocn->density * gravity;
            nodalForces_[i].Fb += b / Scalar(2);
            nodalForces_[i + 1].Fb += b / Scalar(2);
        }
        if (dampingForces && submergedV > Scalar(0)) {
            Vector3 p = (p1 + p2) / Scalar(2);
            Vector3 v1(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]);
            Vector3 v2(vel[3 * i + 3], vel[3 * i + 4], vel[3 * i + 5]);
            Vector3 v = (v1 + v2) / Scalar(2);
            double pp[3] = {p.x(), p.y(), p.z()}, wv[3];
            btc_fluid_velocity(ocn, pp, wv);
            Vector3 waterV(wv[0], wv[1], wv[2]);
            Vector3 relV = v - waterV;
            const Scalar Cd{0.5};
            Scalar f1 = btFabs(segmentVec.dot(relV));
            Scalar f2 = submergedV / (Scalar(M_PI) * radius_ * radius_ * segmentLength);
            {
                Scalar S = Scalar(2) * radius_ * segmentLength;
                Vector3 Fdq = -(Scalar(1) - f1) * f2 * Scalar(0.5) * ocn->density * Scalar(Cd) * S * relV.length() * relV;
                nodalForces_[i].Fdq += Fdq / Scalar(2);
                nodalForces_[i + 1].Fdq += Fdq / Scalar(2);
            }
            {
                Scalar S = f2 * (Scalar(2 * M_PI) * radius_ * segmentLength);
                Vector3 Fdf = -S * relV;
                nodalForces_[i].Fdf += Fdf / Scalar(2);
                nodalForces_[i + 1].Fdf += Fdf / Scalar(2);
            }
        }
    }
    nodalForces_[0].Fb *= Scalar(2);
    nodalForces_[n_nodes - 1].Fb *= Scalar(2);
    nodalForces_[0].Fdq *= Scalar(2);
    nodalForces_[n_nodes - 1].Fdq *= Scalar(2);
    nodalForces_[0].Fdf *= Scalar(2);
    nodalForces_[n_nodes - 1].Fdf *= Scalar(2);
    for (uint32_t i = 0; i < n_nodes; ++i)
        for (int k = 0; k < 3; ++k) {
            out9[9 * i + k] = nodalForces_[i].Fb[k];
            out9[9 * i + 3 + k] = nodalForces_[i].Fdq[k];
            out9[9 * i + 6 + k] = nodalForces_[i].Fdf[k];
        }
}

/* OpenGLDataStructs.cpp:148-156: getOpenGLMatrix then (GLfloat) per element */
void btc_gl_matrix(const double T[12], float out[16])
{
    Scalar glmatrix[16];
    make_transform(T).getOpenGLMatrix(glmatrix);
    for (int k = 0; k < 16; ++k) out[k] = (float)glmatrix[k];
}

/* SolidEntity.cpp:1833-1834: INSIDE analytic buoyancy */
void btc_inside_buoyancy(const double T_CG_[12], const double P_CB_[3], double volume, double density, double g,
                         double Fb_[3], double Tb_[3])
{
    Transform T_CG = make_transform(T_CG_);
    Vector3 P_CB(P_CB_[0], P_CB_[1], P_CB_[2]);
    Vector3 gravity(0, 0, g);
    Vector3 Fb = -volume * density * gravity;
    Vector3 Tb = (T_CG * P_CB - T_CG.getOrigin()).cross(Fb);
    for (int k = 0; k < 3; ++k) { Fb_[k] = Fb[k]; Tb_[k] = Tb[k]; }
}

/* SolidEntity.cpp:1703-1706 (flat-sea branch), fp64 tail only */
void btc_flat_buoyancy(const double T_CG_[12], const double CBsub_[3], double Vsub, double density, double g,
                       double Fb_[3], double Tb_[3])
{
    Transform T_CG = make_transform(T_CG_);
    Vector3 _CBsub(CBsub_[0], CBsub_[1], CBsub_[2]);
    Vector3 gravity(0, 0, g);
    Vector3 _Fb = -Vsub * density * gravity;
    Vector3 _Tb = (_CBsub - T_CG.getOrigin()).cross(_Fb);
    for (int k = 0; k < 3; ++k) { Fb_[k] = _Fb[k]; Tb_[k] = _Tb[k]; }
}

/* helpers to build inputs the way reference scenes do: Transform(Quaternion(yaw,pitch,roll), origin)
 * with BT_EULER_DEFAULT_ZYX (CMakeLists.txt:53), and transform composition T_A * T_B */
void btc_transform_from_euler(double yaw, double pitch, double roll, const double origin[3], double out[12])
{
    btQuaternion q(yaw, pitch, roll);
    Transform T(q, Vector3(origin[0], origin[1], origin[2]));
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) out[3 * r + c] = T.getBasis()[r][c];
    for (int k = 0; k < 3; ++k) out[9 + k] = T.getOrigin()[k];
}
void btc_transform_from_quat(const double q_[4], const double origin[3], double out[12])
{
    btQuaternion q(q_[0], q_[1], q_[2], q_[3]);
    Transform T(q, Vector3(origin[0], origin[1], origin[2]));
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) out[3 * r + c] = T.getBasis()[r][c];
    for (int k = 0; k < 3; ++k) out[9 + k] = T.getOrigin()[k];
}
void btc_transform_mul(const double A[12], const double B[12], double out[12])
{
    Transform T = make_transform(A) * make_transform(B);
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) out[3 * r + c] = T.getBasis()[r][c];
    for (int k = 0; k < 3; ++k) out[9 + k] = T.getOrigin()[k];
}

} // extern "C"
