// sfh_cable.cuh — cables: CableEntity::ComputeHydrodynamicForces (CableEntity.cpp:239-372), fp64.
//
// The reference walks the segments of a cable in order and adds half of each segment's buoyancy / form drag / skin friction
// to its two nodes. Here one thread owns one NODE and evaluates its (at most) two adjacent segments itself — node j receives
// segment j-1's half first and segment j's second, exactly the reference's order — so there is no scatter, no atomics, and the
// doubling of the end nodes (:363-371) is a local multiply. Each segment is evaluated twice; a cable has a few hundred.
#pragma once

#include "sfh_device.cuh"

namespace sfh {

struct SegForces { D3 Fb, Fdq, Fdf; bool wet; };

// CableEntity::circularSegmentArea (:217-237)
__device__ inline double circular_segment_area(double radius, double h)
{
    if (h <= 0.0) return 0.0;
    if (h >= dmul(2.0, radius)) return dmul(dmul(M_PI, radius), radius);
    if (h > radius) {
        h = dsub(dmul(2.0, radius), h);
        const double theta = dmul(2.0, acos(ddiv(dsub(radius, h), radius)));
        return dmul(dmul(radius, radius), dsub(M_PI, ddiv(dsub(theta, sin(theta)), 2.0)));
    }
    const double theta = dmul(2.0, acos(ddiv(dsub(radius, h), radius)));
    return dmul(ddiv(dmul(radius, radius), 2.0), dsub(theta, sin(theta)));
}

// Ocean::GetDepth(const Vector3&) (Ocean.cpp:170-173): the point is narrowed to fp32 first
__device__ inline double cable_depth(bool waves, const WavesDev& wv, D3 p)
{
    const float x = (float)p.x, y = (float)p.y, z = (float)p.z;
    return (double)(waves ? get_depth<true>(wv, x, y, z) : z);
}

// one segment (:249-360): the forces the reference splits between its two nodes
__device__ inline SegForces cable_segment(const OceanDev& oc, const WavesDev& wv, double radius, bool buoyancy, bool damping, D3 p1, D3 p2,
                                          D3 v1, D3 v2)
{
    const double EPS = 2.2204460492503131e-16; // SIMD_EPSILON
    SegForces F = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, false};
    const bool waves = oc.has_waves != 0;
    double d1 = cable_depth(waves, wv, p1), d2 = cable_depth(waves, wv, p2);
    if (d1 > d2) { // the shallower point first; the node velocities keep their order (only their mean is used)
        const D3 tp = p1; p1 = p2; p2 = tp;
        const double td = d1; d1 = d2; d2 = td;
    }
    const double dz = dsub(d2, d1);
    const D3 hv = {dsub(p2.x, p1.x), dsub(p2.y, p1.y), 0.0};
    const double hl2 = ddot(hv, hv);
    const double dxy = hl2 > EPS ? __dsqrt_rn(hl2) : 0.0; // btVector3::safeNorm
    const D3 seg = p2 - p1;
    const double len = dnorm(seg);
    const double alpha = atan(ddiv(dz, dxy));
    const bool vertical = fabs(dxy) < EPS;
    const double ca = cos(alpha), sa = sin(alpha);
    const double margin = vertical ? 0.0 : dmul(radius, ca);
    const double disk = dmul(dmul(M_PI, radius), radius);

    double V = 0.0;
    if (d2 < -margin) return F; // out of the water: `continue`
    if (d1 > margin) {
        V = dmul(disk, len);
    } else if (vertical) {
        V = dmul(disk, fmin(d2, dz));
    } else if (fabs(dz) < EPS) {
        V = dmul(circular_segment_area(radius, dadd(d1, radius)), len);
    } else if (fabs(d1) <= margin && fabs(d2) <= margin) {
        const double A1 = circular_segment_area(radius, dadd(ddiv(d1, ca), radius));
        const double A2 = circular_segment_area(radius, dadd(ddiv(d2, ca), radius));
        V = dmul(ddiv(dadd(A1, A2), 2.0), len);
    } else if (d1 < -margin && d2 > margin) {
        const double l1 = ddiv(dsub(d2, dmul(radius, ca)), sa);
        const double l2 = dmul(dmul(2.0, radius), ca);
        V = dmul(disk, dadd(l1, ddiv(l2, 2.0)));
    } else if (d2 <= margin) {
        const double a = dadd(radius, ddiv(d2, ca));
        const double l = ddiv(a, sa);
        V = ddiv(dmul(circular_segment_area(radius, a), l), 2.0);
    } else if (d1 >= -margin) {
        const double a = dsub(radius, ddiv(d1, ca));
        const double l = ddiv(a, sa);
        V = dsub(dmul(disk, len), ddiv(dmul(circular_segment_area(radius, a), l), 2.0));
    }
    F.wet = true;
    if (buoyancy) { // -submergedV * density * gravity, gravity = (0, 0, g)
        const double s = dmul(-V, oc.density);
        F.Fb = D3{dmul(0.0, s), dmul(0.0, s), dmul(oc.gravity_z, s)};
    }
    if (damping && V > 0.0) {
        const D3 p = (p1 + p2) * 0.5;
        const D3 v = (v1 + v2) * 0.5;
        D3 water = {0.0, 0.0, 0.0}; // Ocean::GetFluidVelocity(const Vector3&) (Ocean.cpp:183-196)
        if (oc.currents_enabled)
            for (uint32_t k = 0; k < oc.n_currents; ++k)
                if (oc.currents[k].enabled) water = water + current_velocity(oc.currents[k], p);
        const D3 rel = v - water;
        const double f1 = fabs(ddot(seg, rel));
        const double f2 = ddiv(V, dmul(disk, len));
        {
            const double S = dmul(dmul(2.0, radius), len);
            const double k = dmul(dmul(dmul(dmul(dmul(dmul(-dsub(1.0, f1), f2), 0.5), oc.density), 0.5), S), dnorm(rel));
            F.Fdq = rel * k;
        }
        {
            const double S = dmul(f2, dmul(dmul(2 * M_PI, radius), len));
            F.Fdf = rel * (-S);
        }
    }
    return F;
}

struct CableParams {
    const sfh_cable* cables;
    uint32_t n_cables;
    const double* pos;
    const double* vel;
    const OceanDev* ocean;
    WavesDev waves;
    sfh_node_forces* out;
};

// grid.y = cable, threads over its nodes
__global__ void __launch_bounds__(128) k_cable_nodes(CableParams P)
{
    const sfh_cable c = P.cables[blockIdx.y];
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= c.n_nodes) return;
    const OceanDev& oc = *P.ocean;
    const bool buoyancy = (c.flags & SFH_F_BUOYANT) && (c.flags & SFH_F_BUOYANCY), damping = (c.flags & SFH_F_DAMPING) != 0;
    auto node = [&](const double* a, uint32_t k) { const double* q = a + 3 * (size_t)(c.first_node + k); return D3{q[0], q[1], q[2]}; };
    D3 Fb = {0, 0, 0}, Fdq = {0, 0, 0}, Fdf = {0, 0, 0};
    auto add_half = [](D3& acc, D3 F) { acc = acc + F * 0.5; }; // X += F / Scalar(2): operator/ multiplies by 1/s
    if (j > 0) {
        const SegForces s = cable_segment(oc, P.waves, c.radius, buoyancy, damping, node(P.pos, j - 1), node(P.pos, j), node(P.vel, j - 1), node(P.vel, j));
        if (s.wet) { add_half(Fb, s.Fb); add_half(Fdq, s.Fdq); add_half(Fdf, s.Fdf); }
    }
    if (j + 1 < c.n_nodes) {
        const SegForces s = cable_segment(oc, P.waves, c.radius, buoyancy, damping, node(P.pos, j), node(P.pos, j + 1), node(P.vel, j), node(P.vel, j + 1));
        if (s.wet) { add_half(Fb, s.Fb); add_half(Fdq, s.Fdq); add_half(Fdf, s.Fdf); }
    }
    if (j == 0 || j + 1 == c.n_nodes) { Fb = Fb * 2.0; Fdq = Fdq * 2.0; Fdf = Fdf * 2.0; } // :363-371
    sfh_node_forces o;
    o.Fb[0] = Fb.x; o.Fb[1] = Fb.y; o.Fb[2] = Fb.z;
    o.Fdq[0] = Fdq.x; o.Fdq[1] = Fdq.y; o.Fdq[2] = Fdq.z;
    o.Fdf[0] = Fdf.x; o.Fdf[1] = Fdf.y; o.Fdf[2] = Fdf.z;
    P.out[c.first_node + j] = o;
}

} // namespace sfh
