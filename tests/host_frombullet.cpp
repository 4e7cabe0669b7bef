// Compile-and-run check of the Bullet entry points of the C++ host mirror (sfb::Transform::FromBullet,
// sfb::Vector3::FromBullet) against the reference's OWN vendored Bullet LinearMath headers, plus the pure host-side helpers
// (field constructors, Ocean::Fill, Atmosphere::Fill, WaveParamsFromSeaState, SumCompoundPartsAero) — no GPU involved.
// Built by tests/test_host_mirror_cpu.py only where /root/reference is mounted; prints "ok".
#include <cmath>
#include <cstdio>

#include "LinearMath/btQuaternion.h"
#include "LinearMath/btTransform.h"

#include "../stonefish_b200/host/StonefishHydro.hpp"

#define CHECK(c) do { if (!(c)) { std::printf("FAILED: %s (line %d)\n", #c, __LINE__); return 1; } } while (0)

int main()
{
    // what SolidEntity::getCGTransform() hands out (double-precision Bullet, CMakeLists.txt:53)
    const btTransform T(btQuaternion(0.3, -0.2, 0.1), btVector3(1.5, -2.5, 3.5));
    const sfb::Transform F = sfb::Transform::FromBullet(T);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) CHECK(F.m[3 * i + j] == T.getBasis()[i][j]);
    CHECK(F.m[9] == 1.5 && F.m[10] == -2.5 && F.m[11] == 3.5);
    const sfb::Vector3 v = sfb::Vector3::FromBullet(btVector3(0.1, 0.2, -0.3));
    CHECK(v.x == 0.1 && v.y == 0.2 && v.z == -0.3);

    // constructor-time normalisation like Jet.cpp:31-37 / Pipe.cpp:31-40, with Bullet doing the same arithmetic
    sfh_ocean o;
    sfb::Ocean ocean(1025.0, 9.81);
    ocean.AddVelocityField(new sfb::Uniform(sfb::Vector3(0.5, -0.3, 0.0)));
    ocean.AddVelocityField(new sfb::Jet(sfb::Vector3(0.0, -3.0, 1.0), sfb::Vector3(0.1, 1.0, 0.05), 0.3, 5.0));
    ocean.AddVelocityField(new sfb::Pipe(sfb::Vector3(-6.0, 0.5, 1.5), sfb::Vector3(6.0, -0.5, 2.0), 2.0, 3.0, 2.0, 2.0));
    ocean.EnableCurrents();
    ocean.Fill(o);
    CHECK(o.n_currents == 3 && o.currents_enabled == 1 && o.has_waves == 0 && o.density == 1025.0);
    const btVector3 jn = btVector3(0.1, 1.0, 0.05).normalized();
    CHECK(o.currents[1].type == SFH_CURRENT_JET && o.currents[1].n[0] == jn.x() && o.currents[1].n[1] == jn.y() && o.currents[1].n[2] == jn.z());
    const btVector3 p1(-6.0, 0.5, 1.5), p2(6.0, -0.5, 2.0);
    const btScalar l = (p2 - p1).length();
    const btVector3 pn = (p2 - p1) / l;
    CHECK(o.currents[2].type == SFH_CURRENT_PIPE && o.currents[2].l == l && o.currents[2].n[0] == pn.x() && o.currents[2].n[2] == pn.z());
    ocean.useGeneratedWaves(true);
    ocean.Fill(o);
    CHECK(o.has_waves == 1);

    sfh_atmosphere a;
    sfb::Atmosphere atm(1.2);
    atm.AddVelocityField(new sfb::Uniform(sfb::Vector3(4.0, 0.0, 0.0)));
    atm.Fill(a);
    CHECK(a.n_winds == 1 && a.density == 1.2 && a.winds[0].a[0] == 4.0);

    const sfh_wave_params wp = sfb::WaveParamsFromSeaState(1.0f);
    CHECK(wp.wind == 7.0f && wp.omega == 5.f * expf(-1.0f) + 0.2f && wp.passes == 8 && wp.seed == 1234 && wp.grid_sizes[0] == 893.f);

    sfb::SolidBody b1, b2, b3;
    b1.Fda = sfb::Vector3(1, 2, 3); b2.Fda = sfb::Vector3(10, 20, 30); b3.Fda = sfb::Vector3(100, 200, 300);
    b1.Tda = sfb::Vector3(-1, 0, 1); b3.Tda = sfb::Vector3(5, 5, 5);
    std::vector<sfb::CompoundPart> parts = {{&b1, true}, {&b2, false}, {&b3, true}};
    sfb::Vector3 Fda, Tda;
    sfb::SumCompoundPartsAero(parts, Fda, Tda);
    CHECK(Fda.x == 101 && Fda.y == 202 && Fda.z == 303 && Tda.x == 4 && Tda.z == 6);
    std::printf("ok\n");
    return 0;
}
