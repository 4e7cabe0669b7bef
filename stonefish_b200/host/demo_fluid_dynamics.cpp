// demo_fluid_dynamics.cpp — the reference's FluidDynamicsTest loop shape on the C++ host mirror:
//   build scenario (meshes, bodies, ocean, currents)  ->  every step: ComputeHydrodynamicForces for all bodies  ->
//   print "Hydrodynamics time: <avg> us" like Tests/FluidDynamicsTest/FluidDynamicsTestManager.cpp:166.
// The scenario is read from a small binary file written by tests/test_host_mirror_gpu.py (so the same inputs can be
// pushed through the Python binding and the oracle), results go to a binary file. Needs a CUDA device at run time.
//   usage: demo_fluid_dynamics scene.bin out.bin [steps [aero_out.bin [sea_state waves_out.bin [cable_out.bin]]]]
// With aero_out.bin the same bodies also go through HydroEngine::ComputeAerodynamicForces once: wind = the scene's velocity
// fields, gas density 1.2, proxy type cycling AUTO/SPHERE/CYLINDER/ELLIPSOID with fdCd as the ellipsoid radii, T_CG2H = T_O.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>

#include "StonefishHydro.hpp"

template <class T>
static T rd(std::istream& in)
{
    T v;
    in.read(reinterpret_cast<char*>(&v), sizeof(T));
    if (!in) { std::cerr << "scene file truncated\n"; std::exit(2); }
    return v;
}
static sfb::Vector3 rd3(std::istream& in) { double x = rd<double>(in), y = rd<double>(in), z = rd<double>(in); return sfb::Vector3(x, y, z); }
static sfb::Transform rdT(std::istream& in) { sfb::Transform t; for (double& e : t.m) e = rd<double>(in); return t; }

int main(int argc, char** argv)
{
    if (argc < 3) { std::cerr << "usage: demo_fluid_dynamics scene.bin out.bin [steps]\n"; return 2; }
    const int steps = argc > 3 ? std::atoi(argv[3]) : 10;
    std::ifstream in(argv[1], std::ios::binary);
    if (!in || rd<uint32_t>(in) != 0x53484653u) { std::cerr << "bad scene file\n"; return 2; }
    const uint32_t nMeshes = rd<uint32_t>(in), nBodies = rd<uint32_t>(in);
    const double density = rd<double>(in), gravity = rd<double>(in);
    const uint32_t currentsEnabled = rd<uint32_t>(in), nCurrents = rd<uint32_t>(in);

    try {
        sfb::HydroEngine engine(0);
        sfb::Ocean ocean(density, gravity);
        sfb::Atmosphere atmosphere(1.2);
        for (uint32_t i = 0; i < nCurrents; ++i) {
            const uint32_t type = rd<uint32_t>(in), enabled = rd<uint32_t>(in);
            double a[10];
            for (double& e : a) e = rd<double>(in);
            sfb::VelocityField* f = nullptr;
            if (type == 0) f = new sfb::Uniform(sfb::Vector3(a[0], a[1], a[2]));
            else if (type == 1) f = new sfb::Jet(sfb::Vector3(a[0], a[1], a[2]), sfb::Vector3(a[3], a[4], a[5]), a[6], a[7]);
            else f = new sfb::Pipe(sfb::Vector3(a[0], a[1], a[2]), sfb::Vector3(a[3], a[4], a[5]), a[6], a[7], a[8], a[9]);
            f->setEnabled(enabled != 0);
            ocean.AddVelocityField(f);
            sfb::VelocityField* w = nullptr; // the atmosphere owns its own copies
            if (type == 0) w = new sfb::Uniform(sfb::Vector3(a[0], a[1], a[2]));
            else if (type == 1) w = new sfb::Jet(sfb::Vector3(a[0], a[1], a[2]), sfb::Vector3(a[3], a[4], a[5]), a[6], a[7]);
            else w = new sfb::Pipe(sfb::Vector3(a[0], a[1], a[2]), sfb::Vector3(a[3], a[4], a[5]), a[6], a[7], a[8], a[9]);
            w->setEnabled(enabled != 0);
            atmosphere.AddVelocityField(w);
        }
        if (currentsEnabled) ocean.EnableCurrents();

        std::vector<uint32_t> meshIds;
        for (uint32_t i = 0; i < nMeshes; ++i) {
            const uint32_t nv = rd<uint32_t>(in), stride = rd<uint32_t>(in), nf = rd<uint32_t>(in);
            std::vector<char> verts((size_t)nv * stride), faces((size_t)nf * 12);
            in.read(verts.data(), verts.size());
            in.read(faces.data(), faces.size());
            meshIds.push_back(engine.UploadMesh(verts.data(), stride, nv, faces.data(), nf)); // as Mesh::getVertexDataPointer() etc.
        }
        std::vector<sfb::SolidBody> bodies(nBodies);
        std::vector<sfb::SolidBody*> ptrs;
        for (auto& b : bodies) {
            b.meshId = meshIds.at(rd<uint32_t>(in));
            b.buoyant = rd<uint32_t>(in) != 0;
            b.T_CG = rdT(in); b.T_C = rdT(in); b.T_O = rdT(in);
            b.v = rd3(in); b.omega = rd3(in); b.aabbMin = rd3(in); b.aabbMax = rd3(in);
            b.fdCd = rd3(in); b.fdCf = rd3(in);
            b.volume = rd<double>(in); b.surface = rd<double>(in);
            b.P_CB = rd3(in);
            ptrs.push_back(&b);
        }

        sfb::HydrodynamicsSettings settings; // {true, true} as Ocean::ApplyFluidForces sets them (Ocean.cpp:250-251)
        double total_us = 0;
        for (int s = 0; s < steps; ++s) {
            auto t0 = std::chrono::steady_clock::now();
            engine.ComputeHydrodynamicForces(settings, ocean, ptrs);
            auto t1 = std::chrono::steady_clock::now();
            if (s > 0) total_us += std::chrono::duration<double, std::micro>(t1 - t0).count();
        }
        std::cout << "Hydrodynamics time: " << (steps > 1 ? total_us / (steps - 1) : 0.0) << " us\n";

        std::ofstream out(argv[2], std::ios::binary);
        for (const auto& b : bodies) {
            const sfb::HydrodynamicsInfo& r = b.info;
            const double d[20] = {r.Fb.x, r.Fb.y, r.Fb.z, r.Tb.x, r.Tb.y, r.Tb.z, r.Fdq.x, r.Fdq.y, r.Fdq.z, r.Tdq.x, r.Tdq.y, r.Tdq.z,
                                  r.Fdf.x, r.Fdf.y, r.Fdf.z, r.Tdf.x, r.Tdf.y, r.Tdf.z, r.Swet, r.Vsub};
            out.write(reinterpret_cast<const char*>(d), sizeof(d));
            const uint32_t u[13] = {(uint32_t)r.position, r.faces, r.classCount[0], r.classCount[1], r.classCount[2], r.classCount[3],
                                    r.classCount[4], r.classCount[5], r.classCount[6], r.classCount[7], r.degenerate, r.quadratic, r.skin};
            out.write(reinterpret_cast<const char*>(u), sizeof(u));
        }
        if (argc > 4) {
            for (size_t i = 0; i < bodies.size(); ++i) {
                sfb::SolidBody& b = bodies[i];
                b.fdApproxType = (sfb::GeometryApproxType)(i % 4);
                b.fdApproxParams[0] = b.fdCd.x; b.fdApproxParams[1] = b.fdCd.y; b.fdApproxParams[2] = b.fdCd.z;
                b.T_CG2H = b.T_O;
            }
            engine.ComputeAerodynamicForces(atmosphere, ptrs);
            std::ofstream aout(argv[4], std::ios::binary);
            for (const auto& b : bodies) {
                const double d[6] = {b.Fda.x, b.Fda.y, b.Fda.z, b.Tda.x, b.Tda.y, b.Tda.z};
                aout.write(reinterpret_cast<const char*>(d), sizeof(d));
            }
        }
        if (argc > 6) { // waves produced on the device (InitWaves / SimulateWaves), then the same hydrodynamic step on them
            engine.InitWaves(sfb::WaveParamsFromSeaState((float)std::atof(argv[5])));
            engine.SimulateWaves(1.5f);
            ocean.useGeneratedWaves(true);
            engine.ComputeHydrodynamicForces(settings, ocean, ptrs);
            std::ofstream wout(argv[6], std::ios::binary);
            for (const auto& b : bodies) {
                const sfb::HydrodynamicsInfo& r = b.info;
                const double d[8] = {r.Fb.x, r.Fb.y, r.Fb.z, r.Tb.x, r.Tb.y, r.Tb.z, r.Swet, r.Vsub};
                wout.write(reinterpret_cast<const char*>(d), sizeof(d));
            }
        }
        if (argc > 7) { // a 40-node cable piercing the (generated) sea surface, through HydroEngine::ComputeCableForces
            sfb::CableBody cable;
            cable.radius = 0.02;
            for (int k = 0; k < 40; ++k) {
                const double s = k / 39.0;
                cable.nodes.push_back(sfb::Vector3(-2.0 + 4.0 * s, -1.0 + 2.0 * s, -0.5 + 2.0 * s));
                cable.velocities.push_back(sfb::Vector3(0.1 * s, -0.2, 0.05 * k));
            }
            std::vector<sfb::CableBody*> cables{&cable};
            engine.ComputeCableForces(settings, ocean, cables);
            std::ofstream cout_(argv[7], std::ios::binary);
            for (const auto& f : cable.nodalForces) {
                const double d[9] = {f.Fb.x, f.Fb.y, f.Fb.z, f.Fdq.x, f.Fdq.y, f.Fdq.z, f.Fdf.x, f.Fdf.y, f.Fdf.z};
                cout_.write(reinterpret_cast<const char*>(d), sizeof(d));
            }
        }
    } catch (const std::exception& e) {
        std::cerr << "error: " << e.what() << "\n";
        return 1;
    }
    return 0;
}
