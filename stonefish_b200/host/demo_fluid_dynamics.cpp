// demo_fluid_dynamics.cpp — the reference's FluidDynamicsTest loop shape on the C++ host mirror:
//   build scenario (meshes, bodies, ocean, currents)  ->  every step: ComputeHydrodynamicForces for all bodies  ->
//   print "Hydrodynamics time: <avg> us" like Tests/FluidDynamicsTest/FluidDynamicsTestManager.cpp:166.
// The scenario is read from a small binary file written by tests/test_host_mirror_gpu.py (so the same inputs can be
// pushed through the Python binding and the oracle), results go to a binary file. Needs a CUDA device at run time.
//   usage: demo_fluid_dynamics scene.bin out.bin [steps [aero_out.bin [sea_state waves_out.bin [cable_out.bin]]]]
//          demo_fluid_dynamics scene.bin out.bin steps --contexts N   the bodies block-partitioned over N contexts (device i % number of
//                                                                     devices; one host thread each: sfb::HydroEngineMulti), results as above
//                                                                     plus a strong-scaling record against ONE context on stdout
// With aero_out.bin the same bodies also go through HydroEngine::ComputeAerodynamicForces once: wind = the scene's velocity
// fields, gas density 1.2, proxy type cycling AUTO/SPHERE/CYLINDER/ELLIPSOID with fdCd as the ellipsoid radii, T_CG2H = T_O.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <functional>
#include <iostream>
#include <string>

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

static void write_results(const char* path, const std::vector<sfb::SolidBody>& bodies)
{
    std::ofstream out(path, std::ios::binary);
    for (const auto& b : bodies) {
        const sfb::HydrodynamicsInfo& r = b.info;
        const double d[20] = {r.Fb.x, r.Fb.y, r.Fb.z, r.Tb.x, r.Tb.y, r.Tb.z, r.Fdq.x, r.Fdq.y, r.Fdq.z, r.Tdq.x, r.Tdq.y, r.Tdq.z,
                              r.Fdf.x, r.Fdf.y, r.Fdf.z, r.Tdf.x, r.Tdf.y, r.Tdf.z, r.Swet, r.Vsub};
        out.write(reinterpret_cast<const char*>(d), sizeof(d));
        const uint32_t u[13] = {(uint32_t)r.position, r.faces, r.classCount[0], r.classCount[1], r.classCount[2], r.classCount[3],
                                r.classCount[4], r.classCount[5], r.classCount[6], r.classCount[7], r.degenerate, r.quadratic, r.skin};
        out.write(reinterpret_cast<const char*>(u), sizeof(u));
    }
}

// The same scene on N contexts (sfb::HydroEngineMulti) and on one: the strong-scaling record of one PROCESS driving several GPUs.
static int run_multi(std::istream& in, const char* outPath, int steps, int nContexts, uint32_t nMeshes, uint32_t nBodies, double density, double gravity,
                     uint32_t currentsEnabled, uint32_t nCurrents)
{
    try {
        int nDev = 0;
        {
            sfb::HydroEngine probe(0); // throws without a device
            while (nDev < 64) { sfh_ctx* c = nullptr; if (sfh_create(nDev, &c) != SFH_OK) break; sfh_destroy(c); ++nDev; }
        }
        if (nContexts < 1) nContexts = 1;
        std::vector<int> devices;
        for (int i = 0; i < nContexts; ++i) devices.push_back(i % nDev);
        sfb::HydroEngineMulti multi(devices);
        sfb::HydroEngine single(0);
        sfb::Ocean ocean(density, gravity);
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
        }
        if (currentsEnabled) ocean.EnableCurrents();
        std::vector<uint32_t> meshIds;
        uint64_t faces = 0;
        std::vector<uint32_t> meshFaces;
        for (uint32_t i = 0; i < nMeshes; ++i) {
            const uint32_t nv = rd<uint32_t>(in), stride = rd<uint32_t>(in), nf = rd<uint32_t>(in);
            std::vector<char> verts((size_t)nv * stride), fc((size_t)nf * 12);
            in.read(verts.data(), verts.size());
            in.read(fc.data(), fc.size());
            const uint32_t a = multi.UploadMesh(verts.data(), stride, nv, fc.data(), nf), b = single.UploadMesh(verts.data(), stride, nv, fc.data(), nf);
            if (a != b) throw std::runtime_error("mesh ids differ");
            meshIds.push_back(a);
            meshFaces.push_back(nf);
        }
        std::vector<sfb::SolidBody> bodies(nBodies);
        std::vector<sfb::SolidBody*> ptrs;
        for (auto& b : bodies) {
            const uint32_t mi = rd<uint32_t>(in);
            b.meshId = meshIds.at(mi);
            faces += meshFaces.at(mi);
            b.buoyant = rd<uint32_t>(in) != 0;
            b.T_CG = rdT(in); b.T_C = rdT(in); b.T_O = rdT(in);
            b.v = rd3(in); b.omega = rd3(in); b.aabbMin = rd3(in); b.aabbMax = rd3(in);
            b.fdCd = rd3(in); b.fdCf = rd3(in);
            b.volume = rd<double>(in); b.surface = rd<double>(in);
            b.P_CB = rd3(in);
            ptrs.push_back(&b);
        }
        sfb::HydrodynamicsSettings settings;
        auto timed = [&](const std::function<void()>& step) {
            double us = 0;
            for (int s = 0; s < steps; ++s) {
                auto t0 = std::chrono::steady_clock::now();
                step();
                auto t1 = std::chrono::steady_clock::now();
                if (s > 0) us += std::chrono::duration<double, std::micro>(t1 - t0).count();
            }
            return steps > 1 ? us / (steps - 1) : 0.0;
        };
        const double us1 = timed([&]() { single.ComputeHydrodynamicForces(settings, ocean, ptrs); });
        std::vector<sfb::HydrodynamicsInfo> ref;
        for (const auto& b : bodies) ref.push_back(b.info);
        const std::vector<sfh_info> refRaw = single.LastInfos();
        const double usN = timed([&]() { multi.ComputeHydrodynamicForces(settings, ocean, ptrs); });
        // same bodies, another partition: counters equal, sums equal up to the order of the fp64 additions
        double worst = 0;
        bool counts = true;
        for (size_t i = 0; i < bodies.size(); ++i) {
            const sfb::HydrodynamicsInfo &a = bodies[i].info, &b = ref[i];
            const double F = std::sqrt(b.Fb.x * b.Fb.x + b.Fb.y * b.Fb.y + b.Fb.z * b.Fb.z) + std::sqrt(b.Fdq.x * b.Fdq.x + b.Fdq.y * b.Fdq.y + b.Fdq.z * b.Fdq.z) +
                             std::sqrt(b.Fdf.x * b.Fdf.x + b.Fdf.y * b.Fdf.y + b.Fdf.z * b.Fdf.z) + 1e-300;
            const sfb::Vector3* va[3] = {&a.Fb, &a.Fdq, &a.Fdf};
            const sfb::Vector3* vb[3] = {&b.Fb, &b.Fdq, &b.Fdf};
            for (int k = 0; k < 3; ++k)
                worst = std::max(worst, std::max(std::fabs(va[k]->x - vb[k]->x), std::max(std::fabs(va[k]->y - vb[k]->y), std::fabs(va[k]->z - vb[k]->z))) / F);
            counts = counts && a.position == b.position && a.faces == b.faces && a.degenerate == b.degenerate && a.quadratic == b.quadratic && a.skin == b.skin;
            for (int k = 0; k < 8; ++k) counts = counts && a.classCount[k] == b.classCount[k];
        }
        write_results(outPath, bodies);
        std::printf("{\"scaling\": \"strong\", \"process\": \"one process, one host thread and one sfh context per shard (sfb::HydroEngineMulti)\", "
                    "\"bodies_total\": %u, \"faces_total_per_step\": %llu, \"contexts\": %d, \"devices\": %d, \"us_per_step_1\": %.2f, \"us_per_step_n\": %.2f, "
                    "\"value_1\": %.4e, \"value_n\": %.4e, \"unit\": \"triangle-evals/s (host to host)\", \"max_rel_diff_vs_one_context\": %.3e, "
                    "\"counts_equal\": %s, \"checksum_one_context\": \"%016llx\", \"checksum_n_contexts\": \"%016llx\"}\n",
                    nBodies, (unsigned long long)faces, nContexts, nDev, us1, usN, faces / (us1 * 1e-6), faces / (usN * 1e-6), worst, counts ? "true" : "false",
                    (unsigned long long)sfb::Checksum(refRaw), (unsigned long long)multi.LastChecksum());
        return counts && worst <= 1e-12 ? 0 : 3;
    } catch (const std::exception& e) {
        std::cerr << "error: " << e.what() << "\n";
        return 1;
    }
}

int main(int argc, char** argv)
{
    if (argc < 3) { std::cerr << "usage: demo_fluid_dynamics scene.bin out.bin [steps]\n"; return 2; }
    const int steps = argc > 3 ? std::atoi(argv[3]) : 10;
    std::ifstream in(argv[1], std::ios::binary);
    if (!in || rd<uint32_t>(in) != 0x53484653u) { std::cerr << "bad scene file\n"; return 2; }
    const uint32_t nMeshes = rd<uint32_t>(in), nBodies = rd<uint32_t>(in);
    const double density = rd<double>(in), gravity = rd<double>(in);
    const uint32_t currentsEnabled = rd<uint32_t>(in), nCurrents = rd<uint32_t>(in);

    const bool multi = argc > 5 && std::string(argv[4]) == "--contexts";
    if (multi) return run_multi(in, argv[2], steps, std::atoi(argv[5]), nMeshes, nBodies, density, gravity, currentsEnabled, nCurrents);

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
