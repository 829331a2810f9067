// TEST INFRASTRUCTURE ONLY (oracle/): driver for the UNMODIFIED reference CUDA code.
//
// This translation unit textually includes the reference's KaminoCore.cu where it
// lies under /root/reference (so that it shares that file's file-static
// __constant__ parameters, KaminoCore.cu:5-9) and drives KaminoSolver the way
// Kamino::run does (KaminoCore.cu:860-872), but instead of the lossy .bgeo writers
// it dumps the raw device state after every phase / step, can start from a state
// read from files, and can time the step loop with the reference's own event
// timers (KaminoSolver.cu:201-218). No reference source is modified or copied.
//
// Usage:
//   kamino_ref dump  <nTheta> <particleDensity> <dt> <radius> <nSteps> <outDir> [inDir] [phaseSteps]
//   kamino_ref bench <nTheta> <particleDensity> <dt> <radius> <nSteps> <stepsPerFrame>
//
// Dump files are raw little-endian float32: <outDir>/<tag>.<field>.f32 with
// field in {velPhi (nT x N), velTheta ((nT-1) x N), density, pressure (nT x N),
// particles (Np x 2, interleaved phi,theta)} and tag in {init, s<k>_adv, s<k>_geo,
// s<k>_proj (= state after step k)}; <outDir>/meta.txt records the sizes.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <chrono>
#include <string>
#include <vector>
#include <fstream>
#include <iostream>
#include <sstream>
#include <map>
#include <algorithm>

#define private public
#include "kernel/KaminoCore.cu"
#undef private

namespace {

bool readFloats(const std::string& path, std::vector<float>& out)
{
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) return false;
    std::streamsize bytes = f.tellg();
    f.seekg(0);
    out.resize((size_t)bytes / sizeof(float));
    f.read(reinterpret_cast<char*>(out.data()), bytes);
    return (bool)f;
}

void writeFloats(const std::string& path, const float* data, size_t n)
{
    std::ofstream f(path, std::ios::binary);
    f.write(reinterpret_cast<const char*>(data), (std::streamsize)(n * sizeof(float)));
    if (!f) { std::fprintf(stderr, "ref_harness: cannot write %s\n", path.c_str()); std::exit(3); }
}

void dumpQuantity(KaminoQuantity* q, const std::string& dir, const std::string& tag, const char* field)
{
    q->copyBackToCPU();
    writeFloats(dir + "/" + tag + "." + field + ".f32", q->cpuBuffer, q->getNPhi() * q->getNTheta());
}

void dumpState(KaminoSolver& s, const std::string& dir, const std::string& tag, bool withPressure)
{
    dumpQuantity(s.velPhi, dir, tag, "velPhi");
    dumpQuantity(s.velTheta, dir, tag, "velTheta");
    dumpQuantity(s.density, dir, tag, "density");
    if (withPressure) dumpQuantity(s.pressure, dir, tag, "pressure");
    s.particles->copyBack2CPU();
    writeFloats(dir + "/" + tag + ".particles.f32", s.particles->coordCPUBuffer, s.particles->numOfParticles * 2);
}

void loadQuantity(KaminoQuantity* q, const std::string& path)
{
    std::vector<float> v;
    if (!readFloats(path, v)) return;
    size_t want = q->getNPhi() * q->getNTheta();
    if (v.size() != want) {
        std::fprintf(stderr, "ref_harness: %s has %zu floats, expected %zu\n", path.c_str(), v.size(), want);
        std::exit(4);
    }
    std::memcpy(q->cpuBuffer, v.data(), want * sizeof(float));
    q->copyToGPU();
}

// Deterministic synthetic density: the reference leaves the
// density buffer uninitialised when no image is given.
void fillSyntheticDensity(KaminoQuantity* rho, float gridLen)
{
    for (size_t j = 0; j < rho->getNTheta(); ++j)
        for (size_t i = 0; i < rho->getNPhi(); ++i) {
            double phi = (double)i * (double)gridLen;
            double theta = ((double)j + 0.5) * (double)gridLen;
            double st = std::sin(theta);
            rho->setCPUValueAt(i, j, (float)(0.5 + 0.5 * std::sin(4.0 * phi) * st * st));
        }
    rho->copyToGPU();
}

void uploadConstants(size_t nPhi, size_t nTheta, float radius, float dt, float gridLen)
{
    // same five uploads as Kamino::run (KaminoCore.cu:868-872)
    checkCudaErrors(cudaMemcpyToSymbol(nPhiGlobal, &nPhi, sizeof(size_t)));
    checkCudaErrors(cudaMemcpyToSymbol(nThetaGlobal, &nTheta, sizeof(size_t)));
    checkCudaErrors(cudaMemcpyToSymbol(radiusGlobal, &radius, sizeof(fReal)));
    checkCudaErrors(cudaMemcpyToSymbol(timeStepGlobal, &dt, sizeof(fReal)));
    checkCudaErrors(cudaMemcpyToSymbol(gridLenGlobal, &gridLen, sizeof(fReal)));
}

int runDump(int argc, char** argv)
{
    if (argc < 8) { std::fprintf(stderr, "dump: too few arguments\n"); return 2; }
    size_t nTheta = (size_t)std::atoll(argv[2]);
    float particleDensity = (float)std::atof(argv[3]);
    float dt = (float)std::atof(argv[4]);
    float radius = (float)std::atof(argv[5]);
    int nSteps = std::atoi(argv[6]);
    std::string outDir = argv[7];
    std::string inDir = argc > 8 ? argv[8] : "";
    if (inDir == "-") inDir = "";
    int phaseSteps = argc > 9 ? std::atoi(argv[9]) : 1;

    size_t nPhi = 2 * nTheta;
    float gridLen = (float)(M_PI / nTheta);   // Kamino::Kamino, KaminoCore.cu:849

    KaminoSolver solver(nPhi, nTheta, radius, dt, 0.0f, 1, 1, 1, 1);
    solver.initDensityfromPic("");
    solver.initParticlesfromPic("", (size_t)particleDensity);
    fillSyntheticDensity(solver.density, gridLen);
    // the pressure buffers are scratch; give them a defined value
    for (size_t k = 0; k < nPhi * nTheta; ++k) solver.pressure->cpuBuffer[k] = 0.0f;
    solver.pressure->copyToGPU();

    if (!inDir.empty()) {
        loadQuantity(solver.velPhi, inDir + "/velPhi.f32");
        loadQuantity(solver.velTheta, inDir + "/velTheta.f32");
        loadQuantity(solver.density, inDir + "/density.f32");
        std::vector<float> p;
        if (readFloats(inDir + "/particles.f32", p)) {
            if (p.size() != solver.particles->numOfParticles * 2) {
                std::fprintf(stderr, "ref_harness: particles.f32 has %zu floats, expected %zu\n",
                             p.size(), solver.particles->numOfParticles * 2);
                return 4;
            }
            std::memcpy(solver.particles->coordCPUBuffer, p.data(), p.size() * sizeof(float));
            solver.particles->copy2GPU();
        }
    }

    uploadConstants(nPhi, nTheta, radius, dt, gridLen);

    {
        std::ofstream meta(outDir + "/meta.txt");
        meta << "nTheta " << nTheta << "\nnPhi " << nPhi << "\nnumParticles " << solver.particles->numOfParticles
             << "\ndt " << dt << "\nradius " << radius << "\ngridLen " << gridLen << "\nnSteps " << nSteps << "\n";
    }

    dumpState(solver, outDir, "init", false);
    for (int s = 1; s <= nSteps; ++s) {
        std::string tag = "s" + std::to_string(s);
        if (s <= phaseSteps) {
            solver.advection();
            checkCudaErrors(cudaDeviceSynchronize());
            dumpState(solver, outDir, tag + "_adv", false);
            solver.geometric();
            checkCudaErrors(cudaDeviceSynchronize());
            dumpState(solver, outDir, tag + "_geo", false);
            solver.projection();
            checkCudaErrors(cudaDeviceSynchronize());
            dumpState(solver, outDir, tag + "_proj", true);
        } else {
            solver.stepForward(dt);
            checkCudaErrors(cudaDeviceSynchronize());
            bool keep = (s == nSteps) || (s % 10 == 0);
            if (keep) dumpState(solver, outDir, tag + "_proj", true);
        }
    }
    return 0;
}

int runBench(int argc, char** argv)
{
    if (argc < 8) { std::fprintf(stderr, "bench: too few arguments\n"); return 2; }
    size_t nTheta = (size_t)std::atoll(argv[2]);
    float particleDensity = (float)std::atof(argv[3]);
    float dt = (float)std::atof(argv[4]);
    float radius = (float)std::atof(argv[5]);
    int nSteps = std::atoi(argv[6]);
    int stepsPerFrame = std::atoi(argv[7]);
    int warmup = argc > 8 ? std::atoi(argv[8]) : 3;

    size_t nPhi = 2 * nTheta;
    float gridLen = (float)(M_PI / nTheta);

    double stepsPerSecKernel = 0.0, stepsPerSecE2E = 0.0, adv = 0.0, geo = 0.0, proj = 0.0;
    size_t numParticles = 0, d2hBytesPerFrame = 0;
    {
        KaminoSolver solver(nPhi, nTheta, radius, dt, 0.0f, 1, 1, 1, 1);
        solver.initDensityfromPic("");
        solver.initParticlesfromPic("", (size_t)particleDensity);
        fillSyntheticDensity(solver.density, gridLen);
        uploadConstants(nPhi, nTheta, radius, dt, gridLen);
        numParticles = solver.particles->numOfParticles;

        for (int s = 0; s < warmup; ++s) solver.stepForward(dt);
        checkCudaErrors(cudaDeviceSynchronize());
        solver.advectionTime = solver.geometricTime = solver.projectionTime = 0.0f;

        // (1) device-resident loop, timed by the reference's own per-phase event timers
        for (int s = 0; s < nSteps; ++s) solver.stepForward(dt);
        checkCudaErrors(cudaDeviceSynchronize());
        adv = solver.advectionTime; geo = solver.geometricTime; proj = solver.projectionTime;
        stepsPerSecKernel = nSteps / (adv + geo + proj);

        // (2) end to end: the frame loop of Kamino::run with the read-backs its writers do
        // (three fields + particle coordinates per frame, KaminoSolver.cu:301-303,375),
        // without the host-side sphere mapping and file output.
        // The timed region starts with the host->device upload of the state (what the
        // reference's constructor / initialisers do once per run).
        solver.velPhi->copyBackToCPU(); solver.velTheta->copyBackToCPU();
        solver.density->copyBackToCPU(); solver.particles->copyBack2CPU();
        checkCudaErrors(cudaDeviceSynchronize());
        auto t0 = std::chrono::steady_clock::now();
        solver.velPhi->copyToGPU();
        solver.velTheta->copyToGPU();
        solver.density->copyToGPU();
        solver.particles->copy2GPU();
        int done = 0;
        while (done < nSteps) {
            int n = std::min(stepsPerFrame, nSteps - done);
            for (int s = 0; s < n; ++s) solver.stepForward(dt);
            solver.velPhi->copyBackToCPU();
            solver.velTheta->copyBackToCPU();
            solver.density->copyBackToCPU();
            solver.particles->copyBack2CPU();
            done += n;
        }
        checkCudaErrors(cudaDeviceSynchronize());
        auto t1 = std::chrono::steady_clock::now();
        stepsPerSecE2E = nSteps / std::chrono::duration<double>(t1 - t0).count();
        d2hBytesPerFrame = (nPhi * nTheta * 2 + nPhi * (nTheta - 1) + numParticles * 2) * sizeof(float);
    }
    std::printf("{\"ref_bench\": true, \"nTheta\": %zu, \"nPhi\": %zu, \"particles\": %zu, \"steps\": %d, "
                "\"advection_s\": %.6f, \"geometric_s\": %.6f, \"projection_s\": %.6f, "
                "\"steps_per_s\": %.3f, \"e2e_steps_per_s\": %.3f, \"steps_per_frame\": %d, \"d2h_bytes_per_frame\": %zu}\n",
                nTheta, nPhi, numParticles, nSteps, adv, geo, proj, stepsPerSecKernel, stepsPerSecE2E,
                stepsPerFrame, d2hBytesPerFrame, d2hBytesPerFrame);
    return 0;
}

} // namespace

int main(int argc, char** argv)
{
    if (argc >= 2 && std::strcmp(argv[1], "dump") == 0) return runDump(argc, argv);
    if (argc >= 2 && std::strcmp(argv[1], "bench") == 0) return runBench(argc, argv);
    std::fprintf(stderr, "usage: kamino_ref dump|bench ... (see ref_harness.cu)\n");
    return 2;
}
