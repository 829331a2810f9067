#include "KaminoSolver.h"
#include "ImageIO.h"

#include <algorithm>
#include <cstdlib>

#include "BgeoWriter.h"
#include "Checkpoint.h"

// kernel/KaminoSolver.cu:12-67: the device allocations, the tridiagonal coefficient
// tables and the FFT plan of the reference's constructor are all inside kamino_create.
// As in Kamino::run (kernel/KaminoCore.cu:862,871) `frameDuration` receives dt and is the
// time step every kernel uses.
KaminoSolver::KaminoSolver(size_t nPhi, size_t nTheta, fReal radius, fReal frameDuration,
    fReal A, int B, int C, int D, int E) :
    ctx(nullptr), nPhi(nPhi), nTheta(nTheta), radius(radius), gridLen((fReal)(M_2PI / nPhi)),
    invGridLen((fReal)(1.0 / (M_2PI / nPhi))), A(A), B(B), C(C), D(D), E(E),
    frameDuration(frameDuration), timeStep(0.0), timeElapsed(0.0),
    phaseTiming(false), stepsTaken(0), particles(nullptr)
{
    if (nPhi != 2 * nTheta) {
        std::cerr << "KaminoSolver: nPhi must be 2 * nTheta" << std::endl;
        std::exit(EXIT_FAILURE);
    }
    const char* env = std::getenv("KAMINO_PHASE_TIMERS");
    phaseTiming = env && env[0] == '1';
    int device = 0;                                        // kernel/KaminoSolver.cu:20
    if (const char* d = std::getenv("KAMINO_DEVICE")) device = std::atoi(d);
    KAMINO_CHECK(nullptr, kamino_create(&ctx, device, (int)nTheta, radius, frameDuration, 1, 0));

    velPhi = new KaminoQuantity("velPhi", nPhi, nTheta, vPhiPhiOffset, vPhiThetaOffset);
    velTheta = new KaminoQuantity("velTheta", nPhi, nTheta - 1, vThetaPhiOffset, vThetaThetaOffset);
    pressure = new KaminoQuantity("p", nPhi, nTheta, centeredPhiOffset, centeredThetaOffset);
    density = new KaminoQuantity("density", nPhi, nTheta, centeredPhiOffset, centeredThetaOffset);
    velPhi->bind(ctx, KAMINO_VEL_PHI);
    velTheta->bind(ctx, KAMINO_VEL_THETA);
    pressure->bind(ctx, KAMINO_PRESSURE);
    density->bind(ctx, KAMINO_DENSITY);

    initialize_velocity();
}

KaminoSolver::~KaminoSolver()
{
    float adv = 0.f, geo = 0.f, proj = 0.f;
    if (ctx) {
        kamino_sync(ctx);
        kamino_phase_times(ctx, &adv, &geo, &proj, 0);
    }
    delete velPhi;
    delete velTheta;
    delete pressure;
    delete density;
    delete particles;
    if (ctx) KAMINO_CHECK(ctx, kamino_destroy(ctx));
    if (phaseTiming) {       // kernel/KaminoSolver.cu:95-103
        const float total = adv + geo + proj;
        std::cout << "Total time used for advection : " << adv << std::endl;
        std::cout << "Total time used for geometric : " << geo << std::endl;
        std::cout << "Total time used for projection : " << proj << std::endl;
        std::cout << "Percentage of advection : " << adv / total * 100.0f << "%" << std::endl;
        std::cout << "Percentage of geometric : " << geo / total * 100.0f << "%" << std::endl;
        std::cout << "Percentage of projection : " << proj / total * 100.0f << "%" << std::endl;
    }
}

// kernel/KaminoInitializer.cu:3-85: FBM curl-noise velocity on the host, then upload.
void KaminoSolver::initialize_velocity()
{
    std::cout << "Initializing velocity..." << std::endl;
    KAMINO_CHECK(ctx, kamino_init_velocity_host((int)nTheta, radius, velPhi->hostData(), velTheta->hostData()));
    velPhi->copyToGPU();
    velTheta->copyToGPU();
}

// kernel/KaminoSolver.cu:243-277: density = mean of the B, G, R channels of the image, mirrored
// and resized to nPhi x nTheta (imread / flip / resize without OpenCV: ImageIO.h).
void KaminoSolver::initDensityfromPic(std::string path)
{
    if (path == "") return;
    ImageBGR imageIn;
    if (!readImageBGR(path, imageIn)) {
        std::cerr << "No density image provided." << std::endl;
        return;
    }
    const ImageBGR resized = resizeLinear(flipHorizontal(imageIn), (int)nPhi, (int)nTheta);
    for (size_t i = 0; i < nPhi; ++i)
        for (size_t j = 0; j < nTheta; ++j) {
            const unsigned char* p = resized.pixel((int)j, (int)i);
            const fReal B = (fReal)(p[0] / 255.0);
            const fReal G = (fReal)(p[1] / 255.0);
            const fReal R = (fReal)(p[2] / 255.0);
            density->setCPUValueAt(i, j, (fReal)((B + G + R) / 3.0));
        }
    density->copyToGPU();
}

// kernel/KaminoSolver.cu:279-282
void KaminoSolver::initParticlesfromPic(std::string path, size_t parPergrid)
{
    delete particles;
    particles = new KaminoParticles(path, (fReal)parPergrid, gridLen, nTheta);
    particles->bind(ctx);
}

void KaminoSolver::advection() { KAMINO_CHECK(ctx, kamino_advect(ctx)); }
void KaminoSolver::geometric() { KAMINO_CHECK(ctx, kamino_geometric(ctx)); }
void KaminoSolver::projection() { KAMINO_CHECK(ctx, kamino_project(ctx)); }

// kernel/KaminoSolver.cu:197-221. The argument is recorded; the kernels use the dt given
// at construction, exactly as the reference's kernels read timeStepGlobal.
void KaminoSolver::stepForward(fReal timeStep)
{
    this->timeStep = timeStep;
    if (phaseTiming) {
        advection();
        geometric();
        projection();
    } else {
        KAMINO_CHECK(ctx, kamino_step(ctx, 1));
    }
    if (particles) particles->swapGPUBuffers();
    ++stepsTaken;
    this->timeElapsed += timeStep;
}

// The steps of one frame of Kamino::run (kernel/KaminoCore.cu:888-894: `nFull` calls stepForward(dt) and
// one stepForward(lastStep)) queued as ONE kamino_step call, i.e. as few CUDA-graph launches as the
// step count allows (10 steps per graph) instead of one launch per step. Same state and same
// bookkeeping as calling stepForward nFull + 1 times.
void KaminoSolver::stepFrame(fReal dt, int nFull, fReal lastStep)
{
    if (phaseTiming) {
        for (int k = 0; k < nFull; ++k) stepForward(dt);
        stepForward(lastStep);
        return;
    }
    const int n = nFull + 1;
    KAMINO_CHECK(ctx, kamino_step(ctx, n));
    if (particles) particles->swapGPUBuffers();
    stepsTaken += (size_t)n;
    for (int k = 0; k < nFull; ++k) this->timeElapsed += dt;
    this->timeStep = lastStep;
    this->timeElapsed += lastStep;
}

void KaminoSolver::synchronize() { KAMINO_CHECK(ctx, kamino_sync(ctx)); }

// Raw checkpoint (no reference counterpart). State = u_phi, u_theta, density, particle
// coordinates; dense host arrays in the field layouts of include/kamino_b200.h.
void KaminoSolver::write_checkpoint(const std::string& path, unsigned frame)
{
    synchronize();
    velPhi->copyBackToCPU();
    velTheta->copyBackToCPU();
    density->copyBackToCPU();
    CheckpointState st;
    st.header.nTheta = (uint32_t)nTheta;
    st.header.nPhi = (uint32_t)nPhi;
    st.header.radius = radius;
    st.header.dt = frameDuration;
    st.header.frame = frame;
    st.header.stepsTaken = stepsTaken;
    st.header.numParticles = particles ? particles->numOfParticles : 0;
    st.velPhi.assign(velPhi->hostData(), velPhi->hostData() + nPhi * nTheta);
    st.velTheta.assign(velTheta->hostData(), velTheta->hostData() + nPhi * (nTheta - 1));
    st.density.assign(density->hostData(), density->hostData() + nPhi * nTheta);
    if (particles && particles->numOfParticles) {
        particles->copyBack2CPU();
        st.particles.assign(particles->coordCPUBuffer, particles->coordCPUBuffer + 2 * particles->numOfParticles);
    }
    std::string error;
    if (!writeCheckpoint(path, st, &error)) {
        std::cerr << error << std::endl;
        std::exit(EXIT_FAILURE);
    }
}

unsigned KaminoSolver::read_checkpoint(const std::string& path)
{
    CheckpointState st;
    std::string error;
    if (!readCheckpoint(path, st, &error)) {
        std::cerr << error << std::endl;
        std::exit(EXIT_FAILURE);
    }
    const size_t have = particles ? particles->numOfParticles : 0;
    if (st.header.nTheta != nTheta || st.header.nPhi != nPhi || st.header.radius != radius
        || st.header.dt != frameDuration || st.header.numParticles != have) {
        std::cerr << "checkpoint: " << path << " was written by a run of a different shape, radius, dt or particle count" << std::endl;
        std::exit(EXIT_FAILURE);
    }
    synchronize();
    std::copy(st.velPhi.begin(), st.velPhi.end(), velPhi->hostData());
    std::copy(st.velTheta.begin(), st.velTheta.end(), velTheta->hostData());
    std::copy(st.density.begin(), st.density.end(), density->hostData());
    velPhi->copyToGPU();
    velTheta->copyToGPU();
    density->copyToGPU();
    if (have) {
        std::copy(st.particles.begin(), st.particles.end(), particles->coordCPUBuffer);
        particles->copy2GPU();
    }
    stepsTaken = (size_t)st.header.stepsTaken;
    return st.header.frame;
}

namespace {

// mapPToSphere / mapVToSphere, kernel/KaminoSolver.cu:403-423 (vec3 stores doubles; the
// angles pass through float)
void pointOnSphere(float radius, double phiIn, double thetaIn, double out[3])
{
    const float theta = (float)thetaIn, phi = (float)phiIn;
    out[0] = radius * std::sin(theta) * std::cos(phi);
    out[2] = radius * std::sin(theta) * std::sin(phi);
    out[1] = radius * std::cos(theta);
}

void velocityOnSphere(double phiIn, double thetaIn, double uTheta, double uPhi, double out[3])
{
    const float theta = (float)thetaIn, phi = (float)phiIn;
    const float ut = (float)uTheta, up = (float)uPhi;
    out[0] = std::cos(theta) * std::cos(phi) * ut - std::sin(phi) * up;
    out[2] = std::cos(theta) * std::sin(phi) * ut + std::cos(phi) * up;
    out[1] = -std::sin(theta) * ut;
}

} // namespace

// kernel/KaminoSolver.cu:284-357: cell-centred velocity mapped to 3-D, plus density.
void KaminoSolver::write_data_bgeo(const std::string& s, const int frame)
{
    const std::string file = s + std::to_string(frame) + ".bgeo";
    std::cout << "Writing to: " << file << std::endl;

    velPhi->copyBackToCPU();
    velTheta->copyBackToCPU();
    density->copyBackToCPU();

    std::vector<float> positions(3 * nPhi * nTheta);
    BgeoAttribute vel{"v", 3, true, std::vector<float>(3 * nPhi * nTheta)};
    BgeoAttribute rho{"density", 1, false, std::vector<float>(nPhi * nTheta)};
    size_t idx = 0;
    for (size_t j = 0; j < nTheta; ++j) {
        for (size_t i = 0; i < nPhi; ++i, ++idx) {
            const fReal uWest = velPhi->getCPUValueAt(i, j);
            const fReal uEast = velPhi->getCPUValueAt(i == nPhi - 1 ? 0 : i + 1, j);
            size_t jNorth, jSouth;
            if (j == 0) jNorth = jSouth = 0;
            else if (j == nTheta - 1) jNorth = jSouth = nTheta - 2;
            else { jNorth = j - 1; jSouth = j; }
            const fReal vNorth = velTheta->getCPUValueAt(i, jNorth);
            const fReal vSouth = velTheta->getCPUValueAt(i, jSouth);
            const fReal velocityPhi = (fReal)((uWest + uEast) / 2.0);
            const fReal velocityTheta = (fReal)((vNorth + vSouth) / 2.0);
            const double phi = (i + centeredPhiOffset) * gridLen;
            const double theta = (j + centeredThetaOffset) * gridLen;
            double p3[3], v3[3];
            velocityOnSphere(phi, theta, velocityTheta, velocityPhi, v3);
            pointOnSphere(radius, phi, theta, p3);
            for (int k = 0; k < 3; ++k) {
                positions[3 * idx + k] = (float)p3[k];
                vel.values[3 * idx + k] = (float)v3[k];
            }
            rho.values[idx] = density->getCPUValueAt(i, j);
        }
    }
    if (!writeBgeo(file, positions, {vel, rho}))
        std::cerr << "Partio: failed to write " << file << std::endl;
}

// kernel/KaminoSolver.cu:359-401: particle positions on the sphere, zero velocity, colour.
// The reference reads colorBGR[3i+1 .. 3i+3] (an off-by-one that is kept; the colour
// buffer has one spare element so the last read stays in bounds).
void KaminoSolver::write_particles_bgeo(const std::string& s, const int frame)
{
    const std::string file = s + std::to_string(frame) + ".bgeo";
    std::cout << "Writing to: " << file << std::endl;
    if (!particles) return;
    particles->copyBack2CPU();
    const size_t n = particles->numOfParticles;
    std::vector<float> positions(3 * n);
    BgeoAttribute vel{"v", 3, true, std::vector<float>(3 * n, 0.0f)};
    BgeoAttribute col{"color", 3, true, std::vector<float>(3 * n)};
    for (size_t i = 0; i < n; ++i) {
        double p3[3];
        pointOnSphere(radius, particles->coordCPUBuffer[2 * i], particles->coordCPUBuffer[2 * i + 1], p3);
        for (int k = 0; k < 3; ++k) {
            positions[3 * i + k] = (float)p3[k];
            col.values[3 * i + k] = particles->colorBGR[3 * i + 1 + k];
        }
    }
    if (!writeBgeo(file, positions, {vel, col}))
        std::cerr << "Partio: failed to write " << file << std::endl;
}
