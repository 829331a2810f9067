#include "Kamino.h"

#include <cstdlib>

#include "KaminoTimer.h"

Kamino::Kamino(fReal radius, size_t nTheta, fReal particleDensity,
    float dt, float DT, int frames,
    fReal A, int B, int C, int D, int E,
    std::string gridPath, std::string particlePath,
    std::string densityImage, std::string solidImage, std::string colorImage) :
    nTheta(nTheta), nPhi(2 * nTheta), gridLen((fReal)(M_PI / nTheta)), radius(radius),
    dt(dt), DT(DT), frames(frames), particleDensity(particleDensity),
    gridPath(gridPath), particlePath(particlePath),
    densityImage(densityImage), solidImage(solidImage), colorImage(colorImage),
    A(A), B(B), C(C), D(D), E(E)
{}

Kamino::~Kamino() {}

// kernel/KaminoCore.cu:860-912. Steps per frame = (iterations of the while loop) + 1: the
// "remainder" step is a full-dt step because the kernels ignore stepForward's argument.
void Kamino::run()
{
    KaminoSolver solver(nPhi, nTheta, radius, dt, A, B, C, D, E);
    solver.initDensityfromPic(densityImage);
    solver.initParticlesfromPic(colorImage, (size_t)this->particleDensity);

    const bool writeGrid = !(gridPath.empty() || gridPath == "null");
    const bool writeParticles = !(particlePath.empty() || particlePath == "null");

    // Checkpoint / restart (additions; the reference cannot resume a run): KAMINO_CHECKPOINT=<file>
    // rewrites <file> after every frame, KAMINO_RESTART=<file> resumes after the frame it holds.
    // A resumed run continues bit for bit (the state is raw fp32, T restarts from i * DT as the loop sets it).
    const char* checkpointPath = std::getenv("KAMINO_CHECKPOINT");
    const char* restartPath = std::getenv("KAMINO_RESTART");
    int firstFrame = 1;
    if (restartPath && restartPath[0]) {
        firstFrame = (int)solver.read_checkpoint(restartPath) + 1;
        std::cout << "Resuming after frame " << firstFrame - 1 << " from " << restartPath << std::endl;
    } else {
        if (writeGrid) solver.write_data_bgeo(gridPath, 0);
        if (writeParticles) solver.write_particles_bgeo(particlePath, 0);
    }

    KaminoTimer timer(solver.context());
    timer.startTimer();

    float T = (firstFrame - 1) * DT;
    for (int i = firstFrame; i <= frames; i++) {
        // the reference's loop (kernel/KaminoCore.cu:888-894) decides how many steps the frame takes; they
        // are queued together so that the device runs them from 10-step graphs
        int nFull = 0;
        while (T < i * DT) {
            ++nFull;
            T += dt;
        }
        solver.stepFrame(dt, nFull, dt + i * DT - T);
        T = i * DT;

        std::cout << "Frame " << i << " is ready" << std::endl;
        if (writeGrid) solver.write_data_bgeo(gridPath, i);
        if (writeParticles) solver.write_particles_bgeo(particlePath, i);
        if (checkpointPath && checkpointPath[0]) solver.write_checkpoint(checkpointPath, (unsigned)i);
    }

    float gpu_time = timer.stopTimer();
    std::cout << "Time spent: " << gpu_time << "ms" << std::endl;
    std::cout << "Performance: " << 1000.0 * (frames - firstFrame + 1) / gpu_time << " frames per second" << std::endl;
}
