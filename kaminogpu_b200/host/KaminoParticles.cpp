#include "KaminoParticles.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>

#include "ImageIO.h"

// kernel/KaminoParticles.cu:3-83. Positions: the rand()-driven jittered lattice (seeded by the C ABI's
// host initialiser, same call order). Colours: the pixel of the mirrored, resized image under the
// particle's cell (:56-57, :64-72: colorBGR = (G, R, B) / 255), black without an image (:73-77).
KaminoParticles::KaminoParticles(std::string path, fReal particleDensity, fReal gridLen, size_t nTheta) :
    nPhi(2 * nTheta), nTheta(nTheta), particlePGrid((size_t)particleDensity), numOfParticles(0),
    coordCPUBuffer(nullptr), colorBGR(nullptr), coordGPUThisStep(nullptr), coordGPUNextStep(nullptr),
    ctx(nullptr), sim(0)
{
    ImageBGR imageIn, imageOut;
    if (!readImageBGR(path, imageIn))
        std::cerr << "No particle color image provided." << std::endl;
    else
        imageOut = resizeLinear(flipHorizontal(imageIn), (int)nPhi, (int)nTheta);
    numOfParticles = (size_t)kamino_particle_count((int)nTheta, particleDensity);
    coordCPUBuffer = new fReal[numOfParticles * 2 + 1]();
    colorBGR = new fReal[numOfParticles * 3 + 1]();
    if (numOfParticles != 0)
        KAMINO_CHECK(nullptr, kamino_seed_particles_host((int)nTheta, particleDensity, coordCPUBuffer));
    if (!imageOut.empty()) {
        for (size_t index = 0; index < numOfParticles; ++index) {
            // size_t x = std::floor(phi / gridLen), y = std::floor(theta / gridLen) (:56-57); the
            // reference indexes the image unchecked, here the cell is clamped into it
            size_t x = (size_t)std::floor(coordCPUBuffer[2 * index] / gridLen);
            size_t y = (size_t)std::floor(coordCPUBuffer[2 * index + 1] / gridLen);
            if (x >= nPhi) x = nPhi - 1;
            if (y >= nTheta) y = nTheta - 1;
            const unsigned char* p = imageOut.pixel((int)y, (int)x);
            colorBGR[3 * index] = (fReal)(p[1] / 255.0);
            colorBGR[3 * index + 1] = (fReal)(p[2] / 255.0);
            colorBGR[3 * index + 2] = (fReal)(p[0] / 255.0);
        }
    }
}

KaminoParticles::~KaminoParticles()
{
    delete[] coordCPUBuffer;
    delete[] colorBGR;
}

void KaminoParticles::refreshViews()
{
    if (!ctx || numOfParticles == 0) return;
    void* p = nullptr;
    KAMINO_CHECK(ctx, kamino_particles_device_ptr(ctx, sim, 0, &p));
    coordGPUThisStep = static_cast<fReal*>(p);
    KAMINO_CHECK(ctx, kamino_particles_device_ptr(ctx, sim, 1, &p));
    coordGPUNextStep = static_cast<fReal*>(p);
}

void KaminoParticles::bind(kamino_ctx* context, int simulation)
{
    ctx = context;
    sim = simulation;
    long have = 0;
    KAMINO_CHECK(ctx, kamino_get_shape(ctx, nullptr, nullptr, nullptr, &have));
    if ((size_t)have != numOfParticles)
        KAMINO_CHECK(ctx, kamino_alloc_particles(ctx, (long)numOfParticles));
    copy2GPU();
    refreshViews();
}

void KaminoParticles::copy2GPU()
{
    if (!ctx) { std::fprintf(stderr, "KaminoParticles::copy2GPU: not attached to a solver context\n"); std::exit(EXIT_FAILURE); }
    if (numOfParticles != 0)
        KAMINO_CHECK(ctx, kamino_upload_particles(ctx, sim, coordCPUBuffer));
}

void KaminoParticles::copyBack2CPU()
{
    if (!ctx) { std::fprintf(stderr, "KaminoParticles::copyBack2CPU: not attached to a solver context\n"); std::exit(EXIT_FAILURE); }
    if (numOfParticles != 0)
        KAMINO_CHECK(ctx, kamino_download_particles(ctx, sim, coordCPUBuffer));
}

void KaminoParticles::swapGPUBuffers()
{
    refreshViews();
}
