#include "KaminoParticles.h"

#include <cstdio>
#include <cstdlib>

// kernel/KaminoParticles.cu:3-83. A non-empty image path (particle colours) is outside the
// solver path; particles are black as in the reference's no-image branch (:73-77).
KaminoParticles::KaminoParticles(std::string path, fReal particleDensity, fReal gridLen, size_t nTheta) :
    nPhi(2 * nTheta), nTheta(nTheta), particlePGrid((size_t)particleDensity), numOfParticles(0),
    coordCPUBuffer(nullptr), colorBGR(nullptr), coordGPUThisStep(nullptr), coordGPUNextStep(nullptr),
    ctx(nullptr), sim(0)
{
    (void)gridLen;
    if (!path.empty())
        std::cerr << "KaminoParticles: colour image '" << path << "' ignored (image input is not supported)" << std::endl;
    numOfParticles = (size_t)kamino_particle_count((int)nTheta, particleDensity);
    coordCPUBuffer = new fReal[numOfParticles * 2 + 1]();
    colorBGR = new fReal[numOfParticles * 3 + 1]();
    if (numOfParticles != 0)
        KAMINO_CHECK(nullptr, kamino_seed_particles_host((int)nTheta, particleDensity, coordCPUBuffer));
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
