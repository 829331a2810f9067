// KaminoParticles: passive tracer particles. Public fields and methods of the reference's
// include/KaminoParticles.cuh:6-24. The device double buffer lives in the solver's
// kamino_b200 context; coordGPUThisStep / coordGPUNextStep are refreshed views of it.
#pragma once

#include "KaminoHeader.h"

class KaminoParticles
{
public:
    size_t nPhi;
    size_t nTheta;
    size_t particlePGrid;
    size_t numOfParticles;

    fReal* coordCPUBuffer;     // (phi, theta) interleaved
    fReal* colorBGR;
    fReal* coordGPUThisStep;
    fReal* coordGPUNextStep;

    KaminoParticles(std::string path, fReal particleDensity, fReal gridLen, size_t nTheta);
    ~KaminoParticles();
    KaminoParticles(const KaminoParticles&) = delete;
    KaminoParticles& operator=(const KaminoParticles&) = delete;

    /* attach to a context: allocates the device buffers there and uploads the seeded set */
    void bind(kamino_ctx* context, int simulation = 0);

    void copy2GPU();
    void copyBack2CPU();
    void swapGPUBuffers();     // device buffers swap inside kamino_advect; refreshes the views

private:
    kamino_ctx* ctx;
    int sim;
    void refreshViews();
};
