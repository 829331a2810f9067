#include "KaminoQuantity.h"

#include <cstdio>
#include <cstdlib>

void kaminoCheck(int code, kamino_ctx* ctx, const char* what, const char* file, int line)
{
    if (code == 0) return;
    std::fprintf(stderr, "kamino_b200 error at %s:%d code=%d \"%s\" : %s\n", file, line, code, what,
                 kamino_last_error(ctx));
    std::exit(EXIT_FAILURE);
}

KaminoQuantity::KaminoQuantity(std::string attributeName, size_t nPhi, size_t nTheta,
    fReal phiOffset, fReal thetaOffset) :
    attrName(attributeName), nPhi(nPhi), nTheta(nTheta), phiOffset(phiOffset), thetaOffset(thetaOffset),
    cpuBuffer(new fReal[nPhi * nTheta]()), ctx(nullptr), field(-1), sim(0)
{}

KaminoQuantity::~KaminoQuantity()
{
    delete[] cpuBuffer;
}

void KaminoQuantity::bind(kamino_ctx* context, int fieldId, int simulation)
{
    ctx = context;
    field = fieldId;
    sim = simulation;
}

void KaminoQuantity::requireBound(const char* what) const
{
    if (!ctx) {
        std::fprintf(stderr, "KaminoQuantity(%s)::%s: not attached to a solver context\n", attrName.c_str(), what);
        std::exit(EXIT_FAILURE);
    }
}

void KaminoQuantity::swapGPUBuffer() {}

void KaminoQuantity::copyToGPU()
{
    requireBound("copyToGPU");
    KAMINO_CHECK(ctx, kamino_upload_field(ctx, field, sim, cpuBuffer));
}

void KaminoQuantity::copyBackToCPU()
{
    requireBound("copyBackToCPU");
    KAMINO_CHECK(ctx, kamino_download_field(ctx, field, sim, cpuBuffer));
}

std::string KaminoQuantity::getName() { return attrName; }
size_t KaminoQuantity::getNPhi() { return nPhi; }
size_t KaminoQuantity::getNTheta() { return nTheta; }
fReal KaminoQuantity::getCPUValueAt(size_t x, size_t y) { return accessCPUValueAt(x, y); }
void KaminoQuantity::setCPUValueAt(size_t x, size_t y, fReal val) { accessCPUValueAt(x, y) = val; }
fReal& KaminoQuantity::accessCPUValueAt(size_t x, size_t y) { return cpuBuffer[y * nPhi + x]; }
fReal KaminoQuantity::getPhiOffset() { return phiOffset; }
fReal KaminoQuantity::getThetaOffset() { return thetaOffset; }

fReal* KaminoQuantity::getGPUThisStep()
{
    requireBound("getGPUThisStep");
    void* p = nullptr;
    KAMINO_CHECK(ctx, kamino_field_device_ptr(ctx, field, sim, 0, &p, nullptr));
    return static_cast<fReal*>(p);
}

fReal* KaminoQuantity::getGPUNextStep()
{
    requireBound("getGPUNextStep");
    void* p = nullptr;
    KAMINO_CHECK(ctx, kamino_field_device_ptr(ctx, field, sim, 1, &p, nullptr));
    return static_cast<fReal*>(p);
}

size_t KaminoQuantity::getThisStepPitchInElements()
{
    requireBound("getThisStepPitchInElements");
    void* p = nullptr;
    size_t pitch = 0;
    KAMINO_CHECK(ctx, kamino_field_device_ptr(ctx, field, sim, 0, &p, &pitch));
    return pitch;
}

size_t KaminoQuantity::getNextStepPitchInElements()
{
    requireBound("getNextStepPitchInElements");
    void* p = nullptr;
    size_t pitch = 0;
    KAMINO_CHECK(ctx, kamino_field_device_ptr(ctx, field, sim, 1, &p, &pitch));
    return pitch;
}
