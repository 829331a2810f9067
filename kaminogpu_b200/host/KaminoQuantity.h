// KaminoQuantity: one staggered field = a host mirror plus a view of the double-buffered
// device field owned by the solver's kamino_b200 context.
// Public surface of the reference's include/KaminoQuantity.cuh:38-69.
#pragma once

#include "KaminoHeader.h"

class KaminoQuantity
{
private:
    std::string attrName;
    size_t nPhi;
    size_t nTheta;
    fReal phiOffset;
    fReal thetaOffset;
    fReal* cpuBuffer;          // [theta * nPhi + phi], kernel/KaminoQuantity.cu:70-73

    kamino_ctx* ctx;           // not owned
    int field;
    int sim;
    void requireBound(const char* what) const;

public:
    KaminoQuantity(std::string attributeName, size_t nPhi, size_t nTheta,
        fReal phiOffset, fReal thetaOffset);
    ~KaminoQuantity();
    KaminoQuantity(const KaminoQuantity&) = delete;
    KaminoQuantity& operator=(const KaminoQuantity&) = delete;

    /* Attach to a field of a context (done by KaminoSolver; the reference allocates its own
       cudaMallocPitch buffers in the constructor, kernel/KaminoQuantity.cu:20-28). */
    void bind(kamino_ctx* context, int fieldId, int simulation = 0);

    /* The device buffers swap inside the kamino_b200 phase calls; kept for source
       compatibility (kernel/KaminoQuantity.cu:53-58). */
    void swapGPUBuffer();
    void copyToGPU();
    void copyBackToCPU();

    std::string getName();
    size_t getNPhi();
    size_t getNTheta();
    fReal getCPUValueAt(size_t x, size_t y);
    void setCPUValueAt(size_t x, size_t y, fReal val);
    fReal& accessCPUValueAt(size_t x, size_t y);
    fReal getPhiOffset();
    fReal getThetaOffset();
    fReal* getGPUThisStep();
    fReal* getGPUNextStep();
    size_t getThisStepPitchInElements();
    size_t getNextStepPitchInElements();

    fReal* hostData() { return cpuBuffer; }
};
