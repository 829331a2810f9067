// KaminoSolver: the per-timestep solver object. Public surface of the reference's
// include/KaminoSolver.cuh:99-111; all device work goes through include/kamino_b200.h.
#pragma once

#include "KaminoQuantity.h"
#include "KaminoParticles.h"

class KaminoSolver
{
private:
    kamino_ctx* ctx;

    size_t nPhi;
    size_t nTheta;
    fReal radius;
    fReal gridLen;
    fReal invGridLen;

    /* stored, never used by the GPU initialiser (kernel/KaminoSolver.cu:12-18) */
    fReal A;
    int B, C, D, E;

    KaminoQuantity* velTheta;
    KaminoQuantity* velPhi;
    KaminoQuantity* pressure;
    KaminoQuantity* density;

    fReal frameDuration;
    fReal timeStep;
    fReal timeElapsed;

    /* per-phase accumulators (kernel/KaminoSolver.cu:201-218); filled only when phase
       timing is on, because timing a phase needs a host sync after it */
    bool phaseTiming;
    size_t stepsTaken;

    void advection();
    void geometric();
    void projection();
    void initialize_velocity();

public:
    KaminoSolver(size_t nPhi, size_t nTheta, fReal radius, fReal frameDuration,
        fReal A, int B, int C, int D, int E);
    ~KaminoSolver();
    KaminoSolver(const KaminoSolver&) = delete;
    KaminoSolver& operator=(const KaminoSolver&) = delete;

    void initDensityfromPic(std::string path);
    void initParticlesfromPic(std::string path, size_t parPergrid);

    /* One step. Asynchronous (one CUDA-graph launch) unless phase timing is on. */
    void stepForward(fReal timeStep);

    void write_data_bgeo(const std::string& s, const int frame);
    void write_particles_bgeo(const std::string& s, const int frame);

    KaminoParticles* particles;

    /* additions over the reference surface */
    /* raw checkpoint / restart (Checkpoint.h): the complete state between two steps; `frame` is the
       caller's position in its frame loop. Both synchronise. read_checkpoint exits (the reference's
       error convention) on an unreadable file or a shape / dt / radius / particle-count mismatch. */
    void write_checkpoint(const std::string& path, unsigned frame);
    unsigned read_checkpoint(const std::string& path);
    /* the steps of one output frame in one call (fewest graph launches): nFull steps of dt, then one of lastStep */
    void stepFrame(fReal dt, int nFull, fReal lastStep);
    void setPhaseTiming(bool on) { phaseTiming = on; }   // default: env KAMINO_PHASE_TIMERS=1
    void synchronize();
    kamino_ctx* context() { return ctx; }
    KaminoQuantity* getVelPhi() { return velPhi; }
    KaminoQuantity* getVelTheta() { return velTheta; }
    KaminoQuantity* getDensity() { return density; }
    KaminoQuantity* getPressure() { return pressure; }
};
