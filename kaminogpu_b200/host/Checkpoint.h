// Raw state checkpoint / restart (the reference only writes lossy .bgeo frames,
// kernel/KaminoSolver.cu:284-401, and cannot resume a run). A checkpoint is the complete state of
// the solver path between two steps: u_phi, u_theta, density and the particle coordinates as raw
// fp32 (the pressure is recomputed by every step and is not state), plus the shape, the parameters
// the kernels use and the position in the frame loop. Resuming from it continues bit for bit.
//
// File layout (little endian): CheckpointHeader | u_phi | u_theta | density | particles | uint32
// Fletcher-style checksum of everything before it.
#ifndef KAMINO_CHECKPOINT_H
#define KAMINO_CHECKPOINT_H

#include <cstdint>
#include <string>
#include <vector>

struct CheckpointHeader {
    char magic[8];            // "KAMINOCK"
    uint32_t version;         // 1
    uint32_t nTheta, nPhi;
    float radius, dt;         // radiusGlobal / timeStepGlobal of the run
    uint32_t frame;           // frames completed (Kamino::run's loop index)
    uint64_t stepsTaken;
    uint64_t numParticles;
};

struct CheckpointState {
    CheckpointHeader header;
    std::vector<float> velPhi;      // nTheta x nPhi
    std::vector<float> velTheta;    // (nTheta - 1) x nPhi
    std::vector<float> density;     // nTheta x nPhi
    std::vector<float> particles;   // numParticles x (phi, theta)
};

// Writes to `path + ".tmp"` and renames, so an interrupted write never leaves a truncated checkpoint.
bool writeCheckpoint(const std::string& path, const CheckpointState& state, std::string* error = nullptr);
// Validates magic, version, sizes and checksum.
bool readCheckpoint(const std::string& path, CheckpointState& state, std::string* error = nullptr);

#endif
