// CPU check of kaminogpu_b200/host/Checkpoint.cpp: round trip, checksum and truncation detection.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>

#include "../../kaminogpu_b200/host/Checkpoint.h"

static int failures = 0;
#define EXPECT(cond) do { if (!(cond)) { std::printf("FAILED: %s (line %d)\n", #cond, __LINE__); ++failures; } } while (0)

int main(int argc, char** argv)
{
    if (argc != 2) return 2;
    const std::string path = argv[1];
    CheckpointState a;
    a.header.nTheta = 16; a.header.nPhi = 32; a.header.radius = 5.0f; a.header.dt = 0.005f;
    a.header.frame = 7; a.header.stepsTaken = 63; a.header.numParticles = 100;
    a.velPhi.resize(16 * 32); a.velTheta.resize(15 * 32); a.density.resize(16 * 32); a.particles.resize(200);
    unsigned seed = 12345;
    auto next = [&seed]() { seed = seed * 1664525u + 1013904223u; return (float)(seed >> 8) / 16777216.0f - 0.5f; };
    for (auto* v : {&a.velPhi, &a.velTheta, &a.density, &a.particles}) for (float& x : *v) x = next();
    a.velPhi[3] = -0.0f;
    std::memset(&a.velTheta[5], 0xff, 4);                       // a NaN payload must survive untouched
    std::string error;
    EXPECT(writeCheckpoint(path, a, &error));
    CheckpointState b;
    EXPECT(readCheckpoint(path, b, &error));
    EXPECT(b.header.nTheta == 16 && b.header.nPhi == 32 && b.header.frame == 7 && b.header.stepsTaken == 63
           && b.header.numParticles == 100 && b.header.radius == 5.0f && b.header.dt == 0.005f);
    EXPECT(b.velPhi.size() == a.velPhi.size() && !std::memcmp(b.velPhi.data(), a.velPhi.data(), 4 * a.velPhi.size()));
    EXPECT(b.velTheta.size() == a.velTheta.size() && !std::memcmp(b.velTheta.data(), a.velTheta.data(), 4 * a.velTheta.size()));
    EXPECT(!std::memcmp(b.density.data(), a.density.data(), 4 * a.density.size()));
    EXPECT(!std::memcmp(b.particles.data(), a.particles.data(), 4 * a.particles.size()));

    // a mismatching array size is refused at write time
    CheckpointState bad = a;
    bad.density.pop_back();
    EXPECT(!writeCheckpoint(path + ".bad", bad, &error));

    // one flipped bit is caught by the checksum; a truncated file is reported as such
    std::ifstream in(path, std::ios::binary);
    std::vector<char> bytes((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    in.close();
    std::vector<char> flipped = bytes;
    flipped[bytes.size() / 2] ^= 0x10;
    { std::ofstream out(path + ".flip", std::ios::binary); out.write(flipped.data(), (std::streamsize)flipped.size()); }
    EXPECT(!readCheckpoint(path + ".flip", b, &error) && error.find("checksum") != std::string::npos);
    { std::ofstream out(path + ".cut", std::ios::binary); out.write(bytes.data(), (std::streamsize)(bytes.size() - 100)); }
    EXPECT(!readCheckpoint(path + ".cut", b, &error) && error.find("truncated") != std::string::npos);
    { std::ofstream out(path + ".junk", std::ios::binary); out << "not a checkpoint at all, just some text that is long enough......"; }
    EXPECT(!readCheckpoint(path + ".junk", b, &error));
    EXPECT(!readCheckpoint(path + ".missing", b, &error));
    std::printf(failures ? "checkpoint check FAILED\n" : "checkpoint check ok\n");
    return failures ? 1 : 0;
}
