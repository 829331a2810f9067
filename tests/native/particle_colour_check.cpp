// CPU check of the image-driven particle colours (kaminogpu_b200/host/KaminoParticles.cpp against
// kernel/KaminoParticles.cu:20-72): construct KaminoParticles with a colour image (no device work
// happens before bind()) and dump coordinates and colours:
//   int64 numOfParticles | 2 n float32 (phi, theta) | 3 n float32 colorBGR
#include <cstdio>
#include <cstdlib>

#include "../../kaminogpu_b200/host/KaminoParticles.h"

int main(int argc, char** argv)
{
    if (argc != 5) return 2;
    const size_t nTheta = (size_t)std::atoi(argv[2]);
    const float density = (float)std::atof(argv[3]);
    const float gridLen = (float)(3.14159265358979323846 / (double)nTheta);
    KaminoParticles particles(argv[1], density, gridLen, nTheta);
    FILE* f = std::fopen(argv[4], "wb");
    if (!f) return 4;
    const long long n = (long long)particles.numOfParticles;
    std::fwrite(&n, sizeof(n), 1, f);
    std::fwrite(particles.coordCPUBuffer, sizeof(float), 2 * (size_t)n, f);
    std::fwrite(particles.colorBGR, sizeof(float), 3 * (size_t)n, f);
    std::fclose(f);
    return 0;
}
