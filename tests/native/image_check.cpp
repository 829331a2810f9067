// CPU check of kaminogpu_b200/host/ImageIO.cpp: read <path>, mirror, resize to nPhi x nTheta and
// derive the density exactly as KaminoSolver::initDensityfromPic does; dump everything as raw bytes:
//   int32 width, height | width*height*3 bytes (BGR as read) | nTheta*nPhi*3 bytes (resized) |
//   nTheta*nPhi float32 (density, row-major [theta][phi])
// Built and run by tests/test_image_init.py (g++, no GPU).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../kaminogpu_b200/host/ImageIO.h"

int main(int argc, char** argv)
{
    if (argc != 4) return 2;
    const int nTheta = std::atoi(argv[2]), nPhi = 2 * nTheta;
    ImageBGR in;
    if (!readImageBGR(argv[1], in)) return 3;
    const ImageBGR out = resizeLinear(flipHorizontal(in), nPhi, nTheta);
    FILE* f = std::fopen(argv[3], "wb");
    if (!f) return 4;
    const int dims[2] = {in.width, in.height};
    std::fwrite(dims, sizeof(int), 2, f);
    std::fwrite(in.data.data(), 1, in.data.size(), f);
    std::fwrite(out.data.data(), 1, out.data.size(), f);
    std::vector<float> density((size_t)nTheta * nPhi);
    for (int j = 0; j < nTheta; ++j)
        for (int i = 0; i < nPhi; ++i) {
            const unsigned char* p = out.pixel(j, i);
            const float B = (float)(p[0] / 255.0), G = (float)(p[1] / 255.0), R = (float)(p[2] / 255.0);
            density[(size_t)j * nPhi + i] = (float)((B + G + R) / 3.0);
        }
    std::fwrite(density.data(), sizeof(float), density.size(), f);
    std::fclose(f);
    return 0;
}
