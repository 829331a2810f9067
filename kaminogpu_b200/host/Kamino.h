// Kamino: the driver object (reference: include/KaminoGPU.cuh:8-49, kernel/KaminoCore.cu:844-912).
#pragma once

#include "KaminoSolver.h"

class Kamino
{
private:
    size_t nTheta;
    size_t nPhi;
    fReal gridLen;
    fReal radius;
    float dt;
    float DT;
    int frames;
    fReal particleDensity;
    std::string gridPath;
    std::string particlePath;
    std::string densityImage;
    std::string solidImage;
    std::string colorImage;
    fReal A;
    int B, C, D, E;

public:
    Kamino(fReal radius = 5.0, size_t nTheta = 128, fReal particleDensity = 200.0,
        float dt = 0.005, float DT = 1.0 / 24.0, int frames = 1000,
        fReal A = 0.0, int B = 1, int C = 1, int D = 1, int E = 1,
        std::string gridPath = "output/frame", std::string particlePath = "particles/frame",
        std::string densityImage = "", std::string solidImage = "", std::string colorImage = "");
    ~Kamino();

    /* run the solver; file output is skipped for a path equal to "null" or "" */
    void run();
};
