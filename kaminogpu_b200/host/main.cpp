// kamino <configKamino.txt> -- the reference's CLI (kernel/main.cu:4-60): sixteen
// whitespace-separated tokens
//   radius nTheta particleDensity dt DT frames A B C D E gridPath particlePath
//   densityImage solidImage colorImage
// with "null" meaning "no image" for the density and colour images (the solid image's
// "null" is left untouched by the reference, kernel/main.cu:37-40, and never used).
#include <fstream>

#include "Kamino.h"

int main(int argc, char** argv)
{
    if (argc != 2) {
        std::cout << "Please provide the path to configKamino.txt as an argument." << std::endl;
        std::cout << "Usage example: ./kamino.exe ./configKamino.txt" << std::endl;
        std::cout << "Configuration file was missing, exiting." << std::endl;
        return -1;
    }
    std::fstream fin;
    fin.open(argv[1], std::ios::in);
    fReal radius = 5.0f; size_t nTheta = 128; fReal particleDensity = 200.0f;
    float dt = 0.005f; float DT = 1.0f / 24.0f; int frames = 1000;
    float A = 0.0f; int B = 1, C = 1, D = 1, E = 1;
    std::string gridPath, particlePath, densityImage, solidImage, colorImage;

    fin >> radius >> nTheta >> particleDensity >> dt >> DT >> frames >> A >> B >> C >> D >> E;
    fin >> gridPath >> particlePath;
    fin >> densityImage;
    if (densityImage == "null") densityImage = "";
    fin >> solidImage;
    fin >> colorImage;
    if (colorImage == "null") colorImage = "";

    Kamino KaminoInstance(radius, nTheta, particleDensity, dt, DT, frames,
        A, B, C, D, E, gridPath, particlePath, densityImage, solidImage, colorImage);
    KaminoInstance.run();
    return 0;
}
