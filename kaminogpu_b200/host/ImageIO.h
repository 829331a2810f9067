// Image input for the initialisers, without OpenCV.
//
// The reference initialises the density field and the particle colours from an image
// (kernel/KaminoSolver.cu:243-277, kernel/KaminoParticles.cu:6-18,64-72) through three OpenCV
// calls: cv::imread(path, IMREAD_COLOR), cv::flip(.., 1) and cv::resize(.., Size(nPhi, nTheta))
// (default INTER_LINEAR). This header provides the same three operations for 8-bit images:
// the readers return BGR pixels as imread does, and resizeLinear reproduces OpenCV's fixed-point
// bilinear resize for CV_8UC3 bit for bit (tests/test_image_init.py checks it against vectors
// produced by cv2.resize, tests/golden/make_resize_goldens.py).
#ifndef KAMINO_IMAGE_IO_H
#define KAMINO_IMAGE_IO_H

#include <string>
#include <vector>

struct ImageBGR {
    int width = 0, height = 0;
    std::vector<unsigned char> data;       // height x width x 3, rows top to bottom, B G R
    bool empty() const { return data.empty(); }
    const unsigned char* pixel(int row, int col) const { return &data[((size_t)row * width + col) * 3]; }
};

// cv::imread(path, IMREAD_COLOR) for binary / ASCII PGM and PPM (maxval <= 255) and for
// non-interlaced 8-bit PNG (grey, grey+alpha, RGB, RGBA, palette; alpha is dropped, grey is
// replicated). Returns false (image left empty) when the file is missing or not decodable, which
// is the condition under which the reference prints "No ... image provided." and carries on.
bool readImageBGR(const std::string& path, ImageBGR& out);

// cv::flip(src, dst, 1): mirror around the vertical axis.
ImageBGR flipHorizontal(const ImageBGR& src);

// cv::resize(src, dst, Size(width, height)) with the default INTER_LINEAR on CV_8UC3.
ImageBGR resizeLinear(const ImageBGR& src, int width, int height);

#endif
