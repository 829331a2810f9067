#include "ImageIO.h"

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iterator>

#ifdef KAMINO_HAVE_ZLIB
#include <zlib.h>
#endif

namespace {

// ---- PNM -----------------------------------------------------------------------------------

struct Cursor {
    const std::vector<unsigned char>& b;
    size_t p;
    bool skipSpaceAndComments()
    {
        while (p < b.size()) {
            if (b[p] == '#') { while (p < b.size() && b[p] != '\n') ++p; }
            else if (b[p] == ' ' || b[p] == '\t' || b[p] == '\r' || b[p] == '\n') ++p;
            else return true;
        }
        return false;
    }
    bool readInt(int& v)
    {
        if (!skipSpaceAndComments() || b[p] < '0' || b[p] > '9') return false;
        long acc = 0;
        while (p < b.size() && b[p] >= '0' && b[p] <= '9') { acc = acc * 10 + (b[p] - '0'); if (acc > 1 << 30) return false; ++p; }
        v = (int)acc;
        return true;
    }
};

bool readPNM(const std::vector<unsigned char>& bytes, ImageBGR& out)
{
    if (bytes.size() < 3 || bytes[0] != 'P') return false;
    const int kind = bytes[1] - '0';                     // 2, 3 ASCII; 5, 6 binary
    if (kind != 2 && kind != 3 && kind != 5 && kind != 6) return false;
    Cursor c{bytes, 2};
    int w = 0, h = 0, maxval = 0;
    if (!c.readInt(w) || !c.readInt(h) || !c.readInt(maxval)) return false;
    if (w <= 0 || h <= 0 || maxval <= 0 || maxval > 255) return false;
    const int channels = (kind == 3 || kind == 6) ? 3 : 1;
    const size_t count = (size_t)w * h * channels;
    std::vector<unsigned char> samples(count);
    if (kind >= 5) {
        ++c.p;                                           // the single whitespace after maxval
        if (bytes.size() < c.p + count) return false;
        std::memcpy(samples.data(), bytes.data() + c.p, count);
    } else {
        for (size_t k = 0; k < count; ++k) { int v; if (!c.readInt(v) || v > maxval) return false; samples[k] = (unsigned char)v; }
    }
    if (maxval != 255)                                   // imread scales to the full 8-bit range
        for (auto& s : samples) s = (unsigned char)((s * 255 + maxval / 2) / maxval);
    out.width = w; out.height = h;
    out.data.resize((size_t)w * h * 3);
    for (size_t px = 0; px < (size_t)w * h; ++px) {
        if (channels == 3) { out.data[3 * px] = samples[3 * px + 2]; out.data[3 * px + 1] = samples[3 * px + 1]; out.data[3 * px + 2] = samples[3 * px]; }
        else out.data[3 * px] = out.data[3 * px + 1] = out.data[3 * px + 2] = samples[px];
    }
    return true;
}

// ---- PNG (8-bit, non-interlaced) --------------------------------------------------------------

#ifdef KAMINO_HAVE_ZLIB
uint32_t be32(const unsigned char* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

int paeth(int a, int b, int c)
{
    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

bool readPNG(const std::vector<unsigned char>& bytes, ImageBGR& out)
{
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (bytes.size() < 8 + 25 || std::memcmp(bytes.data(), sig, 8) != 0) return false;
    size_t p = 8;
    int w = 0, h = 0, colourType = -1;
    std::vector<unsigned char> idat, palette;
    while (p + 12 <= bytes.size()) {
        const uint32_t len = be32(&bytes[p]);
        const unsigned char* type = &bytes[p + 4];
        if (p + 12 + (size_t)len > bytes.size()) return false;
        const unsigned char* body = &bytes[p + 8];
        if (!std::memcmp(type, "IHDR", 4)) {
            if (len != 13) return false;
            w = (int)be32(body); h = (int)be32(body + 4);
            const int depth = body[8]; colourType = body[9];
            if (depth != 8 || body[10] != 0 || body[11] != 0 || body[12] != 0) return false;   // 8-bit, no interlace
        } else if (!std::memcmp(type, "PLTE", 4)) palette.assign(body, body + len);
        else if (!std::memcmp(type, "IDAT", 4)) idat.insert(idat.end(), body, body + len);
        else if (!std::memcmp(type, "IEND", 4)) break;
        p += 12 + (size_t)len;
    }
    int channels;
    switch (colourType) { case 0: channels = 1; break; case 2: channels = 3; break; case 3: channels = 1; break;
                          case 4: channels = 2; break; case 6: channels = 4; break; default: return false; }
    if (w <= 0 || h <= 0 || idat.empty()) return false;
    const size_t stride = (size_t)w * channels;
    std::vector<unsigned char> raw((stride + 1) * h);
    uLongf rawLen = (uLongf)raw.size();
    if (uncompress(raw.data(), &rawLen, idat.data(), (uLong)idat.size()) != Z_OK || rawLen != raw.size()) return false;
    std::vector<unsigned char> img(stride * h);
    for (int y = 0; y < h; ++y) {
        const unsigned char* line = &raw[(stride + 1) * y];
        unsigned char* cur = &img[stride * y];
        const unsigned char* up = y ? &img[stride * (y - 1)] : nullptr;
        const int filter = line[0];
        for (size_t x = 0; x < stride; ++x) {
            const int a = x >= (size_t)channels ? cur[x - channels] : 0;
            const int b = up ? up[x] : 0;
            const int c = (up && x >= (size_t)channels) ? up[x - channels] : 0;
            int v = line[1 + x];
            switch (filter) { case 0: break; case 1: v += a; break; case 2: v += b; break;
                              case 3: v += (a + b) / 2; break; case 4: v += paeth(a, b, c); break; default: return false; }
            cur[x] = (unsigned char)v;
        }
    }
    out.width = w; out.height = h;
    out.data.resize((size_t)w * h * 3);
    for (size_t px = 0; px < (size_t)w * h; ++px) {
        unsigned char r, g, b;
        const unsigned char* s = &img[px * channels];
        if (colourType == 3) {
            if ((size_t)s[0] * 3 + 2 >= palette.size()) return false;
            r = palette[s[0] * 3]; g = palette[s[0] * 3 + 1]; b = palette[s[0] * 3 + 2];
        } else if (channels >= 3) { r = s[0]; g = s[1]; b = s[2]; }
        else r = g = b = s[0];
        out.data[3 * px] = b; out.data[3 * px + 1] = g; out.data[3 * px + 2] = r;
    }
    return true;
}
#endif

// ---- OpenCV's 8-bit bilinear resize ----------------------------------------------------------
// cv::resize, INTER_LINEAR, CV_8U (modules/imgproc/src/resize.cpp): coordinates
// f = (float)((d + 0.5) * scale - 0.5) with scale = 1 / ((double)dst / src), integer part by
// floor, weights (1 - f, f) converted to 11-bit fixed point with round-half-even
// (saturate_cast<short>(w * 2048)); the horizontal pass accumulates S[s] * a0 + S[s + 1] * a1 in
// int (columns at or beyond xmax copy S[s] * 2048), the vertical pass computes
// (((b0 * (R0 >> 4)) >> 16) + ((b1 * (R1 >> 4)) >> 16) + 2) >> 2. Along x the weight is snapped to
// 0 where the source index is clamped; along y only the row indices are clamped (the weights keep
// their fractional split). A 2x2 decimation in both directions is computed as INTER_AREA:
// (a + b + c + d + 2) >> 2.

struct Taps { std::vector<int> ofs; std::vector<short> w0, w1; int limit; };

short fixedWeight(float w) { return (short)std::lrintf(w * 2048.0f); }    // round half to even (default mode)

Taps makeTaps(int dst, int src, bool snap)
{
    Taps t;
    t.ofs.resize(dst); t.w0.resize(dst); t.w1.resize(dst); t.limit = dst;
    const double scale = 1.0 / ((double)dst / (double)src);
    for (int d = 0; d < dst; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)std::floor(f);
        f -= (float)s;
        if (snap) {
            if (s < 0) { f = 0.f; s = 0; }
            if (s + 1 >= src) {
                if (d < t.limit) t.limit = d;
                if (s >= src - 1) { f = 0.f; s = src - 1; }
            }
        }
        t.ofs[d] = s;
        t.w0[d] = fixedWeight(1.f - f);
        t.w1[d] = fixedWeight(f);
    }
    return t;
}

int clampIndex(int v, int n) { return v < 0 ? 0 : (v >= n ? n - 1 : v); }

} // namespace

bool readImageBGR(const std::string& path, ImageBGR& out)
{
    out = ImageBGR();
    if (path.empty()) return false;
    std::ifstream in(path, std::ios::binary);
    if (!in) return false;
    std::vector<unsigned char> bytes((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    if (readPNM(bytes, out)) return true;
#ifdef KAMINO_HAVE_ZLIB
    out = ImageBGR();
    if (readPNG(bytes, out)) return true;
#endif
    out = ImageBGR();
    return false;
}

ImageBGR flipHorizontal(const ImageBGR& src)
{
    ImageBGR dst;
    dst.width = src.width; dst.height = src.height;
    dst.data.resize(src.data.size());
    for (int y = 0; y < src.height; ++y)
        for (int x = 0; x < src.width; ++x)
            std::memcpy(&dst.data[((size_t)y * src.width + x) * 3], src.pixel(y, src.width - 1 - x), 3);
    return dst;
}

ImageBGR resizeLinear(const ImageBGR& src, int width, int height)
{
    ImageBGR dst;
    if (src.empty() || width <= 0 || height <= 0) return dst;
    dst.width = width; dst.height = height;
    dst.data.resize((size_t)width * height * 3);
    if (src.width == 2 * width && src.height == 2 * height) {
        for (int y = 0; y < height; ++y)
            for (int x = 0; x < width; ++x)
                for (int c = 0; c < 3; ++c) {
                    const int sum = src.pixel(2 * y, 2 * x)[c] + src.pixel(2 * y, 2 * x + 1)[c]
                                  + src.pixel(2 * y + 1, 2 * x)[c] + src.pixel(2 * y + 1, 2 * x + 1)[c];
                    dst.data[((size_t)y * width + x) * 3 + c] = (unsigned char)((sum + 2) >> 2);
                }
        return dst;
    }
    const Taps tx = makeTaps(width, src.width, true);
    const Taps ty = makeTaps(height, src.height, false);
    // horizontal pass of every source row (int accumulators)
    std::vector<int> rows((size_t)src.height * width * 3);
    for (int y = 0; y < src.height; ++y)
        for (int x = 0; x < width; ++x) {
            const unsigned char* a = src.pixel(y, tx.ofs[x]);
            int* r = &rows[((size_t)y * width + x) * 3];
            if (x < tx.limit) {
                const unsigned char* b = src.pixel(y, tx.ofs[x] + 1);
                for (int c = 0; c < 3; ++c) r[c] = a[c] * tx.w0[x] + b[c] * tx.w1[x];
            } else {
                for (int c = 0; c < 3; ++c) r[c] = a[c] * 2048;
            }
        }
    for (int y = 0; y < height; ++y) {
        const int* r0 = &rows[(size_t)clampIndex(ty.ofs[y], src.height) * width * 3];
        const int* r1 = &rows[(size_t)clampIndex(ty.ofs[y] + 1, src.height) * width * 3];
        const int b0 = ty.w0[y], b1 = ty.w1[y];
        unsigned char* d = &dst.data[(size_t)y * width * 3];
        for (int k = 0; k < width * 3; ++k) {
            int v = (((b0 * (r0[k] >> 4)) >> 16) + ((b1 * (r1[k] >> 4)) >> 16) + 2) >> 2;
            d[k] = (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v));
        }
    }
    return dst;
}
