#include "BgeoWriter.h"

#include <cstdint>
#include <cstdio>
#include <cstring>

#ifdef KAMINO_HAVE_ZLIB
#include <zlib.h>
#endif

namespace {

struct Buffer
{
    std::vector<unsigned char> bytes;
    void put8(unsigned char v) { bytes.push_back(v); }
    void put16(uint16_t v) { put8((unsigned char)(v >> 8)); put8((unsigned char)v); }
    void put32(uint32_t v) { put16((uint16_t)(v >> 16)); put16((uint16_t)v); }
    void putFloat(float f) { uint32_t u; std::memcpy(&u, &f, 4); put32(u); }
    void putString(const std::string& s) { put16((uint16_t)s.size()); bytes.insert(bytes.end(), s.begin(), s.end()); }
};

} // namespace

bool writeBgeo(const std::string& path, const std::vector<float>& positions,
               const std::vector<BgeoAttribute>& attributes)
{
    const size_t nPoints = positions.size() / 3;
    Buffer out;
    size_t perPoint = 16;
    for (const BgeoAttribute& a : attributes) perPoint += 4 * (size_t)a.count;
    out.bytes.reserve(64 + nPoints * perPoint);

    out.put32(0x4267656Fu);      // "Bgeo"
    out.put8('V');
    out.put32(5);                // version
    out.put32((uint32_t)nPoints);
    out.put32(0);                // primitives
    out.put32(0);                // point groups
    out.put32(0);                // primitive groups
    out.put32((uint32_t)attributes.size());   // point attributes (position is implicit)
    out.put32(0);                // vertex attributes
    out.put32(0);                // primitive attributes
    out.put32(0);                // detail attributes
    for (const BgeoAttribute& a : attributes) {
        out.putString(a.name);
        out.put16((uint16_t)a.count);
        out.put32(a.isVector ? 5u : 0u);
        for (int k = 0; k < a.count; ++k) out.putFloat(0.0f);    // defaults
    }
    for (size_t p = 0; p < nPoints; ++p) {
        out.putFloat(positions[3 * p]);
        out.putFloat(positions[3 * p + 1]);
        out.putFloat(positions[3 * p + 2]);
        out.putFloat(1.0f);      // homogeneous w
        for (const BgeoAttribute& a : attributes)
            for (int k = 0; k < a.count; ++k) out.putFloat(a.values[p * a.count + k]);
    }
    out.put8(0x00);              // end of extra section
    out.put8(0xff);

    // Partio::write(file, parts) as the reference calls it (kernel/KaminoSolver.cu:356,400; forceCompressed
    // defaults to false): compressed only when the name ends in ".gz" (partio_headers/Partio.h:289-290)
    const bool gz = path.size() >= 3 && path.compare(path.size() - 3, 3, ".gz") == 0;
#ifdef KAMINO_HAVE_ZLIB
    if (gz) {
        gzFile f = gzopen(path.c_str(), "wb");
        if (!f) return false;
        size_t done = 0;
        while (done < out.bytes.size()) {
            const unsigned chunk = (unsigned)((out.bytes.size() - done) < (1u << 30) ? (out.bytes.size() - done) : (1u << 30));
            if (gzwrite(f, out.bytes.data() + done, chunk) != (int)chunk) { gzclose(f); return false; }
            done += chunk;
        }
        return gzclose(f) == Z_OK;
    }
#else
    if (gz) return false;        // built without zlib
#endif
    std::FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    const bool ok = std::fwrite(out.bytes.data(), 1, out.bytes.size(), f) == out.bytes.size();
    return (std::fclose(f) == 0) && ok;
}
