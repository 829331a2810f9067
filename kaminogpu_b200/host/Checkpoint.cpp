#include "Checkpoint.h"

#include <cstdio>
#include <cstring>

namespace {

const char kMagic[8] = {'K', 'A', 'M', 'I', 'N', 'O', 'C', 'K'};

struct Checksum {
    uint32_t a = 1, b = 0;
    void add(const void* data, size_t bytes)
    {
        const unsigned char* p = static_cast<const unsigned char*>(data);
        while (bytes) {                                   // Adler-32 in blocks that cannot overflow
            const size_t n = bytes < 5552 ? bytes : 5552;
            for (size_t k = 0; k < n; ++k) { a += p[k]; b += a; }
            a %= 65521u; b %= 65521u;
            p += n; bytes -= n;
        }
    }
    uint32_t value() const { return (b << 16) | a; }
};

bool fail(std::string* error, const std::string& what) { if (error) *error = what; return false; }

bool expectedSizes(const CheckpointState& s)
{
    const size_t nT = s.header.nTheta, nP = s.header.nPhi;
    return s.velPhi.size() == nT * nP && s.velTheta.size() == (nT - 1) * nP && s.density.size() == nT * nP
        && s.particles.size() == 2 * (size_t)s.header.numParticles;
}

} // namespace

bool writeCheckpoint(const std::string& path, const CheckpointState& state, std::string* error)
{
    CheckpointState s = state;
    std::memcpy(s.header.magic, kMagic, 8);
    s.header.version = 1;
    if (s.header.nTheta < 2 || s.header.nPhi != 2 * s.header.nTheta || !expectedSizes(s))
        return fail(error, "checkpoint: array sizes do not match the header");
    const std::string tmp = path + ".tmp";
    std::FILE* f = std::fopen(tmp.c_str(), "wb");
    if (!f) return fail(error, "checkpoint: cannot open " + tmp);
    Checksum sum;
    bool ok = true;
    auto put = [&](const void* data, size_t bytes) {
        sum.add(data, bytes);
        if (bytes && std::fwrite(data, 1, bytes, f) != bytes) ok = false;
    };
    put(&s.header, sizeof(s.header));
    put(s.velPhi.data(), sizeof(float) * s.velPhi.size());
    put(s.velTheta.data(), sizeof(float) * s.velTheta.size());
    put(s.density.data(), sizeof(float) * s.density.size());
    put(s.particles.data(), sizeof(float) * s.particles.size());
    const uint32_t check = sum.value();
    if (std::fwrite(&check, sizeof(check), 1, f) != 1) ok = false;
    if (std::fclose(f) != 0) ok = false;
    if (!ok) { std::remove(tmp.c_str()); return fail(error, "checkpoint: write to " + tmp + " failed"); }
    if (std::rename(tmp.c_str(), path.c_str()) != 0) return fail(error, "checkpoint: cannot rename " + tmp);
    return true;
}

bool readCheckpoint(const std::string& path, CheckpointState& s, std::string* error)
{
    std::FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return fail(error, "checkpoint: cannot open " + path);
    Checksum sum;
    bool ok = true;
    auto get = [&](void* data, size_t bytes) {
        if (bytes && std::fread(data, 1, bytes, f) != bytes) { ok = false; return; }
        sum.add(data, bytes);
    };
    get(&s.header, sizeof(s.header));
    if (!ok || std::memcmp(s.header.magic, kMagic, 8) != 0 || s.header.version != 1) {
        std::fclose(f);
        return fail(error, "checkpoint: " + path + " is not a kamino checkpoint (version 1)");
    }
    const size_t nT = s.header.nTheta, nP = s.header.nPhi;
    if (nT < 2 || nT > 65536 || nP != 2 * nT || s.header.numParticles > (1ull << 36)) {
        std::fclose(f);
        return fail(error, "checkpoint: implausible shape in " + path);
    }
    s.velPhi.resize(nT * nP);
    s.velTheta.resize((nT - 1) * nP);
    s.density.resize(nT * nP);
    s.particles.resize(2 * (size_t)s.header.numParticles);
    get(s.velPhi.data(), sizeof(float) * s.velPhi.size());
    get(s.velTheta.data(), sizeof(float) * s.velTheta.size());
    get(s.density.data(), sizeof(float) * s.density.size());
    get(s.particles.data(), sizeof(float) * s.particles.size());
    uint32_t check = 0;
    if (ok && std::fread(&check, sizeof(check), 1, f) != 1) ok = false;
    std::fclose(f);
    if (!ok) return fail(error, "checkpoint: " + path + " is truncated");
    if (check != sum.value()) return fail(error, "checkpoint: checksum mismatch in " + path);
    return true;
}
