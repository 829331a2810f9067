// TEST INFRASTRUCTURE ONLY (oracle/): link shim that lets the UNMODIFIED reference
// translation units under /root/reference/KaminoGPU/kernel link on Linux.
//
// The reference links against OpenCV 2.4.13 and Partio, of which only Windows
// binaries are vendored (KaminoGPU/lib/*.lib). With "null" image paths none of the
// image functions does real work, and the benchmark configs turn file output off,
// so the ten unresolved third-party symbols are provided
// here as inert stand-ins:
//   cv::imread  -> always an empty Mat ("no image"), which is the path the
//                  reference takes for a "" file name (KaminoParticles.cu:7-11,
//                  KaminoSolver.cu:243-254)
//   cv::flip / cv::resize -> never reached with an empty Mat; abort if they are
//   Partio::create -> an in-memory particle set that accepts attributes/particles
//   Partio::write  -> no-op ("output off")
// Nothing here is part of the product; nothing here is copied from the reference.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "opencv2/opencv.hpp"
#include "Partio.h"

namespace cv {

Mat imread(const string&, int) { return Mat(); }

void flip(InputArray, OutputArray, int)
{
    std::fprintf(stderr, "link_shim: cv::flip reached (image input is out of scope)\n");
    std::abort();
}

void resize(InputArray, OutputArray, Size, double, double, int)
{
    std::fprintf(stderr, "link_shim: cv::resize reached (image input is out of scope)\n");
    std::abort();
}

void fastFree(void* ptr) { std::free(ptr); }

void Mat::deallocate()
{
    // An empty Mat owns nothing; a non-empty one is never created by this shim.
}

void Mat::copySize(const Mat& m)
{
    dims = m.dims;
    rows = m.rows;
    cols = m.cols;
}

} // namespace cv

// cv::_InputArray / cv::_OutputArray have a dozen virtuals whose vtables would have to
// be emitted by any real constructor definition. The two constructors the reference
// objects import are only reachable when an image was loaded (never, with this shim),
// so they are provided under their mangled names as plain functions that abort.
extern "C" {
void shim_cv_InputArray_ctor(void*, const void*) __asm__("_ZN2cv11_InputArrayC1ERKNS_3MatE");
void shim_cv_OutputArray_ctor(void*, void*) __asm__("_ZN2cv12_OutputArrayC1ERNS_3MatE");
void shim_cv_InputArray_ctor(void*, const void*)
{
    std::fprintf(stderr, "link_shim: cv::_InputArray constructed (image input is out of scope)\n");
    std::abort();
}
void shim_cv_OutputArray_ctor(void*, void*)
{
    std::fprintf(stderr, "link_shim: cv::_OutputArray constructed (image input is out of scope)\n");
    std::abort();
}
}

namespace Partio {

namespace {

// Minimal in-memory particle container: enough for addAttribute / addParticle /
// dataWrite / release, which is all the reference's writers touch.
class ShimParticles : public ParticlesDataMutable
{
    struct Column { ParticleAttribute attr; std::vector<char> bytes; int stride; };
    mutable std::vector<Column> columns;
    int count = 0;
    std::vector<std::string> noStrings;

public:
    void release() const override { delete this; }
    int numParticles() const override { return count; }
    int numAttributes() const override { return (int)columns.size(); }
    int numFixedAttributes() const override { return 0; }
    bool attributeInfo(const char* name, ParticleAttribute& a) const override
    {
        for (auto& c : columns) if (c.attr.name == name) { a = c.attr; return true; }
        return false;
    }
    bool fixedAttributeInfo(const char*, FixedAttribute&) const override { return false; }
    bool attributeInfo(const int idx, ParticleAttribute& a) const override
    {
        if (idx < 0 || idx >= (int)columns.size()) return false;
        a = columns[idx].attr; return true;
    }
    bool fixedAttributeInfo(const int, FixedAttribute&) const override { return false; }

    const std::vector<std::string>& indexedStrs(const ParticleAttribute&) const override { return noStrings; }
    const std::vector<std::string>& fixedIndexedStrs(const FixedAttribute&) const override { return noStrings; }
    int lookupIndexedStr(const ParticleAttribute&, const char*) const override { return -1; }
    int lookupFixedIndexedStr(const FixedAttribute&, const char*) const override { return -1; }
    void dataAsFloat(const ParticleAttribute&, const int, const ParticleIndex*, const bool, float*) const override {}
    void findPoints(const float[3], const float[3], std::vector<ParticleIndex>&) const override {}
    float findNPoints(const float[3], int, const float, std::vector<ParticleIndex>&, std::vector<float>&) const override { return 0.f; }
    int findNPoints(const float[3], int, const float, ParticleIndex*, float*, float*) const override { return 0; }
    const_iterator setupConstIterator(const int) const override { return const_iterator(); }

    int registerIndexedStr(const ParticleAttribute&, const char*) override { return -1; }
    int registerFixedIndexedStr(const FixedAttribute&, const char*) override { return -1; }
    void setIndexedStr(const ParticleAttribute&, int, const char*) override {}
    void setFixedIndexedStr(const FixedAttribute&, int, const char*) override {}
    void sort() override {}
    ParticleAttribute addAttribute(const char* name, ParticleAttributeType type, const int n) override
    {
        Column c;
        c.attr.type = type; c.attr.count = n; c.attr.name = name;
        c.attr.attributeIndex = (int)columns.size();
        c.stride = TypeSize(type) * n;
        c.bytes.resize((size_t)c.stride * (size_t)count);
        columns.push_back(c);
        return c.attr;
    }
    FixedAttribute addFixedAttribute(const char* name, ParticleAttributeType type, const int n) override
    {
        FixedAttribute f; f.type = type; f.count = n; f.name = name; f.attributeIndex = -1; return f;
    }
    ParticleIndex addParticle() override
    {
        ++count;
        for (auto& c : columns) {
            size_t need = (size_t)c.stride * (size_t)count;
            if (c.bytes.size() < need) c.bytes.resize(need < 4096 ? 4096 : need * 2);
        }
        return (ParticleIndex)(count - 1);
    }
    iterator addParticles(const int n) override
    {
        for (int i = 0; i < n; ++i) addParticle();
        return iterator();
    }
    iterator setupIterator(const int) override { return iterator(); }

private:
    void* dataInternal(const ParticleAttribute& a, const ParticleIndex i) const override
    {
        Column& c = columns[a.attributeIndex];
        return c.bytes.data() + (size_t)c.stride * (size_t)i;
    }
    void* fixedDataInternal(const FixedAttribute&) const override { return nullptr; }
    void dataInternalMultiple(const ParticleAttribute&, const int, const ParticleIndex*, const bool, char*) const override {}
};

} // namespace

ParticlesDataMutable* create() { return new ShimParticles(); }

void write(const char*, const ParticlesData&, const bool, bool, std::ostream&) {}

} // namespace Partio
