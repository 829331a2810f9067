// Minimal writer for Houdini "classic" binary geometry (.bgeo, version 5) point clouds: the
// file format Partio's BGEO back end emits for the reference's write_data_bgeo /
// write_particles_bgeo (kernel/KaminoSolver.cu:284-401). Big-endian; gzip-compressed like
// Partio::write's default when built with zlib (-DKAMINO_HAVE_ZLIB), plain otherwise --
// Partio's reader sniffs the gzip magic and accepts both.
#pragma once

#include <string>
#include <vector>

struct BgeoAttribute
{
    std::string name;     // "v", "density", "color"
    int count;            // floats per point
    bool isVector;        // Houdini type 5 (vector) when true, 0 (float) otherwise
    std::vector<float> values;   // count floats per point
};

// positions: 3 floats per point. Returns false if the file cannot be written.
bool writeBgeo(const std::string& path, const std::vector<float>& positions,
               const std::vector<BgeoAttribute>& attributes);
