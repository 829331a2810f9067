// Minimal writer for Houdini "classic" binary geometry (.bgeo, version 5) point clouds: the
// file format Partio's BGEO back end emits for the reference's write_data_bgeo /
// write_particles_bgeo (kernel/KaminoSolver.cu:284-401). Big-endian, uncompressed -- Partio::write
// compresses only when the file name ends in ".gz" or forceCompressed is set
// (partio_headers/Partio.h:289-290), and the reference writes "<prefix><frame>.bgeo" with the
// default arguments. A ".gz" name is gzip-compressed here too (needs -DKAMINO_HAVE_ZLIB).
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
