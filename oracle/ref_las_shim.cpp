// TEST INFRASTRUCTURE. C entry point around the reference's OWN LAS loader, which is compiled from
// where it lies (/root/reference/modules/progressive_octree/LasLoader.cpp, unmodified) together with
// this shim into oracle/_ref/libref_las.so. Nothing of the reference is copied here: the two reference
// functions are only declared through the reference's own header.
#include <cstring>
#include <string>
#include "LasLoader.h"      // -I /root/reference/modules/progressive_octree: LasHeader, loadHeader(), loadLasNative()

// unsuck.hpp declares getMemoryData() and calls it only on its out-of-memory error path; its definition
// lives in include/unsuck_platform_specific.cpp, which needs OpenGL headers that do not exist here.
// This stand-in is never reached by the loader.
MemoryData getMemoryData() { return MemoryData(); }

extern "C" {
// header fields as the reference parses them (LasLoader.h:21-55)
int ref_las_header(const char* path, uint64_t* numPoints, uint64_t* bytesPerPoint, uint64_t* format, uint64_t* offsetToPointData,
                   double* scale, double* offset, double* min, double* max) {
    LasHeader h = loadHeader(std::string(path));
    *numPoints = h.numPoints; *bytesPerPoint = h.bytesPerPoint; *format = h.format; *offsetToPointData = h.offsetToPointData;
    memcpy(scale, h.scale, 24); memcpy(offset, h.offset, 24); memcpy(min, h.min, 24); memcpy(max, h.max, 24);
    return 0;
}
// loadLasNative(file, header, firstPoint, numPoints, target, translation) — LasLoader.cpp:169-226
int ref_las_load(const char* path, uint64_t firstPoint, uint64_t numPoints, void* target, const double* translation) {
    LasHeader h = loadHeader(std::string(path));
    double t[3] = {translation[0], translation[1], translation[2]};
    loadLasNative(std::string(path), h, firstPoint, numPoints, target, t);
    return 0;
}
}
