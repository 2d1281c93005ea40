// TEST INFRASTRUCTURE. C entry point around the reference's OWN LAS loader, which is compiled from
// where it lies (/root/reference/modules/progressive_octree/LasLoader.cpp, unmodified) together with
// this shim into oracle/_ref/libref_las.so. Nothing of the reference is copied here: the two reference
// functions are only declared through the reference's own header.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
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
// the reference's loader threads (spawnLoader, main_progressive_octree.cpp:811-958, one loadLasNative call per
// batch per thread): `threads` threads take batches of `batchPoints` points round-robin
int ref_las_load_parallel(const char* path, uint64_t firstPoint, uint64_t numPoints, uint64_t batchPoints, void* target,
                          const double* translation, int threads) {
    LasHeader h = loadHeader(std::string(path));
    std::string file(path);
    uint64_t numBatches = (numPoints + batchPoints - 1) / batchPoints;
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) {
        pool.emplace_back([=]() {
            double tr[3] = {translation[0], translation[1], translation[2]};
            for (uint64_t b = t; b < numBatches; b += threads) {
                uint64_t first = b * batchPoints, n = std::min(batchPoints, numPoints - first);
                loadLasNative(file, h, firstPoint + first, n, (char*)target + first * 16, tr);
            }
        });
    }
    for (auto& th : pool) th.join();
    return 0;
}
// timing harness: long-lived loader threads as in the reference (their thread_local read buffers are
// allocated once); every thread loads its batches once untimed, all threads meet, then the timed pass.
// Returns the wall time of the timed pass in seconds.
double ref_las_bench(const char* path, uint64_t numPoints, uint64_t batchPoints, void* target, int threads) {
    LasHeader h = loadHeader(std::string(path));
    std::string file(path);
    uint64_t numBatches = (numPoints + batchPoints - 1) / batchPoints;
    std::atomic<int> arrived{0}, finished{0};
    std::chrono::steady_clock::time_point t0, t1;
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) {
        pool.emplace_back([&, t]() {
            double tr[3] = {0.0, 0.0, 0.0};
            auto pass = [&]() {
                for (uint64_t b = t; b < numBatches; b += threads) {
                    uint64_t first = b * batchPoints, n = std::min(batchPoints, numPoints - first);
                    loadLasNative(file, h, first, n, (char*)target + first * 16, tr);
                }
            };
            pass();
            if (arrived.fetch_add(1) + 1 == threads) t0 = std::chrono::steady_clock::now();
            while (arrived.load() < threads) std::this_thread::yield();
            pass();
            if (finished.fetch_add(1) + 1 == threads) t1 = std::chrono::steady_clock::now();
        });
    }
    for (auto& th : pool) th.join();
    return std::chrono::duration<double>(t1 - t0).count();
}
}
