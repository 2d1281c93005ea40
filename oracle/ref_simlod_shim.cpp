// TEST INFRASTRUCTURE. C entry points around the reference's OWN .simlod loader, compiled from where it lies
// (/root/reference/modules/progressive_octree/SimlodLoader.cpp, unmodified, with `-include cstdint` because the file
// relies on a transitive include) together with this shim into oracle/_ref/libref_simlod.so.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <string>
#include <thread>
#include <vector>
#include "unsuck.hpp"
#include "SimlodLoader.h"      // -I /root/reference/modules/progressive_octree: loadFileNative()

// see ref_las_shim.cpp: only reached on unsuck.hpp's out-of-memory path
MemoryData getMemoryData() { return MemoryData(); }

extern "C" {
// loadFileNative(file, firstByte, numBytes, target, &padding) — SimlodLoader.cpp:147-157
int ref_simlod_load(const char* path, uint64_t firstByte, uint64_t numBytes, void* target) {
    uint64_t padding = 0;
    loadFileNative(std::string(path), firstByte, numBytes, target, &padding);
    return (int)padding;
}
// the reference's loader threads (spawnLoader, main_progressive_octree.cpp:811-958): long-lived threads, one
// loadFileNative call per 1 M-point batch; one untimed pass, then the timed pass. Returns seconds.
double ref_simlod_bench(const char* path, uint64_t numPoints, uint64_t batchPoints, void* target, int threads) {
    std::string file(path);
    uint64_t numBatches = (numPoints + batchPoints - 1) / batchPoints;
    std::atomic<int> arrived{0}, finished{0};
    std::chrono::steady_clock::time_point t0, t1;
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) {
        pool.emplace_back([&, t]() {
            auto pass = [&]() {
                for (uint64_t b = t; b < numBatches; b += threads) {
                    uint64_t first = b * batchPoints, n = std::min(batchPoints, numPoints - first);
                    uint64_t padding = 0;
                    loadFileNative(file, 24 + first * 16, n * 16, (char*)target + first * 16, &padding);
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
