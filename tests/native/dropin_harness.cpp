// dropin_harness.cpp — drives the three SHIPPED cubins (simlod_b200/cubin/simlod_{reset,construct,render}.cubin) exactly the way
// the reference host does, WITHOUT going through simlod_b200/csrc/host.cpp: the buffers of initCudaProgram
// (main_progressive_octree.cpp:552-586, including the 300 000 000-byte momentary buffer that is NOT cleared — it is
// filled with a garbage pattern here to make that explicit), the argument arrays of resetCUDA / updateOctree /
// renderCUDA (main.cpp:337-345, 374-382, 499-507), the launch shapes (reset 1 x 1, construct numSMs x 256, render
// occupancy x numSMs x 256, always cooperative), and the uploader step (main.cpp:1040-1050).
//
//   dropin_harness <cubin dir> <points.bin (16-byte XYZRGBA)> <uniforms.bin (480 bytes)> <out prefix>
// writes <out>.stats (112 bytes, Stats after the frame), <out>.nodes, <out>.heap (octree image for the
// canonicaliser), <out>.fb (W*H u64) and prints one summary line. Test infrastructure; links only libcuda.
#include <cuda.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/simlod_abi.h"

#define CU(call) do { CUresult _r = (call); if (_r != CUDA_SUCCESS) { const char* s = nullptr; cuGetErrorString(_r, &s); \
    fprintf(stderr, "%s failed: %s (%d) at line %d\n", #call, s ? s : "?", (int)_r, __LINE__); return 2; } } while (0)

static std::vector<char> slurp(const std::string& path) {
    std::vector<char> out;
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return out;
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    out.resize((size_t)n);
    if (n && fread(out.data(), 1, (size_t)n, f) != (size_t)n) out.clear();
    fclose(f);
    return out;
}
static bool dump(const std::string& path, const void* p, size_t n) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    bool ok = fwrite(p, 1, n, f) == n;
    fclose(f);
    return ok;
}

int main(int argc, char** argv) {
    if (argc < 5) { fprintf(stderr, "usage: %s <cubin dir> <points.bin> <uniforms.bin> <out prefix>\n", argv[0]); return 1; }
    const std::string dir = argv[1], out = argv[4];
    std::vector<char> pts = slurp(argv[2]), uni = slurp(argv[3]);
    if (pts.empty() || uni.size() != sizeof(SimlodUniforms)) { fprintf(stderr, "bad input files\n"); return 1; }
    const uint64_t numPoints = pts.size() / 16;
    SimlodUniforms uniforms;
    memcpy(&uniforms, uni.data(), sizeof(uniforms));
    const int width = (int)uniforms.width, height = (int)uniforms.height;

    CU(cuInit(0));
    CUdevice dev; CU(cuDeviceGet(&dev, 0));
    CUcontext ctx; CU(cuDevicePrimaryCtxRetain(&ctx, dev)); CU(cuCtxSetCurrent(ctx));
    int numSMs = 0; CU(cuDeviceGetAttribute(&numSMs, CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT, dev));

    // the three programs, looked up by the reference's kernel names (CudaModularProgram.h:245-252)
    CUmodule mods[3]; CUfunction fns[3];
    const char* files[3] = {"simlod_reset.cubin", "simlod_construct.cubin", "simlod_render.cubin"};
    const char* names[3] = {"kernel", "kernel_construct", "kernel_render"};
    for (int i = 0; i < 3; i++) {
        std::vector<char> image = slurp(dir + "/" + files[i]);
        if (image.empty()) { fprintf(stderr, "cannot read %s/%s\n", dir.c_str(), files[i]); return 1; }
        image.push_back(0);
        CU(cuModuleLoadData(&mods[i], image.data()));
        CU(cuModuleGetFunction(&fns[i], mods[i], names[i]));
    }

    // initCudaProgram (main.cpp:552-586)
    const size_t MOMENTARY = 300000000ull, NODES = 40000000ull, RENDER = 200000000ull, RING = 50ull * 1000000ull * 16ull;
    const size_t PERSISTENT = 6ull << 30;
    CUdeviceptr buffer, nodes, renderbuffer, stats, numBatchesUploaded, batchSizes, frameStart, cudaprint, ring, persistent;
    CU(cuMemAlloc(&buffer, MOMENTARY)); CU(cuMemAlloc(&nodes, NODES)); CU(cuMemAlloc(&renderbuffer, RENDER));
    CU(cuMemAlloc(&stats, sizeof(SimlodStats))); CU(cuMemAlloc(&numBatchesUploaded, 4)); CU(cuMemAlloc(&batchSizes, 4 * 50));
    CU(cuMemAlloc(&frameStart, 8)); CU(cuMemAlloc(&cudaprint, 1024 * 1000 + 16)); CU(cuMemAlloc(&ring, RING)); CU(cuMemAlloc(&persistent, PERSISTENT));
    CU(cuMemsetD8(buffer, 0xCD, MOMENTARY));             // "not cleared": whatever the allocation holds; make it hostile
    CU(cuMemsetD8(renderbuffer, 0xCD, RENDER));
    CU(cuMemsetD8(persistent, 0xCD, 64ull << 20));
    CU(cuMemsetD8(nodes, 0, NODES));                     // the reference relies on a zeroed nodes[] (reset.cu:55-69)
    CU(cuMemsetD8(stats, 0, sizeof(SimlodStats)));
    CU(cuMemsetD8(cudaprint, 0, 16));
    uniforms.persistentBufferCapacity = PERSISTENT;      // main.cpp:325-326
    uniforms.momentaryBufferCapacity = MOMENTARY;

    // GL colour attachment stand-in (main.cpp:472-486)
    CUDA_ARRAY3D_DESCRIPTOR ad{}; ad.Width = (size_t)width; ad.Height = (size_t)height; ad.Depth = 0;
    ad.Format = CU_AD_FORMAT_UNSIGNED_INT8; ad.NumChannels = 4; ad.Flags = CUDA_ARRAY3D_SURFACE_LDST;
    CUarray array; CU(cuArray3DCreate(&array, &ad));
    CUDA_RESOURCE_DESC rd{}; rd.resType = CU_RESOURCE_TYPE_ARRAY; rd.res.array.hArray = array;
    CUsurfObject surface; CU(cuSurfObjectCreate(&surface, &rd));

    // resetCUDA (main.cpp:333-361): 1 block x 1 thread, cooperative
    {
        void* args[] = {&uniforms, &persistent, &nodes, &stats, &cudaprint, &numBatchesUploaded, &batchSizes};
        CU(cuLaunchCooperativeKernel(fns[0], 1, 1, 1, 1, 1, 1, 0, 0, args));
        CU(cuCtxSynchronize());
    }
    // uploader (main.cpp:1040-1050) + updateOctree (main.cpp:364-428), numSMs blocks of 256 threads
    SimlodStats hs{};
    uint32_t uploaded = 0;
    int launches = 0;
    const uint64_t numBatches = (numPoints + 999999) / 1000000;
    for (uint64_t b = 0; b < numBatches; b++) {
        const uint32_t count = (uint32_t)std::min<uint64_t>(1000000, numPoints - b * 1000000);
        const uint32_t slot = uploaded % 50;
        CU(cuMemcpyHtoD(ring + (size_t)slot * 16000000ull, pts.data() + b * 16000000ull, (size_t)count * 16));
        CU(cuMemcpyHtoD(batchSizes + 4 * slot, &count, 4));
        uploaded++;
        CU(cuMemcpyHtoD(numBatchesUploaded, &uploaded, 4));
        // one launch per frame while batches stream in; here: a launch after every second batch, then drain
        if ((b & 1) == 1 || b + 1 == numBatches) {
            for (int guard = 0; guard < 100; guard++) {
                void* args[] = {&uniforms, &ring, &buffer, &persistent, &nodes, &stats, &frameStart, &cudaprint, &numBatchesUploaded, &batchSizes};
                CU(cuLaunchCooperativeKernel(fns[1], (unsigned)numSMs, 1, 1, 256, 1, 1, 0, 0, args));
                CU(cuCtxSynchronize());
                launches++;
                CU(cuMemcpyDtoH(&hs, stats, sizeof(hs)));
                if (hs.batchletIndex >= uploaded || hs.memCapacityReached) break;
            }
        }
    }
    // renderCUDA (main.cpp:465-546)
    int occ = 0;
    CU(cuOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fns[2], 256, 0));
    {
        void* args[] = {&renderbuffer, &uniforms, &nodes, &surface, &stats, &frameStart, &cudaprint};
        CU(cuLaunchCooperativeKernel(fns[2], (unsigned)(occ * numSMs), 1, 1, 256, 1, 1, 0, 0, args));
        CU(cuCtxSynchronize());
    }
    CU(cuMemcpyDtoH(&hs, stats, sizeof(hs)));

    std::vector<char> hnodes((size_t)hs.numNodes * sizeof(SimlodNode));
    CU(cuMemcpyDtoH(hnodes.data(), nodes, hnodes.size()));
    uint64_t heapUsed = 0;
    CU(cuMemcpyDtoH(&heapUsed, persistent + 8, 8));
    std::vector<char> heap((size_t)heapUsed);
    CU(cuMemcpyDtoH(heap.data(), persistent, heap.size()));
    std::vector<uint64_t> fb((size_t)width * height);
    CU(cuMemcpyDtoH(fb.data(), renderbuffer + 31200144ull, fb.size() * 8));
    const uint64_t addrs[2] = {(uint64_t)nodes, (uint64_t)persistent};
    if (!dump(out + ".stats", &hs, sizeof(hs)) || !dump(out + ".nodes", hnodes.data(), hnodes.size()) || !dump(out + ".heap", heap.data(), heap.size()) ||
        !dump(out + ".fb", fb.data(), fb.size() * 8) || !dump(out + ".addrs", addrs, sizeof(addrs))) { fprintf(stderr, "cannot write outputs\n"); return 1; }
    printf("dropin_harness: %llu points, %d construct launches on %d blocks, numNodes %u numPoints %u numVoxels %u dbg %u, render %d blocks, visible nodes %u\n",
           (unsigned long long)numPoints, launches, numSMs, hs.numNodes, hs.numPoints, hs.numVoxels, hs.dbg, occ * numSMs, hs.numVisibleNodes);
    return 0;
}
