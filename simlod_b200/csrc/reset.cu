// reset.cu — re-initialises the octree state; drop-in for the reference's reset `kernel`
// (modules/progressive_octree/reset.cu:20-86), launched 1 block x 1 thread, cooperatively,
// by resetCUDA() (main_progressive_octree.cpp:333-361). Works for any launch shape.
#include <cooperative_groups.h>
#include <stdint.h>
#include "../../include/simlod_abi.h"

namespace cg = cooperative_groups;
struct CudaPrint;

extern "C" __global__ void kernel(const SimlodUniforms uniforms, uint8_t* buffer_octree, SimlodNode* nodes,
                                  SimlodStats* stats, CudaPrint* cudaprint,
                                  uint32_t* numBatchesUploaded_volatile, uint32_t* batchSizes) {
    cg::grid_group grid = cg::this_grid();
    SimlodHeapHeader* heap = reinterpret_cast<SimlodHeapHeader*>(buffer_octree);
    SimlodNode* root = &nodes[0];

    if (grid.thread_rank() == 0) {
        heap->buffer = buffer_octree;
        heap->offset = 16;                                   // reset.cu:42-43: header occupies the first 16 bytes

        uint64_t* s = reinterpret_cast<uint64_t*>(stats);    // *stats = Stats() (reset.cu:45)
        for (int i = 0; i < (int)(sizeof(SimlodStats) / 8); i++) s[i] = 0;
        stats->numNodes = 1;
        stats->frameID = (uint32_t)uniforms.frameCounter;

        for (int i = 0; i < 8; i++) root->children[i] = nullptr;
        root->isFiltered = 0;
        root->counter = 0;
        root->numPoints = 0;
        root->level = 0;
        root->X = 0; root->Y = 0; root->Z = 0;
        root->countIteration = 0;
        for (int i = 0; i < 20; i++) root->name[i] = 0;
        root->name[0] = 'r';
        root->numVoxels = 0;
        root->numVoxelsStored = 0;
        root->voxelChunks = nullptr;
        // The reference leaves root->points untouched and relies on a zeroed allocation
        // (reset.cu:55-69). A second reset would then link new chunks behind a dangling one;
        // clearing it is indistinguishable on every valid sequence.
        root->points = nullptr;
        // first heap allocation: the root's occupancy grid (reset.cu:69, utils.h.cu:185-197)
        root->grid = reinterpret_cast<SimlodOccupancyGrid*>(buffer_octree + heap->offset);
        heap->offset += SIMLOD_GRID_STRIDE;

        *numBatchesUploaded_volatile = 0;
        for (int i = 0; i < SIMLOD_BATCH_STREAM_SIZE; i++) batchSizes[i] = 0;
    }
    grid.sync();

    uint4* words = reinterpret_cast<uint4*>(root->grid->values);
    for (uint64_t i = grid.thread_rank(); i < SIMLOD_GRID_WORDS / 4; i += grid.size()) words[i] = make_uint4(0, 0, 0, 0);
}
