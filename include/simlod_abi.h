// simlod_abi.h — host/device ABI of the SimLOD hot path, restated byte-for-byte.
//
// These are the structs that cross the reference's launch boundary. The layouts are
// re-declared here (not copied) and every size/offset that the reference's kernels and
// host rely on is pinned with a static_assert, so that our kernels can be launched by
// the reference host (modules/progressive_octree/main_progressive_octree.cpp:333-546)
// and the reference kernels can be launched by our host, on the same buffers.
//
//   Point          modules/progressive_octree/structures.cuh:30-35   (16 B)
//   Chunk          modules/progressive_octree/structures.cuh:62-67   (16016 B, heap stride 16032)
//   OccupancyGrid  modules/progressive_octree/structures.cuh:69-72   (262144 B, heap stride 262160)
//   Node           modules/progressive_octree/structures.cuh:74-144  (152 B)
//   Uniforms/Stats modules/progressive_octree/HostDeviceInterface.h:10-71 (480 B / 112 B)
//   constants      modules/progressive_octree/structures.cuh:21-28
#pragma once
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
#define SIMLOD_STATIC_ASSERT(c, m) static_assert(c, m)
#else
#define SIMLOD_STATIC_ASSERT(c, m) _Static_assert(c, m)
#endif

enum {
    SIMLOD_MAX_POINTS_PER_NODE = 50000,     // leaf capacity, inclusive (structures.cuh:21, voxels.cu:211-212)
    SIMLOD_POINTS_PER_CHUNK    = 1000,      // structures.cuh:22
    SIMLOD_GRID_SIZE           = 128,       // structures.cuh:23
    SIMLOD_GRID_WORDS          = 128 * 128 * 128 / 32,
    SIMLOD_MAX_DEPTH           = 20,        // structures.cuh:25
    SIMLOD_BATCH_STREAM_SIZE   = 50,        // ring slots (structures.cuh:28, main.cpp:36)
    SIMLOD_MAX_BATCH_SIZE      = 1000000,   // points per ring slot (main.cpp:37)
    SIMLOD_CHUNK_STRIDE        = 16032,     // 16*((16016+16)/16)   (utils.h.cu:185-197)
    SIMLOD_GRID_STRIDE         = 262160,    // 16*((262144+16)/16)
};

typedef struct SimlodPoint {
    float    x, y, z;
    uint32_t color;   // 0xAABBGGRR
} SimlodPoint;

typedef struct SimlodChunk {
    SimlodPoint        points[SIMLOD_POINTS_PER_CHUNK];
    int32_t            size;        // never written by the reference
    int32_t            padding_0;
    struct SimlodChunk* next;
} SimlodChunk;

typedef struct SimlodOccupancyGrid {
    uint32_t values[SIMLOD_GRID_WORDS];   // bit index = x + 128*y + 128*128*z
} SimlodOccupancyGrid;

typedef struct SimlodNode {
    struct SimlodNode*   children[8];   //   0  child index = x<<2 | y<<1 | z
    uint32_t             counter;       //  64  points ever counted into this node while it was a leaf
    uint32_t             numPoints;     //  68  points stored in `points`
    uint32_t             level;         //  72
    uint32_t             X, Y, Z;       //  76  node coordinate at `level`
    uint32_t             countIteration;//  88
    uint32_t             countFlag;     //  92
    uint8_t              name[20];      //  96  'r' + one digit per level
    uint8_t              visible;       // 116  written by kernel_render
    uint8_t              isFiltered;    // 117
    uint8_t              isLeaf;        // 118  stays 1 forever in the reference (never updated)
    uint8_t              isLarge;       // 119  written by kernel_render
    SimlodOccupancyGrid* grid;          // 120
    SimlodChunk*         points;        // 128
    SimlodChunk*         voxelChunks;   // 136
    uint32_t             numVoxels;     // 144
    uint32_t             numVoxelsStored;//148
} SimlodNode;

typedef struct SimlodFloat4 { float x, y, z, w; } SimlodFloat4;
typedef struct SimlodMat4   { SimlodFloat4 rows[4]; } SimlodMat4;   // HostDeviceInterface.h:6-8 (row-major: host stores glm::transpose)

typedef struct SimlodUniforms {
    float      width;                      //   0
    float      height;                     //   4
    float      time;                       //   8
    float      fovy_rad;                   //  12
    SimlodMat4 world;                      //  16
    SimlodMat4 view;                       //  80
    SimlodMat4 proj;                       // 144
    SimlodMat4 transform;                  // 208
    SimlodMat4 transform_updateBound;      // 272
    SimlodMat4 transformInv_updateBound;   // 336
    uint64_t   persistentBufferCapacity;   // 400
    uint64_t   momentaryBufferCapacity;    // 408
    uint64_t   frameCounter;               // 416
    float      boxMin[3];                  // 424
    float      boxMax[3];                  // 436
    uint8_t    showBoundingBox;            // 448
    uint8_t    showPoints;                 // 449
    uint8_t    colorByNode;                // 450
    uint8_t    colorByLOD;                 // 451
    uint8_t    colorWhite;                 // 452
    uint8_t    doUpdateVisibility;         // 453
    uint8_t    doProgressive;              // 454
    uint8_t    _pad0;
    float      LOD;                        // 456
    uint8_t    useHighQualityShading;      // 460
    uint8_t    _pad1[3];
    float      minNodeSize;                // 464
    int32_t    pointSize;                  // 468
    uint8_t    updateStats;                // 472
    uint8_t    enableEDL;                  // 473
    uint8_t    _pad2[2];
    float      edlStrength;                // 476
} SimlodUniforms;

typedef struct SimlodStats {
    uint32_t frameID;                     //   0
    uint32_t numNodes;                    //   4  allocation cursor into nodes[] (1 after reset, +8 per split)
    uint32_t numInner;                    //   8
    uint32_t numLeaves;                   //  12
    uint32_t numNonemptyLeaves;           //  16
    uint32_t numPoints;                   //  20
    uint32_t numVoxels;                   //  24
    uint32_t _pad0;
    uint64_t allocatedBytes_momentary;    //  32
    uint64_t allocatedBytes_persistent;   //  40
    uint32_t numVisibleNodes;             //  48
    uint32_t numVisibleInner;             //  52
    uint32_t numVisibleLeaves;            //  56
    uint32_t numVisiblePoints;            //  60
    uint32_t numVisibleVoxels;            //  64
    uint32_t numChunksPoints;             //  68
    uint32_t numChunksVoxels;             //  72
    uint32_t batchletIndex;               //  76
    uint64_t numPointsProcessed;          //  80
    uint64_t numAllocatedChunks;          //  88
    uint64_t chunkPoolSize;               //  96
    uint32_t dbg;                         // 104
    uint8_t  memCapacityReached;          // 108
    uint8_t  _pad1[3];
} SimlodStats;

// header of the persistent heap (utils.h.cu:180-227, reset.cu:40-43): lives at heap byte 0
typedef struct SimlodHeapHeader {
    uint8_t* buffer;
    uint64_t offset;    // starts at 16; every alloc advances by 16*((size+16)/16)
} SimlodHeapHeader;

SIMLOD_STATIC_ASSERT(sizeof(SimlodPoint) == 16, "Point");
SIMLOD_STATIC_ASSERT(sizeof(SimlodChunk) == 16016, "Chunk");
SIMLOD_STATIC_ASSERT(offsetof(SimlodChunk, size) == 16000, "Chunk.size");
SIMLOD_STATIC_ASSERT(offsetof(SimlodChunk, next) == 16008, "Chunk.next");
SIMLOD_STATIC_ASSERT(sizeof(SimlodOccupancyGrid) == 262144, "OccupancyGrid");
SIMLOD_STATIC_ASSERT(sizeof(SimlodNode) == 152, "Node");
SIMLOD_STATIC_ASSERT(offsetof(SimlodNode, counter) == 64, "Node.counter");
SIMLOD_STATIC_ASSERT(offsetof(SimlodNode, numPoints) == 68, "Node.numPoints");
SIMLOD_STATIC_ASSERT(offsetof(SimlodNode, level) == 72, "Node.level");
SIMLOD_STATIC_ASSERT(offsetof(SimlodNode, X) == 76, "Node.X");
SIMLOD_STATIC_ASSERT(offsetof(SimlodNode, countIteration) == 88, "Node.countIteration");
SIMLOD_STATIC_ASSERT(offsetof(SimlodNode, name) == 96, "Node.name");
SIMLOD_STATIC_ASSERT(offsetof(SimlodNode, visible) == 116, "Node.visible");
SIMLOD_STATIC_ASSERT(offsetof(SimlodNode, isLarge) == 119, "Node.isLarge");
SIMLOD_STATIC_ASSERT(offsetof(SimlodNode, grid) == 120, "Node.grid");
SIMLOD_STATIC_ASSERT(offsetof(SimlodNode, points) == 128, "Node.points");
SIMLOD_STATIC_ASSERT(offsetof(SimlodNode, voxelChunks) == 136, "Node.voxelChunks");
SIMLOD_STATIC_ASSERT(offsetof(SimlodNode, numVoxels) == 144, "Node.numVoxels");
SIMLOD_STATIC_ASSERT(offsetof(SimlodNode, numVoxelsStored) == 148, "Node.numVoxelsStored");
SIMLOD_STATIC_ASSERT(sizeof(SimlodUniforms) == 480, "Uniforms");
SIMLOD_STATIC_ASSERT(offsetof(SimlodUniforms, transform) == 208, "Uniforms.transform");
SIMLOD_STATIC_ASSERT(offsetof(SimlodUniforms, transform_updateBound) == 272, "Uniforms.transform_updateBound");
SIMLOD_STATIC_ASSERT(offsetof(SimlodUniforms, persistentBufferCapacity) == 400, "Uniforms.persistentBufferCapacity");
SIMLOD_STATIC_ASSERT(offsetof(SimlodUniforms, boxMin) == 424, "Uniforms.boxMin");
SIMLOD_STATIC_ASSERT(offsetof(SimlodUniforms, boxMax) == 436, "Uniforms.boxMax");
SIMLOD_STATIC_ASSERT(offsetof(SimlodUniforms, showBoundingBox) == 448, "Uniforms.showBoundingBox");
SIMLOD_STATIC_ASSERT(offsetof(SimlodUniforms, LOD) == 456, "Uniforms.LOD");
SIMLOD_STATIC_ASSERT(offsetof(SimlodUniforms, useHighQualityShading) == 460, "Uniforms.useHighQualityShading");
SIMLOD_STATIC_ASSERT(offsetof(SimlodUniforms, minNodeSize) == 464, "Uniforms.minNodeSize");
SIMLOD_STATIC_ASSERT(offsetof(SimlodUniforms, pointSize) == 468, "Uniforms.pointSize");
SIMLOD_STATIC_ASSERT(offsetof(SimlodUniforms, edlStrength) == 476, "Uniforms.edlStrength");
SIMLOD_STATIC_ASSERT(sizeof(SimlodStats) == 112, "Stats");
SIMLOD_STATIC_ASSERT(offsetof(SimlodStats, numNodes) == 4, "Stats.numNodes");
SIMLOD_STATIC_ASSERT(offsetof(SimlodStats, allocatedBytes_momentary) == 32, "Stats.allocatedBytes_momentary");
SIMLOD_STATIC_ASSERT(offsetof(SimlodStats, numVisibleNodes) == 48, "Stats.numVisibleNodes");
SIMLOD_STATIC_ASSERT(offsetof(SimlodStats, batchletIndex) == 76, "Stats.batchletIndex");
SIMLOD_STATIC_ASSERT(offsetof(SimlodStats, numPointsProcessed) == 80, "Stats.numPointsProcessed");
SIMLOD_STATIC_ASSERT(offsetof(SimlodStats, numAllocatedChunks) == 88, "Stats.numAllocatedChunks");
SIMLOD_STATIC_ASSERT(offsetof(SimlodStats, chunkPoolSize) == 96, "Stats.chunkPoolSize");
SIMLOD_STATIC_ASSERT(offsetof(SimlodStats, memCapacityReached) == 108, "Stats.memCapacityReached");
