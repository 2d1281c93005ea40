// Compile-time check of include/simlod_abi.h against the reference's OWN headers, field by field (run by
// tests/test_abi_and_library.py with nvcc when /root/reference is present; nothing is executed).
//   HostDeviceInterface.h  Uniforms, Stats, mat4     (host <-> device interface of the three kernels)
//   structures.cuh         Point, Chunk, OccupancyGrid, Node and the constants the kernels are built with
#include <cstddef>
#include <cstdint>
#include "HostDeviceInterface.h"      // reference (modules/progressive_octree)
#include "helper_math.h"              // reference: dot() used by structures.cuh
#include "structures.cuh"             // reference
#include "simlod_abi.h"               // ours

#define SAME_FIELD(OURS, REF, f) \
    static_assert(offsetof(OURS, f) == offsetof(REF, f) && sizeof(((OURS*)nullptr)->f) == sizeof(((REF*)nullptr)->f), "layout of " #REF "::" #f)

static_assert(sizeof(SimlodPoint) == sizeof(Point) && sizeof(SimlodChunk) == sizeof(Chunk) && sizeof(SimlodNode) == sizeof(Node), "record sizes");
static_assert(sizeof(SimlodOccupancyGrid) == sizeof(OccupancyGrid) && sizeof(SimlodUniforms) == sizeof(Uniforms) && sizeof(SimlodStats) == sizeof(Stats), "record sizes");
static_assert(sizeof(SimlodMat4) == sizeof(mat4), "mat4");
static_assert(SIMLOD_MAX_POINTS_PER_NODE == MAX_POINTS_PER_NODE && SIMLOD_POINTS_PER_CHUNK == POINTS_PER_CHUNK && SIMLOD_MAX_DEPTH == MAX_DEPTH, "constants");
static_assert(SIMLOD_GRID_WORDS == GRID_NUM_CELLS / 32u && SIMLOD_BATCH_STREAM_SIZE == BATCH_STREAM_SIZE, "constants");

SAME_FIELD(SimlodPoint, Point, x); SAME_FIELD(SimlodPoint, Point, y); SAME_FIELD(SimlodPoint, Point, z); SAME_FIELD(SimlodPoint, Point, color);
SAME_FIELD(SimlodChunk, Chunk, points); SAME_FIELD(SimlodChunk, Chunk, size); SAME_FIELD(SimlodChunk, Chunk, next);
SAME_FIELD(SimlodOccupancyGrid, OccupancyGrid, values);

SAME_FIELD(SimlodNode, Node, children); SAME_FIELD(SimlodNode, Node, counter); SAME_FIELD(SimlodNode, Node, numPoints);
SAME_FIELD(SimlodNode, Node, level); SAME_FIELD(SimlodNode, Node, X); SAME_FIELD(SimlodNode, Node, Y); SAME_FIELD(SimlodNode, Node, Z);
SAME_FIELD(SimlodNode, Node, countIteration); SAME_FIELD(SimlodNode, Node, countFlag); SAME_FIELD(SimlodNode, Node, name);
SAME_FIELD(SimlodNode, Node, visible); SAME_FIELD(SimlodNode, Node, isFiltered); SAME_FIELD(SimlodNode, Node, isLeaf); SAME_FIELD(SimlodNode, Node, isLarge);
SAME_FIELD(SimlodNode, Node, grid); SAME_FIELD(SimlodNode, Node, points); SAME_FIELD(SimlodNode, Node, voxelChunks);
SAME_FIELD(SimlodNode, Node, numVoxels); SAME_FIELD(SimlodNode, Node, numVoxelsStored);

SAME_FIELD(SimlodUniforms, Uniforms, width); SAME_FIELD(SimlodUniforms, Uniforms, height); SAME_FIELD(SimlodUniforms, Uniforms, time);
SAME_FIELD(SimlodUniforms, Uniforms, fovy_rad); SAME_FIELD(SimlodUniforms, Uniforms, world); SAME_FIELD(SimlodUniforms, Uniforms, view);
SAME_FIELD(SimlodUniforms, Uniforms, proj); SAME_FIELD(SimlodUniforms, Uniforms, transform); SAME_FIELD(SimlodUniforms, Uniforms, transform_updateBound);
SAME_FIELD(SimlodUniforms, Uniforms, transformInv_updateBound); SAME_FIELD(SimlodUniforms, Uniforms, persistentBufferCapacity);
SAME_FIELD(SimlodUniforms, Uniforms, momentaryBufferCapacity); SAME_FIELD(SimlodUniforms, Uniforms, frameCounter);
SAME_FIELD(SimlodUniforms, Uniforms, boxMin); SAME_FIELD(SimlodUniforms, Uniforms, boxMax);
SAME_FIELD(SimlodUniforms, Uniforms, showBoundingBox); SAME_FIELD(SimlodUniforms, Uniforms, showPoints); SAME_FIELD(SimlodUniforms, Uniforms, colorByNode);
SAME_FIELD(SimlodUniforms, Uniforms, colorByLOD); SAME_FIELD(SimlodUniforms, Uniforms, colorWhite); SAME_FIELD(SimlodUniforms, Uniforms, doUpdateVisibility);
SAME_FIELD(SimlodUniforms, Uniforms, doProgressive); SAME_FIELD(SimlodUniforms, Uniforms, LOD); SAME_FIELD(SimlodUniforms, Uniforms, useHighQualityShading);
SAME_FIELD(SimlodUniforms, Uniforms, minNodeSize); SAME_FIELD(SimlodUniforms, Uniforms, pointSize); SAME_FIELD(SimlodUniforms, Uniforms, updateStats);
SAME_FIELD(SimlodUniforms, Uniforms, enableEDL); SAME_FIELD(SimlodUniforms, Uniforms, edlStrength);

SAME_FIELD(SimlodStats, Stats, frameID); SAME_FIELD(SimlodStats, Stats, numNodes); SAME_FIELD(SimlodStats, Stats, numInner); SAME_FIELD(SimlodStats, Stats, numLeaves);
SAME_FIELD(SimlodStats, Stats, numNonemptyLeaves); SAME_FIELD(SimlodStats, Stats, numPoints); SAME_FIELD(SimlodStats, Stats, numVoxels);
SAME_FIELD(SimlodStats, Stats, allocatedBytes_momentary); SAME_FIELD(SimlodStats, Stats, allocatedBytes_persistent);
SAME_FIELD(SimlodStats, Stats, numVisibleNodes); SAME_FIELD(SimlodStats, Stats, numVisibleInner); SAME_FIELD(SimlodStats, Stats, numVisibleLeaves);
SAME_FIELD(SimlodStats, Stats, numVisiblePoints); SAME_FIELD(SimlodStats, Stats, numVisibleVoxels); SAME_FIELD(SimlodStats, Stats, numChunksPoints);
SAME_FIELD(SimlodStats, Stats, numChunksVoxels); SAME_FIELD(SimlodStats, Stats, batchletIndex); SAME_FIELD(SimlodStats, Stats, numPointsProcessed);
SAME_FIELD(SimlodStats, Stats, numAllocatedChunks); SAME_FIELD(SimlodStats, Stats, chunkPoolSize); SAME_FIELD(SimlodStats, Stats, dbg);
SAME_FIELD(SimlodStats, Stats, memCapacityReached);

int main() { return 0; }
