// construct.cu — incremental octree/LOD builder for sm_100a (B200).
//
// Drop-in for the reference's `kernel_construct`
// (modules/progressive_octree/progressive_octree_voxels.cu:804-1010): same extern "C"
// name, same argument list, same Node/Chunk/OccupancyGrid/Stats contents afterwards
// (canonical form: DESIGN.md §3), launched cooperatively with 256-thread blocks by
// updateOctree() (main_progressive_octree.cpp:364-428). It is NOT a translation of that
// kernel; what is preserved is the observable state, what is new is how it is computed:
//
//   reference                                           here
//   ------------------------------------------------    ------------------------------------------
//   3 full passes over the batch (count, voxel-         1 streaming pass (count + voxel-sample fused),
//   sample, insert), each re-descending the tree        leaf id + slot cached per point (8 B), 1 insert pass
//   via 64-byte children[] pointer arrays               descent through a 4 B/node first-child table
//   contiguous-per-thread ranges (uncoalesced)          grid-stride 128-bit coalesced loads
//   atomicAdd(numPoints) per point at insert            slot = warp-aggregated counter add (no atomics at insert)
//   chunk lists walked i/1000 hops per point/voxel      per-batch chunk directory: O(1) address per element
//   one thread walks/extends each node's list           tail pointers kept per node; only dirty nodes visited
//   1 global atomic per created voxel (backlog)         one atomic per warp per level (ballot-aggregated)
//   >= 24 grid-wide barriers per batch                  3 (+2 per split round)
//
// The scratch ("momentary") buffer is carved with our own layout (Scratch below); it fits in
// the 300 000 000 bytes the unmodified host allocates (main_progressive_octree.cpp:554),
// unlike the reference's carve-out which needs 408 800 192 (voxels.cu:834-856).
#include <cooperative_groups.h>
#include <stdint.h>
#include "../../include/simlod_abi.h"
#include "fpmath.cuh"

namespace cg = cooperative_groups;

typedef SimlodPoint Point;
typedef SimlodChunk Chunk;
typedef SimlodNode Node;
typedef SimlodStats Stats;
typedef SimlodUniforms Uniforms;
typedef SimlodHeapHeader Heap;
struct CudaPrint;   // opaque: the reference's debug channel is a dead parameter (CudaPrint.cuh:51)

// ------------------------------------------------------------------------------------------
// scratch layout inside the momentary buffer (all offsets 256-byte aligned)
// ------------------------------------------------------------------------------------------
namespace scratch {
constexpr uint64_t NODE_CAP       = 263168;            // >= floor(40 000 000 / 152) nodes the host allocates
constexpr uint64_t MAX_BATCH      = SIMLOD_MAX_BATCH_SIZE;
constexpr uint64_t SPILL_CAP      = 3ull << 20;        // spilled points per batch (reference re-inserts <= 3 000 001)
constexpr uint64_t ITEM_CAP       = MAX_BATCH + SPILL_CAP;
constexpr uint64_t VOXEL_CAP      = 8ull << 20;        // voxels created per batch (reference backlog: 10 M)
constexpr uint64_t DIR_CAP        = 1ull << 20;        // chunk directory entries per batch
constexpr uint64_t QUEUE_CAP      = 4ull << 20;        // free-chunk stack (reference: 1 M)
constexpr uint64_t SPILLNODE_CAP  = 100000;            // voxels.cu:847

constexpr uint64_t ROW_CAP        = 65536;             // leaves that hold points at the same time (x 64 chunk slots)
constexpr uint64_t ROW_SLOTS      = 64;                // chunk pointers per leaf row (a leaf holds <= 50 chunks)
constexpr uint64_t VOXEL_SHARED   = 1ull << 20;        // tail of the voxel backlog shared by all blocks (overflow of a block's own segment)
constexpr uint64_t BLOCK_CAP      = 4096;              // per-block cursor slots (grid sizes up to 4096 blocks)

constexpr uint64_t align256(uint64_t x) { return (x + 255) & ~255ull; }
constexpr uint64_t OFF_CTL        = 0;
constexpr uint64_t OFF_FIRSTCHILD = 4096;
constexpr uint64_t OFF_GRIDPTR    = align256(OFF_FIRSTCHILD + NODE_CAP * 4);
constexpr uint64_t OFF_LEAFROW    = align256(OFF_GRIDPTR + NODE_CAP * 8);
constexpr uint64_t OFF_VTAIL      = align256(OFF_LEAFROW + NODE_CAP * 4);
constexpr uint64_t OFF_VDIR       = align256(OFF_VTAIL + NODE_CAP * 8);
constexpr uint64_t OFF_DIRTYLEAF  = align256(OFF_VDIR + NODE_CAP * 8);
constexpr uint64_t OFF_DIRTYVOX   = align256(OFF_DIRTYLEAF + NODE_CAP * 4);
constexpr uint64_t OFF_SPILLINFO  = align256(OFF_DIRTYVOX + NODE_CAP * 4);
constexpr uint64_t OFF_BLOCKCUR   = align256(OFF_SPILLINFO + SPILLNODE_CAP * 32);
constexpr uint64_t OFF_ROWFREE    = align256(OFF_BLOCKCUR + BLOCK_CAP * 4);
constexpr uint64_t OFF_ROWS       = align256(OFF_ROWFREE + ROW_CAP * 4);
constexpr uint64_t OFF_CHUNKDIR   = align256(OFF_ROWS + ROW_CAP * ROW_SLOTS * 8);
constexpr uint64_t OFF_QUEUE      = align256(OFF_CHUNKDIR + DIR_CAP * 8);
constexpr uint64_t OFF_LEAFOF     = align256(OFF_QUEUE + QUEUE_CAP * 8);
constexpr uint64_t OFF_SLOTOF     = align256(OFF_LEAFOF + ITEM_CAP * 4);
constexpr uint64_t OFF_SPILLED    = align256(OFF_SLOTOF + ITEM_CAP * 4);
constexpr uint64_t OFF_VKEY       = align256(OFF_SPILLED + SPILL_CAP * 16);
constexpr uint64_t OFF_VCOLOR     = align256(OFF_VKEY + VOXEL_CAP * 8);
constexpr uint64_t TOTAL          = align256(OFF_VCOLOR + VOXEL_CAP * 4);
static_assert(TOTAL <= 300000000ull, "scratch must fit the host's 300 MB momentary buffer (main.cpp:554)");
}  // namespace scratch

enum : uint32_t {   // Ctl::errorFlags, mirrored into Stats::dbg
    ERR_SPILL_OVERFLOW  = 1u << 0,   // more than SPILL_CAP spilled points in one batch (reference-undefined regime)
    ERR_VOXEL_OVERFLOW  = 1u << 1,   // more than VOXEL_CAP voxels created in one batch
    ERR_DIR_OVERFLOW    = 1u << 2,
    ERR_NODE_OVERFLOW   = 1u << 3,   // nodes[] capacity exceeded
    ERR_QUEUE_OVERFLOW  = 1u << 4,
    ERR_SPILLNODE_OVERFLOW = 1u << 5,
    ERR_ROW_OVERFLOW    = 1u << 6,   // more than ROW_CAP non-empty leaves, or a leaf with more than 64 chunks
};

struct BatchCounters {              // one set per batch parity: batch b uses set b & 1, the other one is cleared meanwhile
    uint32_t numSpillTotal;        // spilling nodes found so far in this batch (monotonic)
    uint32_t numSpilled;           // spilled points in this batch
    uint32_t numBacklog;           // voxels of this batch that went to the shared overflow part of the backlog
    uint32_t numDirtyLeaves;
    uint32_t numDirtyVox;
    uint32_t dirCursor;
    uint32_t _pad[2];
};

struct Ctl {
    uint32_t numBatchesUploaded;   // snapshot of the volatile host-updated counter (voxels.cu:872-876)
    uint32_t errorFlags;
    uint64_t elapsedNanos;
    uint64_t memUsed;              // heap offset snapshot for the capacity guard
    uint32_t rowBump;              // leaf rows handed out so far (persistent across launches)
    uint32_t rowFreeCount;         // entries on the row free stack (persistent)
    uint32_t statCounters[8];      // @32
    uint64_t voxelsByPass[2];      // @64 voxels created in first-visit passes / in re-walk passes since the last reset
    uint64_t spilledTotal;         // @80 spilled (re-inserted) points since the last reset: the `s` of the roofline's 32*s bytes
    uint64_t voxelsTotal;          // @88 voxels created since the last reset (incl. leaf-root voxels)
    BatchCounters batch[2];        // @96
    uint64_t phaseNanos[8];        // @160 time per phase since reset, by the grid's first thread (%globaltimer):
                                   //      0 count+sample, 1 split round, 2 re-walk, 3 deferred sampling, 4 allocate, 5 insert, 6 stats, 7 launch prologue
};
static_assert(offsetof(Ctl, spilledTotal) == 80, "bench.py reads Ctl::spilledTotal at byte 80");

// what the lane that sees a leaf cross 50 000 records about it (everything the split round needs)
struct SpillInfo {
    uint32_t node;
    uint32_t stored;       // points the leaf held before this batch
    uint32_t base;         // where they go in the spill buffer
    uint32_t row;          // the leaf's chunk row (+1)
    uint32_t childBase;    // index of child 0
    uint32_t level;
    uint64_t grid;         // occupancy grid of the new inner node
};
static_assert(sizeof(SpillInfo) == 32, "SpillInfo");

struct DirEntry { uint32_t base; uint32_t k0; };   // chunkDir[base + (slot/1000 - k0)] holds element `slot`

struct Ctx {
    Node*      nodes;
    Stats*     stats;
    Heap*      heap;
    uint8_t*   heapBytes;
    Ctl*       ctl;
    uint32_t*  firstChild;    // node -> index of child 0 (children are 8 consecutive nodes); 0 = leaf
    uint64_t*  gridPtr;       // node -> OccupancyGrid* (0 = none)
    BatchCounters* bc;        // counters of the batch in flight
    uint32_t*  leafRow;       // leaf -> row of its chunk pointers (+1; 0 = leaf holds no chunk)
    uint64_t*  rows;          // [ROW_CAP][64] chunk pointers of leaves, in list order
    uint32_t*  rowFree;       // stack of recycled rows
    uint64_t*  voxelTail;     // node -> last Chunk* of voxel list  (valid iff node.voxelChunks != 0)
    DirEntry*  voxelDir;
    uint32_t*  dirtyLeaves;
    uint32_t*  dirtyVox;
    SpillInfo* spill;
    uint32_t*  blockCursor;   // per block: entries of its own backlog segment filled in this batch
    uint32_t   segCap;        // entries per block segment
    uint64_t*  chunkDir;
    uint64_t*  chunkQueue;
    uint32_t*  leafOf;        // item -> leaf node | level << 24
    uint32_t*  slotOf;        // item -> index inside the leaf
    Point*     spilled;
    uint64_t*  vkey;          // cell | node << 21 | slot << 41
    uint32_t*  vcolor;
    float minx, miny, minz, size, rcpSize;
};

// ------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ldv(const uint32_t* p) { return *(volatile const uint32_t*)p; }
__device__ __forceinline__ uint64_t ldv(const uint64_t* p) { return *(volatile const uint64_t*)p; }
__device__ __forceinline__ uint32_t laneId() { return threadIdx.x & 31; }
__device__ __forceinline__ uint32_t lanemaskLt() { uint32_t m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }
__device__ __forceinline__ uint64_t globaltimer() { uint64_t t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

__device__ __forceinline__ uint4 ldPoint(const Point* p) {     // streaming 128-bit load
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void stPoint(Point* p, uint4 v) {
    asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

struct Coords { uint32_t X, Y, Z, pX, pY, pZ; };

// voxels.cu:148-155 — X = u32(2^20 * (p - min) / size), pX = u32(2^28 * (p - min) / size)
__device__ __forceinline__ Coords quantize(const Ctx& c, uint4 pt) {
    float dx = fpx::add(__uint_as_float(pt.x), -c.minx);
    float dy = fpx::add(__uint_as_float(pt.y), -c.miny);
    float dz = fpx::add(__uint_as_float(pt.z), -c.minz);
    Coords q;
    q.X  = fpx::f2u(fpx::mul_ftz(fpx::mul(dx, 1048576.0f), c.rcpSize));
    q.Y  = fpx::f2u(fpx::mul_ftz(fpx::mul(dy, 1048576.0f), c.rcpSize));
    q.Z  = fpx::f2u(fpx::mul_ftz(fpx::mul(dz, 1048576.0f), c.rcpSize));
    q.pX = fpx::f2u(fpx::mul_ftz(fpx::mul(dx, 268435456.0f), c.rcpSize));
    q.pY = fpx::f2u(fpx::mul_ftz(fpx::mul(dy, 268435456.0f), c.rcpSize));
    q.pZ = fpx::f2u(fpx::mul_ftz(fpx::mul(dz, 268435456.0f), c.rcpSize));
    return q;
}
// voxels.cu:171-179
__device__ __forceinline__ uint32_t childIndexAt(const Coords& q, uint32_t level) {
    uint32_t sh = SIMLOD_MAX_DEPTH - 1 - level;
    return (((q.X >> sh) & 1u) << 2) | (((q.Y >> sh) & 1u) << 1) | ((q.Z >> sh) & 1u);
}
// voxels.cu:78-88
__device__ __forceinline__ uint32_t cellAt(const Coords& q, uint32_t level) {
    uint32_t sh = SIMLOD_MAX_DEPTH + 1 - level;
    uint32_t cx = (q.pX >> sh) & 127u, cy = (q.pY >> sh) & 127u, cz = (q.pZ >> sh) & 127u;
    return cx | (cy << 7) | (cz << 14);
}

template <typename T>
__device__ __forceinline__ T* carve(uint32_t* buffer, uint64_t off) { return reinterpret_cast<T*>(reinterpret_cast<uint8_t*>(buffer) + off); }

// ------------------------------------------------------------------------------------------
// block-local voxel bookkeeping. A created voxel needs (a) a slot in its node's voxel list =
// numVoxels++ and (b) a backlog entry. For a coherent scan these are two very hot global
// addresses (the upper nodes' counters, the backlog cursor); same-address atomics serialise in L2
// and were 2/3 of the counting pass. Instead every block owns a segment of the backlog and a
// 64-entry node -> count table in shared memory; winners take a block-local rank, and when the
// block has finished its pass it adds each node's count to numVoxels ONCE and patches the entries
// it wrote with the returned base.
// ------------------------------------------------------------------------------------------
#ifndef SIMLOD_VOXTAB_SIZE
#define SIMLOD_VOXTAB_SIZE 64          // tuning knob (tools/exp_variants.py): power of two, <= 256
#endif
constexpr uint32_t VOXTAB_SIZE = SIMLOD_VOXTAB_SIZE;
static_assert((VOXTAB_SIZE & (VOXTAB_SIZE - 1)) == 0 && VOXTAB_SIZE <= 256, "table size must be a power of two that one block can sweep");
constexpr uint32_t VOXTAB_EMPTY = 0xffffffffu;
__shared__ uint32_t sh_tabKey[VOXTAB_SIZE];
__shared__ uint32_t sh_tabCount[VOXTAB_SIZE];
__shared__ uint32_t sh_tabBase[VOXTAB_SIZE];
__shared__ uint32_t sh_cursor;        // next free entry of this block's backlog segment
__shared__ uint32_t sh_passStart;     // first entry written in the current pass

__device__ __forceinline__ void voxelPassBegin(const Ctx& c, bool firstPassOfBatch) {
    if (threadIdx.x < VOXTAB_SIZE) { sh_tabKey[threadIdx.x] = VOXTAB_EMPTY; sh_tabCount[threadIdx.x] = 0; }
    if (threadIdx.x == 0) { sh_cursor = firstPassOfBatch ? 0u : c.blockCursor[blockIdx.x]; sh_passStart = sh_cursor; }
    __syncthreads();
}

// shared fall-back (block table or block segment full): global atomics, final key at once
__device__ __noinline__ void recordVoxelShared(const Ctx& c, uint32_t node, uint32_t cell, uint32_t color) {
    uint32_t vslot = atomicAdd(&c.nodes[node].numVoxels, 1u);
    if (vslot == ldv(&c.nodes[node].numVoxelsStored)) { uint32_t d = atomicAdd(&c.bc->numDirtyVox, 1u); c.dirtyVox[d] = node; }
    uint32_t b = atomicAdd(&c.bc->numBacklog, 1u);
    if (b < scratch::VOXEL_SHARED) {
        uint64_t at = scratch::VOXEL_CAP - scratch::VOXEL_SHARED + b;
        c.vkey[at] = (uint64_t)cell | ((uint64_t)node << 21) | ((uint64_t)vslot << 41);
        c.vcolor[at] = color;
    } else {
        atomicOr(&c.ctl->errorFlags, ERR_VOXEL_OVERFLOW);
    }
}

__device__ __forceinline__ void recordVoxel(const Ctx& c, uint32_t node, uint32_t cell, uint32_t color) {
    uint32_t h = (node * 0x9E3779B1u) >> 26;
    uint32_t slot = VOXTAB_EMPTY;
    for (uint32_t probe = 0; probe < VOXTAB_SIZE; probe++) {
        uint32_t s = (h + probe) & (VOXTAB_SIZE - 1);
        uint32_t k = atomicCAS(&sh_tabKey[s], VOXTAB_EMPTY, node);
        if (k == VOXTAB_EMPTY || k == node) { slot = s; break; }
    }
    if (slot == VOXTAB_EMPTY) { recordVoxelShared(c, node, cell, color); return; }
    uint32_t idx = atomicAdd(&sh_cursor, 1u);
    if (idx >= c.segCap) { recordVoxelShared(c, node, cell, color); return; }
    uint32_t rank = atomicAdd(&sh_tabCount[slot], 1u);
    uint64_t at = (uint64_t)blockIdx.x * c.segCap + idx;
    c.vkey[at] = (uint64_t)cell | ((uint64_t)slot << 21) | ((uint64_t)rank << 41);      // node/slot patched in voxelPassEnd
    c.vcolor[at] = color;
}

__device__ __forceinline__ void voxelPassEnd(const Ctx& c, bool freshPass) {
    __syncthreads();
    if (threadIdx.x < VOXTAB_SIZE) {
        uint32_t node = sh_tabKey[threadIdx.x], cnt = sh_tabCount[threadIdx.x];
        if (node != VOXTAB_EMPTY && cnt > 0) {
            uint32_t base = atomicAdd(&c.nodes[node].numVoxels, cnt);
            if (base == ldv(&c.nodes[node].numVoxelsStored)) {           // first voxels of this node in this batch
                uint32_t d = atomicAdd(&c.bc->numDirtyVox, 1u);
                c.dirtyVox[d] = node;
            }
            sh_tabBase[threadIdx.x] = base;
        }
    }
    __syncthreads();
    const uint32_t endIdx = min(sh_cursor, c.segCap);
    for (uint32_t e = sh_passStart + threadIdx.x; e < endIdx; e += blockDim.x) {
        uint64_t at = (uint64_t)blockIdx.x * c.segCap + e;
        uint64_t k = c.vkey[at];
        uint32_t slot = (uint32_t)(k >> 21) & (VOXTAB_SIZE - 1);
        uint32_t rank = (uint32_t)(k >> 41);
        c.vkey[at] = (k & 0x1fffffull) | ((uint64_t)sh_tabKey[slot] << 21) | ((uint64_t)(sh_tabBase[slot] + rank) << 41);
    }
    if (threadIdx.x == 0) {
        c.blockCursor[blockIdx.x] = endIdx;
        if (endIdx > sh_passStart) {
            atomicAdd(reinterpret_cast<unsigned long long*>(&c.ctl->voxelsTotal), (unsigned long long)(endIdx - sh_passStart));
            atomicAdd(reinterpret_cast<unsigned long long*>(&c.ctl->voxelsByPass[freshPass ? 0 : 1]), (unsigned long long)(endIdx - sh_passStart));
        }
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------
// block-local leaf counting. A scan is spatially coherent, so at any moment most of the grid
// counts into the same handful of leaves, and one atomicAdd(counter) per warp per iteration
// (what the reference does, voxels.cu:203-218) makes those counters the hottest addresses of the
// pass — with the returned value on every warp's critical path. Here a warp takes a block-local
// rank from a shared-memory table instead; when the block has finished its pass it adds each
// leaf's total to the global counter once (where the spill / first-touch detection now happens)
// and turns the provisional ranks of its items into slots.
// ------------------------------------------------------------------------------------------
constexpr uint32_t PROVISIONAL = 0x80000000u;
__shared__ uint32_t sh_leafKey[VOXTAB_SIZE];
__shared__ uint32_t sh_leafCount[VOXTAB_SIZE];
__shared__ uint32_t sh_leafBase[VOXTAB_SIZE];

// the global step: add `cnt` points to a leaf's counter; first-touch and spill detection (voxels.cu:203-218)
__device__ __noinline__ uint32_t countGlobal(const Ctx& c, uint32_t node, uint32_t level, uint32_t cnt) {
    Node* leaf = &c.nodes[node];
    uint32_t old = atomicAdd(&leaf->counter, cnt);
    uint32_t stored = ldv(&leaf->numPoints);
    if (old == stored) {                                     // first points of this leaf in this batch
        uint32_t d = atomicAdd(&c.bc->numDirtyLeaves, 1u);
        c.dirtyLeaves[d] = node;
    }
    if (old <= SIMLOD_MAX_POINTS_PER_NODE && old + cnt > SIMLOD_MAX_POINTS_PER_NODE) {
        // this leaf spills (voxels.cu:211-217). Reserve everything its split needs right here, so the
        // split round is one phase: room in the spill buffer, 8 node slots, the occupancy grid.
        uint32_t s = atomicAdd(&c.bc->numSpillTotal, 1u);
        if (s < scratch::SPILLNODE_CAP) {
            SpillInfo info;
            info.node = node;
            info.stored = stored;
            info.level = level;
            info.row = c.leafRow[node];
            info.base = stored ? atomicAdd(&c.bc->numSpilled, stored) : 0u;
            if ((uint64_t)info.base + stored > scratch::SPILL_CAP) atomicOr(&c.ctl->errorFlags, ERR_SPILL_OVERFLOW);
            info.childBase = atomicAdd(&c.stats->numNodes, 8u);                                // voxels.cu:317
            if (info.childBase + 8 > scratch::NODE_CAP) atomicOr(&c.ctl->errorFlags, ERR_NODE_OVERFLOW);
            uint64_t g = c.gridPtr[node];
            if (g == 0) g = (uint64_t)(c.heapBytes + atomicAdd(reinterpret_cast<unsigned long long*>(&c.heap->offset), (unsigned long long)SIMLOD_GRID_STRIDE));   // voxels.cu:363-365
            info.grid = g;
            c.spill[s] = info;
        } else {
            atomicOr(&c.ctl->errorFlags, ERR_SPILLNODE_OVERFLOW);
        }
    }
    return old;
}

__device__ __forceinline__ uint32_t tabFind(const uint32_t* keys, uint32_t key) {
    uint32_t h = (key * 0x9E3779B1u) >> 26;
    for (uint32_t probe = 0; probe < VOXTAB_SIZE; probe++) {
        uint32_t s = (h + probe) & (VOXTAB_SIZE - 1);
        uint32_t k = keys[s];
        if (k == key) return s;
        if (k == VOXTAB_EMPTY) break;
    }
    return VOXTAB_EMPTY;
}
__device__ __forceinline__ uint32_t tabInsert(uint32_t* keys, uint32_t key) {
    uint32_t h = (key * 0x9E3779B1u) >> 26;
    for (uint32_t probe = 0; probe < VOXTAB_SIZE; probe++) {
        uint32_t s = (h + probe) & (VOXTAB_SIZE - 1);
        uint32_t k = atomicCAS(&keys[s], VOXTAB_EMPTY, key);
        if (k == VOXTAB_EMPTY || k == key) return s;
    }
    return VOXTAB_EMPTY;
}

// ------------------------------------------------------------------------------------------
// the per-point walk: descend from (node, level) to the leaf; optionally voxel-sample every
// node on the way that owns an occupancy grid; optionally count the point into the leaf.
// Warp-synchronous: all 32 lanes call it together, `valid` masks lanes without an item.
//   count : voxels.cu:145-220 (doCounting::countPoint)     sample: voxels.cu:426-470 + 50-121
// ------------------------------------------------------------------------------------------
template <bool SAMPLE, bool COUNT>
__device__ __forceinline__ void walk(const Ctx& c, bool valid, uint4 pt, uint32_t node, uint32_t level,
                                     uint32_t& leafPacked, uint32_t& slot) {
    const uint32_t FULL = 0xffffffffu;
    const uint32_t lane = laneId();
    const uint32_t ltmask = lanemaskLt();
    Coords q = quantize(c, pt);
    // per-lane descent (no warp-level synchronisation inside: voxel bookkeeping is block-local).
    // atomicOr results are not needed to continue the descent, so up to 3 of them stay in flight
    // per lane and are only looked at when the leaf has been reached (or a 4th one is issued).
    uint32_t pending = 0;
    uint32_t old0 = 0, old1 = 0, old2 = 0, key0 = 0, key1 = 0, key2 = 0;      // key = node | (cell & 31) << 20 ... see below
    uint32_t cel0 = 0, cel1 = 0, cel2 = 0;
    auto settle = [&]() {
        if (pending > 0 && (old0 & (1u << (cel0 & 31u))) == 0) recordVoxel(c, key0, cel0, pt.w);
        if (pending > 1 && (old1 & (1u << (cel1 & 31u))) == 0) recordVoxel(c, key1, cel1, pt.w);
        if (pending > 2 && (old2 & (1u << (cel2 & 31u))) == 0) recordVoxel(c, key2, cel2, pt.w);
        pending = 0;
    };
    if (valid) {
        for (;;) {
            if (level >= SIMLOD_MAX_DEPTH) break;                       // voxels.cu:169 loop bound: a level-20 node is the leaf
            if (SAMPLE) {
                uint64_t g = c.gridPtr[node];
                if (g != 0) {
                    uint32_t cell = cellAt(q, level);
                    uint32_t* word = reinterpret_cast<uint32_t*>(g) + (cell >> 5);
                    uint32_t bit = 1u << (cell & 31u);
                    // non-atomic pre-test (voxels.cu:93-94): a set bit seen through the (non-coherent) L1 is final
                    uint32_t seen = *word;
                    if ((seen & bit) == 0) {
                        // neighbouring points hit the same cell: one atomic per distinct cell among the converged lanes
                        uint32_t active = __activemask();
                        uint32_t peers = __match_any_sync(active, (uint64_t)(uintptr_t)word * 32ull + (cell & 31u));
                        if (lane == (uint32_t)__ffs(peers) - 1u) {
                            if (pending == 3) settle();
                            uint32_t old = atomicOr(word, bit);
                            if (pending == 0) { old0 = old; key0 = node; cel0 = cell; }
                            else if (pending == 1) { old1 = old; key1 = node; cel1 = cell; }
                            else { old2 = old; key2 = node; cel2 = cell; }
                            pending++;
                        }
                    }
                }
            }
            uint32_t fc = c.firstChild[node];
            if (fc == 0) break;
            node = fc + childIndexAt(q, level);
            level++;
        }
        if (SAMPLE) settle();
    }
    __syncwarp();

    leafPacked = node | (level << 24);
    if (COUNT) {
        uint32_t vmask = __ballot_sync(FULL, valid);
        if (valid) {
            uint32_t peers = __match_any_sync(vmask, node);
            uint32_t leader = __ffs(peers) - 1;
            uint32_t cnt = __popc(peers);
            uint32_t r = 0;
            if (lane == leader) {
                uint32_t t = tabInsert(sh_leafKey, node);
                if (t != VOXTAB_EMPTY) r = atomicAdd(&sh_leafCount[t], cnt) | PROVISIONAL;     // block-local rank
                else                   r = countGlobal(c, node, level, cnt);                    // table full: final slot at once
            }
            r = __shfl_sync(peers, r, leader);
            slot = r + __popc(peers & ltmask);
        }
    }
}

// ------------------------------------------------------------------------------------------
// TMA staging of the batch. In a first-visit pass every block streams its contiguous run of the
// batch through two 8 KB shared-memory stages with 1-D bulk copies (cp.async.bulk ... mbarrier::
// complete_tx, SASS: UBLKCP): one elected thread issues the copy of the next 512-point tile while
// the block walks the current one, so the HBM latency of the batch read leaves the critical path
// and no registers or LSU slots are spent on it. Threads then read their point with one LDS.128.
// ------------------------------------------------------------------------------------------
#ifndef SIMLOD_NO_TMA
#define SIMLOD_TMA 1
#else
#define SIMLOD_TMA 0
#endif
#ifndef SIMLOD_TILE_POINTS
#define SIMLOD_TILE_POINTS 512         // tuning knob (tools/exp_variants.py): a multiple of 256
#endif
constexpr uint32_t TILE_POINTS = SIMLOD_TILE_POINTS;
static_assert(TILE_POINTS % 256 == 0 && TILE_POINTS >= 256, "a tile is walked in 256-point steps");
#if SIMLOD_TMA
__shared__ __align__(128) uint4 sh_tile[2][TILE_POINTS];
__shared__ __align__(8) uint64_t sh_tileBar[2];
#if defined(SIMLOD_DYNAMIC_TILES)
constexpr uint32_t MY_TILES_CAP = 64;
__shared__ uint32_t sh_tileIdx[2];
__shared__ uint32_t sh_myTiles[MY_TILES_CAP];
__shared__ uint32_t sh_numMyTiles;
#endif

__device__ __forceinline__ void tileBarInit() {
    if (threadIdx.x == 0) {
        uint32_t b0 = (uint32_t)__cvta_generic_to_shared(&sh_tileBar[0]), b1 = (uint32_t)__cvta_generic_to_shared(&sh_tileBar[1]);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(b0) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(b1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
}
__device__ __forceinline__ void tileLoad(uint32_t stage, const Point* src, uint32_t numPoints) {
    uint32_t bar = (uint32_t)__cvta_generic_to_shared(&sh_tileBar[stage]);
    uint32_t dst = (uint32_t)__cvta_generic_to_shared(&sh_tile[stage][0]);
    uint32_t bytes = numPoints * 16u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tileWait(uint32_t stage, uint32_t parity) {
    uint32_t bar = (uint32_t)__cvta_generic_to_shared(&sh_tileBar[stage]);
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
#endif

// one pass over batch points (ring slot) followed by the spilled points of this batch
//   FRESH  : items start at the root (first visit); otherwise only items whose cached leaf has
//            been split since are walked on, starting at that (now inner) node
template <bool SAMPLE, bool COUNT, bool FRESH>
__device__ void itemPass(const Ctx& c, const Point* batch, uint32_t numBatch, uint32_t numSpilled) {
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (COUNT && threadIdx.x < VOXTAB_SIZE) { sh_leafKey[threadIdx.x] = VOXTAB_EMPTY; sh_leafCount[threadIdx.x] = 0; }
    if (SAMPLE) voxelPassBegin(c, FRESH); else __syncthreads();
    // batch points: every block owns one contiguous run of the batch (scans are spatially coherent, so
    // all iterations of a block revisit the same upper-level nodes and occupancy words: L1 hits)
    // (re-walk passes touch a few contiguous runs of items, so they stay grid-strided to spread those runs)
    const uint32_t perBlock = FRESH ? (((numBatch + gridDim.x - 1) / gridDim.x + 31u) & ~31u) : numBatch;
    const uint32_t blockFirst = FRESH ? blockIdx.x * perBlock : blockIdx.x * blockDim.x;
    const uint32_t blockEnd = FRESH ? min(numBatch, blockFirst + perBlock) : numBatch;
    const uint32_t step = FRESH ? blockDim.x : stride;
#if SIMLOD_TMA
#if defined(SIMLOD_DYNAMIC_TILES)
    // EXPERIMENTAL variant (tools/exp_variants.py, not the shipped configuration): tiles are handed out from a global
    // cursor instead of fixed runs, so a block that draws cheap tiles takes more of them. The block remembers its
    // tiles for the rank -> slot pass below.
    if (FRESH) {
        const uint32_t totalTiles = (numBatch + TILE_POINTS - 1) / TILE_POINTS;
        uint32_t* cursor = &c.bc->_pad[(SAMPLE && !COUNT) ? 1 : 0];
        tileBarInit();
        if (threadIdx.x == 0) {
            sh_numMyTiles = 0;
            const uint32_t t0 = atomicAdd(cursor, 1u);
            sh_tileIdx[0] = t0;
            if (t0 < totalTiles) tileLoad(0, batch + t0 * TILE_POINTS, min(TILE_POINTS, numBatch - t0 * TILE_POINTS));
        }
        for (uint32_t n = 0;; n++) {
            __syncthreads();                               // sh_tileIdx[n & 1] published; stage (n+1)&1 drained
            const uint32_t cur = sh_tileIdx[n & 1];
            if (cur >= totalTiles) break;                  // block-uniform
            if (threadIdx.x == 0) {
                const uint32_t nxt = atomicAdd(cursor, 1u);
                sh_tileIdx[(n + 1) & 1] = nxt;
                if (nxt < totalTiles) tileLoad((n + 1) & 1, batch + nxt * TILE_POINTS, min(TILE_POINTS, numBatch - nxt * TILE_POINTS));
                if (sh_numMyTiles < MY_TILES_CAP) sh_myTiles[sh_numMyTiles] = cur; else atomicOr(&c.ctl->errorFlags, ERR_DIR_OVERFLOW);
                sh_numMyTiles++;
            }
            tileWait(n & 1, (n >> 1) & 1);
            const uint32_t tileFirst = cur * TILE_POINTS, tileEnd = min(numBatch, tileFirst + TILE_POINTS);
#pragma unroll 1
            for (uint32_t k = 0; k < TILE_POINTS / 256; k++) {
                const uint32_t idx = k * 256 + threadIdx.x;
                const uint32_t i = tileFirst + idx;
                const bool valid = i < tileEnd;
                uint4 pt = valid ? sh_tile[n & 1][idx] : make_uint4(0, 0, 0, 0);
                uint32_t lp = 0, slot = 0;
                walk<SAMPLE, COUNT>(c, valid, pt, 0, 0, lp, slot);
                if (valid && COUNT) { c.leafOf[i] = lp; c.slotOf[i] = slot; }
            }
        }
    } else
#else
    if (FRESH) {
        const uint32_t runLen = blockFirst < blockEnd ? blockEnd - blockFirst : 0u;
        const uint32_t numTiles = (runLen + TILE_POINTS - 1) / TILE_POINTS;
        tileBarInit();
        if (threadIdx.x == 0 && numTiles > 0) tileLoad(0, batch + blockFirst, min(TILE_POINTS, runLen));
        for (uint32_t t = 0; t < numTiles; t++) {
            const uint32_t tileFirst = blockFirst + t * TILE_POINTS;
            if (threadIdx.x == 0 && t + 1 < numTiles)      // stage (t+1)&1 was drained at the barrier that ended iteration t-1
                tileLoad((t + 1) & 1, batch + tileFirst + TILE_POINTS, min(TILE_POINTS, blockEnd - (tileFirst + TILE_POINTS)));
            tileWait(t & 1, (t >> 1) & 1);
#pragma unroll 1
            for (uint32_t k = 0; k < TILE_POINTS / 256; k++) {
                const uint32_t idx = k * 256 + threadIdx.x;
                const uint32_t i = tileFirst + idx;
                const bool valid = i < blockEnd;
                uint4 pt = valid ? sh_tile[t & 1][idx] : make_uint4(0, 0, 0, 0);
                uint32_t lp = 0, slot = 0;
                walk<SAMPLE, COUNT>(c, valid, pt, 0, 0, lp, slot);
                if (valid && COUNT) { c.leafOf[i] = lp; c.slotOf[i] = slot; }
            }
            __syncthreads();
        }
    } else
#endif
#endif
    for (uint32_t base = blockFirst + (threadIdx.x - laneId()); base < blockEnd; base += step) {
        uint32_t i = base + laneId();
        bool valid = i < blockEnd;
        uint32_t node = 0, level = 0;
        if (!FRESH && valid) {
            uint32_t lp = c.leafOf[i];
            node = lp & 0xffffffu; level = lp >> 24;
            valid = c.firstChild[node] != 0 && level < SIMLOD_MAX_DEPTH;
        }
        if (!FRESH && !__any_sync(0xffffffffu, valid)) continue;
        uint4 pt = make_uint4(0, 0, 0, 0);
        if (valid) pt = ldPoint(batch + i);
        uint32_t lp = 0, slot = 0;
        walk<SAMPLE, COUNT>(c, valid, pt, node, level, lp, slot);
        if (valid && COUNT) { c.leafOf[i] = lp; c.slotOf[i] = slot; }
    }
    // spilled points (always carry a cached start node: the leaf they were spilled from)
    for (uint32_t base = tid - laneId(); base < numSpilled; base += stride) {
        uint32_t j = base + laneId();
        bool valid = j < numSpilled;
        uint32_t node = 0, level = 0;
        if (valid && !(FRESH && !COUNT)) {       // sampling-only fresh pass restarts at the root
            uint32_t lp = c.leafOf[scratch::MAX_BATCH + j];
            node = lp & 0xffffffu; level = lp >> 24;
            valid = c.firstChild[node] != 0 && level < SIMLOD_MAX_DEPTH;
        }
        uint4 pt = make_uint4(0, 0, 0, 0);
        if (valid) pt = *reinterpret_cast<const uint4*>(c.spilled + j);
        uint32_t lp = 0, slot = 0;
        walk<SAMPLE, COUNT>(c, valid, pt, node, level, lp, slot);
        if (valid && COUNT) { c.leafOf[scratch::MAX_BATCH + j] = lp; c.slotOf[scratch::MAX_BATCH + j] = slot; }
    }
    if (SAMPLE) voxelPassEnd(c, FRESH);
    if (COUNT) {
        // flush the block's leaf table: one global add per distinct leaf, then provisional ranks -> slots
        __syncthreads();
        if (threadIdx.x < VOXTAB_SIZE) {
            uint32_t leaf = sh_leafKey[threadIdx.x], cnt = sh_leafCount[threadIdx.x];
            if (leaf != VOXTAB_EMPTY && cnt > 0) sh_leafBase[threadIdx.x] = countGlobal(c, leaf, c.nodes[leaf].level, cnt);
        }
        __syncthreads();
#if SIMLOD_TMA && defined(SIMLOD_DYNAMIC_TILES)
        if (FRESH) {
            const uint32_t mine = min(sh_numMyTiles, MY_TILES_CAP);
            for (uint32_t m = 0; m < mine; m++) {
                const uint32_t tileFirst = sh_myTiles[m] * TILE_POINTS, tileEnd = min(numBatch, tileFirst + TILE_POINTS);
                for (uint32_t i = tileFirst + threadIdx.x; i < tileEnd; i += blockDim.x) {
                    uint32_t sl = c.slotOf[i];
                    if (sl & PROVISIONAL) c.slotOf[i] = sh_leafBase[tabFind(sh_leafKey, c.leafOf[i] & 0xffffffu)] + (sl & ~PROVISIONAL);
                }
            }
        } else
#endif
        for (uint32_t base = blockFirst + (threadIdx.x - laneId()); base < blockEnd; base += step) {
            uint32_t i = base + laneId();
            if (i < blockEnd) {
                uint32_t sl = c.slotOf[i];
                if (sl & PROVISIONAL) c.slotOf[i] = sh_leafBase[tabFind(sh_leafKey, c.leafOf[i] & 0xffffffu)] + (sl & ~PROVISIONAL);
            }
        }
        for (uint32_t j = tid; j < numSpilled; j += stride) {
            uint32_t sl = c.slotOf[scratch::MAX_BATCH + j];
            if (sl & PROVISIONAL) c.slotOf[scratch::MAX_BATCH + j] = sh_leafBase[tabFind(sh_leafKey, c.leafOf[scratch::MAX_BATCH + j] & 0xffffffu)] + (sl & ~PROVISIONAL);
        }
    }
}

// ------------------------------------------------------------------------------------------
// split round, ONE phase (voxels.cu:245-289 spill copy, :308-383 doSplitting). Everything a split
// needs was reserved by the lane that detected it, so for every spilling leaf the three jobs are
// independent and spread over the whole grid, one warp per item:
//   parts 0..63  copy chunk k of the leaf's stored points into the spill buffer (16 KB, 128-bit)
//   parts 64..79 clear 1/16 of the new inner node's occupancy grid (sic: the root's populated grid too)
//   part  80     create the 8 children, return the chunks to the free stack, publish the node as inner
// ------------------------------------------------------------------------------------------
constexpr uint32_t SPLIT_PARTS = 81;

__device__ void splitRound(const Ctx& c, uint32_t begin, uint32_t end) {
    const uint32_t numWarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = laneId();
    const uint32_t numItems = (end - begin) * SPLIT_PARTS;
    for (uint32_t item = warp; item < numItems; item += numWarps) {
        const SpillInfo info = c.spill[begin + item / SPLIT_PARTS];
        const uint32_t part = item % SPLIT_PARTS;
        const uint32_t numChunks = (info.stored + SIMLOD_POINTS_PER_CHUNK - 1) / SIMLOD_POINTS_PER_CHUNK;
        if (info.childBase + 8 > scratch::NODE_CAP) continue;
        if (part < 64) {
            if (part >= numChunks || info.row == 0 || (uint64_t)info.base + info.stored > scratch::SPILL_CAP) continue;
            const Chunk* chunk = reinterpret_cast<const Chunk*>(c.rows[(uint64_t)(info.row - 1) * scratch::ROW_SLOTS + part]);
            const uint32_t first = part * SIMLOD_POINTS_PER_CHUNK;
            const uint32_t n = min((uint32_t)SIMLOD_POINTS_PER_CHUNK, info.stored - first);
            const uint32_t tag = info.node | (info.level << 24);
            for (uint32_t i = lane; i < n; i += 32) {
                uint4 v = *reinterpret_cast<const uint4*>(&chunk->points[i]);
                *reinterpret_cast<uint4*>(c.spilled + info.base + first + i) = v;
                c.leafOf[scratch::MAX_BATCH + info.base + first + i] = tag;
            }
        } else if (part < 80) {
            uint4* g = reinterpret_cast<uint4*>(info.grid) + (uint64_t)(part - 64) * (SIMLOD_GRID_WORDS / 4 / 16);
            for (uint32_t i = lane; i < SIMLOD_GRID_WORDS / 4 / 16; i += 32) g[i] = make_uint4(0, 0, 0, 0);
        } else {
            Node* parent = &c.nodes[info.node];
            const uint32_t pX = parent->X, pY = parent->Y, pZ = parent->Z;
            if (lane < 8) {
                // default-constructed Node + the fields doSplitting sets (voxels.cu:324-342)
                Node* child = &c.nodes[info.childBase + lane];
                uint64_t* raw = reinterpret_cast<uint64_t*>(child);
#pragma unroll
                for (int w = 0; w < 19; w++) raw[w] = 0;
                child->level = info.level + 1;
                child->X = 2 * pX + ((lane >> 2) & 1);
                child->Y = 2 * pY + ((lane >> 1) & 1);
                child->Z = 2 * pZ + (lane & 1);
                for (int b = 0; b < 20; b++) child->name[b] = parent->name[b];
                reinterpret_cast<uint8_t*>(child)[offsetof(Node, name) + info.level + 1] = (uint8_t)('0' + lane);   // name[level] (sic: level 20 lands on `visible`)
                child->isLeaf = 1;
                parent->children[lane] = child;
                c.firstChild[info.childBase + lane] = 0;
                c.gridPtr[info.childBase + lane] = 0;
                c.leafRow[info.childBase + lane] = 0;
            }
            // return the leaf's chunks to the free stack (voxels.cu:345-357)
            if (numChunks > 0 && info.row != 0) {
                uint64_t a0 = 0;
                if (lane == 0) a0 = atomicAdd(reinterpret_cast<unsigned long long*>(&c.stats->numAllocatedChunks), (unsigned long long)(0ull - numChunks));
                a0 = __shfl_sync(0xffffffffu, a0, 0);
                for (uint32_t k = lane; k < numChunks && k < scratch::ROW_SLOTS; k += 32) {
                    Chunk* chunk = reinterpret_cast<Chunk*>(c.rows[(uint64_t)(info.row - 1) * scratch::ROW_SLOTS + k]);
                    chunk->next = nullptr;
                    uint64_t qi = a0 - 1 - k;
                    if (qi < scratch::QUEUE_CAP) c.chunkQueue[qi] = (uint64_t)chunk;
                    else atomicOr(&c.ctl->errorFlags, ERR_QUEUE_OVERFLOW);
                }
            }
            __syncwarp();
            if (lane == 0) {
                if (info.row != 0) {                    // the row goes back to the row pool; its contents stay readable for this phase
                    uint32_t f = atomicAdd(&c.ctl->rowFreeCount, 1u);
                    c.rowFree[f] = info.row;
                    c.leafRow[info.node] = 0;
                }
                parent->numPoints = 0;
                parent->points = nullptr;
                parent->grid = reinterpret_cast<SimlodOccupancyGrid*>(info.grid);
                c.gridPtr[info.node] = info.grid;
                c.firstChild[info.node] = info.childBase;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// chunk allocation for the nodes touched by this batch (voxels.cu:485-538, 641-672)
// ------------------------------------------------------------------------------------------
// block-wide exclusive prefix sum of one value per thread (256 threads); returns the block total
__device__ __forceinline__ uint32_t blockExclusiveScan(uint32_t v, uint32_t& total) {
    __shared__ uint32_t sh_warpSum[8];
    __shared__ uint32_t sh_total;
    const uint32_t lane = laneId(), warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += t; }
    __syncthreads();                       // protects sh_* against the previous call
    if (lane == 31) sh_warpSum[warp] = incl;
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t run = 0; for (int w = 0; w < 8; w++) { uint32_t t = sh_warpSum[w]; sh_warpSum[w] = run; run += t; } sh_total = run; }
    __syncthreads();
    total = sh_total;
    return sh_warpSum[warp] + incl - v;
}

// The reference lets every node thread bump numAllocatedChunks / the heap offset once per chunk
// (voxels.cu:505-511). All dirty nodes of a batch sit in one or two blocks here, so a block adds its
// whole demand with ONE atomic per counter and hands out sub-ranges by prefix sum: the same totals,
// the same pooled-vs-fresh split (indices >= chunkPoolSize are fresh), a handful of atomics.
__device__ void allocateChunks(const Ctx& c, uint64_t poolSize) {
    __shared__ uint64_t sh_a0, sh_freshOff, sh_firstFresh;
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t numDirtyLeaves = ldv(&c.bc->numDirtyLeaves);
    const uint32_t numDirtyVox = ldv(&c.bc->numDirtyVox);

    for (uint32_t first = blockIdx.x * blockDim.x; first < numDirtyLeaves; first += stride) {      // block-uniform trip count
        const uint32_t d = first + threadIdx.x;
        uint32_t n = 0, cnt = 0, have = 0, existing = 0, needed = 0;
        Node* node = nullptr;
        bool live = false;
        if (d < numDirtyLeaves) {
            n = c.dirtyLeaves[d];
            if (c.firstChild[n] == 0) {                       // else: became an inner node in this batch
                node = &c.nodes[n];
                cnt = node->counter; have = node->numPoints;
                if (cnt > have) {
                    live = true;
                    existing = (have + SIMLOD_POINTS_PER_CHUNK - 1) / SIMLOD_POINTS_PER_CHUNK;
                    uint32_t required = (cnt + SIMLOD_POINTS_PER_CHUNK - 1) / SIMLOD_POINTS_PER_CHUNK;
                    needed = required - existing;
                    if (required > scratch::ROW_SLOTS) { atomicOr(&c.ctl->errorFlags, ERR_ROW_OVERFLOW); needed = 0; live = false; }
                }
            }
        }
        uint32_t total = 0;
        const uint32_t offset = blockExclusiveScan(needed, total);
        if (threadIdx.x == 0 && total > 0) {
            uint64_t a0 = atomicAdd(reinterpret_cast<unsigned long long*>(&c.stats->numAllocatedChunks), (unsigned long long)total);
            uint64_t firstFresh = a0 > poolSize ? a0 : poolSize;          // indices >= poolSize are new heap chunks (voxels.cu:509-515)
            uint64_t numFresh = a0 + total > firstFresh ? a0 + total - firstFresh : 0;
            sh_a0 = a0; sh_firstFresh = firstFresh;
            sh_freshOff = numFresh ? atomicAdd(reinterpret_cast<unsigned long long*>(&c.heap->offset), (unsigned long long)(numFresh * SIMLOD_CHUNK_STRIDE)) : 0;
        }
        __syncthreads();
        if (live) {
            if (needed > 0) {
                uint32_t row = c.leafRow[n];
                if (row == 0) {                                 // first chunk of this leaf: take a row (recycled first)
                    uint32_t f = atomicSub(&c.ctl->rowFreeCount, 1u);
                    if (f >= 1 && f <= scratch::ROW_CAP) {
                        row = c.rowFree[f - 1];
                    } else {
                        atomicAdd(&c.ctl->rowFreeCount, 1u);
                        uint32_t r = atomicAdd(&c.ctl->rowBump, 1u);
                        if (r >= scratch::ROW_CAP) { atomicOr(&c.ctl->errorFlags, ERR_ROW_OVERFLOW); row = 0; }
                        else row = r + 1;
                    }
                    c.leafRow[n] = row;
                }
                if (row != 0) {
                    uint64_t* slots = c.rows + (uint64_t)(row - 1) * scratch::ROW_SLOTS;
                    Chunk* tail = existing ? reinterpret_cast<Chunk*>(slots[existing - 1]) : nullptr;
                    const uint64_t a0 = sh_a0 + offset, firstFresh = sh_firstFresh, freshOff = sh_freshOff;
                    for (uint32_t t = 0; t < needed; t++) {
                        uint64_t idx = a0 + t;
                        Chunk* chunk = idx < poolSize ? reinterpret_cast<Chunk*>(c.chunkQueue[idx])
                                                      : reinterpret_cast<Chunk*>(c.heapBytes + freshOff + (idx - firstFresh) * SIMLOD_CHUNK_STRIDE);
                        chunk->next = nullptr;
                        if (tail) tail->next = chunk; else node->points = chunk;
                        tail = chunk;
                        slots[existing + t] = (uint64_t)chunk;
                    }
                }
            }
            node->numPoints = cnt;      // slots [have, cnt) were handed out by the counting pass; filled by insertAll
        }
    }

    // (served from the other end of the grid, so that leaf and voxel-node allocation run side by side)
    for (uint32_t first = (gridDim.x - 1 - blockIdx.x) * blockDim.x; first < numDirtyVox; first += stride) {
        const uint32_t d = first + threadIdx.x;
        uint32_t n = 0, cnt = 0, have = 0, existing = 0, needed = 0, nseg = 0, k0 = 0;
        Node* node = nullptr;
        bool live = false;
        if (d < numDirtyVox) {
            n = c.dirtyVox[d];
            node = &c.nodes[n];
            cnt = node->numVoxels; have = node->numVoxelsStored;
            if (cnt > have) {
                live = true;
                k0 = have / SIMLOD_POINTS_PER_CHUNK;
                nseg = (cnt - 1) / SIMLOD_POINTS_PER_CHUNK - k0 + 1;
                existing = (have + SIMLOD_POINTS_PER_CHUNK - 1) / SIMLOD_POINTS_PER_CHUNK;
                needed = (cnt + SIMLOD_POINTS_PER_CHUNK - 1) / SIMLOD_POINTS_PER_CHUNK - existing;
            }
        }
        uint32_t totalNeeded = 0, totalSeg = 0;
        const uint32_t offNeeded = blockExclusiveScan(needed, totalNeeded);
        const uint32_t offSeg = blockExclusiveScan(nseg, totalSeg);
        if (threadIdx.x == 0) {
            // voxel chunks are never recycled: always fresh heap memory (voxels.cu:652-666)
            sh_freshOff = totalNeeded ? atomicAdd(reinterpret_cast<unsigned long long*>(&c.heap->offset), (unsigned long long)((uint64_t)totalNeeded * SIMLOD_CHUNK_STRIDE)) : 0;
            sh_a0 = totalSeg ? atomicAdd(&c.bc->dirCursor, totalSeg) : 0;
        }
        __syncthreads();
        if (live) {
            const uint32_t base = (uint32_t)sh_a0 + offSeg;
            if ((uint64_t)base + nseg > scratch::DIR_CAP) { atomicOr(&c.ctl->errorFlags, ERR_DIR_OVERFLOW); }
            else {
                c.voxelDir[n] = DirEntry{base, k0};
                Chunk* tail = node->voxelChunks ? reinterpret_cast<Chunk*>(c.voxelTail[n]) : nullptr;
                uint32_t j = 0;
                if (have % SIMLOD_POINTS_PER_CHUNK != 0) c.chunkDir[base + j++] = (uint64_t)tail;
                const uint64_t freshOff = sh_freshOff + (uint64_t)offNeeded * SIMLOD_CHUNK_STRIDE;
                for (uint32_t t = 0; t < needed; t++) {
                    Chunk* chunk = reinterpret_cast<Chunk*>(c.heapBytes + freshOff + (uint64_t)t * SIMLOD_CHUNK_STRIDE);
                    chunk->next = nullptr;
                    if (tail) tail->next = chunk; else node->voxelChunks = chunk;
                    tail = chunk;
                    c.chunkDir[base + j++] = (uint64_t)chunk;
                }
                if (needed > 0) c.voxelTail[n] = (uint64_t)tail;
                node->numVoxelsStored = cnt;
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// insertion: every point/voxel already owns (leaf, slot); the leaf's chunk row (points) or the
// per-batch chunk directory (voxels) turns that into an address with two cached lookups
// (voxels.cu:540-639 insertPoints, 674-698 insertVoxels walk slot/1000 list links instead)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ Point* pointSlotAddress(const Ctx& c, uint32_t row, uint32_t slot) {
    Chunk* chunk = reinterpret_cast<Chunk*>(c.rows[(uint64_t)(row - 1) * scratch::ROW_SLOTS + slot / SIMLOD_POINTS_PER_CHUNK]);
    return &chunk->points[slot % SIMLOD_POINTS_PER_CHUNK];
}

__device__ __forceinline__ void insertVoxel(const Ctx& c, uint64_t at) {
    uint64_t key = c.vkey[at];
    uint32_t cell = (uint32_t)(key & 0x1fffffu);
    uint32_t node = (uint32_t)((key >> 21) & 0xfffffu);
    uint32_t vslot = (uint32_t)(key >> 41);
    const Node* nd = &c.nodes[node];
    uint32_t level = nd->level, X = nd->X, Y = nd->Y, Z = nd->Z;
    // cell centre in world space (voxels.cu:103-114; instruction sequence: see fpmath.cuh)
    float nodeSize = fpx::mul_ftz(fpx::ex2(-fpx::u2f(level)), c.size);
    float vx = fpx::add(fpx::fma(nodeSize, fpx::u2f(X), c.minx),
                        fpx::mul_ftz(fpx::mul(nodeSize, fpx::add(fpx::u2f(cell & 127u), 0.5f)), 0.0078125f));
    float vy = fpx::add(fpx::fma(nodeSize, fpx::u2f(Y), c.miny),
                        fpx::mul_ftz(fpx::mul(nodeSize, fpx::add(fpx::u2f((cell >> 7) & 127u), 0.5f)), 0.0078125f));
    float vz = fpx::add(fpx::fma(nodeSize, fpx::u2f(Z), c.minz),
                        fpx::mul_ftz(fpx::mul(nodeSize, fpx::add(fpx::u2f((cell >> 14) & 127u), 0.5f)), 0.0078125f));
    uint4 v = make_uint4(__float_as_uint(vx), __float_as_uint(vy), __float_as_uint(vz), c.vcolor[at]);
    DirEntry d = c.voxelDir[node];
    Chunk* chunk = reinterpret_cast<Chunk*>(c.chunkDir[d.base + (vslot / SIMLOD_POINTS_PER_CHUNK - d.k0)]);
    stPoint(&chunk->points[vslot % SIMLOD_POINTS_PER_CHUNK], v);
}

__device__ void insertAll(const Ctx& c, const Point* batch, uint32_t numBatch, uint32_t numSpilled, uint32_t numSharedVoxels) {
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    // two independent items per iteration: the three dependent lookups (item -> leaf row -> chunk) of one
    // overlap with those of the other
    for (uint32_t i = tid; i < numBatch; i += 2 * stride) {
        const uint32_t i2 = i + stride;
        const bool has2 = i2 < numBatch;
        uint4 p1 = ldPoint(batch + i);
        uint4 p2 = has2 ? ldPoint(batch + i2) : make_uint4(0, 0, 0, 0);
        uint32_t n1 = c.leafOf[i] & 0xffffffu, s1 = c.slotOf[i];
        uint32_t n2 = has2 ? (c.leafOf[i2] & 0xffffffu) : 0u, s2 = has2 ? c.slotOf[i2] : 0u;
        uint32_t r1 = c.leafRow[n1], r2 = has2 ? c.leafRow[n2] : 0u;
        if (r1) stPoint(pointSlotAddress(c, r1, s1), p1);
        if (r2) stPoint(pointSlotAddress(c, r2, s2), p2);
    }
    for (uint32_t j = tid; j < numSpilled; j += stride) {
        uint4 pt = *reinterpret_cast<const uint4*>(c.spilled + j);
        uint32_t node = c.leafOf[scratch::MAX_BATCH + j] & 0xffffffu;
        uint32_t row = c.leafRow[node];
        if (row) stPoint(pointSlotAddress(c, row, c.slotOf[scratch::MAX_BATCH + j]), pt);
    }
    // voxels: the per-block backlog segments are uneven, so every block first builds the prefix sums of
    // the segment fills in shared memory and the whole grid then strides over the concatenation
    __shared__ uint32_t sh_segStart[1025];
    if (gridDim.x <= 1024) {
        uint32_t total = 0;
        for (uint32_t b0 = 0; b0 < gridDim.x; b0 += blockDim.x) {
            uint32_t b = b0 + threadIdx.x;
            uint32_t v = b < gridDim.x ? c.blockCursor[b] : 0u;
            uint32_t chunkTotal = 0;
            uint32_t off = blockExclusiveScan(v, chunkTotal);
            if (b < gridDim.x) sh_segStart[b] = total + off;
            total += chunkTotal;
        }
        if (threadIdx.x == 0) sh_segStart[gridDim.x] = total;
        __syncthreads();
        for (uint32_t v = tid; v < total; v += stride) {
            uint32_t lo = 0, hi = gridDim.x;                     // last segment with start <= v
            while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (sh_segStart[mid] <= v) lo = mid; else hi = mid; }
            insertVoxel(c, (uint64_t)lo * c.segCap + (v - sh_segStart[lo]));
        }
    } else {
        const uint32_t own = c.blockCursor[blockIdx.x];
        for (uint32_t e = threadIdx.x; e < own; e += blockDim.x) insertVoxel(c, (uint64_t)blockIdx.x * c.segCap + e);
    }
    for (uint32_t b = tid; b < numSharedVoxels; b += stride) insertVoxel(c, scratch::VOXEL_CAP - scratch::VOXEL_SHARED + b);
}

__device__ __forceinline__ void clearBatchCounters(BatchCounters* b) {
    b->numSpillTotal = 0; b->numSpilled = 0; b->numBacklog = 0; b->numDirtyLeaves = 0; b->numDirtyVox = 0; b->dirCursor = 0;
#if defined(SIMLOD_DYNAMIC_TILES)
    b->_pad[0] = 0; b->_pad[1] = 0;            // tile cursors of the experimental dynamic hand-out
#endif
}

// ------------------------------------------------------------------------------------------
// kernel_construct — voxels.cu:804-1010
// ------------------------------------------------------------------------------------------
extern "C" __global__ void __launch_bounds__(256, 4)
kernel_construct(const Uniforms uniforms, Point* points, uint32_t* buffer, uint8_t* buffer_persistent, Node* nodes,
                 Stats* stats, uint64_t* frameStartTimestamp, CudaPrint* cudaprint,
                 uint32_t* numBatchesUploaded_volatile, uint32_t* batchSizes) {
    cg::grid_group grid = cg::this_grid();
    const bool first = grid.thread_rank() == 0;
    const uint64_t tStart = globaltimer();

    Ctx c;
    c.nodes = nodes;
    c.stats = stats;
    c.heap = reinterpret_cast<Heap*>(buffer_persistent);
    c.heapBytes = buffer_persistent;
    c.ctl = carve<Ctl>(buffer, scratch::OFF_CTL);
    c.bc = &c.ctl->batch[0];
    c.firstChild = carve<uint32_t>(buffer, scratch::OFF_FIRSTCHILD);
    c.gridPtr = carve<uint64_t>(buffer, scratch::OFF_GRIDPTR);
    c.leafRow = carve<uint32_t>(buffer, scratch::OFF_LEAFROW);
    c.rows = carve<uint64_t>(buffer, scratch::OFF_ROWS);
    c.rowFree = carve<uint32_t>(buffer, scratch::OFF_ROWFREE);
    c.voxelTail = carve<uint64_t>(buffer, scratch::OFF_VTAIL);
    c.voxelDir = carve<DirEntry>(buffer, scratch::OFF_VDIR);
    c.dirtyLeaves = carve<uint32_t>(buffer, scratch::OFF_DIRTYLEAF);
    c.dirtyVox = carve<uint32_t>(buffer, scratch::OFF_DIRTYVOX);
    c.spill = carve<SpillInfo>(buffer, scratch::OFF_SPILLINFO);
    c.blockCursor = carve<uint32_t>(buffer, scratch::OFF_BLOCKCUR);
    c.segCap = gridDim.x <= scratch::BLOCK_CAP ? (uint32_t)((scratch::VOXEL_CAP - scratch::VOXEL_SHARED) / gridDim.x) : 0u;
    c.chunkDir = carve<uint64_t>(buffer, scratch::OFF_CHUNKDIR);
    c.chunkQueue = carve<uint64_t>(buffer, scratch::OFF_QUEUE);
    c.leafOf = carve<uint32_t>(buffer, scratch::OFF_LEAFOF);
    c.slotOf = carve<uint32_t>(buffer, scratch::OFF_SLOTOF);
    c.spilled = carve<Point>(buffer, scratch::OFF_SPILLED);
    c.vkey = carve<uint64_t>(buffer, scratch::OFF_VKEY);
    c.vcolor = carve<uint32_t>(buffer, scratch::OFF_VCOLOR);

    // octree cube = boxMin + max extent on every axis (voxels.cu:860-863)
    float sx = fpx::sub(uniforms.boxMax[0], uniforms.boxMin[0]);
    float sy = fpx::sub(uniforms.boxMax[1], uniforms.boxMin[1]);
    float sz = fpx::sub(uniforms.boxMax[2], uniforms.boxMin[2]);
    c.size = fmaxf(fmaxf(sx, sy), sz);
    c.rcpSize = fpx::rcp(c.size);
    c.minx = uniforms.boxMin[0]; c.miny = uniforms.boxMin[1]; c.minz = uniforms.boxMin[2];

    if (first) {
        *frameStartTimestamp = tStart;
        c.ctl->numBatchesUploaded = *(volatile uint32_t*)numBatchesUploaded_volatile;   // one snapshot for all threads
        c.ctl->errorFlags = 0;
        c.ctl->elapsedNanos = 0;
        c.ctl->memUsed = c.heap->offset;
        clearBatchCounters(&c.ctl->batch[0]);
        clearBatchCounters(&c.ctl->batch[1]);
        for (int i = 0; i < 8; i++) c.ctl->statCounters[i] = 0;
        if (stats->batchletIndex == 0) {       // fresh after the reset kernel: the tree is the root alone
            c.ctl->spilledTotal = 0; c.ctl->voxelsTotal = 0; c.ctl->voxelsByPass[0] = 0; c.ctl->voxelsByPass[1] = 0;
            for (int i = 0; i < 8; i++) c.ctl->phaseNanos[i] = 0;
            c.ctl->rowBump = 0; c.ctl->rowFreeCount = 0;
            c.firstChild[0] = 0;
            c.leafRow[0] = 0;
            c.gridPtr[0] = (uint64_t)nodes[0].grid;
        }
    }
    grid.sync();
    uint64_t tPhase = tStart;
#define PHASE_DONE(k) do { if (first) { uint64_t _t = globaltimer(); c.ctl->phaseNanos[k] += _t - tPhase; tPhase = _t; } } while (0)
    PHASE_DONE(7);

    const uint32_t numBatchesUploaded = ldv(&c.ctl->numBatchesUploaded);
    const uint32_t firstBatch = ldv(&stats->batchletIndex);
    const uint32_t numBatches = min(numBatchesUploaded - firstBatch, 20u);     // voxels.cu:883
    const uint32_t lastBatch = firstBatch + numBatches;

    for (uint32_t batchIndex = firstBatch; batchIndex < lastBatch; batchIndex++) {
        const uint32_t ringSlot = batchIndex % SIMLOD_BATCH_STREAM_SIZE;
        const uint32_t batchSize = min(ldv(&batchSizes[ringSlot]), (uint32_t)SIMLOD_MAX_BATCH_SIZE);
        const Point* batch = points + (uint64_t)ringSlot * SIMLOD_MAX_BATCH_SIZE;
        c.bc = &c.ctl->batch[batchIndex & 1];
        BatchCounters* other = &c.ctl->batch[(batchIndex + 1) & 1];

        // capacity guard (voxels.cu:896-912): stop consuming batches 200 MB before the heap is full
        const bool memCapacityReached = ldv(&c.ctl->memUsed) + 200000000ull >= uniforms.persistentBufferCapacity;
        if (first) stats->memCapacityReached = memCapacityReached ? 1 : 0;
        if (memCapacityReached) break;

        const bool deferSampling = ldv(&c.firstChild[0]) == 0;    // root still a leaf: see DESIGN.md §4 (root grid is wiped when it splits)
        const uint64_t poolSize = ldv(&stats->chunkPoolSize);

        // ---- pass 1: count (+ sample) every batch point ------------------------------------
        if (deferSampling) itemPass<false, true, true>(c, batch, batchSize, 0);
        else               itemPass<true, true, true>(c, batch, batchSize, 0);
        grid.sync();
        PHASE_DONE(0);

        // ---- split rounds (voxels.cu:385-415 expand): 2 barriers each ------------------------
        uint32_t spillBegin = 0;
        for (int round = 0; round < 20; round++) {
            const uint32_t spillEnd = min(ldv(&c.bc->numSpillTotal), (uint32_t)scratch::SPILLNODE_CAP);
            if (spillEnd == spillBegin) break;
            splitRound(c, spillBegin, spillEnd);
            grid.sync();
            PHASE_DONE(1);
            const uint32_t numSpilled = min(ldv(&c.bc->numSpilled), (uint32_t)scratch::SPILL_CAP);
            if (deferSampling) itemPass<false, true, false>(c, batch, batchSize, numSpilled);
            else               itemPass<true, true, false>(c, batch, batchSize, numSpilled);
            grid.sync();
            PHASE_DONE(2);
            spillBegin = spillEnd;
        }
        const uint32_t numSpilled = min(ldv(&c.bc->numSpilled), (uint32_t)scratch::SPILL_CAP);
        if (deferSampling) {
            // the root was a leaf when the batch started: sample along the final paths, as the
            // reference does after expand() (voxels.cu:738-742)
            itemPass<true, false, true>(c, batch, batchSize, numSpilled);
            grid.sync();
            PHASE_DONE(3);
        }

        // ---- chunk allocation for touched nodes --------------------------------------------
        allocateChunks(c, poolSize);
        if (first) clearBatchCounters(other);           // the next batch's counter set is idle during this phase
        grid.sync();
        PHASE_DONE(4);

        // ---- insertion + bookkeeping (voxels.cu:925-949) --------------------------------------
        const uint32_t numVoxels = min(ldv(&c.bc->numBacklog), (uint32_t)scratch::VOXEL_SHARED);     // shared overflow part only
        if (first) {
            uint64_t allocated = ldv(&stats->numAllocatedChunks);
            if (allocated > poolSize) stats->chunkPoolSize = allocated;        // voxels.cu:535-537
            stats->batchletIndex = batchIndex + 1;
            stats->numPointsProcessed += batchSize;
            c.ctl->spilledTotal += numSpilled; c.ctl->voxelsTotal += numVoxels;
            c.ctl->memUsed = ldv(&c.heap->offset);
            c.ctl->elapsedNanos = globaltimer() - tStart;
        }
        insertAll(c, batch, batchSize, numSpilled, numVoxels);
        grid.sync();
        PHASE_DONE(5);
        const float elapsedMs = float(ldv(&c.ctl->elapsedNanos)) / 1000000.0f;
        if (elapsedMs > 10.0f) break;          // MAX_PROCESSING_TIME (voxels.cu:22,940)
    }

    // ---- octree statistics (voxels.cu:958-1009) --------------------------------------------
    {
        const uint32_t numNodes = ldv(&stats->numNodes);
        const uint32_t stride = gridDim.x * blockDim.x;
        uint32_t inner = 0, leaves = 0, nonempty = 0, pts = 0, vox = 0, chP = 0, chV = 0;
        for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < numNodes; n += stride) {
            const Node* node = &nodes[n];
            if (c.firstChild[n] == 0) {
                uint32_t np = node->numPoints;
                leaves++; pts += np; chP += (np + SIMLOD_POINTS_PER_CHUNK - 1) / SIMLOD_POINTS_PER_CHUNK;
                if (np > 0) nonempty++;
            } else {
                uint32_t nv = node->numVoxels;
                inner++; vox += nv; chV += (nv + SIMLOD_POINTS_PER_CHUNK - 1) / SIMLOD_POINTS_PER_CHUNK;
            }
        }
        uint32_t vals[7] = {inner, leaves, nonempty, pts, vox, chP, chV};
#pragma unroll
        for (int k = 0; k < 7; k++) {
            uint32_t v = vals[k];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (laneId() == 0 && v) atomicAdd(&c.ctl->statCounters[k], v);
        }
    }
    grid.sync();
    PHASE_DONE(6);
    if (first) {
        stats->numInner = ldv(&c.ctl->statCounters[0]);
        stats->numLeaves = ldv(&c.ctl->statCounters[1]);
        stats->numNonemptyLeaves = ldv(&c.ctl->statCounters[2]);
        stats->numPoints = ldv(&c.ctl->statCounters[3]);
        stats->numVoxels = ldv(&c.ctl->statCounters[4]);
        stats->numChunksPoints = ldv(&c.ctl->statCounters[5]);
        stats->numChunksVoxels = ldv(&c.ctl->statCounters[6]);
        stats->allocatedBytes_momentary = scratch::TOTAL;
        stats->allocatedBytes_persistent = ldv(&c.heap->offset);
        stats->frameID = (uint32_t)uniforms.frameCounter;
        stats->dbg = ldv(&c.ctl->errorFlags);
    }
}
