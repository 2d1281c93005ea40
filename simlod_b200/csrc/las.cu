// las.cu — LAS point-record decode on the GPU (the first "next" row after the hot path).
//
// The reference decodes LAS records on 2 x cores CPU threads before uploading 16-byte points
// (modules/progressive_octree/LasLoader.cpp:169-226 `loadLasNative`: int32 XYZ * scale + offset in
// double -> float, 16-bit RGB -> 8-bit), which caps its LAS path at 200-300 Mpoints/s (README.md:10).
// Here the raw records are uploaded as they are in the file and one kernel turns a batch into the
// ring slot's XYZRGBA points: every block stages 256 records (256 x bytesPerPoint bytes, always a
// multiple of 16) into shared memory with ONE 1-D TMA bulk copy, threads pick their record apart with
// 16-bit shared-memory loads (records are only 2-byte aligned), and the 16-byte results are stored
// fully coalesced. Arithmetic is the reference's, operation for operation, in IEEE double without
// contraction (the host compiler emits mul + add on x86-64): bit-identical xyz.
#include <stdint.h>
#include "../../include/simlod_abi.h"

constexpr uint32_t LAS_TILE = 256;          // records per block iteration
constexpr uint32_t LAS_MAX_BPP = 96;        // largest supported record (LAS 1.4 format 10 = 67 bytes + extra bytes)

// Records are 2-byte aligned when bytesPerPoint and the RGB offset are even (formats 0-3, 6-8 without odd extra bytes):
// 16-bit shared loads. Odd record sizes (format 5 = 63 bytes, formats 4 / 9 / 10, odd extra bytes) put every other
// record on an odd address: fields are assembled from byte loads there (the reference memcpy's, LasLoader.cpp:206-217).
template <bool EVEN> __device__ __forceinline__ uint32_t smemU16(const uint8_t* p) {
    if (EVEN) return *reinterpret_cast<const uint16_t*>(p);
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8);
}
template <bool EVEN> __device__ __forceinline__ int32_t smemI32(const uint8_t* p) { return (int32_t)(smemU16<EVEN>(p) | (smemU16<EVEN>(p + 2) << 16)); }

template <bool EVEN>
__device__ __forceinline__ uint4 decodeRecord(const uint8_t* r, uint32_t offsetRgb, double scaleX, double scaleY, double scaleZ, double offsetX, double offsetY, double offsetZ) {
    // LasLoader.cpp:206-210: point.x = double(XYZ[0]) * scale_x + offset_x  (double mul, double add, then to float)
    float x = __double2float_rn(__dadd_rn(__dmul_rn((double)smemI32<EVEN>(r + 0), scaleX), offsetX));
    float y = __double2float_rn(__dadd_rn(__dmul_rn((double)smemI32<EVEN>(r + 4), scaleY), offsetY));
    float z = __double2float_rn(__dadd_rn(__dmul_rn((double)smemI32<EVEN>(r + 8), scaleZ), offsetZ));
    uint32_t color = 0xff000000u;          // the reference leaves alpha (and, without RGB, the colour) uninitialised
    if (offsetRgb > 0) {
        // LasLoader.cpp:212-217: 16-bit channels above 255 are scaled down by 256
        uint32_t cr = smemU16<EVEN>(r + offsetRgb), cg = smemU16<EVEN>(r + offsetRgb + 2), cb = smemU16<EVEN>(r + offsetRgb + 4);
        cr = cr > 255 ? cr / 256 : cr; cg = cg > 255 ? cg / 256 : cg; cb = cb > 255 ? cb / 256 : cb;
        color |= cr | (cg << 8) | (cb << 16);
    }
    return make_uint4(__float_as_uint(x), __float_as_uint(y), __float_as_uint(z), color);
}

extern "C" __global__ void __launch_bounds__(256)
simlod_las_decode(const uint8_t* __restrict__ records, uint64_t numPoints, uint32_t bytesPerPoint, uint32_t offsetRgb,
                  double scaleX, double scaleY, double scaleZ, double offsetX, double offsetY, double offsetZ,
                  SimlodPoint* __restrict__ out) {
    __shared__ __align__(128) uint8_t sh_rec[LAS_TILE * LAS_MAX_BPP];
    __shared__ __align__(8) uint64_t sh_bar;
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&sh_bar);
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sh_rec);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint64_t numTiles = (numPoints + LAS_TILE - 1) / LAS_TILE;
    uint32_t parity = 0;
    for (uint64_t tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
        const uint64_t first = tile * LAS_TILE;
        const uint32_t n = (uint32_t)min((uint64_t)LAS_TILE, numPoints - first);
        const uint32_t bytes = n * bytesPerPoint;
        const uint8_t* src = records + first * bytesPerPoint;
        const bool bulk = (bytes % 16u) == 0 && (((uintptr_t)src) % 16u) == 0;      // full tiles of an aligned buffer always qualify
        if (bulk) {
            if (threadIdx.x == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
            }
            uint32_t done;
            do {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(bar), "r"(parity) : "memory");
            } while (!done);
            parity ^= 1;
        } else {                                  // ragged tail or odd alignment: plain byte copy
            for (uint32_t b = threadIdx.x; b < bytes; b += blockDim.x) sh_rec[b] = src[b];
            __syncthreads();
        }
        if (threadIdx.x < n) {
            const uint8_t* r = sh_rec + threadIdx.x * bytesPerPoint;
            const bool even = ((bytesPerPoint | offsetRgb) & 1u) == 0;
            uint4 v = even ? decodeRecord<true>(r, offsetRgb, scaleX, scaleY, scaleZ, offsetX, offsetY, offsetZ)
                           : decodeRecord<false>(r, offsetRgb, scaleX, scaleY, scaleZ, offsetX, offsetY, offsetZ);
            *reinterpret_cast<uint4*>(out + first + threadIdx.x) = v;
        }
        __syncthreads();                          // the stage is reused by the next tile
    }
}
