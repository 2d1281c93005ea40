// gen.cu — synthetic point streams generated on the device (bench / test infrastructure of the
// launch surface; not part of the reference's ABI). The benchmark configurations are 36 M - 2 G
// points (BASELINE.json configs 2-5); producing them with numpy on the host takes
// minutes, here it takes milliseconds, and any sub-range can be produced independently because
// every generator is counter-based (splitmix64 of the point index).
//
// These kernels restate simlod_b200/data.py (uniform_cube, terrain, shell) operation by operation
// in IEEE double / float arithmetic. The file is compiled with --fmad=false, so terrain and
// uniform_cube are bit-identical to the numpy generators (tests/test_generators.py); shell uses the
// device's sin/cos/sqrt, which can differ from glibc's in the last ulp of the double before the
// narrowing to float (compared with a tolerance).
#include <stdint.h>

struct GenPoint { float x, y, z; uint32_t color; };

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// 24-bit uniform in [0, 1), exactly representable in float32 (data.py:_uniform24)
__device__ __forceinline__ float uniform24(uint64_t counter) { return (float)(uint32_t)(splitmix64(counter) >> 40) * 5.9604644775390625e-08f; }

// ---- config 1: uniform cube (data.py:uniform_cube) ------------------------------------------------------
extern "C" __global__ void simlod_gen_uniform(GenPoint* out, uint64_t first, uint64_t count, uint64_t seed, float size) {
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t c = (seed << 32) + (first + k) * 4ull;
        GenPoint p;
        p.x = uniform24(c) * size;
        p.y = uniform24(c + 1) * size;
        p.z = uniform24(c + 2) * size;
        p.color = (uint32_t)(splitmix64(c + 3) & 0xFFFFFFull) | 0xFF000000u;
        out[k] = p;
    }
}

// ---- configs 2/3/5: fBm terrain in flight strips (data.py:terrain) -----------------------------------------
__device__ __forceinline__ double hash2(int64_t ix, int64_t iy, uint64_t seed) {
    uint64_t k = ((uint64_t)ix * 0x9E3779B1ull) ^ ((uint64_t)iy * 0x85EBCA77ull) ^ seed;
    return (double)(splitmix64(k) >> 40) * 5.9604644775390625e-08;
}
__device__ __forceinline__ double valueNoise(double x, double y, uint64_t seed) {
    double fix = floor(x), fiy = floor(y);
    double fx = x - fix, fy = y - fiy;
    int64_t ix = (int64_t)fix, iy = (int64_t)fiy;
    double sx = fx * fx * (3.0 - 2.0 * fx), sy = fy * fy * (3.0 - 2.0 * fy);
    double v00 = hash2(ix, iy, seed), v10 = hash2(ix + 1, iy, seed);
    double v01 = hash2(ix, iy + 1, seed), v11 = hash2(ix + 1, iy + 1, seed);
    return (v00 * (1.0 - sx) + v10 * sx) * (1.0 - sy) + (v01 * (1.0 - sx) + v11 * sx) * sy;
}
__device__ __forceinline__ double terrainHeight(double x, double y, uint64_t seed) {
    double h = 0.0, amp = 1.0, freq = 1.0 / 1600.0, norm = 0.0;
#pragma unroll 1
    for (int octave = 0; octave < 5; octave++) {
        h += amp * valueNoise(x * freq, y * freq, seed + 101ull * (uint64_t)octave);
        norm += amp;
        amp *= 0.5;
        freq *= 2.0;
    }
    return (h / norm) * (300.0 - 1.0);
}
__device__ __forceinline__ double clampd(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }

extern "C" __global__ void simlod_gen_terrain(GenPoint* out, uint64_t nTotal, uint64_t first, uint64_t count, uint64_t seed) {
    const double EX = 4800.0, EY = 4300.0, EZ = 300.0, STRIP = 50.0;
    const uint64_t numStrips = 96;                                   // int(4800 / 50)
    const uint64_t perStrip = (nTotal + numStrips - 1) / numStrips;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = first + k;
        const uint64_t stripI = i / perStrip;
        const double strip = (double)stripI;
        const double t = (double)(i % perStrip) / (double)perStrip;
        const uint64_t c = (seed << 40) + i * 4ull;
        const double u0 = (double)uniform24(c), u1 = (double)uniform24(c + 1), u2 = (double)uniform24(c + 2);
        const double x = fmin((strip + u0) * STRIP, EX - 0.01);
        const double along = (stripI & 1ull) == 0 ? t : 1.0 - t;     // serpentine flight lines
        const double y = clampd(along * EY + (u1 - 0.5) * 4.0, 0.0, EY - 0.01);
        const double z = clampd(terrainHeight(x, y, seed) + (u2 - 0.5) * 0.4, 0.0, EZ - 0.01);
        const double hn = clampd(z / EZ, 0.0, 1.0);
        const uint32_t r = (uint32_t)(40.0 + 200.0 * hn);
        const uint32_t g = (uint32_t)(90.0 + 140.0 * (1.0 - fabs(hn - 0.5) * 2.0));
        const uint32_t b = (uint32_t)(60.0 + 120.0 * (1.0 - hn));
        GenPoint p;
        p.x = (float)x; p.y = (float)y; p.z = (float)z;
        p.color = r | (g << 8) | (b << 16) | 0xFF000000u;
        out[k] = p;
    }
}

// ---- config 4: sphere shell in latitude / longitude tile order (data.py:shell) -------------------------------
extern "C" __global__ void simlod_gen_shell(GenPoint* out, uint64_t nTotal, uint64_t first, uint64_t count, uint64_t seed) {
    const uint64_t tilesLat = 64, tilesLon = 128, numTiles = tilesLat * tilesLon;
    const uint64_t perTile = (nTotal + numTiles - 1) / numTiles;
    const double PI = 3.141592653589793, CTR = 2048.0;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = first + k;
        const uint64_t tile = i / perTile;
        const double tlat = (double)(tile / tilesLon), tlon = (double)(tile % tilesLon);
        const uint64_t c = (seed << 40) + i * 4ull;
        const double u0 = (double)uniform24(c), u1 = (double)uniform24(c + 1), u2 = (double)uniform24(c + 2);
        const uint32_t col = (uint32_t)(splitmix64(c + 3) & 0xFFFFFFull);
        const double cz = -1.0 + 2.0 * (tlat + u0) / (double)tilesLat;          // equal-area in z
        const double phi = 2.0 * PI * (tlon + u1) / (double)tilesLon;
        const double r = 1800.0 + (u2 - 0.5) * 0.5;
        const double s = sqrt(fmax(0.0, 1.0 - cz * cz));
        GenPoint p;
        p.x = (float)(CTR + r * s * cos(phi));
        p.y = (float)(CTR + r * s * sin(phi));
        p.z = (float)(CTR + r * cz);
        p.color = col | 0xFF000000u;
        out[k] = p;
    }
}
