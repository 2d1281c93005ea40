// util.cu — two tiny helper kernels of the launch surface (not part of the reference's ABI).
#include <stdint.h>

// MUFU.RCP(x): lets a CPU restatement reproduce a / size for a cube size that is not a power of two
extern "C" __global__ void simlod_util_rcp(float x, float* out) {
    float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); *out = r;
}
// streaming fill used to evict L2 between timed runs
extern "C" __global__ void simlod_util_fill(uint4* dst, uint64_t count, uint32_t value) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x)
        dst[i] = make_uint4(value, value, value, value);
}
