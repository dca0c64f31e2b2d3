// common.cuh — shared device/host helpers for the sm_100a filter kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace swg {

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int64_t i64;

constexpr u32 NONE32 = 0xFFFFFFFFu;
constexpr u64 NONE64 = ~(u64)0;

struct CudaError {
    cudaError_t code;
    const char *file;
    int line;
};

#define SWG_CUDA(expr)                                                     \
    do {                                                                   \
        cudaError_t _e = (expr);                                           \
        if (_e != cudaSuccess) throw swg::CudaError{_e, __FILE__, __LINE__}; \
    } while (0)

static inline u32 cdiv(u64 a, u64 b) { return (u32)((a + b - 1) / b); }
static inline int bits_for(u64 max_value) { // number of bits needed to represent max_value (>=1)
    int b = 1;
    while (b < 64 && (max_value >> b)) b++;
    return b;
}

__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ u32 lanemask_lt() {
    u32 m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
__device__ __forceinline__ u32 ld_volatile_u32(const u32 *p) {
    u32 v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u32(u32 *p, u32 v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// launch counter (the bench's "gpu_launches" claim): every kernel launch goes through LAUNCH.
struct LaunchCounter {
    u64 n = 0;
};

} // namespace swg
