// common.cuh -- shared host/device helpers for libotgan.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/otgan.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libotgan is written for sm_100a (B200) only"
#endif

namespace otgan {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

// ---- thread-local error string + launch counter (the only mutable state in the library) -------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define OTGAN_REQUIRE(cond, ...)                                  \
    do {                                                          \
        if (!(cond)) {                                            \
            ::otgan::set_error(__VA_ARGS__);                      \
            return OTGAN_EINVAL;                                  \
        }                                                         \
    } while (0)

// call after every kernel launch: captures launch errors without synchronising
#define OTGAN_CHECK_LAUNCH(name)                                                          \
    do {                                                                                  \
        cudaError_t e__ = cudaGetLastError();                                             \
        if (e__ != cudaSuccess) {                                                         \
            ::otgan::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));  \
            return OTGAN_ECUDA;                                                           \
        }                                                                                 \
        ::otgan::count_launch();                                                          \
    } while (0)

#define OTGAN_CUDA(call)                                                                  \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            ::otgan::set_error("%s failed: %s", #call, cudaGetErrorString(e__));          \
            return OTGAN_ECUDA;                                                           \
        }                                                                                 \
    } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: set it once per (kernel, device).  `done` is the
// caller's static bit mask (one bit per device ordinal < 64; higher ordinals set the attribute on every launch).
#define OTGAN_SET_MAX_SMEM(kernel, bytes)                                                              \
    do {                                                                                               \
        static unsigned long long done__ = 0ull;                                                       \
        int dev__ = 0;                                                                                 \
        OTGAN_CUDA(cudaGetDevice(&dev__));                                                             \
        if (dev__ >= 64 || !((done__ >> dev__) & 1ull)) {                                              \
            OTGAN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
            if (dev__ < 64) done__ |= 1ull << dev__;                                                   \
        }                                                                                              \
    } while (0)
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- device helpers ------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// 16-byte async copy global->shared; bytes beyond src_bytes are zero-filled (src_bytes in {0,4,8,12,16})
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
    uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc, int src_bytes) {
    uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N));
}
#endif

}  // namespace otgan
