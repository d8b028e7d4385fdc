// Shared helpers for libgsr_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/gsr_sm100.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libgsr_sm100 is written for sm_100a (B200) only"
#endif

namespace gsr {

extern thread_local char g_last_error[256];
extern std::atomic<uint64_t> g_launches;

inline int fail(int code, const char* msg) {
    snprintf(g_last_error, sizeof(g_last_error), "%s", msg);
    return code;
}

// Call after every launch: counts it and converts the launch status.
inline int launched(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_last_error, sizeof(g_last_error), "%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

constexpr int kNumSMs = 148;   // B200
constexpr int kMaxDynSmem = 232448;   // dynamic shared memory one CTA may opt in to (227 KB)

// csrc/eigh.cu: Jacobi pass over the graphs whose info is -2 (the fallback behind csrc/eigh_tridiag.cu)
int eigh_jacobi_redo_failed(const float* A, float* w, float* U, int32_t* info, float* scratch, long long stride,
                            int batch, int n, int canonical_sign, cudaStream_t st);

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float sgnf(float x) {       // torch.sgn: sgn(0) = 0, NaN -> NaN... (sign(NaN)=0 in torch; abs' grad uses sgn)
    return (float)((x > 0.f) - (x < 0.f));
}

// streaming (read-once) 128-bit load that does not pollute L1
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

}  // namespace gsr
