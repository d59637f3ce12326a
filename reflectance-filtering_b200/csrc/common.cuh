// Shared host/device helpers for librf_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <mutex>
#include <cstdarg>
#include <cstdio>

#include "../../include/rf_b200.h"

namespace rf {

// ---- error reporting (thread-local message, C status codes) -----------------------------
char *err_buf();
int fail(int code, const char *fmt, ...);
extern std::atomic<unsigned long long> g_launches;

#define RF_CUDA_TRY(expr)                                                                    \
    do {                                                                                     \
        cudaError_t e__ = (expr);                                                            \
        if (e__ != cudaSuccess)                                                              \
            return rf::fail(RF_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                            __FILE__, __LINE__);                                             \
    } while (0)

// call right after a <<<>>> launch
#define RF_LAUNCH_CHECK(name)                                                                \
    do {                                                                                     \
        rf::g_launches.fetch_add(1, std::memory_order_relaxed);                              \
        cudaError_t e__ = cudaGetLastError();                                                \
        if (e__ != cudaSuccess)                                                              \
            return rf::fail(RF_ECUDA, "launch of %s failed: %s", name, cudaGetErrorString(e__)); \
    } while (0)

int sm_count();  // of the current device (cached per device)

// One-time per-device setup (cudaFuncSetAttribute ...) that several host threads may reach at once:
//   static DeviceOnce once;  std::lock_guard<std::mutex> lock(once.mu);  if (!once.done[dev & 63]) { ...; once.done[dev & 63] = true; }
struct DeviceOnce {
    std::mutex mu;
    bool done[64] = {};
};

// ---- borders (OpenCV borderInterpolate) ---------------------------------------------------
// REFLECT_101: gfedcb|abcdefgh|gfedcba      (joint bilateral)
__host__ __device__ __forceinline__ int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;
    return p;
}
// REFLECT: fedcba|abcdefgh|hgfedcb           (guided filter box means)
__host__ __device__ __forceinline__ int reflect(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p - 1 : 2 * len - 1 - p;
    return p;
}

// saturate_cast<uchar>(float): round-half-even, clamp
__device__ __forceinline__ uint8_t sat_u8(float v)
{
    int r = __float2int_rn(v);
    return (uint8_t)min(max(r, 0), 255);
}

// packed two-lane FP32 FMA (sm_100 FFMA2): d = a * b + c on both halves of a 64-bit register
__device__ __forceinline__ unsigned long long pack2(float lo, float hi)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ void ffma2(unsigned long long &acc, unsigned long long a, unsigned long long b)
{
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}

}  // namespace rf
