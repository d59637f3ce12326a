// Library plumbing (errors, device selection, launch counter) and the small layout / statistics
// kernels around the three hot kernels.
#include <cstring>

#include "common.cuh"

namespace rf {

std::atomic<unsigned long long> g_launches{0};

char *err_buf()
{
    static thread_local char buf[512] = "";
    return buf;
}

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}

int sm_count()
{
    static int cached[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    int &c = cached[dev & 63];
    if (c == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        c = v;
    }
    return c;
}

// gray [n] -> bgr [n][3]
__global__ void replicate_gray_kernel(const uint8_t *__restrict__ gray, uint8_t *__restrict__ bgr, size_t n)
{
    // each thread expands 4 gray pixels (one 32-bit load) into 12 bytes (three 32-bit stores)
    const size_t n4 = n / 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t g = reinterpret_cast<const uint32_t *>(gray)[i];
        const uint32_t a = g & 0xff, b = (g >> 8) & 0xff, c = (g >> 16) & 0xff, d = g >> 24;
        uint32_t *o = reinterpret_cast<uint32_t *>(bgr) + 3 * i;
        o[0] = a | (a << 8) | (a << 16) | (b << 24);
        o[1] = b | (b << 8) | (c << 16) | (c << 24);
        o[2] = c | (d << 8) | (d << 16) | (d << 24);
    }
    if (blockIdx.x == 0 && threadIdx.x < n % 4) {
        const size_t i = n4 * 4 + threadIdx.x;
        bgr[3 * i] = bgr[3 * i + 1] = bgr[3 * i + 2] = gray[i];
    }
}

__global__ void extract_gray_kernel(const uint8_t *__restrict__ bgr, uint8_t *__restrict__ gray, size_t n,
                                    int *all_equal)
{
    bool eq = true;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint8_t b = bgr[3 * i], g = bgr[3 * i + 1], r = bgr[3 * i + 2];
        gray[i] = b;
        eq &= (b == g) & (g == r);
    }
    if (all_equal && !__all_sync(0xffffffffu, eq) && (threadIdx.x & 31) == 0) *all_equal = 0;
}

__global__ void stats_kernel(const uint8_t *__restrict__ in, const uint8_t *__restrict__ out, size_t n, double *stats)
{
    double s1 = 0, s2 = 0, sd = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int o = out[i], a = in[i];
        s1 += o;
        s2 += o * o;
        sd += abs(o - a);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        s1 += __shfl_down_sync(0xffffffffu, s1, d);
        s2 += __shfl_down_sync(0xffffffffu, s2, d);
        sd += __shfl_down_sync(0xffffffffu, sd, d);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(stats + 1, s1);
        atomicAdd(stats + 2, s2);
        atomicAdd(stats + 3, sd);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(stats + 0, (double)n);
}

}  // namespace rf

using namespace rf;

extern "C" int rf_version(void) { return 100; }  // 0.1.0

extern "C" const char *rf_last_error(void) { return err_buf(); }

extern "C" unsigned long long rf_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int rf_set_device(int device)
{
    RF_CUDA_TRY(cudaSetDevice(device));
    return RF_OK;
}

extern "C" int rf_device_info(int *sm, int *cc_major, int *cc_minor, size_t *total_mem)
{
    int dev = 0;
    RF_CUDA_TRY(cudaGetDevice(&dev));
    cudaDeviceProp p;
    RF_CUDA_TRY(cudaGetDeviceProperties(&p, dev));
    if (sm) *sm = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (total_mem) *total_mem = p.totalGlobalMem;
    return RF_OK;
}

static unsigned grid_for(size_t n, int per_thread)
{
    size_t blocks = (n / per_thread + 255) / 256 + 1;
    const size_t cap = (size_t)sm_count() * 16;
    return (unsigned)(blocks < cap ? blocks : cap);
}

extern "C" int rf_replicate_gray_u8(const uint8_t *gray, uint8_t *bgr, size_t n_px, void *stream)
{
    if (!gray || !bgr) return fail(RF_EINVAL, "rf_replicate_gray_u8: NULL pointer");
    if (n_px == 0) return RF_OK;
    if (((uintptr_t)gray | (uintptr_t)bgr) % 4) return fail(RF_EINVAL, "rf_replicate_gray_u8: pointers must be 4-byte aligned");
    replicate_gray_kernel<<<grid_for(n_px, 4), 256, 0, (cudaStream_t)stream>>>(gray, bgr, n_px);
    RF_LAUNCH_CHECK("replicate_gray_kernel");
    return RF_OK;
}

extern "C" int rf_extract_gray_u8(const uint8_t *bgr, uint8_t *gray, size_t n_px, int *all_equal_flag, void *stream)
{
    if (!gray || !bgr) return fail(RF_EINVAL, "rf_extract_gray_u8: NULL pointer");
    if (n_px == 0) return RF_OK;
    extract_gray_kernel<<<grid_for(n_px, 1), 256, 0, (cudaStream_t)stream>>>(bgr, gray, n_px, all_equal_flag);
    RF_LAUNCH_CHECK("extract_gray_kernel");
    return RF_OK;
}

extern "C" int rf_accumulate_stats_u8(const uint8_t *in, const uint8_t *out, size_t n_bytes, double *stats4,
                                      void *stream)
{
    if (!in || !out || !stats4) return fail(RF_EINVAL, "rf_accumulate_stats_u8: NULL pointer");
    if (n_bytes == 0) return RF_OK;
    stats_kernel<<<grid_for(n_bytes, 8), 256, 0, (cudaStream_t)stream>>>(in, out, n_bytes, stats4);
    RF_LAUNCH_CHECK("stats_kernel");
    return RF_OK;
}
