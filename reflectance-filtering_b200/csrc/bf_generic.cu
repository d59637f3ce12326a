// Joint bilateral filter, generic path: any radius, any OpenCV border type, 8-bit and CV_32F images.
//
// The rest of cv2.ximgproc.jointBilateralFilter's surface (the call site
// /root/reference/filter_reflectance.py:60-64 passes whatever ndarrays it is given): the borderType argument, CV_32F
// joint / src (jointBilateralFilter_32f: range weight interpolated in a 4096-bins-per-channel exp table scaled to the
// joint image's own value range, SURVEY A.2 last bullet), and radii beyond what the tiled kernels of bf.cu / bf2.cu
// hold in shared memory.  None of this is on the measured path, so the kernel is the plain form of the algorithm:
// one thread per output pixel walks the disc in raster order, reads joint and source through L1 and uses the SAME
// tables and the SAME separately rounded float operations, in the same order, as the CPU restatement
// (oracle/rf_oracle.c) -- results are bit-equal to it for 8-bit images and for CV_32F up to the last bit of the
// device's double-precision exp() in the table.
#include <cfloat>
#include <cmath>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace rf {
namespace bfg {

struct Args {
    const void *joint;
    const void *src;
    void *dst;
    const int *tap_ofs;    // [ntaps] (di << 16) | (dj & 0xffff), raster order over the disc
    const float *tap_w;    // [ntaps] spatial weights
    const float *lut;      // u8: [256 * lut_scale_channels] colour weights; f32: per image [nbins + 2], after the header
    const float *scale;    // f32: per image scale_index (0 for a constant joint image)
    int n, h, w, jc, sc, ntaps, border, nbins;
    int alpha_scale;       // u8: 3 for a gray plane that stands for three equal channels, else 1
};

// cv::borderInterpolate; -1 = outside with BORDER_CONSTANT (value 0)
__device__ __forceinline__ int border_index(int p, int len, int border)
{
    if (p >= 0 && p < len) return p;
    switch (border) {
        case 0: return -1;
        case 1: return p < 0 ? 0 : len - 1;
        case 2: return reflect(p, len);
        case 3: {
            const int q = p % len;
            return q < 0 ? q + len : q;
        }
        default: return reflect101(p, len);
    }
}

template <typename T, int JC, int SC>
__global__ void __launch_bounds__(256) bf_generic_kernel(const Args a)
{
    extern __shared__ float cw[];  // u8: the colour-weight table
    const int img = blockIdx.z;
    if (sizeof(T) == 1) {
        for (int i = threadIdx.x + threadIdx.y * blockDim.x; i < 256 * JC * a.alpha_scale; i += blockDim.x * blockDim.y)
            cw[i] = a.lut[i];
        __syncthreads();
    }
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= a.w || y >= a.h) return;
    const size_t npx = (size_t)a.h * a.w;
    const T *J = static_cast<const T *>(a.joint) + img * npx * JC;
    const T *S = static_cast<const T *>(a.src) + img * npx * SC;
    const float *lut = sizeof(T) == 1 ? cw : a.lut + (size_t)img * (a.nbins + 2);
    const float scale_index = sizeof(T) == 1 ? 0.0f : a.scale[img];
    T j0[JC];
#pragma unroll
    for (int c = 0; c < JC; ++c) j0[c] = J[((size_t)y * a.w + x) * JC + c];
    float wsum = 0.0f, s[SC];
#pragma unroll
    for (int c = 0; c < SC; ++c) s[c] = 0.0f;
    for (int k = 0; k < a.ntaps; ++k) {
        const int ofs = __ldg(a.tap_ofs + k);
        const int yy = border_index(y + (ofs >> 16), a.h, a.border);
        const int xx = border_index(x + (int)(short)(ofs & 0xffff), a.w, a.border);
        const bool inside = yy >= 0 && xx >= 0;
        const size_t q = inside ? (size_t)yy * a.w + xx : 0;
        float wt;
        if (sizeof(T) == 1) {
            int alpha = 0;
#pragma unroll
            for (int c = 0; c < JC; ++c) alpha += abs((int)j0[c] - (inside ? (int)J[q * JC + c] : 0));
            wt = __fmul_rn(__ldg(a.tap_w + k), lut[alpha * a.alpha_scale]);
        } else {
            float alpha = 0.0f;
#pragma unroll
            for (int c = 0; c < JC; ++c) alpha = __fadd_rn(alpha, fabsf(__fsub_rn((float)j0[c], inside ? (float)J[q * JC + c] : 0.0f)));
            alpha = __fmul_rn(alpha, scale_index);
            int idx = (int)alpha;
            if (idx > a.nbins) idx = a.nbins;  // a zero-padded border can exceed the image's own range
            alpha = __fsub_rn(alpha, (float)idx);
            const float l0 = __ldg(lut + idx), l1 = __ldg(lut + idx + 1);
            wt = __fmul_rn(__ldg(a.tap_w + k), __fadd_rn(l0, __fmul_rn(alpha, __fsub_rn(l1, l0))));
        }
#pragma unroll
        for (int c = 0; c < SC; ++c) s[c] = __fadd_rn(s[c], __fmul_rn(wt, inside ? (float)S[q * SC + c] : 0.0f));
        wsum = __fadd_rn(wsum, wt);
    }
    T *o = static_cast<T *>(a.dst) + (img * npx + (size_t)y * a.w + x) * SC;
#pragma unroll
    for (int c = 0; c < SC; ++c) {
        const float v = __fdiv_rn(s[c], wsum);
        if (sizeof(T) == 1)
            o[c] = (T)sat_u8(v);
        else
            o[c] = (T)v;
    }
}

// ---- CV_32F set-up: value range of every joint image, then its exp table --------------------------------------
// header per image: scale[img]; tables follow
__global__ void __launch_bounds__(256) range_kernel(const float *joint, size_t per_image, int jc, int nbins, float *scale)
{
    __shared__ float smn[8], smx[8];
    const float *J = joint + (size_t)blockIdx.x * per_image;
    float mn = FLT_MAX, mx = -FLT_MAX;
    for (size_t i = threadIdx.x; i < per_image; i += blockDim.x) {
        const float v = J[i];
        mn = fminf(mn, v);
        mx = fmaxf(mx, v);
    }
    for (int d = 16; d > 0; d >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, d));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    }
    if ((threadIdx.x & 31) == 0) {
        smn[threadIdx.x >> 5] = mn;
        smx[threadIdx.x >> 5] = mx;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) {
            mn = fminf(mn, smn[i]);
            mx = fmaxf(mx, smx[i]);
        }
        const bool flat = fabs((double)mx - (double)mn) < 1.1920928955078125e-7;
        const float color_range = (float)((double)mx - (double)mn) * jc;
        scale[blockIdx.x] = flat ? 0.0f : nbins / color_range;
    }
}

__global__ void exp_table_kernel(const float *scale, float *lut, int nbins, double gcc)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nbins + 2) return;
    const float sc = scale[blockIdx.y];
    const double val = sc == 0.0f ? 0.0 : i / (double)sc;
    lut[(size_t)blockIdx.y * (nbins + 2) + i] = (float)exp(val * val * gcc);
}

// ---- host -------------------------------------------------------------------------------------------------------
struct Taps {
    int device, r;
    double sigma_space;
    int ntaps;
    int *d_ofs;
    float *d_w;
};
struct ColorLut {
    int device, entries;
    double sigma_color;
    float *d_lut;
};
static std::mutex g_mu;
static std::vector<Taps> g_taps;
static std::vector<ColorLut> g_luts;

// tables of one (sigma_space, radius) / (sigma_color, entries); call with g_mu held until the kernel is launched
static int get_taps(double sigma_space, int r, Taps *out)
{
    int dev = 0;
    RF_CUDA_TRY(cudaGetDevice(&dev));
    for (const Taps &t : g_taps)
        if (t.device == dev && t.r == r && t.sigma_space == sigma_space) {
            *out = t;
            return RF_OK;
        }
    std::vector<int> ofs;
    std::vector<float> wt;
    const double gsc = -0.5 / (sigma_space * sigma_space);
    for (int i = -r; i <= r; ++i)
        for (int j = -r; j <= r; ++j) {
            const double rr = std::sqrt((double)i * i + (double)j * j);
            if (rr > r) continue;
            ofs.push_back((int)(((unsigned)i << 16) | ((unsigned)j & 0xffffu)));
            wt.push_back((float)std::exp(rr * rr * gsc));
        }
    Taps t{dev, r, sigma_space, (int)ofs.size(), nullptr, nullptr};
    RF_CUDA_TRY(cudaMalloc(&t.d_ofs, ofs.size() * sizeof(int)));
    RF_CUDA_TRY(cudaMalloc(&t.d_w, wt.size() * sizeof(float)));
    RF_CUDA_TRY(cudaMemcpy(t.d_ofs, ofs.data(), ofs.size() * sizeof(int), cudaMemcpyHostToDevice));
    RF_CUDA_TRY(cudaMemcpy(t.d_w, wt.data(), wt.size() * sizeof(float), cudaMemcpyHostToDevice));
    if (g_taps.size() >= 16) {
        cudaFree(g_taps.front().d_ofs);
        cudaFree(g_taps.front().d_w);
        g_taps.erase(g_taps.begin());
    }
    g_taps.push_back(t);
    *out = t;
    return RF_OK;
}

static int get_color_lut(double sigma_color, int entries, const float **out)
{
    int dev = 0;
    RF_CUDA_TRY(cudaGetDevice(&dev));
    for (const ColorLut &l : g_luts)
        if (l.device == dev && l.entries == entries && l.sigma_color == sigma_color) {
            *out = l.d_lut;
            return RF_OK;
        }
    std::vector<float> h(entries);
    const double gcc = -0.5 / (sigma_color * sigma_color);
    for (int i = 0; i < entries; ++i) h[i] = (float)std::exp((double)i * i * gcc);
    ColorLut l{dev, entries, sigma_color, nullptr};
    RF_CUDA_TRY(cudaMalloc(&l.d_lut, entries * sizeof(float)));
    RF_CUDA_TRY(cudaMemcpy(l.d_lut, h.data(), entries * sizeof(float), cudaMemcpyHostToDevice));
    if (g_luts.size() >= 16) {
        cudaFree(g_luts.front().d_lut);
        g_luts.erase(g_luts.begin());
    }
    g_luts.push_back(l);
    *out = l.d_lut;
    return RF_OK;
}

template <typename T>
static int launch(const Args &a, cudaStream_t st)
{
    const dim3 block(32, 8), grid((a.w + 31) / 32, (a.h + 7) / 8, a.n);
    const size_t smem = sizeof(T) == 1 ? (size_t)256 * a.jc * a.alpha_scale * sizeof(float) : 0;
#define RF_BFG(JC, SC)                                                       \
    if (a.jc == JC && a.sc == SC) {                                          \
        bf_generic_kernel<T, JC, SC><<<grid, block, smem, st>>>(a);          \
        RF_LAUNCH_CHECK("bf_generic_kernel");                                \
        return RF_OK;                                                        \
    }
    RF_BFG(1, 1)
    RF_BFG(1, 3)
    RF_BFG(3, 1)
    RF_BFG(3, 3)
#undef RF_BFG
    return fail(RF_EINVAL, "bf_generic: channels must be 1 or 3");
}

int radius_of(double sigma_space, int d)
{
    int r = d <= 0 ? (int)std::nearbyint(sigma_space * 1.5) : d / 2;  // cvRound: round-half-even
    return r < 1 ? 1 : r;
}

constexpr int MAX_RADIUS = 2048;  // (di, dj) are packed into 16 bits each; the tap table of r = 2048 is 100 MB
extern const int MAX_RADIUS_PUBLIC = MAX_RADIUS;

int run_u8(const uint8_t *joint, int jc, const uint8_t *src, int sc, uint8_t *dst, int n, int h, int w,
           double sigma_color, double sigma_space, int d, int alpha_scale, int border, cudaStream_t st)
{
    Args a{};
    a.joint = joint, a.src = src, a.dst = dst;
    a.n = n, a.h = h, a.w = w, a.jc = jc, a.sc = sc, a.border = border, a.alpha_scale = alpha_scale;
    const int r = radius_of(sigma_space, d);
    if (r > MAX_RADIUS) return fail(RF_EUNSUPPORTED, "joint bilateral: radius %d exceeds the supported maximum %d", r, MAX_RADIUS);
    std::lock_guard<std::mutex> lk(g_mu);
    Taps t;
    int rc = get_taps(sigma_space, r, &t);
    if (rc != RF_OK) return rc;
    rc = get_color_lut(sigma_color, 256 * jc * alpha_scale, &a.lut);
    if (rc != RF_OK) return rc;
    a.tap_ofs = t.d_ofs, a.tap_w = t.d_w, a.ntaps = t.ntaps;
    for (int i0 = 0; i0 < n; i0 += 65535) {  // gridDim.z limit
        Args b = a;
        b.n = n - i0 < 65535 ? n - i0 : 65535;
        b.joint = joint + (size_t)i0 * h * w * jc, b.src = src + (size_t)i0 * h * w * sc, b.dst = dst + (size_t)i0 * h * w * sc;
        rc = launch<uint8_t>(b, st);
        if (rc != RF_OK) return rc;
    }
    return RF_OK;
}

// per-image scale (padded to 4 floats) + per-image exp table
size_t workspace_f32(int n, int jc) { return ((size_t)((n + 3) & ~3) + (size_t)n * (4096 * jc + 2)) * sizeof(float); }

int run_f32(const float *joint, int jc, const float *src, int sc, float *dst, int n, int h, int w, double sigma_color,
            double sigma_space, int d, int border, void *ws, size_t ws_bytes, cudaStream_t st)
{
    if (n > 65535) return fail(RF_EUNSUPPORTED, "rf_joint_bilateral_f32: more than 65535 images per call");
    if (ws_bytes < workspace_f32(n, jc) || !ws)
        return fail(RF_EINVAL, "rf_joint_bilateral_f32: workspace of %zu bytes needed, %zu given", workspace_f32(n, jc), ws_bytes);
    Args a{};
    a.joint = joint, a.src = src, a.dst = dst;
    a.n = n, a.h = h, a.w = w, a.jc = jc, a.sc = sc, a.border = border, a.alpha_scale = 1;
    a.nbins = 4096 * jc;
    const int r = radius_of(sigma_space, d);
    if (r > MAX_RADIUS) return fail(RF_EUNSUPPORTED, "joint bilateral: radius %d exceeds the supported maximum %d", r, MAX_RADIUS);
    float *scale = static_cast<float *>(ws);
    float *lut = scale + ((n + 3) & ~3);
    a.scale = scale, a.lut = lut;
    range_kernel<<<n, 256, 0, st>>>(joint, (size_t)h * w * jc, jc, a.nbins, scale);
    RF_LAUNCH_CHECK("bfg::range_kernel");
    exp_table_kernel<<<dim3((a.nbins + 2 + 255) / 256, n), 256, 0, st>>>(scale, lut, a.nbins, -0.5 / (sigma_color * sigma_color));
    RF_LAUNCH_CHECK("bfg::exp_table_kernel");
    std::lock_guard<std::mutex> lk(g_mu);
    Taps t;
    const int rc = get_taps(sigma_space, r, &t);
    if (rc != RF_OK) return rc;
    a.tap_ofs = t.d_ofs, a.tap_w = t.d_w, a.ntaps = t.ntaps;
    return launch<float>(a, st);
}

}  // namespace bfg
}  // namespace rf
