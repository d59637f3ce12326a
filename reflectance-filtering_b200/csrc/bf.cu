// Joint bilateral filter, 8-bit, for sm_100a.
//
// Replaces cv2.ximgproc.jointBilateralFilter at /root/reference/filter_reflectance.py:60-64
// (semantics: OpenCV-contrib 3.1.0 jointBilateralFilter_8u).
//
// Design (DESIGN.md "K3"): the filter is ALU/SFU bound (3,409 taps per pixel at sigma_space 22,
// a handful of bytes per pixel), so the kernel is organised around issue slots, not bytes:
//   * a CTA stages the joint (and, when distinct, the src) window of its tile in shared memory
//     once, REFLECT_101 resolved at fill time, pixels packed to one 32-bit word (B|G<<8|R<<16);
//   * each thread owns P = 8 consecutive output pixels of one row and walks the disc row by row;
//     neighbours arrive four at a time with one conflict-free LDS.128 and are reused by all 8
//     outputs, so loads and byte->float conversions amortise to < 1 instruction per tap;
//   * the range distance is one VABSDIFF4.ACC (its accumulator operand is 0x4B000000, so the
//     result *is* the float 2^23 + alpha), both Gaussian weights are one MUFU.EX2 of
//     alpha^2 * kc + (e(dy) + e(dx)), the spatial exponents coming from a small per-row table whose
//     entries outside the disc are -inf (weight exactly 0, no per-tap predicate);
//   * the four accumulators (3 channels + weight sum) are two packed FFMA2.
// A second instantiation handles the BF(CNN, CNN) case, where joint and src are gray planes that
// stand for three equal channels: alpha = 3|dJ|, one channel to accumulate.
#include <cmath>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace rf {
namespace bf {

constexpr int P = 8;            // output pixels per thread (along x)
constexpr int TW = 32;          // tile width: 4 lanes x P
constexpr int ROWS_PER_WARP = 8;
constexpr int MAX_RADIUS = 64;  // shared-memory bound for the two-tile (joint != src) case

struct Args {
    const uint8_t *joint;
    const uint8_t *src;
    uint8_t *dst;
    const float *tab;  // device: (r+1) x tabw exponents, then (r+1) ints (row widths, ceil4)
    int jc, sc, dc;
    int n, h, w;
    int r, rpad, pitch, tabw;
    float ksqrt;  // sqrt(0.5 / sigma_color^2 * log2(e)) (* alpha_scale for replicated gray): the range
                  // weight is exp2(-(alpha * ksqrt)^2)
};

__host__ __device__ inline int rows_of(int wy, int r) { return ROWS_PER_WARP * wy + 2 * r; }

// ---- shared-memory window fill ---------------------------------------------------------------
template <int WY>
__device__ __forceinline__ void fill_packed(uint32_t *tile, const uint8_t *img, int cn, const Args &a,
                                            int tx0, int ty0)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rows = rows_of(WY, a.r), cols = TW + 2 * a.rpad;
    for (int yy = warp; yy < rows; yy += WY) {
        const int gy = reflect101(ty0 - a.r + yy, a.h);
        const uint8_t *row = img + (size_t)gy * a.w * cn;
        uint32_t *trow = tile + yy * a.pitch;
        for (int cc = lane; cc < cols; cc += 32) {
            const int gx = reflect101(tx0 - a.rpad + cc, a.w);
            const uint8_t *px = row + (size_t)gx * cn;
            // byte 3 is 1 in every packed pixel: it cancels in the SAD and converts to the 1.0f that
            // the weight-sum lane of the second FFMA2 multiplies
            uint32_t v = px[0] | 0x01000000u;
            if (cn == 3) v |= ((uint32_t)px[1] << 8) | ((uint32_t)px[2] << 16);
            trow[cc] = v;
        }
    }
}

template <int WY>
__device__ __forceinline__ void fill_float(float *tile, const uint8_t *img, const Args &a, int tx0, int ty0)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rows = rows_of(WY, a.r), cols = TW + 2 * a.rpad;
    for (int yy = warp; yy < rows; yy += WY) {
        const int gy = reflect101(ty0 - a.r + yy, a.h);
        const uint8_t *row = img + (size_t)gy * a.w;
        float *trow = tile + yy * a.pitch;
        for (int cc = lane; cc < cols; cc += 32) {
            const int gx = reflect101(tx0 - a.rpad + cc, a.w);
            trow[cc] = (float)row[gx];
        }
    }
}

__device__ __forceinline__ void load_table(float *tab_s, const Args &a)
{
    const int n = (a.r + 1) * a.tabw + (a.r + 1);
    for (int i = threadIdx.x; i < n; i += blockDim.x) tab_s[i] = a.tab[i];
}

__device__ __forceinline__ float byte_to_float(uint32_t word, int byte)
{
    // PRMT builds the float 2^23 + b from the byte, one FADD removes the bias: ALU + FMA pipes,
    // keeps the XU pipe for MUFU.EX2 (I2F.U8 is an XU instruction).
    const uint32_t sel = 0x7440u | (uint32_t)byte;  // bytes: {b, 0x00(4), 0x00(4), 0x4B(7)} of {word, magic}
    return __uint_as_float(__byte_perm(word, 0x4B000000u, sel)) - 8388608.0f;
}

// ---- colour kernel: packed 3-channel (or 1-channel) joint and src ------------------------------
template <int WY, bool SEP>
__global__ void __launch_bounds__(32 * WY) bf_color_kernel(const Args a)
{
    extern __shared__ __align__(16) uint32_t smem[];
    const int rows = rows_of(WY, a.r);
    uint32_t *tj = smem;
    uint32_t *ts = SEP ? tj + rows * a.pitch : tj;
    float *tab = reinterpret_cast<float *>(ts + rows * a.pitch);
    const int *roww4 = reinterpret_cast<const int *>(tab + (a.r + 1) * a.tabw);

    const int img = blockIdx.z;
    const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * (ROWS_PER_WARP * WY);
    const size_t npx = (size_t)a.h * a.w;
    fill_packed<WY>(tj, a.joint + img * npx * a.jc, a.jc, a, tx0, ty0);
    if (SEP) fill_packed<WY>(ts, a.src + img * npx * a.sc, a.sc, a, tx0, ty0);
    load_table(tab, a);
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x0 = (lane & 3) * P;
    const int ty = warp * ROWS_PER_WARP + (lane >> 2);

    unsigned long long acc01[P], acc23[P];
    uint32_t jc[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
        acc01[p] = 0ull;
        acc23[p] = 0ull;
        jc[p] = tj[(ty + a.r) * a.pitch + a.rpad + x0 + p];
    }
    const float sc = a.ksqrt, scb = -8388608.0f * a.ksqrt;
    const int r = a.r;
    for (int dyi = 0; dyi <= 2 * r; ++dyi) {
        const int ady = dyi < r ? r - dyi : dyi - r;
        const int w4 = roww4[ady];
        const uint32_t *jrow = tj + (ty + dyi) * a.pitch + a.rpad + x0;
        const uint32_t *srow = ts + (ty + dyi) * a.pitch + a.rpad + x0;
        const float *trow = tab + ady * a.tabw + a.rpad;
        for (int qb = -w4; qb < P + w4; qb += 4) {
            const uint4 jn = *reinterpret_cast<const uint4 *>(jrow + qb);
            const uint4 sn = SEP ? *reinterpret_cast<const uint4 *>(srow + qb) : jn;
            const float4 t0 = *reinterpret_cast<const float4 *>(trow + qb);
            const float4 t1 = *reinterpret_cast<const float4 *>(trow + qb + 4);
            const float4 t2 = *reinterpret_cast<const float4 *>(trow + qb + 8);
            const float T[12] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w, t2.x, t2.y, t2.z, t2.w};
            const uint32_t jv[4] = {jn.x, jn.y, jn.z, jn.w};
            const uint32_t sv[4] = {sn.x, sn.y, sn.z, sn.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const unsigned long long bg = pack2(byte_to_float(sv[j], 0), byte_to_float(sv[j], 1));
                const unsigned long long r1 = pack2(byte_to_float(sv[j], 2), byte_to_float(sv[j], 3));
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    uint32_t d;
                    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;"
                        : "=r"(d)
                        : "r"(jc[p]), "r"(jv[j]), "r"(0x4B000000u));
                    // d is the float 2^23 + alpha; one FMA gives RN(alpha * sc) exactly as a multiply would
                    const float u = fmaf(__uint_as_float(d), sc, scb);
                    const float wgt = ex2_approx(fmaf(-u, u, T[j - p + 7]));
                    const unsigned long long ww = pack2(wgt, wgt);
                    ffma2(acc01[p], ww, bg);
                    ffma2(acc23[p], ww, r1);
                }
            }
        }
    }

    const int gy = ty0 + ty;
    if (gy >= a.h) return;
    uint8_t *drow = a.dst + (img * npx + (size_t)gy * a.w) * a.dc;
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const int gx = tx0 + x0 + p;
        if (gx >= a.w) break;
        float s0, s1, s2, ws;
        unpack2(acc01[p], s0, s1);
        unpack2(acc23[p], s2, ws);
        uint8_t *o = drow + (size_t)gx * a.dc;
        o[0] = sat_u8(__fdiv_rn(s0, ws));
        if (a.dc == 3) {
            o[1] = sat_u8(__fdiv_rn(s1, ws));
            o[2] = sat_u8(__fdiv_rn(s2, ws));
        }
    }
}

// ---- gray kernel: 1-channel joint and src held as floats ---------------------------------------
// alpha = scale * |dJ| with scale folded into ksqrt (scale = 3 for replicated gray, 1 for true 1-channel).
template <int WY, bool SEP>
__global__ void __launch_bounds__(32 * WY) bf_gray_kernel(const Args a)
{
    extern __shared__ __align__(16) uint32_t smem[];
    const int rows = rows_of(WY, a.r);
    float *tj = reinterpret_cast<float *>(smem);
    float *ts = SEP ? tj + rows * a.pitch : tj;
    float *tab = ts + rows * a.pitch;
    const int *roww4 = reinterpret_cast<const int *>(tab + (a.r + 1) * a.tabw);

    const int img = blockIdx.z;
    const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * (ROWS_PER_WARP * WY);
    const size_t npx = (size_t)a.h * a.w;
    fill_float<WY>(tj, a.joint + img * npx, a, tx0, ty0);
    if (SEP) fill_float<WY>(ts, a.src + img * npx, a, tx0, ty0);
    load_table(tab, a);
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x0 = (lane & 3) * P;
    const int ty = warp * ROWS_PER_WARP + (lane >> 2);

    float sum[P], wsum[P];
    float jc[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
        sum[p] = 0.0f;
        wsum[p] = 0.0f;
        jc[p] = tj[(ty + a.r) * a.pitch + a.rpad + x0 + p];
    }
    const float sc = a.ksqrt;
    const int r = a.r;
    for (int dyi = 0; dyi <= 2 * r; ++dyi) {
        const int ady = dyi < r ? r - dyi : dyi - r;
        const int w4 = roww4[ady];
        const float *jrow = tj + (ty + dyi) * a.pitch + a.rpad + x0;
        const float *srow = ts + (ty + dyi) * a.pitch + a.rpad + x0;
        const float *trow = tab + ady * a.tabw + a.rpad;
        for (int qb = -w4; qb < P + w4; qb += 4) {
            const float4 jn = *reinterpret_cast<const float4 *>(jrow + qb);
            const float4 sn = SEP ? *reinterpret_cast<const float4 *>(srow + qb) : jn;
            const float4 t0 = *reinterpret_cast<const float4 *>(trow + qb);
            const float4 t1 = *reinterpret_cast<const float4 *>(trow + qb + 4);
            const float4 t2 = *reinterpret_cast<const float4 *>(trow + qb + 8);
            const float T[12] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w, t2.x, t2.y, t2.z, t2.w};
            const float jv[4] = {jn.x, jn.y, jn.z, jn.w};
            const float sv[4] = {sn.x, sn.y, sn.z, sn.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    // this path is MUFU-bound (one EX2 per tap), so the plain FFMA + FADD pair costs
                    // nothing over a packed FFMA2 and needs no (s, 1.0f) register pairs
                    const float u = (jc[p] - jv[j]) * sc;
                    const float wgt = ex2_approx(fmaf(-u, u, T[j - p + 7]));
                    sum[p] = fmaf(wgt, sv[j], sum[p]);
                    wsum[p] += wgt;
                }
            }
        }
    }

    const int gy = ty0 + ty;
    if (gy >= a.h) return;
    uint8_t *drow = a.dst + (img * npx + (size_t)gy * a.w) * a.dc;
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const int gx = tx0 + x0 + p;
        if (gx >= a.w) break;
        const uint8_t v = sat_u8(__fdiv_rn(sum[p], wsum[p]));
        uint8_t *o = drow + (size_t)gx * a.dc;
        o[0] = v;
        if (a.dc == 3) {
            o[1] = v;
            o[2] = v;
        }
    }
}

// ---- host side -----------------------------------------------------------------------------------
struct Geometry {
    int r, rpad, pitch, tabw, taps;
};

static int ceil4(int v) { return (v + 3) & ~3; }

static Geometry geometry(double sigma_space, int d)
{
    Geometry g;
    // cvRound(sigma_space * 1.5): round-half-even, as OpenCV; at least 1
    int r = d <= 0 ? (int)std::nearbyint(sigma_space * 1.5) : d / 2;
    g.r = r < 1 ? 1 : r;
    g.rpad = ceil4(g.r);
    int cols = TW + 2 * g.rpad;
    g.pitch = cols + ((cols % 8 == 4) ? 0 : 4);  // pitch = 4 (mod 8) words: conflict-free LDS.128
    g.tabw = 2 * g.rpad + 16;
    g.taps = 0;
    for (int i = -g.r; i <= g.r; ++i)
        for (int j = -g.r; j <= g.r; ++j)
            if (i * i + j * j <= g.r * g.r) ++g.taps;
    return g;
}

struct TableEntry {
    int device;
    double sigma_space;
    int r;
    float *d_tab;
};
static std::mutex g_tab_mu;
static std::vector<TableEntry> g_tabs;

// exponent table in the log2 domain: tab[ady][i] = (dx^2 + ady^2) * (-0.5 / sigma_space^2) * log2(e),
// dx = i - rpad - 7, or -inf outside the disc; followed by ceil4(half width) per |dy|.
static int get_table(double sigma_space, const Geometry &g, const float **out)
{
    int dev = 0;
    RF_CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_tab_mu);
    for (const TableEntry &e : g_tabs)
        if (e.device == dev && e.sigma_space == sigma_space && e.r == g.r) {
            *out = e.d_tab;
            return RF_OK;
        }
    const int n = (g.r + 1) * g.tabw + (g.r + 1);
    std::vector<float> h(n);
    const double gsc = -0.5 / (sigma_space * sigma_space) * 1.4426950408889634074;
    for (int ady = 0; ady <= g.r; ++ady) {
        for (int i = 0; i < g.tabw; ++i) {
            const int dx = i - g.rpad - 7;
            const int d2 = dx * dx + ady * ady;
            h[ady * g.tabw + i] = d2 <= g.r * g.r ? (float)(d2 * gsc) : -INFINITY;
        }
        const int hw = (int)std::floor(std::sqrt((double)(g.r * g.r - ady * ady)));
        int w4 = ceil4(hw);
        reinterpret_cast<int *>(h.data())[(g.r + 1) * g.tabw + ady] = w4;
    }
    float *d = nullptr;
    RF_CUDA_TRY(cudaMalloc(&d, n * sizeof(float)));
    RF_CUDA_TRY(cudaMemcpy(d, h.data(), n * sizeof(float), cudaMemcpyHostToDevice));
    if (g_tabs.size() >= 64) {  // bounded cache: drop the oldest entry of this device
        for (size_t i = 0; i < g_tabs.size(); ++i)
            if (g_tabs[i].device == dev) {
                cudaFree(g_tabs[i].d_tab);
                g_tabs.erase(g_tabs.begin() + i);
                break;
            }
    }
    g_tabs.push_back({dev, sigma_space, g.r, d});
    *out = d;
    return RF_OK;
}

static size_t smem_bytes(int wy, const Geometry &g, bool sep)
{
    const size_t tile = (size_t)rows_of(wy, g.r) * g.pitch * 4;
    return tile * (sep ? 2 : 1) + ((size_t)(g.r + 1) * g.tabw + (g.r + 1)) * 4;
}

template <typename K>
static int launch(K kernel, int wy, const Args &a, size_t smem, cudaStream_t st, const char *name)
{
    // opt in to large dynamic shared memory once per (kernel instantiation, device); K is the same
    // function-pointer type for every instantiation, so the record is keyed by the pointer value
    static std::mutex mu;
    static std::vector<std::pair<const void *, int>> configured;
    int dev = 0;
    RF_CUDA_TRY(cudaGetDevice(&dev));
    {
        std::lock_guard<std::mutex> lk(mu);
        bool done = false;
        for (const auto &c : configured) done |= (c.first == (const void *)kernel && c.second == dev);
        if (!done) {
            RF_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            configured.emplace_back((const void *)kernel, dev);
        }
    }
    dim3 grid((a.w + TW - 1) / TW, (a.h + ROWS_PER_WARP * wy - 1) / (ROWS_PER_WARP * wy), a.n);
    kernel<<<grid, 32 * wy, smem, st>>>(a);
    RF_LAUNCH_CHECK(name);
    return RF_OK;
}

static int pick_wy(const Args &a, const Geometry &g, bool sep)
{
    // Larger CTAs amortise the window fill; smaller ones balance a small grid over 148 SMs.
    const long warp_tiles = (long)((a.w + TW - 1) / TW) * ((a.h + ROWS_PER_WARP - 1) / ROWS_PER_WARP) * a.n;
    const long sms = sm_count();
    int wy = 4;
    if (warp_tiles < sms * 4 * 8) wy = 2;
    if (warp_tiles < sms * 2 * 8) wy = 1;
    while (wy > 1 && smem_bytes(wy, g, sep) > 200 * 1024) wy >>= 1;
    return wy;
}

}  // namespace bf
}  // namespace rf

namespace rf {
namespace bf2 {  // bf2.cu: packed two-output kernel for single-channel joint + src
int run(const uint8_t *joint, const uint8_t *src, uint8_t *dst, int n, int h, int w, int r, double sigma_color,
        double sigma_space, double alpha_scale, cudaStream_t st);
}
}  // namespace rf

using namespace rf;

// RF_BF_V1=1 keeps the first-generation gray kernel (used to cross-check the two implementations)
static bool use_v1_gray()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("RF_BF_V1");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

extern "C" int rf_joint_bilateral_max_radius(void) { return bf::MAX_RADIUS; }

extern "C" int rf_joint_bilateral_geometry(double sigma_space, int d, int *radius, int *taps)
{
    if (sigma_space <= 0) sigma_space = 1;
    bf::Geometry g = bf::geometry(sigma_space, d);
    if (radius) *radius = g.r;
    if (taps) *taps = g.taps;
    return RF_OK;
}

extern "C" int rf_joint_bilateral_u8(const uint8_t *joint, int jc, const uint8_t *src, int sc, uint8_t *dst,
                                     int n, int h, int w, double sigma_color, double sigma_space, int d,
                                     unsigned flags, void *stream)
{
    if (!joint || !src || !dst) return fail(RF_EINVAL, "rf_joint_bilateral_u8: NULL image pointer");
    if (!(jc == 1 || jc == 3) || !(sc == 1 || sc == 3))
        return fail(RF_EINVAL, "rf_joint_bilateral_u8: channels must be 1 or 3 (joint %d, src %d)", jc, sc);
    if (n < 0 || h < 1 || w < 1) return fail(RF_EINVAL, "rf_joint_bilateral_u8: bad shape n=%d h=%d w=%d", n, h, w);
    if (n == 0) return RF_OK;
    if (dst == joint || dst == src) return fail(RF_EINVAL, "rf_joint_bilateral_u8: dst must not alias an input");
    if (joint == src && jc != sc) return fail(RF_EINVAL, "rf_joint_bilateral_u8: aliased joint/src with different channel counts");
    const bool gray_rep = (flags & RF_BF_GRAY_REPLICATED) != 0;
    if (gray_rep && !(jc == 1 && sc == 1))
        return fail(RF_EINVAL, "rf_joint_bilateral_u8: RF_BF_GRAY_REPLICATED needs 1-channel planes");
    if (sigma_color <= 0) sigma_color = 1;
    if (sigma_space <= 0) sigma_space = 1;
    const bf::Geometry g = bf::geometry(sigma_space, d);
    if (g.r > bf::MAX_RADIUS)
        return fail(RF_EUNSUPPORTED, "rf_joint_bilateral_u8: radius %d exceeds the supported maximum %d", g.r,
                    bf::MAX_RADIUS);
    if (n > 65535) return fail(RF_EUNSUPPORTED, "rf_joint_bilateral_u8: more than 65535 images per call");

    bf::Args a;
    a.joint = joint;
    a.src = src;
    a.dst = dst;
    a.jc = jc;
    a.sc = sc;
    a.dc = sc;
    a.n = n;
    a.h = h;
    a.w = w;
    a.r = g.r;
    a.rpad = g.rpad;
    a.pitch = g.pitch;
    a.tabw = g.tabw;
    int rc = bf::get_table(sigma_space, g, &a.tab);
    if (rc != RF_OK) return rc;
    const bool sep = joint != src;
    cudaStream_t st = (cudaStream_t)stream;
    const double ksq = std::sqrt(0.5 / (sigma_color * sigma_color) * 1.4426950408889634074);
    const bool gray = (jc == 1 && sc == 1);
    if (gray && !use_v1_gray())
        return bf2::run(joint, src, dst, n, h, w, g.r, sigma_color, sigma_space, gray_rep ? 3.0 : 1.0,
                        (cudaStream_t)stream);
    if (gray) {
        const double scale = gray_rep ? 3.0 : 1.0;
        a.ksqrt = (float)(ksq * scale);
        a.dc = 1;
        const int wy = bf::pick_wy(a, g, sep);
        const size_t smem = bf::smem_bytes(wy, g, sep);
#define RF_BF_GRAY(WY)                                                                                   \
    case WY:                                                                                             \
        return sep ? bf::launch(bf::bf_gray_kernel<WY, true>, WY, a, smem, st, "bf_gray_kernel")         \
                   : bf::launch(bf::bf_gray_kernel<WY, false>, WY, a, smem, st, "bf_gray_kernel")
        switch (wy) {
            RF_BF_GRAY(1);
            RF_BF_GRAY(2);
            RF_BF_GRAY(4);
        }
#undef RF_BF_GRAY
    } else {
        a.ksqrt = (float)ksq;
        const int wy = bf::pick_wy(a, g, sep);
        const size_t smem = bf::smem_bytes(wy, g, sep);
#define RF_BF_COLOR(WY)                                                                                  \
    case WY:                                                                                             \
        return sep ? bf::launch(bf::bf_color_kernel<WY, true>, WY, a, smem, st, "bf_color_kernel")       \
                   : bf::launch(bf::bf_color_kernel<WY, false>, WY, a, smem, st, "bf_color_kernel")
        switch (wy) {
            RF_BF_COLOR(1);
            RF_BF_COLOR(2);
            RF_BF_COLOR(4);
        }
#undef RF_BF_COLOR
    }
    return fail(RF_EINVAL, "rf_joint_bilateral_u8: internal dispatch error");
}
