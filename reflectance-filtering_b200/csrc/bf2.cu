// Joint bilateral filter, single-channel (gray) path, second generation.
//
// Same semantics as bf.cu (jointBilateralFilter_8u; call site
// /root/reference/filter_reflectance.py:60-64) for the case where joint and src are one plane each:
// the BF(CNN, CNN) configuration (alpha = 3|dJ| for the three equal channels cv2.imread makes of the
// CNN's gray PNG) and true 1-channel images (alpha = |dJ|).
//
// What ncu showed for the first kernel (profiles/r01_bf_gray_v1_ncu_full.txt): the XU pipe (MUFU.EX2, one
// per tap, 16 lanes/clk/SM) is 86 % busy, but 15.5 % of the executed taps lie outside the disc because a
// thread's 8 outputs x 4 neighbours form a coarse block.  This version makes the block 2 x 2:
//   * a thread owns TWO adjacent outputs, held as the halves of 64-bit registers; every FP32 step of a tap
//     pair is one packed instruction (FFMA2 / FADD2 issue at half rate but carry two taps), so the
//     issue slots per tap drop from 6.3 to ~4.9 and leave the XU pipe as the only limiter;
//   * neighbours arrive two at a time (LDS.64), executed taps / useful taps = 1.035 instead of 1.18;
//   * the per-row spatial exponents for the four taps of a chunk come from one broadcast LDS.128 of a
//     table laid out per chunk: (e(qb), e(qb-1), e(qb+1), e(qb)), +inf outside the disc (weight 0).
#include <cmath>
#include <cstdlib>
#include <mutex>
#include <type_traits>
#include <vector>

#include "common.cuh"

namespace rf {
namespace bf2 {

constexpr int TW = 64;  // tile width: 32 lanes x 2 outputs

struct Args {
    const uint8_t *joint;
    const uint8_t *src;
    uint8_t *dst;
    const float *tab;  // device: (r+1) x nchunk float4, then (r+1) ints (even half widths)
    int n, h, w, dc;
    int r, rpad, pitch, nchunk;
    float ksqrt;  // sqrt(0.5 / sigma_color^2 * log2 e) * alpha_scale: the joint tile is stored pre-multiplied by it
    float inv_ksqrt;
};

__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b)
{
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long ffma2r(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ float ex2_neg(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(-x));
    return y;
}

// 2^-a for both halves on the FMA pipe: a = n - g with n = round(a), g in [-0.5, 0.5];
// 2^g by a degree-5 polynomial (max relative error 2.0e-7 in FP32 Horner form, i.e. as good as MUFU.EX2),
// 2^-n by subtracting n from the exponent field.  a is clamped to 126, which also maps the +inf of
// out-of-disc taps to a weight of 2^-126 (1e-38 against a weight sum >= 1).
__device__ __forceinline__ unsigned long long exp2neg_poly2(unsigned long long a2)
{
    float a0, a1;
    unpack2(a2, a0, a1);
    const unsigned long long ac = pack2(fminf(a0, 126.0f), fminf(a1, 126.0f));
    const unsigned long long nm = fadd2(ac, pack2(12582912.0f, 12582912.0f));    // 1.5 * 2^23 + n
    const unsigned long long nf = fadd2(nm, pack2(-12582912.0f, -12582912.0f));  // n as float
    const unsigned long long g = ffma2r(ac, pack2(-1.0f, -1.0f), nf);            // n - a
    unsigned long long p = pack2(0.0013280353741720319f, 0.0013280353741720319f);
    p = ffma2r(p, g, pack2(0.009675574488937855f, 0.009675574488937855f));
    p = ffma2r(p, g, pack2(0.05550701171159744f, 0.05550701171159744f));
    p = ffma2r(p, g, pack2(0.24022118747234344f, 0.24022118747234344f));
    p = ffma2r(p, g, pack2(0.6931470036506653f, 0.6931470036506653f));
    p = ffma2r(p, g, pack2(1.0000001192092896f, 1.0000001192092896f));
    float p0, p1, n0, n1;
    unpack2(p, p0, p1);
    unpack2(nm, n0, n1);
    // the low bits of (1.5 * 2^23 + n) are n; shifting by 23 drops the magic constant entirely
    const float r0 = __uint_as_float(__float_as_uint(p0) - (__float_as_uint(n0) << 23));
    const float r1 = __uint_as_float(__float_as_uint(p1) - (__float_as_uint(n1) << 23));
    return pack2(r0, r1);
}

// POLY: every POLY-th chunk evaluates the weights of its first neighbour with exp2neg_poly2 instead of
// MUFU.EX2 (0 = never).  The kernel is bound by the XU pipe (one EX2 per tap); every packed instruction taken out of
// the MUFU path (round 2: the window is stored pre-multiplied by ksqrt, which removes the FMUL2 of a tap pair) leaves
// FMA-pipe and issue room to move more weight pairs over: one pair in eight (POLY = 4) is the measured optimum,
// 98.7 % of the 1-MUFU-per-tap roofline.
template <int WY, bool SEP, int POLY>
__global__ void __launch_bounds__(32 * WY) bf_gray2_kernel(const Args a)
{
    extern __shared__ __align__(16) float smem_f[];
    const int rows = WY + 2 * a.r;
    const int tile = (rows * a.pitch + 3) & ~3;  // keeps the float4 table 16-byte aligned
    float *tj = smem_f;
    float *ts = SEP ? tj + tile : tj;
    float4 *tab = reinterpret_cast<float4 *>(ts + tile);
    const int *roww2 = reinterpret_cast<const int *>(tab + (a.r + 1) * a.nchunk);

    const int img = blockIdx.z;
    const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * WY;
    const size_t npx = (size_t)a.h * a.w;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    {
        const int cols = TW + 2 * a.rpad + 2;
        const uint8_t *J = a.joint + img * npx;
        const uint8_t *S = a.src + img * npx;
        for (int yy = warp; yy < rows; yy += WY) {
            const int gy = reflect101(ty0 - a.r + yy, a.h);
            const uint8_t *jrow = J + (size_t)gy * a.w;
            const uint8_t *srow = S + (size_t)gy * a.w;
            for (int cc = lane; cc < cols; cc += 32) {
                const int gx = reflect101(tx0 - a.rpad + cc, a.w);
                // the joint window is stored as ksqrt * J: the range term of a tap is then (dJ')^2, one packed
                // instruction less per tap pair.  With joint == src the sums come out scaled and are divided at the
                // end; a distinct source is scaled the same way so that equal inputs give equal bytes on both paths
                tj[yy * a.pitch + cc] = (float)jrow[gx] * a.ksqrt;
                if (SEP) ts[yy * a.pitch + cc] = (float)srow[gx] * a.ksqrt;
            }
        }
        const int n4 = (a.r + 1) * a.nchunk * 4 + (a.r + 1);
        float *t = reinterpret_cast<float *>(tab);
        for (int i = threadIdx.x; i < n4; i += 32 * WY) t[i] = a.tab[i];
    }
    __syncthreads();

    const int x0 = lane * 2;
    const int ty = warp;
    const float *jc_p = tj + (ty + a.r) * a.pitch + a.rpad + x0;
    const unsigned long long jc2 = pack2(jc_p[0], jc_p[1]);
    const unsigned long long neg1 = pack2(-1.0f, -1.0f);
    unsigned long long sum2 = 0ull, wsum2 = 0ull;
    const int r = a.r;
    for (int dyi = 0; dyi <= 2 * r; ++dyi) {
        const int ady = dyi < r ? r - dyi : dyi - r;
        const int w2 = roww2[ady];
        // chunk t covers neighbours qb = 2t - rpad and qb + 1
        const int t0 = (a.rpad - w2) >> 1, t1 = (a.rpad + w2) >> 1;
        const float2 *jrow = reinterpret_cast<const float2 *>(tj + (ty + dyi) * a.pitch + x0);
        const float2 *srow = reinterpret_cast<const float2 *>(ts + (ty + dyi) * a.pitch + x0);
        const float4 *trow = tab + ady * a.nchunk;
        // one chunk = neighbours (qb, qb+1) x outputs (p0, p1); USE_POLY is a compile-time choice so the
        // unrolled groups below are straight-line code
        auto chunk = [&](int t, auto use_poly) {
            const float2 jn = jrow[t];
            const float2 sn = SEP ? srow[t] : jn;
            const float4 e = trow[t];
            {   // neighbour qb: taps (p0: dx = qb, p1: dx = qb - 1)
                const unsigned long long u2 = ffma2r(pack2(jn.x, jn.x), neg1, jc2);
                const unsigned long long a2 = ffma2r(u2, u2, pack2(e.x, e.y));
                unsigned long long w;
                if (decltype(use_poly)::value) {
                    w = exp2neg_poly2(a2);
                } else {
                    float a0, a1;
                    unpack2(a2, a0, a1);
                    w = pack2(ex2_neg(a0), ex2_neg(a1));
                }
                sum2 = ffma2r(pack2(sn.x, sn.x), w, sum2);
                wsum2 = fadd2(wsum2, w);
            }
            {   // neighbour qb + 1: taps (p0: dx = qb + 1, p1: dx = qb)
                const unsigned long long u2 = ffma2r(pack2(jn.y, jn.y), neg1, jc2);
                const unsigned long long a2 = ffma2r(u2, u2, pack2(e.z, e.w));
                float a0, a1;
                unpack2(a2, a0, a1);
                const unsigned long long w = pack2(ex2_neg(a0), ex2_neg(a1));
                sum2 = ffma2r(pack2(sn.y, sn.y), w, sum2);
                wsum2 = fadd2(wsum2, w);
            }
        };
        int t = t0;
        if (POLY > 0) {
            // groups of POLY chunks: the first one takes the polynomial path for its first neighbour
            for (; t + POLY - 1 <= t1; t += POLY) {
                chunk(t, std::true_type{});
#pragma unroll
                for (int u = 1; u < (POLY > 0 ? POLY : 1); ++u) chunk(t + u, std::false_type{});
            }
        }
        for (; t <= t1; ++t) chunk(t, std::false_type{});
    }

    const int gy = ty0 + ty;
    if (gy >= a.h) return;
    float s0, s1, w0, w1;
    unpack2(sum2, s0, s1);
    unpack2(wsum2, w0, w1);
    uint8_t *drow = a.dst + (img * npx + (size_t)gy * a.w) * a.dc;
    const float sv[2] = {s0, s1}, wv[2] = {w0, w1};
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const int gx = tx0 + x0 + p;
        if (gx >= a.w) break;
        const float q = __fdiv_rn(sv[p], wv[p]);
        const uint8_t v = sat_u8(q * a.inv_ksqrt);
        uint8_t *o = drow + (size_t)gx * a.dc;
        o[0] = v;
        if (a.dc == 3) {
            o[1] = v;
            o[2] = v;
        }
    }
}

// ---- host ------------------------------------------------------------------------------------------
struct Geometry {
    int r, rpad, pitch, nchunk;
};

static Geometry geometry(int r)
{
    Geometry g;
    g.r = r;
    g.rpad = (r + 1) & ~1;               // even
    g.pitch = TW + 2 * g.rpad + 2;       // even: LDS.64 of a lane's two neighbours is always aligned
    g.nchunk = g.rpad + 2;
    return g;
}

struct TableEntry {
    int device;
    double sigma_space;
    int r;
    float *d_tab;
};
static std::mutex g_mu;
static std::vector<TableEntry> g_tabs;

// tab[ady][t] = (E(qb), E(qb-1), E(qb+1), E(qb)) with qb = 2t - rpad and
// E(dx) = (dx^2 + ady^2) * 0.5 / sigma_space^2 * log2(e) inside the disc, +inf outside;
// followed by the even-rounded half width of every row.
// Call with g_mu held; the caller keeps it until its kernel is in the stream: an eviction cudaFree()s (which waits for
// running kernels) and must not hit a table between another thread's lookup and its launch.
static int get_table(double sigma_space, const Geometry &g, const float **out)
{
    int dev = 0;
    RF_CUDA_TRY(cudaGetDevice(&dev));
    for (const TableEntry &e : g_tabs)
        if (e.device == dev && e.sigma_space == sigma_space && e.r == g.r) {
            *out = e.d_tab;
            return RF_OK;
        }
    const int n = (g.r + 1) * g.nchunk * 4 + (g.r + 1);
    std::vector<float> h(n);
    const double gs = 0.5 / (sigma_space * sigma_space) * 1.4426950408889634074;
    auto E = [&](int dx, int ady) -> float {
        const int d2 = dx * dx + ady * ady;
        return d2 <= g.r * g.r ? (float)(d2 * gs) : INFINITY;
    };
    for (int ady = 0; ady <= g.r; ++ady) {
        for (int t = 0; t < g.nchunk; ++t) {
            const int qb = 2 * t - g.rpad;
            float *e = &h[(ady * g.nchunk + t) * 4];
            e[0] = E(qb, ady);
            e[1] = E(qb - 1, ady);
            e[2] = E(qb + 1, ady);
            e[3] = E(qb, ady);
        }
        const int hw = (int)std::floor(std::sqrt((double)(g.r * g.r - ady * ady)));
        reinterpret_cast<int *>(h.data())[(g.r + 1) * g.nchunk * 4 + ady] = (hw + 1) & ~1;
    }
    float *d = nullptr;
    RF_CUDA_TRY(cudaMalloc(&d, n * sizeof(float)));
    RF_CUDA_TRY(cudaMemcpy(d, h.data(), n * sizeof(float), cudaMemcpyHostToDevice));
    if (g_tabs.size() >= 64) {
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
    const size_t tile = (((size_t)(wy + 2 * g.r) * g.pitch + 3) & ~(size_t)3) * 4;
    return tile * (sep ? 2 : 1) + ((size_t)(g.r + 1) * g.nchunk * 4 + (g.r + 1)) * 4;
}

// RF_BF_POLY=<n> overrides how often the polynomial path is used (0 = never, n = first neighbour of every n-th chunk).
// Measured on B200, 64x512x384, c20 s22 (profiles/r02_bf_poly_sweep.txt): with the pre-scaled window (round 2) off
// 10.19 ms, n = 5: 9.59, 4: 9.34, 3: 9.46, 2: 10.43; before it (round 1) off 10.33, 5: 9.87, 4: 9.99, 3: 10.57.
static int poly_every()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("RF_BF_POLY");
        v = e ? atoi(e) : 4;
        if (v != 0 && (v < 2 || v > 5)) v = 4;
    }
    return v;
}

template <int WY, bool SEP, int POLY>
static int launch(const Args &a, size_t smem, cudaStream_t st)
{
    static DeviceOnce once;  // one-time per-device kernel attribute; safe from several host threads
    int dev = 0;
    RF_CUDA_TRY(cudaGetDevice(&dev));
    {
        std::lock_guard<std::mutex> lock(once.mu);
        if (!once.done[dev & 63]) {
            RF_CUDA_TRY(cudaFuncSetAttribute(bf_gray2_kernel<WY, SEP, POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            once.done[dev & 63] = true;
        }
    }
    dim3 grid((a.w + TW - 1) / TW, (a.h + WY - 1) / WY, a.n);
    bf_gray2_kernel<WY, SEP, POLY><<<grid, 32 * WY, smem, st>>>(a);
    RF_LAUNCH_CHECK("bf_gray2_kernel");
    return RF_OK;
}

int run(const uint8_t *joint, const uint8_t *src, uint8_t *dst, int n, int h, int w, int r, double sigma_color,
        double sigma_space, double alpha_scale, cudaStream_t st)
{
    const Geometry g = geometry(r);
    Args a;
    a.joint = joint;
    a.src = src;
    a.dst = dst;
    a.n = n;
    a.h = h;
    a.w = w;
    a.dc = 1;
    a.r = g.r;
    a.rpad = g.rpad;
    a.pitch = g.pitch;
    a.nchunk = g.nchunk;
    const double ks = std::sqrt(0.5 / (sigma_color * sigma_color) * 1.4426950408889634074) * alpha_scale;
    a.ksqrt = (float)ks;
    a.inv_ksqrt = (float)(1.0 / (double)a.ksqrt);
    const bool sep = joint != src;
    // rows per CTA: big tiles amortise the window fill, small ones balance small grids over the SMs
    const long tiles16 = (long)((w + TW - 1) / TW) * ((h + 15) / 16) * n;
    int wy = 16;
    if (tiles16 < 4L * sm_count()) wy = 8;
    if (tiles16 < 1L * sm_count()) wy = 4;
    while (wy > 4 && smem_bytes(wy, g, sep) > 100 * 1024) wy >>= 1;
    const size_t smem = smem_bytes(wy, g, sep);
    if (smem > 227 * 1024) return fail(RF_EUNSUPPORTED, "bf_gray2: radius %d needs %zu bytes of shared memory", r, smem);
    std::lock_guard<std::mutex> lk(g_mu);  // held until the kernel that reads the table is in the stream
    const int rc = get_table(sigma_space, g, &a.tab);
    if (rc != RF_OK) return rc;
#define RF_BF2P(WY, PL) (sep ? launch<WY, true, PL>(a, smem, st) : launch<WY, false, PL>(a, smem, st))
#define RF_BF2(WY)                                                                            \
    case WY:                                                                                  \
        switch (poly_every()) {                                                               \
            case 0: return RF_BF2P(WY, 0);                                                    \
            case 2: return RF_BF2P(WY, 2);                                                    \
            case 3: return RF_BF2P(WY, 3);                                                    \
            case 4: return RF_BF2P(WY, 4);                                                    \
            case 5: return RF_BF2P(WY, 5);                                                    \
            default: return RF_BF2P(WY, 4);                                                   \
        }
    switch (wy) {
        RF_BF2(4);
        RF_BF2(8);
        RF_BF2(16);
    }
#undef RF_BF2
#undef RF_BF2P
    return fail(RF_EINVAL, "bf_gray2: internal dispatch error");
}

}  // namespace bf2
}  // namespace rf
