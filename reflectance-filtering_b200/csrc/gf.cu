// Guided filter (colour guide, 8-bit) for sm_100a.
//
// Replaces cv2.ximgproc.guidedFilter(guide, src, radius, eps) at
// /root/reference/filter_reflectance.py:67-70 (semantics: OpenCV-contrib 3.1.0 guided_filter.cpp,
// SURVEY.md Appendix A.3): 33 normalised (2r+1)^2 box means with BORDER_REFLECT, cov(I) + eps*Id on
// the 0..255 scale, per-pixel 3x3 inverse by cofactors, q = mean(a).I + mean(b), round-half-even.
//
// Two streaming passes (DESIGN.md "K4"):
//   pass A  reads guide + src (u8), forms the 9 + 4*SC window sums I, I*I', p, p*I as EXACT integers
//           (products <= 65,025: a window of (2r+1)^2 <= 255^2 pixels stays below 2^32, so uint32 up to
//           r = 127 and uint64 above), converts each to the box
//           mean exactly as OpenCV does -- float(double(sum) * (1.0 / (2r+1)^2)) -- and solves
//           for (a0, a1, a2, b) per source channel; writes one float4 per pixel per channel.
//   pass B  box-means the four coefficient planes with FP64 running sums (OpenCV's box filter
//           accumulates CV_32F input in double), applies q = b + a.I and writes uint8.
// Both passes give every thread one image column of a vertical strip: the vertical window sum
// lives in registers and slides down the strip (add the entering row, subtract the leaving
// one); the horizontal window sum is a block-wide prefix scan (warp shuffles + one shared-memory
// hop) followed by P[x+r] - P[x-r-1].  All elementwise FP32 math uses the non-contracting
// __fmul_rn/__fadd_rn forms so pass A reproduces the oracle's separate multiply/add bit for bit.
#include <cmath>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"

namespace rf {
namespace gf {

constexpr int NT = 512;      // threads per CTA = columns per strip including the 2r halo
constexpr int NW = NT / 32;  // warps
constexpr int MAX_RADIUS = (NT - 32) / 2;

struct Args {
    const void *guide;     // [n][h][w][GC]  uint8 or float (template parameter TI of the kernels)
    const void *src;       // [n][h][w][SC]
    float4 *ab;            // [n][SC][h][w] (a0, a1, a2, b)
    void *dst;             // [n][h][w][SC]  depth of src
    int n, h, w, r;
    int twa;       // output columns per strip
    int seg_rows;  // output rows per CTA
    float eps;
    double scale;  // 1 / (2r+1)^2
};

// ---- block-wide inclusive scan of Q per-thread values, result left in v[] -----------------------
// pref[q][0] must be zero; on return pref[q][tid + 1] holds the inclusive prefix of column tid.
template <typename T, int Q>
__device__ __forceinline__ void block_scan_to_smem(T (&v)[Q], T *pref, T *wtot)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        T x = v[q];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const T t = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += t;
        }
        v[q] = x;
        if (lane == 31) wtot[q * NW + warp] = x;
    }
    __syncthreads();
    if (warp < (Q * NW + 31) / 32) {
        const bool live = tid < Q * NW;
        const T own = live ? wtot[tid] : T(0);
        T inc = own;
#pragma unroll
        for (int d = 1; d < NW; d <<= 1) {
            const T t = __shfl_up_sync(0xffffffffu, inc, d, NW);
            if ((tid & (NW - 1)) >= d) inc += t;
        }
        if (live) wtot[tid] = inc - own;  // exclusive offset of each warp
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        v[q] += wtot[q * NW + warp];
        pref[q * (NT + 1) + tid + 1] = v[q];
    }
    __syncthreads();
}

// ---- pass A --------------------------------------------------------------------------------------
template <int GC>
__host__ __device__ constexpr int n_quant(int sc) { return GC == 3 ? 9 + 4 * sc : 2 + 2 * sc; }

// window quantities of one pixel.  Colour guide: I (3), I*I' (6), then per source channel p, p*I0, p*I1, p*I2.
// 1-channel guide: I, I*I, then per source channel p, p*I.
// TI = uint8_t: exact integer products.  TI = float (CV_32F images): float products, each rounded once as
// OpenCV's multiply() of two CV_32F planes does, summed in double like its box filter.
template <int SC, int GC, typename TI>
struct PixA {
    typedef typename std::conditional<std::is_same<TI, float>::value, float, uint32_t>::type FT;
    FT f[n_quant<GC>(SC)];
    __device__ __forceinline__ static FT mul(FT a, FT b)
    {
        if constexpr (std::is_same<TI, float>::value)
            return __fmul_rn(a, b);
        else
            return a * b;
    }
    __device__ __forceinline__ void load(const TI *g, const TI *s)
    {
        if constexpr (GC == 3) {
            const FT i0 = g[0], i1 = g[1], i2 = g[2];
            f[0] = i0; f[1] = i1; f[2] = i2;
            f[3] = mul(i0, i0); f[4] = mul(i0, i1); f[5] = mul(i0, i2);
            f[6] = mul(i1, i1); f[7] = mul(i1, i2); f[8] = mul(i2, i2);
#pragma unroll
            for (int c = 0; c < SC; ++c) {
                const FT p = s[c];
                f[9 + 4 * c] = p;
                f[10 + 4 * c] = mul(p, i0);
                f[11 + 4 * c] = mul(p, i1);
                f[12 + 4 * c] = mul(p, i2);
            }
        } else {
            const FT i0 = g[0];
            f[0] = i0;
            f[1] = mul(i0, i0);
#pragma unroll
            for (int c = 0; c < SC; ++c) {
                const FT p = s[c];
                f[2 + 2 * c] = p;
                f[3 + 2 * c] = mul(p, i0);
            }
        }
    }
};

// ST: type of the window sums (uint32_t for r <= MAX_RADIUS_U32, where 65,025 * (2r+1)^2 < 2^32; else 64-bit)
constexpr int MAX_RADIUS_U32 = 127;

template <int SC, int GC, typename ST, typename TI>
__global__ void __launch_bounds__(NT) gf_pass_a(const Args g)
{
    constexpr int Q = n_quant<GC>(SC);
    extern __shared__ __align__(16) unsigned char sm_raw[];
    ST *pref = reinterpret_cast<ST *>(sm_raw);  // [Q][NT + 1]
    ST *wtot = pref + Q * (NT + 1);             // [Q][NW]
    const int tid = threadIdx.x;
    const int img = blockIdx.z;
    const int sx0 = blockIdx.x * g.twa;
    const int y0 = blockIdx.y * g.seg_rows;
    const int y1 = min(g.h, y0 + g.seg_rows);
    const int r = g.r;
    const int col = sx0 - r + tid;
    const bool col_active = tid < g.twa + 2 * r;
    const int xin = reflect(col, g.w);
    const bool is_out = tid >= r && tid < r + g.twa && col < g.w;
    const size_t img_px = (size_t)g.h * g.w;
    const TI *G = static_cast<const TI *>(g.guide) + img * img_px * GC + (size_t)xin * GC;
    const TI *S = static_cast<const TI *>(g.src) + img * img_px * SC + (size_t)xin * SC;

    for (int q = tid; q < Q; q += NT) pref[q * (NT + 1)] = 0u;

    ST V[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) V[q] = 0u;
    if (col_active) {
        for (int dy = -r; dy < r; ++dy) {
            const size_t yy = (size_t)reflect(y0 + dy, g.h);
            PixA<SC, GC, TI> px;
            px.load(G + yy * g.w * GC, S + yy * g.w * SC);
#pragma unroll
            for (int q = 0; q < Q; ++q) V[q] += px.f[q];
        }
    }
    for (int y = y0; y < y1; ++y) {
        if (col_active) {
            const size_t yy = (size_t)reflect(y + r, g.h);
            PixA<SC, GC, TI> px;
            px.load(G + yy * g.w * GC, S + yy * g.w * SC);
#pragma unroll
            for (int q = 0; q < Q; ++q) V[q] += px.f[q];
        }
        ST Pq[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) Pq[q] = V[q];
        block_scan_to_smem<ST, Q>(Pq, pref, wtot);
        if (is_out) {
            float m[Q];
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const ST s = pref[q * (NT + 1) + tid + r + 1] - pref[q * (NT + 1) + tid - r];
                m[q] = (float)((double)s * g.scale);  // == cv::boxFilter's float(sum * scale)
            }
            if constexpr (GC == 1) {
                // 1-channel guide: the 1x1 inverse is a reciprocal, alpha = cov(I, p) * inv (oracle: guided_gray_guide)
                const float var = __fadd_rn(__fsub_rn(m[1], __fmul_rn(m[0], m[0])), g.eps);
                const float inv = __fdiv_rn(1.0f, var);
#pragma unroll
                for (int c = 0; c < SC; ++c) {
                    const float mp = m[2 + 2 * c];
                    const float k0 = __fsub_rn(m[3 + 2 * c], __fmul_rn(mp, m[0]));
                    const float a0 = __fmul_rn(k0, inv);
                    const float b = __fsub_rn(mp, __fmul_rn(a0, m[0]));
                    g.ab[((size_t)(img * SC + c) * g.h + y) * g.w + col] = make_float4(a0, 0.0f, 0.0f, b);
                }
            } else {
            // cov(I) + eps on the diagonal; symmetric storage 0:(0,0) 1:(0,1) 2:(0,2) 3:(1,1) 4:(1,2) 5:(2,2)
            float c00 = __fadd_rn(__fsub_rn(m[3], __fmul_rn(m[0], m[0])), g.eps);
            float c01 = __fsub_rn(m[4], __fmul_rn(m[0], m[1]));
            float c02 = __fsub_rn(m[5], __fmul_rn(m[0], m[2]));
            float c11 = __fadd_rn(__fsub_rn(m[6], __fmul_rn(m[1], m[1])), g.eps);
            float c12 = __fsub_rn(m[7], __fmul_rn(m[1], m[2]));
            float c22 = __fadd_rn(__fsub_rn(m[8], __fmul_rn(m[2], m[2])), g.eps);
            // cofactors with cyclic indices: cof[k][l] = c[k+1][l+1]*c[k+2][l+2] - c[k+1][l+2]*c[k+2][l+1]
            const float f00 = __fsub_rn(__fmul_rn(c11, c22), __fmul_rn(c12, c12));
            const float f01 = __fsub_rn(__fmul_rn(c12, c02), __fmul_rn(c01, c22));
            const float f02 = __fsub_rn(__fmul_rn(c01, c12), __fmul_rn(c11, c02));
            const float f11 = __fsub_rn(__fmul_rn(c22, c00), __fmul_rn(c02, c02));
            const float f12 = __fsub_rn(__fmul_rn(c02, c01), __fmul_rn(c12, c00));
            const float f22 = __fsub_rn(__fmul_rn(c00, c11), __fmul_rn(c01, c01));
            float det = __fmul_rn(c00, f00);
            det = __fadd_rn(det, __fmul_rn(c01, f01));
            det = __fadd_rn(det, __fmul_rn(c02, f02));
            if (g.eps < 1e-2f && fabsf(det) < 1e-6f) det = 1e-6f;
            const float i00 = __fdiv_rn(f00, det), i01 = __fdiv_rn(f01, det), i02 = __fdiv_rn(f02, det);
            const float i11 = __fdiv_rn(f11, det), i12 = __fdiv_rn(f12, det), i22 = __fdiv_rn(f22, det);
#pragma unroll
            for (int c = 0; c < SC; ++c) {
                const float mp = m[9 + 4 * c];
                const float k0 = __fsub_rn(m[10 + 4 * c], __fmul_rn(mp, m[0]));
                const float k1 = __fsub_rn(m[11 + 4 * c], __fmul_rn(mp, m[1]));
                const float k2 = __fsub_rn(m[12 + 4 * c], __fmul_rn(mp, m[2]));
                float a0 = __fmul_rn(i00, k0);
                a0 = __fadd_rn(a0, __fmul_rn(i01, k1));
                a0 = __fadd_rn(a0, __fmul_rn(i02, k2));
                float a1 = __fmul_rn(i01, k0);
                a1 = __fadd_rn(a1, __fmul_rn(i11, k1));
                a1 = __fadd_rn(a1, __fmul_rn(i12, k2));
                float a2 = __fmul_rn(i02, k0);
                a2 = __fadd_rn(a2, __fmul_rn(i12, k1));
                a2 = __fadd_rn(a2, __fmul_rn(i22, k2));
                float b = __fsub_rn(mp, __fmul_rn(a0, m[0]));
                b = __fsub_rn(b, __fmul_rn(a1, m[1]));
                b = __fsub_rn(b, __fmul_rn(a2, m[2]));
                g.ab[((size_t)(img * SC + c) * g.h + y) * g.w + col] = make_float4(a0, a1, a2, b);
            }
            }  // GC == 3
        }
        if (col_active) {
            const size_t yy = (size_t)reflect(y - r, g.h);
            PixA<SC, GC, TI> px;
            px.load(G + yy * g.w * GC, S + yy * g.w * SC);
#pragma unroll
            for (int q = 0; q < Q; ++q) V[q] -= px.f[q];
        }
    }
}

// ---- pass B --------------------------------------------------------------------------------------
template <int SC, int GC, typename TI>
__global__ void __launch_bounds__(NT) gf_pass_b(const Args g)
{
    constexpr int Q = 4 * SC;
    extern __shared__ __align__(16) double sm_f64[];
    double *pref = sm_f64;               // [Q][NT + 1]
    double *wtot = pref + Q * (NT + 1);  // [Q][NW]
    const int tid = threadIdx.x;
    const int img = blockIdx.z;
    const int sx0 = blockIdx.x * g.twa;
    const int y0 = blockIdx.y * g.seg_rows;
    const int y1 = min(g.h, y0 + g.seg_rows);
    const int r = g.r;
    const int col = sx0 - r + tid;
    const bool col_active = tid < g.twa + 2 * r;
    const int xin = reflect(col, g.w);
    const bool is_out = tid >= r && tid < r + g.twa && col < g.w;
    const size_t img_px = (size_t)g.h * g.w;
    const float4 *AB = g.ab + (size_t)img * SC * img_px + xin;

    for (int q = tid; q < Q; q += NT) pref[q * (NT + 1)] = 0.0;

    double V[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) V[q] = 0.0;
    auto add_row = [&](int yy, double sign) {
#pragma unroll
        for (int c = 0; c < SC; ++c) {
            const float4 v = AB[(size_t)c * img_px + (size_t)yy * g.w];
            V[4 * c + 0] += sign * (double)v.x;
            V[4 * c + 1] += sign * (double)v.y;
            V[4 * c + 2] += sign * (double)v.z;
            V[4 * c + 3] += sign * (double)v.w;
        }
    };
    if (col_active)
        for (int dy = -r; dy < r; ++dy) add_row(reflect(y0 + dy, g.h), 1.0);
    for (int y = y0; y < y1; ++y) {
        if (col_active) add_row(reflect(y + r, g.h), 1.0);
        double Pq[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) Pq[q] = V[q];
        block_scan_to_smem<double, Q>(Pq, pref, wtot);
        if (is_out) {
            const TI *gp = static_cast<const TI *>(g.guide) + (img * img_px + (size_t)y * g.w + col) * GC;
            const float i0 = gp[0], i1 = GC == 3 ? gp[GC - 2] : 0.0f, i2 = GC == 3 ? gp[GC - 1] : 0.0f;
            TI *o = static_cast<TI *>(g.dst) + (img * img_px + (size_t)y * g.w + col) * SC;
#pragma unroll
            for (int c = 0; c < SC; ++c) {
                float m[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int q = 4 * c + k;
                    const double s = pref[q * (NT + 1) + tid + r + 1] - pref[q * (NT + 1) + tid - r];
                    m[k] = (float)(s * g.scale);
                }
                float v = m[3];
                v = __fadd_rn(v, __fmul_rn(m[0], i0));
                v = __fadd_rn(v, __fmul_rn(m[1], i1));
                v = __fadd_rn(v, __fmul_rn(m[2], i2));
                if constexpr (std::is_same<TI, float>::value)
                    o[c] = v;
                else
                    o[c] = sat_u8(v);
            }
        }
        if (col_active) add_row(reflect(y - r, g.h), -1.0);
    }
}

static size_t per_image_ws(int sc, int h, int w) { return (size_t)sc * h * w * sizeof(float4); }

template <int SC, int GC, typename TI = uint8_t>
static int run(Args a, cudaStream_t st)
{
    constexpr bool F32 = std::is_same<TI, float>::value;
    const bool wide = F32 || a.r > MAX_RADIUS_U32;  // 64-bit window sums (double for CV_32F images)
    const size_t smem_a = ((size_t)n_quant<GC>(SC) * (NT + 1 + NW)) * (wide ? sizeof(unsigned long long) : sizeof(uint32_t));
    const size_t smem_b = ((size_t)(4 * SC) * (NT + 1 + NW)) * sizeof(double);
    static DeviceOnce once;
    int dev = 0;
    RF_CUDA_TRY(cudaGetDevice(&dev));
    {
        std::lock_guard<std::mutex> lock(once.mu);
        if (!once.done[dev & 63]) {
            if constexpr (F32) {
                RF_CUDA_TRY(cudaFuncSetAttribute(gf_pass_a<SC, GC, double, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            } else {
                RF_CUDA_TRY(cudaFuncSetAttribute(gf_pass_a<SC, GC, uint32_t, uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
                RF_CUDA_TRY(cudaFuncSetAttribute(gf_pass_a<SC, GC, unsigned long long, uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            }
            RF_CUDA_TRY(cudaFuncSetAttribute(gf_pass_b<SC, GC, TI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            once.done[dev & 63] = true;
        }
    }
    const int max_twa = NT - 2 * a.r;
    const int strips = (a.w + max_twa - 1) / max_twa;
    a.twa = (a.w + strips - 1) / strips;
    // split rows only when the grid would leave most SMs idle; each segment pays 2r rows of warm-up
    int segs = 1;
    const long ctas = (long)strips * a.n;
    const int sms = sm_count();
    if (ctas < sms) {
        segs = (int)((sms + ctas - 1) / ctas);
        const int max_segs = a.h / (2 * a.r + 1) > 1 ? a.h / (2 * a.r + 1) : 1;
        if (segs > max_segs) segs = max_segs;
    }
    a.seg_rows = (a.h + segs - 1) / segs;
    dim3 grid(strips, (a.h + a.seg_rows - 1) / a.seg_rows, a.n);
    if constexpr (F32) {
        gf_pass_a<SC, GC, double, float><<<grid, NT, smem_a, st>>>(a);
    } else {
        if (wide)
            gf_pass_a<SC, GC, unsigned long long, uint8_t><<<grid, NT, smem_a, st>>>(a);
        else
            gf_pass_a<SC, GC, uint32_t, uint8_t><<<grid, NT, smem_a, st>>>(a);
    }
    RF_LAUNCH_CHECK("gf_pass_a");
    gf_pass_b<SC, GC, TI><<<grid, NT, smem_b, st>>>(a);
    RF_LAUNCH_CHECK("gf_pass_b");
    return RF_OK;
}

}  // namespace gf
}  // namespace rf

namespace rf {
namespace gf2 {  // gf2.cu: strip kernels for radius <= 64
bool supported(int r, int h, int w);
size_t workspace_per_image(int sc, int h, int w, int r, int iterations);
int run(const uint8_t *guide, const uint8_t *src, int sc, uint8_t *dst, void *ws, int n, int h, int w, int r,
        double eps, int iterations, cudaStream_t st);
}  // namespace gf2
}  // namespace rf

using namespace rf;

// RF_GF_GENERIC=1 in the environment forces the generic (any radius, FP64 box mean) kernels: used by the
// tests to cross-check the two implementations against each other.
static int flags_generic_path()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("RF_GF_GENERIC");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v;
}

// workspace one image needs: the larger of the two implementations' planes; iterated calls add the guide
// statistics (fast path) or one uint8 image to ping-pong through (generic path)
static size_t per_image_bytes(int sc, int h, int w, int radius, int iterations)
{
    size_t per = gf::per_image_ws(sc, h, w) + (iterations > 1 ? (((size_t)h * w * sc + 15) & ~(size_t)15) : 0);
    if (gf2::supported(radius, h, w)) {
        const size_t p2 = gf2::workspace_per_image(sc, h, w, radius, iterations);
        if (p2 > per) per = p2;
    }
    return per;
}

static size_t workspace_bytes(int sc, int n, int h, int w, int radius, int iterations)
{
    if (!(sc == 1 || sc == 3) || n < 1 || h < 1 || w < 1 || iterations < 1) return 0;
    // the call processes the batch in chunks if given less; never ask for more than 8 GiB
    const size_t per = per_image_bytes(sc, h, w, radius, iterations);
    size_t want = per * (size_t)n;
    const size_t cap = (size_t)8 << 30;
    if (want > cap) want = (cap / per > 0 ? cap / per : 1) * per;
    return want;
}

static int guided_impl(const char *fn, const uint8_t *guide, int gc, const uint8_t *src, int sc, uint8_t *dst, int n,
                       int h, int w, int radius, double eps, int iterations, void *ws, size_t ws_bytes, void *stream)
{
    if (!guide || !src || !dst || !ws) return fail(RF_EINVAL, "%s: NULL pointer", fn);
    if (!(gc == 1 || gc == 3)) return fail(RF_EINVAL, "%s: guide channels must be 1 or 3 (got %d)", fn, gc);
    if (!(sc == 1 || sc == 3)) return fail(RF_EINVAL, "%s: src channels must be 1 or 3 (got %d)", fn, sc);
    if (n < 0 || h < 1 || w < 1) return fail(RF_EINVAL, "%s: bad shape n=%d h=%d w=%d", fn, n, h, w);
    if (radius < 0) return fail(RF_EINVAL, "%s: negative radius", fn);
    if (iterations < 1) return fail(RF_EINVAL, "%s: iterations must be >= 1", fn);
    if (radius > gf::MAX_RADIUS)
        return fail(RF_EUNSUPPORTED, "%s: radius %d exceeds the supported maximum %d", fn, radius, gf::MAX_RADIUS);
    if (n == 0) return RF_OK;
    if (dst == src || dst == guide) return fail(RF_EINVAL, "%s: dst must not alias an input", fn);
    const size_t per = per_image_bytes(sc, h, w, radius, iterations);
    if (ws_bytes < per) return fail(RF_EINVAL, "%s: workspace too small (%zu < %zu bytes)", fn, ws_bytes, per);
    if ((uintptr_t)ws % 16) return fail(RF_EINVAL, "%s: workspace must be 16-byte aligned", fn);
    const int k = 2 * radius + 1;
    int chunk = (int)(ws_bytes / per < (size_t)n ? ws_bytes / per : (size_t)n);
    if (chunk > 65535) chunk = 65535;
    const size_t img_px = (size_t)h * w;
    // 1-channel guides (not reachable from the reference CLI, which reads 3 channels) take the generic kernels
    const bool generic = gc == 1 || !gf2::supported(radius, h, w) || (flags_generic_path() != 0);
    for (int i0 = 0; i0 < n; i0 += chunk) {
        const int nn = n - i0 < chunk ? n - i0 : chunk;
        if (!generic) {
            int rc = gf2::run(guide + i0 * img_px * 3, src + i0 * img_px * sc, sc, dst + i0 * img_px * sc, ws, nn, h, w,
                              radius, eps, iterations, (cudaStream_t)stream);
            if (rc != RF_OK) return rc;
            continue;
        }
        // generic kernels: every iteration is a full filter; intermediate images alternate between a scratch
        // image at the end of the workspace and dst, arranged so that the last one lands in dst
        uint8_t *tmp = (uint8_t *)ws + gf::per_image_ws(sc, h, w) * (size_t)nn;
        const uint8_t *in = src + i0 * img_px * sc;
        for (int it = 0; it < iterations; ++it) {
            uint8_t *out = ((iterations - 1 - it) & 1) ? tmp : dst + i0 * img_px * sc;
            gf::Args a;
            a.n = nn;
            a.guide = guide + i0 * img_px * gc;
            a.src = in;
            a.dst = out;
            a.ab = (float4 *)ws;
            a.h = h;
            a.w = w;
            a.r = radius;
            a.eps = (float)eps;
            a.scale = 1.0 / ((double)k * k);
            a.twa = 0;
            a.seg_rows = 0;
            int rc = gc == 3 ? (sc == 1 ? gf::run<1, 3>(a, (cudaStream_t)stream) : gf::run<3, 3>(a, (cudaStream_t)stream))
                             : (sc == 1 ? gf::run<1, 1>(a, (cudaStream_t)stream) : gf::run<3, 1>(a, (cudaStream_t)stream));
            if (rc != RF_OK) return rc;
            in = out;
        }
    }
    return RF_OK;
}

extern "C" int rf_guided_max_radius(void) { return gf::MAX_RADIUS; }

extern "C" size_t rf_guided_workspace_bytes(int sc, int n, int h, int w, int radius)
{
    return workspace_bytes(sc, n, h, w, radius, 1);
}

extern "C" int rf_guided_u8(const uint8_t *guide, int gc, const uint8_t *src, int sc, uint8_t *dst, int n, int h,
                            int w, int radius, double eps, void *ws, size_t ws_bytes, void *stream)
{
    return guided_impl("rf_guided_u8", guide, gc, src, sc, dst, n, h, w, radius, eps, 1, ws, ws_bytes, stream);
}

extern "C" size_t rf_guided_iterated_workspace_bytes(int sc, int n, int h, int w, int radius, int iterations)
{
    return workspace_bytes(sc, n, h, w, radius, iterations);
}

extern "C" int rf_guided_iterated_u8(const uint8_t *guide, int gc, const uint8_t *src, int sc, uint8_t *dst, int n,
                                     int h, int w, int radius, double eps, int iterations, void *ws, size_t ws_bytes,
                                     void *stream)
{
    return guided_impl("rf_guided_iterated_u8", guide, gc, src, sc, dst, n, h, w, radius, eps, iterations, ws, ws_bytes,
                       stream);
}

// ---- CV_32F images (the rest of cv2.ximgproc.guidedFilter's surface; not reachable from the reference CLI):
// float guide and source (values as they are, no scaling), float result.  Generic kernels only: float products, double
// window sums, the same solve.
extern "C" size_t rf_guided_f32_workspace_bytes(int sc, int n, int h, int w)
{
    if (!(sc == 1 || sc == 3) || n < 1 || h < 1 || w < 1) return 0;
    const size_t per = gf::per_image_ws(sc, h, w);
    size_t want = per * (size_t)n;
    const size_t cap = (size_t)8 << 30;
    if (want > cap) want = (cap / per > 0 ? cap / per : 1) * per;
    return want;
}

extern "C" int rf_guided_f32(const float *guide, int gc, const float *src, int sc, float *dst, int n, int h, int w,
                             int radius, double eps, void *ws, size_t ws_bytes, void *stream)
{
    const char *fn = "rf_guided_f32";
    if (!guide || !src || !dst || !ws) return fail(RF_EINVAL, "%s: NULL pointer", fn);
    if (!(gc == 1 || gc == 3)) return fail(RF_EINVAL, "%s: guide channels must be 1 or 3 (got %d)", fn, gc);
    if (!(sc == 1 || sc == 3)) return fail(RF_EINVAL, "%s: src channels must be 1 or 3 (got %d)", fn, sc);
    if (n < 0 || h < 1 || w < 1) return fail(RF_EINVAL, "%s: bad shape n=%d h=%d w=%d", fn, n, h, w);
    if (radius < 0) return fail(RF_EINVAL, "%s: negative radius", fn);
    if (radius > gf::MAX_RADIUS)
        return fail(RF_EUNSUPPORTED, "%s: radius %d exceeds the supported maximum %d", fn, radius, gf::MAX_RADIUS);
    if (n == 0) return RF_OK;
    if (dst == src || dst == guide) return fail(RF_EINVAL, "%s: dst must not alias an input", fn);
    const size_t per = gf::per_image_ws(sc, h, w);
    if (ws_bytes < per) return fail(RF_EINVAL, "%s: workspace too small (%zu < %zu bytes)", fn, ws_bytes, per);
    if ((uintptr_t)ws % 16) return fail(RF_EINVAL, "%s: workspace must be 16-byte aligned", fn);
    int chunk = (int)(ws_bytes / per < (size_t)n ? ws_bytes / per : (size_t)n);
    if (chunk > 65535) chunk = 65535;
    const size_t img_px = (size_t)h * w;
    const int k = 2 * radius + 1;
    for (int i0 = 0; i0 < n; i0 += chunk) {
        gf::Args a;
        a.n = n - i0 < chunk ? n - i0 : chunk;
        a.guide = guide + i0 * img_px * gc;
        a.src = src + i0 * img_px * sc;
        a.dst = dst + i0 * img_px * sc;
        a.ab = (float4 *)ws;
        a.h = h;
        a.w = w;
        a.r = radius;
        a.eps = (float)eps;
        a.scale = 1.0 / ((double)k * k);
        a.twa = 0;
        a.seg_rows = 0;
        cudaStream_t st = (cudaStream_t)stream;
        const int rc = gc == 3 ? (sc == 1 ? gf::run<1, 3, float>(a, st) : gf::run<3, 3, float>(a, st))
                               : (sc == 1 ? gf::run<1, 1, float>(a, st) : gf::run<3, 1, float>(a, st));
        if (rc != RF_OK) return rc;
    }
    return RF_OK;
}
