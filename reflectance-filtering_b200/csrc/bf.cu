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
#include <type_traits>
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
    const float *tab;  // device: (r+1) x 2 x tabw exponents, then (r+1) ints (row half widths, ceil4), then (r+1) exact ones
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

__device__ __forceinline__ void load_table(float *tab_s, const Args &a)
{
    const int n = (a.r + 1) * 2 * a.tabw + 2 * (a.r + 1);
    for (int i = threadIdx.x; i < n; i += blockDim.x) tab_s[i] = a.tab[i];
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
    const int *roww4 = reinterpret_cast<const int *>(tab + (a.r + 1) * 2 * a.tabw);

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
    const unsigned long long sc2 = pack2(a.ksqrt, a.ksqrt);
    const unsigned long long scb2 = pack2(-8388608.0f * a.ksqrt, -8388608.0f * a.ksqrt);
    const int r = a.r;
    for (int dyi = 0; dyi <= 2 * r; ++dyi) {
        const int ady = dyi < r ? r - dyi : dyi - r;
        const int w4 = roww4[ady];
        const int hw = roww4[a.r + 1 + ady];  // exact half width of this row of the disc
        const uint32_t *jrow = tj + (ty + dyi) * a.pitch + a.rpad + x0;
        const uint32_t *srow = ts + (ty + dyi) * a.pitch + a.rpad + x0;
        const float *trowA = tab + ady * 2 * a.tabw + a.rpad;  // E(dx), dx = index - 7 relative to qb
        const float *trowB = trowA + a.tabw;                   // the same row shifted by one entry
        // one group of four neighbours (qb .. qb+3) against the outputs [PLO, PHI).  The first group of a row
        // (qb = -w4) lies left of every window of outputs 4..7 and the last one (qb = P + w4 - 4) right of every
        // window of outputs 0..3 (their table entries are +inf = weight exactly 0), so those halves are skipped:
        // ~6 % fewer taps at r = 33, identical results.
        // EDGE 1 / 2 (the two quads at the left / right end of a row): a tap pair that lies outside the window of its
        // output for every lane -- dx = qb + jp - p beyond -hw / +hw, a warp-uniform test -- is skipped (weights exactly 0):
        // eight outputs share one neighbour range, which costs ~13 % of the taps at r = 33 without it.
        auto quad = [&](const int qb, auto plo_c, auto phi_c, auto edge_c) {
            constexpr int PLO = decltype(plo_c)::value, PHI = decltype(phi_c)::value, EDGE = decltype(edge_c)::value;
            const uint4 jn = *reinterpret_cast<const uint4 *>(jrow + qb);
            const uint4 sn = SEP ? *reinterpret_cast<const uint4 *>(srow + qb) : jn;
            const float4 a0 = *reinterpret_cast<const float4 *>(trowA + qb);
            const float4 a1 = *reinterpret_cast<const float4 *>(trowA + qb + 4);
            const float4 a2v = *reinterpret_cast<const float4 *>(trowA + qb + 8);
            const float4 b0 = *reinterpret_cast<const float4 *>(trowB + qb);
            const float4 b1 = *reinterpret_cast<const float4 *>(trowB + qb + 4);
            const float4 b2 = *reinterpret_cast<const float4 *>(trowB + qb + 8);
            const float TA[12] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2v.x, a2v.y, a2v.z, a2v.w};
            const float TB[12] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w, b2.x, b2.y, b2.z, b2.w};
            const uint32_t jv[4] = {jn.x, jn.y, jn.z, jn.w};
            const uint32_t sv[4] = {sn.x, sn.y, sn.z, sn.w};
            unsigned long long bg[4], r1[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                bg[j] = pack2(byte_to_float(sv[j], 0), byte_to_float(sv[j], 1));
                r1[j] = pack2(byte_to_float(sv[j], 2), byte_to_float(sv[j], 3));
            }
            // taps are paired over two adjacent neighbours (jp, jp+1) of the same output p: their spatial
            // exponents are adjacent table entries, so the range/space arithmetic is one packed FFMA2 each
#pragma unroll
            for (int jp = 0; jp < 4; jp += 2) {
#pragma unroll
                for (int p = PLO; p < PHI; ++p) {
                    if (EDGE == 1 && qb + hw < p - jp - 1) continue;  // both taps left of output p's window
                    if (EDGE == 2 && qb - hw > p - jp) continue;      // both taps right of it
                    uint32_t d0, d1;
                    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d0) : "r"(jc[p]), "r"(jv[jp]), "r"(0x4B000000u));
                    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d1) : "r"(jc[p]), "r"(jv[jp + 1]), "r"(0x4B000000u));
                    // d is the float 2^23 + alpha; the FMA gives RN(alpha * ksqrt) exactly as a multiply would
                    const unsigned long long u2 = ffma2r(pack2(__uint_as_float(d0), __uint_as_float(d1)), sc2, scb2);
                    const int i = jp - p + 7;  // table entry of tap (jp, p); tap (jp+1, p) is entry i + 1
                    const unsigned long long e2 = (i & 1) ? pack2(TB[i - 1], TB[i]) : pack2(TA[i], TA[i + 1]);
                    const unsigned long long x2 = ffma2r(u2, u2, e2);
                    float x0f, x1f;
                    unpack2(x2, x0f, x1f);
                    const float w0 = ex2_neg(x0f), w1 = ex2_neg(x1f);
                    const unsigned long long ww0 = pack2(w0, w0), ww1 = pack2(w1, w1);
                    ffma2(acc01[p], ww0, bg[jp]);
                    ffma2(acc23[p], ww0, r1[jp]);
                    ffma2(acc01[p], ww1, bg[jp + 1]);
                    ffma2(acc23[p], ww1, r1[jp + 1]);
                }
            }
        };
        using I0 = std::integral_constant<int, 0>;
        using I4 = std::integral_constant<int, P / 2>;
        using I8 = std::integral_constant<int, P>;
        using E0 = std::integral_constant<int, 0>;
        using EL = std::integral_constant<int, 1>;
        using ER = std::integral_constant<int, 2>;
        if (w4 >= 4) {
            quad(-w4, I0{}, I4{}, EL{});
            quad(-w4 + 4, I0{}, I8{}, EL{});
            for (int qb = -w4 + 8; qb < P + w4 - 8; qb += 4) quad(qb, I0{}, I8{}, E0{});
            quad(P + w4 - 8, I0{}, I8{}, ER{});
            quad(P + w4 - 4, I4{}, I8{}, ER{});
        } else {  // the one-pixel rows at the top and bottom of the disc: two quads in all
            quad(-w4, I0{}, I4{}, E0{});
            quad(P + w4 - 4, I4{}, I8{}, E0{});
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

// exponent table in the log2 domain, two copies per |dy| row:
//   A[ady][i] = (dx^2 + ady^2) * 0.5 / sigma_space^2 * log2(e), dx = i - rpad - 7, +inf outside the disc
//   B[ady][i] = A[ady][i + 1]   (so that any two adjacent entries are an aligned register pair in A or in B)
// followed by ceil4(half width) per |dy|.  The weight of a tap is exp2(-(alpha*ksqrt)^2 - A).
// Call with g_tab_mu held, and keep it until the kernel that reads the table is in the stream: an eviction cudaFree()s
// (which waits for running kernels) and must not hit a table between another thread's lookup and its launch.
static int get_table(double sigma_space, const Geometry &g, const float **out)
{
    int dev = 0;
    RF_CUDA_TRY(cudaGetDevice(&dev));
    for (const TableEntry &e : g_tabs)
        if (e.device == dev && e.sigma_space == sigma_space && e.r == g.r) {
            *out = e.d_tab;
            return RF_OK;
        }
    const int n = (g.r + 1) * 2 * g.tabw + 2 * (g.r + 1);
    std::vector<float> h(n);
    const double gsc = 0.5 / (sigma_space * sigma_space) * 1.4426950408889634074;
    auto E = [&](int i, int ady) -> float {
        const int dx = i - g.rpad - 7;
        const int d2 = dx * dx + ady * ady;
        return (i < g.tabw && d2 <= g.r * g.r) ? (float)(d2 * gsc) : INFINITY;
    };
    for (int ady = 0; ady <= g.r; ++ady) {
        for (int i = 0; i < g.tabw; ++i) {
            h[(size_t)ady * 2 * g.tabw + i] = E(i, ady);
            h[(size_t)ady * 2 * g.tabw + g.tabw + i] = E(i + 1, ady);
        }
        const int hw = (int)std::floor(std::sqrt((double)(g.r * g.r - ady * ady)));
        reinterpret_cast<int *>(h.data())[(g.r + 1) * 2 * g.tabw + ady] = ceil4(hw);
        reinterpret_cast<int *>(h.data())[(g.r + 1) * 2 * g.tabw + (g.r + 1) + ady] = hw;
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
    return tile * (sep ? 2 : 1) + ((size_t)(g.r + 1) * 2 * g.tabw + 2 * (g.r + 1)) * 4;
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
    // CTA = 32 x 8*WY pixels.  A warp's work is fixed (32 x 8 pixels), an SM runs its resident warps concurrently
    // and saturates at about a dozen of them, and shared memory (window + tables) decides how many CTAs fit.
    // Cost model: waves x max(1, resident warps / 12); ties go to the variant with more resident warps
    // (tall CTAs amortise the halo), then to the one that spreads over more SMs.
    const long sms = sm_count();
    const long cols = (a.w + TW - 1) / TW;
    int best = 1;
    double best_cost = 1e30;
    long best_warps = 0;
    for (int wy = 1; wy <= 16; wy <<= 1) {
        const size_t smem = smem_bytes(wy, g, sep);
        if (smem > 227 * 1024) break;
        long res = (long)((227 * 1024) / smem);
        if (res * wy > 48) res = 48 / wy > 0 ? 48 / wy : 1;
        const long ctas = cols * ((a.h + ROWS_PER_WARP * wy - 1) / (ROWS_PER_WARP * wy)) * a.n;
        const long waves = (ctas + sms * res - 1) / (sms * res);
        const long resident = res * wy < (ctas * wy + sms - 1) / sms ? res * wy : (ctas * wy + sms - 1) / sms;
        const double cost = (double)waves * (resident > 12 ? resident / 12.0 : 1.0);
        if (cost < best_cost * 0.97 || (cost < best_cost * 1.03 && resident > best_warps)) {
            best_cost = cost < best_cost ? cost : best_cost;
            best_warps = resident;
            best = wy;
        }
    }
    return best;
}

}  // namespace bf
}  // namespace rf

namespace rf {
namespace bf2 {  // bf2.cu: packed two-output kernel for single-channel joint + src
int run(const uint8_t *joint, const uint8_t *src, uint8_t *dst, int n, int h, int w, int r, double sigma_color,
        double sigma_space, double alpha_scale, cudaStream_t st);
}
}  // namespace rf

namespace rf {
namespace bfg {  // bf_generic.cu: any radius, any border type, 8-bit and CV_32F
int run_u8(const uint8_t *joint, int jc, const uint8_t *src, int sc, uint8_t *dst, int n, int h, int w,
           double sigma_color, double sigma_space, int d, int alpha_scale, int border, cudaStream_t st);
int run_f32(const float *joint, int jc, const float *src, int sc, float *dst, int n, int h, int w, double sigma_color,
            double sigma_space, int d, int border, void *ws, size_t ws_bytes, cudaStream_t st);
size_t workspace_f32(int n, int jc);
extern const int MAX_RADIUS_PUBLIC;
}
}  // namespace rf

using namespace rf;

extern "C" int rf_joint_bilateral_max_radius(void) { return bfg::MAX_RADIUS_PUBLIC; }
extern "C" int rf_joint_bilateral_fast_max_radius(void) { return bf::MAX_RADIUS; }

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
    return rf_joint_bilateral_u8_border(joint, jc, src, sc, dst, n, h, w, sigma_color, sigma_space, d, flags,
                                        RF_BORDER_REFLECT_101, stream);
}

extern "C" size_t rf_joint_bilateral_f32_workspace_bytes(int n, int jc)
{
    if (n < 0 || !(jc == 1 || jc == 3)) return 0;
    return bfg::workspace_f32(n, jc);
}

extern "C" int rf_joint_bilateral_f32(const float *joint, int jc, const float *src, int sc, float *dst, int n, int h,
                                      int w, double sigma_color, double sigma_space, int d, int border_type,
                                      void *workspace, size_t workspace_bytes, void *stream)
{
    if (!joint || !src || !dst) return fail(RF_EINVAL, "rf_joint_bilateral_f32: NULL image pointer");
    if (!(jc == 1 || jc == 3) || !(sc == 1 || sc == 3))
        return fail(RF_EINVAL, "rf_joint_bilateral_f32: channels must be 1 or 3 (joint %d, src %d)", jc, sc);
    if (n < 0 || h < 1 || w < 1) return fail(RF_EINVAL, "rf_joint_bilateral_f32: bad shape n=%d h=%d w=%d", n, h, w);
    if (border_type < 0 || border_type > 4) return fail(RF_EINVAL, "rf_joint_bilateral_f32: border type %d", border_type);
    if (n == 0) return RF_OK;
    if (dst == joint || dst == src) return fail(RF_EINVAL, "rf_joint_bilateral_f32: dst must not alias an input");
    if (sigma_color <= 0) sigma_color = 1;
    if (sigma_space <= 0) sigma_space = 1;
    return bfg::run_f32(joint, jc, src, sc, dst, n, h, w, sigma_color, sigma_space, d, border_type, workspace,
                        workspace_bytes, (cudaStream_t)stream);
}

extern "C" int rf_joint_bilateral_u8_border(const uint8_t *joint, int jc, const uint8_t *src, int sc, uint8_t *dst,
                                            int n, int h, int w, double sigma_color, double sigma_space, int d,
                                            unsigned flags, int border_type, void *stream)
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
    if (border_type < 0 || border_type > 4) return fail(RF_EINVAL, "rf_joint_bilateral_u8: border type %d", border_type);
    cudaStream_t st0 = (cudaStream_t)stream;
    const int alpha_scale = gray_rep ? 3 : 1;
    // the tiled kernels hold the REFLECT_101 window of a tile in shared memory: other borders, radii beyond
    // rf_joint_bilateral_fast_max_radius() (or beyond what fits for this channel combination) and batches beyond the
    // grid limit go through the generic kernel (bf_generic.cu), which is bit-equal to the CPU restatement
    {
        int r0 = d <= 0 ? (int)std::nearbyint(sigma_space * 1.5) : d / 2;
        r0 = r0 < 1 ? 1 : r0;
        if (border_type != RF_BORDER_REFLECT_101 || r0 > bf::MAX_RADIUS || n > 65535)
            return bfg::run_u8(joint, jc, src, sc, dst, n, h, w, sigma_color, sigma_space, d, alpha_scale, border_type, st0);
    }
    const bf::Geometry g = bf::geometry(sigma_space, d);

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
    const bool sep = joint != src;
    cudaStream_t st = (cudaStream_t)stream;
    const double ksq = std::sqrt(0.5 / (sigma_color * sigma_color) * 1.4426950408889634074);
    const bool gray = (jc == 1 && sc == 1);
    if (gray) {
        const int rc = bf2::run(joint, src, dst, n, h, w, g.r, sigma_color, sigma_space, gray_rep ? 3.0 : 1.0,
                                (cudaStream_t)stream);
        if (rc != RF_EUNSUPPORTED) return rc;
        return bfg::run_u8(joint, jc, src, sc, dst, n, h, w, sigma_color, sigma_space, d, alpha_scale, border_type, st0);
    }
    {
        a.ksqrt = (float)ksq;
        const int wy = bf::pick_wy(a, g, sep);
        const size_t smem = bf::smem_bytes(wy, g, sep);
        if (smem > 227 * 1024)  // r = 61..64 with a distinct colour joint: the two tiles do not fit
            return bfg::run_u8(joint, jc, src, sc, dst, n, h, w, sigma_color, sigma_space, d, alpha_scale, border_type, st0);
        std::lock_guard<std::mutex> lk(bf::g_tab_mu);  // held until the kernel that reads the table is in the stream
        const int rc = bf::get_table(sigma_space, g, &a.tab);
        if (rc != RF_OK) return rc;
#define RF_BF_COLOR(WY)                                                                                  \
    case WY:                                                                                             \
        return sep ? bf::launch(bf::bf_color_kernel<WY, true>, WY, a, smem, st, "bf_color_kernel")       \
                   : bf::launch(bf::bf_color_kernel<WY, false>, WY, a, smem, st, "bf_color_kernel")
        switch (wy) {
            RF_BF_COLOR(1);
            RF_BF_COLOR(2);
            RF_BF_COLOR(4);
            RF_BF_COLOR(8);
            RF_BF_COLOR(16);
        }
#undef RF_BF_COLOR
    }
    return fail(RF_EINVAL, "rf_joint_bilateral_u8: internal dispatch error");
}
