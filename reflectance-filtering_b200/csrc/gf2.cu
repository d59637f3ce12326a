// Guided filter, fast path for radius <= 64 (and images at least one halo wide): strip kernels with
// per-lane column chunks over column-padded planes.
//
// Same semantics as gf.cu (ximgproc guidedFilter; call site
// /root/reference/filter_reflectance.py:67-70).  The work decomposition follows the measured pipe rates
// on B200 (profiles/r01_microbench_pipes.txt): SHFL, I2F and anything FP64 issue at 1/4 to 1/8 of the FP32
// rate, byte loads with per-pixel border logic are poison, so
//   * a pack kernel writes guide + source once as 32-bit pixels (B | G<<8 | R<<16 | p<<24) into planes
//     that are padded by the halo with BORDER_REFLECT columns already resolved: every later load is an
//     aligned 128-bit load of four pixels, with no border case in x (rows reflect by index);
//   * every lane owns C consecutive columns of a 32*C wide strip and keeps the vertical window sums of
//     its quantities in FP32 registers -- integers below 2^23 ((2r+1)*255^2 for r <= 64), so the FFMA
//     updates (add the entering row, subtract the leaving one) are exact;
//   * the horizontal direction is a per-lane serial prefix over its C columns plus ONE 5-step warp scan
//     of the lane totals per quantity (uint32, wrap-around safe): 5/C shuffles per column;
//   * the 9 + 4*SC quantities are split over the warps of the CTA (each warp scans the whole strip for
//     its 3-4 quantities); prefixes go to a double-buffered shared-memory row, and after ONE
//     __syncthreads per row all threads turn P[x+r] - P[x-r-1] into means, solve the 3x3 system with
//     the oracle's non-contracted multiply/add order and store the coefficient planes (again padded,
//     border pixels mirrored into the halo so pass B has no border case either);
//   * no FP64: the box mean is float(S) * float(1/k^2) (<= 1 ulp from OpenCV's float(double(S)/k^2);
//     measured effect ~2.5e-5 of the output bytes move by 1 LSB); pass B accumulates the
//     coefficient planes in FP32 the same way;
//   * iterated filtering with one guide (createGuidedFilter(guide, r, eps) reused for several filter() calls,
//     SURVEY 8f-4; the "3 x GF" configuration): the first pass A also stores mean(I) and the inverse
//     covariance per pixel, later iterations only accumulate the 4*SC source quantities and read those nine
//     floats back; pass B writes its uint8 output straight into the packed planes, so later iterations need
//     no pack kernel.  The arithmetic per pixel is the same function in both modes: results are byte-identical
//     to repeated single calls.
#include <cstdlib>

#include "common.cuh"

namespace rf {
namespace gf2 {

constexpr int MAX_RADIUS = 64;  // (2r+1) * 65025 < 2^23
constexpr int QMAX = 21;

struct Args {
    const uint8_t *guide;  // [n][h][w][3]
    const uint8_t *src;    // [n][h][w][SC]
    uint32_t *packed;      // [n][NP][h][wp]  NP = 1 (SC == 1: B,G,R,p) or 2 (SC == 3: B,G,R,0 / p0,p1,p2,0)
    float *ab;             // [n][SC][4][h][wp]  (a0, a1, a2, b), columns padded like `packed`
    float *gstat;          // [n][9][h][wp]  guide statistics kept for iterated filtering: mean I (3) and the
                           // inverse of cov(I) + eps*Id (00, 01, 02, 11, 12, 22); NULL when not wanted
    uint8_t *dst;          // [n][h][w][SC]
    int store_dst;         // pass B writes dst (last iteration)
    int store_packed;      // pass B writes its output into the source bytes of `packed` (input of the next iteration)
    int n, h, w, r;
    int rh;          // halo columns on each side: round_up(r + 1, 4)
    int wp;          // padded row pitch in pixels: round_up(w + 2 * rh + 16, 4)
    int twe;         // output columns per strip (multiple of 4)
    int seg_rows;    // output rows per CTA
    float eps;
    float inv_area;  // 1 / (2r+1)^2
};

__device__ __forceinline__ float b2f(uint32_t word, int byte)
{
    return __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7440u | (uint32_t)byte)) - 8388608.0f;
}

// ---- pack: u8 interleaved -> padded 32-bit pixels ------------------------------------------------
constexpr int PACK_ROWS = 8;  // rows per thread: enough independent byte loads in flight, 8x fewer CTAs
template <int SC>
__global__ void pack_kernel(const Args g)
{
    const int xp = blockIdx.x * blockDim.x + threadIdx.x;
    const int img = blockIdx.z;
    if (xp >= g.wp) return;
    const int x = reflect(xp - g.rh, g.w);
    const size_t img_px = (size_t)g.h * g.w;
    const size_t plane = (size_t)g.h * g.wp;
    constexpr int NP = SC == 1 ? 1 : 2;
    const int ya = blockIdx.y * PACK_ROWS;
    uint32_t bgr[PACK_ROWS], sw[PACK_ROWS];
#pragma unroll
    for (int k = 0; k < PACK_ROWS; ++k) {
        const int y = min(ya + k, g.h - 1);
        const uint8_t *gp = g.guide + (img * img_px + (size_t)y * g.w + x) * 3;
        const uint8_t *sp = g.src + (img * img_px + (size_t)y * g.w + x) * SC;
        bgr[k] = gp[0] | ((uint32_t)gp[1] << 8) | ((uint32_t)gp[2] << 16);
        sw[k] = SC == 1 ? (uint32_t)sp[0] << 24 : sp[0] | ((uint32_t)sp[1 % SC] << 8) | ((uint32_t)sp[2 % SC] << 16);
    }
#pragma unroll
    for (int k = 0; k < PACK_ROWS; ++k) {
        const int y = ya + k;
        if (y >= g.h) break;
        uint32_t *o = g.packed + (size_t)img * NP * plane + (size_t)y * g.wp + xp;
        if (SC == 1) {
            o[0] = bgr[k] | sw[k];
        } else {
            o[0] = bgr[k];
            o[plane] = sw[k];
        }
    }
}

// channel ids: 0..2 guide, 3..5 source, 6 = the constant 1.  Quantity q is ch[qa(q)] * ch[qb(q)].
// Order = the m[] layout of the solve: I (3), I*I' (6), then per source channel p, p*I0, p*I1, p*I2.
__host__ __device__ constexpr int qa(int q)
{
    constexpr int t[QMAX] = {0, 1, 2, 0, 0, 0, 1, 1, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5};
    return t[q];
}
__host__ __device__ constexpr int qb(int q)
{
    constexpr int t[QMAX] = {6, 6, 6, 0, 1, 2, 1, 2, 2, 6, 0, 1, 2, 6, 0, 1, 2, 6, 0, 1, 2};
    return t[q];
}
__host__ __device__ constexpr bool q_is_linear(int q) { return qb(q) == 6; }

constexpr int NQG = 4;  // most quantities any warp owns

// one padded row's chunk of this lane (C packed pixels per plane), in flight or landed
template <int SC, int C>
struct RowRaw {
    uint32_t w0[C];
    uint32_t w1[SC == 1 ? 1 : C];
};

template <int SC, int C>
__device__ __forceinline__ RowRaw<SC, C> prefetch_row(const uint32_t *plane0, size_t plane_stride, int wp, int yy, int xp0)
{
    RowRaw<SC, C> r;
    const uint4 *p = reinterpret_cast<const uint4 *>(plane0 + (size_t)yy * wp + xp0);
#pragma unroll
    for (int c = 0; c < C; c += 4) {
        const uint4 t = __ldg(p + c / 4);
        r.w0[c] = t.x;
        r.w0[c + 1] = t.y;
        r.w0[c + 2] = t.z;
        r.w0[c + 3] = t.w;
    }
    if (SC == 3) {
        const uint4 *q = reinterpret_cast<const uint4 *>(plane0 + plane_stride + (size_t)yy * wp + xp0);
#pragma unroll
        for (int c = 0; c < (SC == 1 ? 0 : C); c += 4) {
            const uint4 t = __ldg(q + c / 4);
            r.w1[c] = t.x;
            r.w1[c + 1] = t.y;
            r.w1[c + 2] = t.z;
            r.w1[c + 3] = t.w;
        }
    } else {
        r.w1[0] = 0u;
    }
    return r;
}

// (byte of pixel c, byte of pixel c+1) -> packed float pair, optionally negated: PRMT builds 2^23 + b in each half,
// one packed add (or fused negate-add) removes the bias for both
__device__ __forceinline__ unsigned long long b2f2(uint32_t w0, uint32_t w1, int byte, bool negate)
{
    const uint32_t sel = 0x7440u | (uint32_t)byte;
    const unsigned long long p = pack2(__uint_as_float(__byte_perm(w0, 0x4B000000u, sel)),
                                       __uint_as_float(__byte_perm(w1, 0x4B000000u, sel)));
    unsigned long long r;
    if (negate)
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(p), "l"(pack2(-1.0f, -1.0f)), "l"(pack2(8388608.0f, 8388608.0f)));
    else
        asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(p), "l"(pack2(-8388608.0f, -8388608.0f)));
    return r;
}

// Adds (sign = +1) or removes (sign = -1) one row: V[q][c] += sign * ch[qa] * ch[qb].  Two adjacent columns per packed
// instruction (FADD2 / FFMA2 halve the issue slots of the conversion and of the accumulation; the FMA pipe does the
// same work); the sign rides on the first factor, which is converted negated when a row leaves.  Exact: all values are
// integers below 2^23.
template <int SC, int C, int Q0, int NQ, int QBASE = 0>
__device__ __forceinline__ void accumulate(float (&V)[NQG][C], const RowRaw<SC, C> &row, float sign)
{
    static_assert(C % 2 == 0, "column pairs");
    const bool neg = sign < 0.0f;
#pragma unroll
    for (int c = 0; c < C; c += 2) {
        // which channels this group needs as first factor (signed) and as second factor (plain)
        bool need_a[6] = {false, false, false, false, false, false}, need_b[6] = {false, false, false, false, false, false};
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            need_a[qa(Q0 + q)] = true;
            if (qb(Q0 + q) != 6) need_b[qb(Q0 + q)] = true;
        }
        unsigned long long sa[6], pb[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const uint32_t w0 = k < 3 ? row.w0[c] : row.w1[SC == 1 ? 0 : c];
            const uint32_t w1 = k < 3 ? row.w0[c + 1] : row.w1[SC == 1 ? 0 : c + 1];
            const int byte = SC == 1 ? (k < 3 ? k : 3) : (k < 3 ? k : k - 3);
            const bool live = SC == 3 || k <= 3;  // SC == 1: channels 4, 5 do not exist
            const uint32_t x0 = (SC == 1 && k == 3) ? row.w0[c] : w0, x1 = (SC == 1 && k == 3) ? row.w0[c + 1] : w1;
            sa[k] = pb[k] = 0ull;
            if (live && need_a[k]) sa[k] = b2f2(x0, x1, byte, neg);
            if (live && need_b[k]) pb[k] = (need_a[k] && !neg) ? sa[k] : b2f2(x0, x1, byte, false);
        }
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            unsigned long long v = pack2(V[q][c], V[q][c + 1]);
            if (qb(Q0 + q) == 6)
                asm("add.rn.f32x2 %0, %1, %2;" : "=l"(v) : "l"(v), "l"(sa[qa(Q0 + q)]));
            else
                asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(v) : "l"(sa[qa(Q0 + q)]), "l"(pb[qb(Q0 + q)]));
            unpack2(v, V[q][c], V[q][c + 1]);
        }
    }
}

// per-lane serial prefix + warp scan of the lane totals; inclusive strip-wide prefixes to shared memory
template <int SC, int C, int Q0, int NQ, int QBASE = 0>
__device__ __forceinline__ void scan_store(const float (&V)[NQG][C], uint32_t *P, int nx, int lane)
{
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        uint32_t pre[C];
        uint32_t run = 0;
#pragma unroll
        for (int c = 0; c < C; c += 2) {
            // integer value + 0x4B000000 per element; the float add on two columns at once
            unsigned long long b2;
            asm("add.rn.f32x2 %0, %1, %2;" : "=l"(b2) : "l"(pack2(V[q][c], V[q][c + 1])), "l"(pack2(8388608.0f, 8388608.0f)));
            float b0, b1;
            unpack2(b2, b0, b1);
            run += __float_as_uint(b0);
            pre[c] = run;
            run += __float_as_uint(b1);
            pre[c + 1] = run;
        }
        uint32_t incl = run;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        const uint32_t excl = incl - run;
        uint32_t *dst = P + (Q0 + q - QBASE) * nx + lane * C;
#pragma unroll
        for (int c = 0; c < C; c += 4)
            *reinterpret_cast<uint4 *>(dst + c) =
                make_uint4(pre[c] + excl, pre[c + 1] + excl, pre[c + 2] + excl, pre[c + 3] + excl);
    }
}

// warp-uniform dispatch of the per-group templates (gray: 4 warps, colour source: 6 warps)
#define RF_GF2_DISPATCH(FN, ...)                              \
    if (SC == 1) {                                            \
        switch (group) {                                      \
            case 0: FN<SC, C, 0, 3>(__VA_ARGS__); break;      \
            case 1: FN<SC, C, 3, 3>(__VA_ARGS__); break;      \
            case 2: FN<SC, C, 6, 3>(__VA_ARGS__); break;      \
            default: FN<SC, C, 9, 4>(__VA_ARGS__); break;     \
        }                                                     \
    } else {                                                  \
        switch (group) {                                      \
            case 0: FN<SC, C, 0, 3>(__VA_ARGS__); break;      \
            case 1: FN<SC, C, 3, 3>(__VA_ARGS__); break;      \
            case 2: FN<SC, C, 6, 3>(__VA_ARGS__); break;      \
            case 3: FN<SC, C, 9, 4>(__VA_ARGS__); break;      \
            case 4: FN<SC, C, 13, 4>(__VA_ARGS__); break;     \
            default: FN<SC, C, 17, 4>(__VA_ARGS__); break;    \
        }                                                     \
    }

// iterations >= 2: only the source quantities 9 .. 9+4*SC-1, spread over the same number of warps
#define RF_GF2_DISPATCH_SRC(FN, ...)                          \
    if (SC == 1) {                                            \
        switch (group) {                                      \
            case 0: FN<SC, C, 9, 1, 9>(__VA_ARGS__); break;   \
            case 1: FN<SC, C, 10, 1, 9>(__VA_ARGS__); break;  \
            case 2: FN<SC, C, 11, 1, 9>(__VA_ARGS__); break;  \
            default: FN<SC, C, 12, 1, 9>(__VA_ARGS__); break; \
        }                                                     \
    } else {                                                  \
        switch (group) {                                      \
            case 0: FN<SC, C, 9, 2, 9>(__VA_ARGS__); break;   \
            case 1: FN<SC, C, 11, 2, 9>(__VA_ARGS__); break;  \
            case 2: FN<SC, C, 13, 2, 9>(__VA_ARGS__); break;  \
            case 3: FN<SC, C, 15, 2, 9>(__VA_ARGS__); break;  \
            case 4: FN<SC, C, 17, 2, 9>(__VA_ARGS__); break;  \
            default: FN<SC, C, 19, 2, 9>(__VA_ARGS__); break; \
        }                                                     \
    }

template <int SC>
__host__ __device__ constexpr int n_groups() { return SC == 1 ? 4 : 6; }

enum Mode { FULL = 0, FULL_STORE = 1, SRC_ONLY = 2 };

// window sum -> box mean.  Sums of single channels stay below 2^23: exact integer->float on the FMA pipe.
__device__ __forceinline__ float box_mean(uint32_t s, bool linear, float inv_area)
{
    const float sf = linear ? __uint_as_float(s | 0x4B000000u) - 8388608.0f : (float)s;
    return __fmul_rn(sf, inv_area);
}

// inverse of cov(I) + eps*Id from the nine guide means m[0..8]; out: inv 00, 01, 02, 11, 12, 22
__device__ __forceinline__ void guide_inverse(const float *m, float eps, float *inv)
{
    // cov(I) + eps*Id, symmetric storage 0:(0,0) 1:(0,1) 2:(0,2) 3:(1,1) 4:(1,2) 5:(2,2)
    const float c00 = __fadd_rn(__fsub_rn(m[3], __fmul_rn(m[0], m[0])), eps);
    const float c01 = __fsub_rn(m[4], __fmul_rn(m[0], m[1]));
    const float c02 = __fsub_rn(m[5], __fmul_rn(m[0], m[2]));
    const float c11 = __fadd_rn(__fsub_rn(m[6], __fmul_rn(m[1], m[1])), eps);
    const float c12 = __fsub_rn(m[7], __fmul_rn(m[1], m[2]));
    const float c22 = __fadd_rn(__fsub_rn(m[8], __fmul_rn(m[2], m[2])), eps);
    const float f00 = __fsub_rn(__fmul_rn(c11, c22), __fmul_rn(c12, c12));
    const float f01 = __fsub_rn(__fmul_rn(c12, c02), __fmul_rn(c01, c22));
    const float f02 = __fsub_rn(__fmul_rn(c01, c12), __fmul_rn(c11, c02));
    const float f11 = __fsub_rn(__fmul_rn(c22, c00), __fmul_rn(c02, c02));
    const float f12 = __fsub_rn(__fmul_rn(c02, c01), __fmul_rn(c12, c00));
    const float f22 = __fsub_rn(__fmul_rn(c00, c11), __fmul_rn(c01, c01));
    float det = __fmul_rn(c00, f00);
    det = __fadd_rn(det, __fmul_rn(c01, f01));
    det = __fadd_rn(det, __fmul_rn(c02, f02));
    if (eps < 1e-2f && fabsf(det) < 1e-6f) det = 1e-6f;
    // one correctly rounded reciprocal instead of six divisions: each entry is within 1 ulp of cof/det
    const float rdet = __frcp_rn(det);
    inv[0] = __fmul_rn(f00, rdet);
    inv[1] = __fmul_rn(f01, rdet);
    inv[2] = __fmul_rn(f02, rdet);
    inv[3] = __fmul_rn(f11, rdet);
    inv[4] = __fmul_rn(f12, rdet);
    inv[5] = __fmul_rn(f22, rdet);
}

// a = inv * (mean(I p) - mean(I) mean(p)),  b = mean(p) - a . mean(I);  ms = (mean p, mean p*I0, p*I1, p*I2)
__device__ __forceinline__ void solve_source(const float *ms, const float *mi, const float *inv, float *v)
{
    const float mp = ms[0];
    const float k0 = __fsub_rn(ms[1], __fmul_rn(mp, mi[0]));
    const float k1 = __fsub_rn(ms[2], __fmul_rn(mp, mi[1]));
    const float k2 = __fsub_rn(ms[3], __fmul_rn(mp, mi[2]));
    float a0 = __fmul_rn(inv[0], k0);
    a0 = __fadd_rn(a0, __fmul_rn(inv[1], k1));
    a0 = __fadd_rn(a0, __fmul_rn(inv[2], k2));
    float a1 = __fmul_rn(inv[1], k0);
    a1 = __fadd_rn(a1, __fmul_rn(inv[3], k1));
    a1 = __fadd_rn(a1, __fmul_rn(inv[4], k2));
    float a2 = __fmul_rn(inv[2], k0);
    a2 = __fadd_rn(a2, __fmul_rn(inv[4], k1));
    a2 = __fadd_rn(a2, __fmul_rn(inv[5], k2));
    float b = __fsub_rn(mp, __fmul_rn(a0, mi[0]));
    b = __fsub_rn(b, __fmul_rn(a1, mi[1]));
    b = __fsub_rn(b, __fmul_rn(a2, mi[2]));
    v[0] = a0;
    v[1] = a1;
    v[2] = a2;
    v[3] = b;
}

// ---- pass A ---------------------------------------------------------------------------------------
// MODE FULL: all 9 + 4*SC quantities.  FULL_STORE: the same, and the guide statistics go to g.gstat.
// SRC_ONLY: only the 4*SC source quantities; mean(I) and the inverse covariance come from g.gstat.
template <int SC, int C, int MODE>
__global__ void __launch_bounds__(32 * n_groups<SC>()) pass_a_kernel(const Args g)
{
    constexpr int QB = MODE == SRC_ONLY ? 9 : 0;           // first quantity this kernel accumulates
    constexpr int Q = 9 + 4 * SC - QB, NX = 32 * C, NT = 32 * n_groups<SC>(), NP = SC == 1 ? 1 : 2;
    constexpr int MAXPX = (NX + NT - 1) / NT;  // output pixels a thread solves per row, at most
    extern __shared__ __align__(16) uint32_t pbuf[];  // [2][Q][NX]
    const int tid = threadIdx.x, lane = tid & 31, group = tid >> 5;
    const int img = blockIdx.z;
    const int sx0 = blockIdx.x * g.twe;  // first output column of the strip (image coordinates)
    const int y0 = blockIdx.y * g.seg_rows;
    const int y1 = min(g.h, y0 + g.seg_rows);
    const size_t plane = (size_t)g.h * g.wp;
    const uint32_t *PK = g.packed + (size_t)img * NP * plane;
    const int r = g.r;
    const int n_out = min(g.twe, g.w - sx0);
    const uint32_t bias = (uint32_t)(2 * r + 1) * 0x4B000000u;  // what the biased elements add to a window
    // padded coordinate of this lane's first column is sx0 + lane*C (strip origin sx0 - rh, plus the rh
    // offset of the padding).  Lanes whose chunk would start beyond the padded row re-read the last
    // chunk: their prefixes lie right of every window of this strip and are never used.
    const int xp0c = min(sx0 + lane * C, g.wp - C);
    float *GS = MODE == FULL ? nullptr : g.gstat + (size_t)img * 9 * plane;

    float V[NQG][C];
#pragma unroll
    for (int q = 0; q < NQG; ++q)
#pragma unroll
        for (int c = 0; c < C; ++c) V[q][c] = 0.0f;

    // the per-pixel solve of one output row from its prefixes P (and, in SRC_ONLY mode, the statistics gs)
    auto math_row = [&](const int y, const uint32_t *P, const float (*gs)[9]) {
#pragma unroll(MODE == SRC_ONLY ? MAXPX : 1)
        for (int kk = 0; kk < MAXPX; ++kk) {
            const int idx = tid + kk * NT;
            if (idx >= n_out) break;
            const int i = g.rh + idx;
            const int x = sx0 + idx;
            const size_t row_off = (size_t)y * g.wp;
            float mi[3], inv[6];
            if (MODE == SRC_ONLY) {
#pragma unroll
                for (int k = 0; k < 3; ++k) mi[k] = gs[MODE == SRC_ONLY ? kk : 0][k];
#pragma unroll
                for (int k = 0; k < 6; ++k) inv[k] = gs[MODE == SRC_ONLY ? kk : 0][3 + k];
            } else {
                float m[9];
#pragma unroll
                for (int q = 0; q < 9; ++q)
                    m[q] = box_mean(P[q * NX + i + r] - P[q * NX + i - r - 1] - bias, q_is_linear(q), g.inv_area);
                guide_inverse(m, g.eps, inv);
                mi[0] = m[0];
                mi[1] = m[1];
                mi[2] = m[2];
                if (MODE == FULL_STORE) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) GS[k * plane + row_off + g.rh + x] = mi[k];
#pragma unroll
                    for (int k = 0; k < 6; ++k) GS[(3 + k) * plane + row_off + g.rh + x] = inv[k];
                }
            }
            // mirrored copies for the halo columns pass B will read (BORDER_REFLECT: -1-j <-> j)
            const int xl = x < g.rh ? g.rh - 1 - x : -1;                   // padded column of the left mirror
            const int xr = x >= g.w - g.rh ? g.rh + 2 * g.w - 1 - x : -1;  // padded column of the right mirror
#pragma unroll
            for (int c = 0; c < SC; ++c) {
                float ms[4], v[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int q = 9 + 4 * c + k;
                    ms[k] = box_mean(P[(q - QB) * NX + i + r] - P[(q - QB) * NX + i - r - 1] - bias, k == 0, g.inv_area);
                }
                solve_source(ms, mi, inv, v);
                float *o = g.ab + ((size_t)(img * SC + c) * 4) * plane + row_off;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    o[k * plane + g.rh + x] = v[k];
                    if (xl >= 0) o[k * plane + xl] = v[k];
                    if (xr >= 0 && xr < g.wp) o[k * plane + xr] = v[k];
                }
            }
        }
    };
    auto load_stats = [&](const int y, float (*gs)[9]) {
        const size_t row_off = (size_t)y * g.wp + g.rh + sx0;
#pragma unroll
        for (int k = 0; k < MAXPX; ++k) {
            const int idx = min(tid + k * NT, n_out - 1);
#pragma unroll
            for (int j = 0; j < 9; ++j) gs[k][j] = __ldg(GS + j * plane + row_off + idx);
        }
    };

    // One loop over the rows entering the window: steps 0..2r-1 only warm the vertical sums up, every later
    // step also emits a row.  Software-pipelined by one row: between two barriers a warp scans row y AND
    // solves row y-1 (whose prefixes all warps stored before the previous barrier), so the shuffle-latency-
    // bound scan overlaps the arithmetic of the solve.  Rows are requested one step ahead of their use.
    RowRaw<SC, C> cur_in = prefetch_row<SC, C>(PK, plane, g.wp, reflect(y0 - r, g.h), xp0c);
    RowRaw<SC, C> cur_out = cur_in;
    const int n_steps = 2 * r + (y1 - y0);
    float gs[MODE == SRC_ONLY ? MAXPX : 1][9];
    for (int t = 0; t < n_steps; ++t) {
        const RowRaw<SC, C> nxt_in = prefetch_row<SC, C>(PK, plane, g.wp, reflect(y0 - r + t + 1, g.h), xp0c);
        // cached guide statistics of the pixels this thread solves in this step (row y-1): requested before
        // the accumulate / scan work so that their DRAM latency is hidden
        if (MODE == SRC_ONLY && t > 2 * r) load_stats(y0 + t - 2 * r - 1, gs);
        if (MODE == SRC_ONLY) {
            RF_GF2_DISPATCH_SRC(accumulate, V, cur_in, 1.0f)
        } else {
            RF_GF2_DISPATCH(accumulate, V, cur_in, 1.0f)
        }
        cur_in = nxt_in;
        if (t < 2 * r) continue;
        const int y = y0 + t - 2 * r;
        const RowRaw<SC, C> nxt_out = prefetch_row<SC, C>(PK, plane, g.wp, reflect(y + 1 - r, g.h), xp0c);
        uint32_t *P = pbuf + ((y - y0) & 1) * (Q * NX);
        if (MODE == SRC_ONLY) {
            RF_GF2_DISPATCH_SRC(scan_store, V, P, NX, lane)
            RF_GF2_DISPATCH_SRC(accumulate, V, cur_out, -1.0f)
        } else {
            RF_GF2_DISPATCH(scan_store, V, P, NX, lane)
            RF_GF2_DISPATCH(accumulate, V, cur_out, -1.0f)
        }
        cur_out = nxt_out;
        if (y > y0) math_row(y - 1, pbuf + ((y - 1 - y0) & 1) * (Q * NX), gs);
        // One barrier per row: the prefixes of row y are visible to everyone after it, and everyone has
        // finished reading the other buffer (row y-1), which the next step overwrites.
        __syncthreads();
    }
    if (MODE == SRC_ONLY) load_stats(y1 - 1, gs);
    math_row(y1 - 1, pbuf + ((y1 - 1 - y0) & 1) * (Q * NX), gs);
}

// ---- pass B ---------------------------------------------------------------------------------------
// one warp per coefficient plane (4 * SC warps); FP32 vertical sliding sums and FP32 prefixes
template <int SC, int C>
__global__ void __launch_bounds__(32 * 4 * SC) pass_b_kernel(const Args g)
{
    constexpr int Q = 4 * SC, NX = 32 * C, NT = 32 * Q, NP = SC == 1 ? 1 : 2;
    extern __shared__ __align__(16) float fbuf[];  // [2][Q][NX]
    const int tid = threadIdx.x, lane = tid & 31, pl = tid >> 5;
    const int img = blockIdx.z;
    const int sx0 = blockIdx.x * g.twe;
    const int y0 = blockIdx.y * g.seg_rows;
    const int y1 = min(g.h, y0 + g.seg_rows);
    const int xp0c = min(sx0 + lane * C, g.wp - C);
    const size_t plane = (size_t)g.h * g.wp;
    const size_t img_px = (size_t)g.h * g.w;
    const float *A = g.ab + ((size_t)img * Q + pl) * plane;
    const uint32_t *PK = g.packed + (size_t)img * NP * plane;
    const int r = g.r;
    const int n_out = min(g.twe, g.w - sx0);

    float V[C];
#pragma unroll
    for (int c = 0; c < C; ++c) V[c] = 0.0f;
    struct Chunk {
        float v[C];
    };
    auto fetch = [&](int yy) {
        Chunk k;
        const float4 *p = reinterpret_cast<const float4 *>(A + (size_t)yy * g.wp + xp0c);
#pragma unroll
        for (int c = 0; c < C; c += 4) {
            const float4 t = __ldg(p + c / 4);
            k.v[c] = t.x;
            k.v[c + 1] = t.y;
            k.v[c + 2] = t.z;
            k.v[c + 3] = t.w;
        }
        return k;
    };
    auto add = [&](const Chunk &k, float sign) {  // exact either way: sign * x is x or -x
#pragma unroll
        for (int c = 0; c < C; c += 2) {
            unsigned long long v = pack2(V[c], V[c + 1]);
            ffma2(v, pack2(sign, sign), pack2(k.v[c], k.v[c + 1]));
            unpack2(v, V[c], V[c + 1]);
        }
    };
    // Rows are loaded into registers PF steps ahead of their use.  Measured on B200 (64 x 512 x 384, r = 45):
    // PF = 3 costs occupancy (128 registers) and is 35 % slower; an additional prefetch.global.L2 eight rows
    // ahead is 15 % slower (the kernel already moves 3.4x the compulsory bytes: the leaving-row stream and the
    // strip overlap are re-read from DRAM, so extra requests only add pressure).
    constexpr int PF = 1;
    Chunk qin[PF], qout[PF];
#pragma unroll
    for (int i = 0; i < PF; ++i) qin[i] = qout[i] = fetch(reflect(y0 - r + i, g.h));
    // output row y from the prefixes P of the four (twelve) coefficient planes
    auto math_row = [&](const int y, const float *P) {
        for (int idx = tid; idx < n_out; idx += NT) {
            const int i = g.rh + idx;
            const int x = sx0 + idx;
            const uint32_t gw = PK[(size_t)y * g.wp + g.rh + x];
            const float i0 = b2f(gw, 0), i1 = b2f(gw, 1), i2 = b2f(gw, 2);
            uint8_t res[SC];
#pragma unroll
            for (int c = 0; c < SC; ++c) {
                float m[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float *Pq = P + (4 * c + k) * NX;
                    m[k] = __fmul_rn(Pq[i + r] - Pq[i - r - 1], g.inv_area);
                }
                float v = m[3];
                v = __fadd_rn(v, __fmul_rn(m[0], i0));
                v = __fadd_rn(v, __fmul_rn(m[1], i1));
                v = __fadd_rn(v, __fmul_rn(m[2], i2));
                res[c] = sat_u8(v);
            }
            if (g.store_dst) {
                uint8_t *o = g.dst + (img * img_px + (size_t)y * g.w + x) * SC;
#pragma unroll
                for (int c = 0; c < SC; ++c) o[c] = res[c];
            }
            if (g.store_packed) {
                // the next iteration filters this output: put it where pack_kernel would have put it, mirrored
                // halo columns included.  Other CTAs of this launch read only the guide bytes of these words.
                const int xl = x < g.rh ? g.rh - 1 - x : -1;
                const int xr = x >= g.w - g.rh ? g.rh + 2 * g.w - 1 - x : -1;
                uint32_t *row = g.packed + (size_t)img * NP * plane + (size_t)y * g.wp;
                if (SC == 1) {
                    uint8_t *rb = reinterpret_cast<uint8_t *>(row);
                    rb[(size_t)(g.rh + x) * 4 + 3] = res[0];
                    if (xl >= 0) rb[(size_t)xl * 4 + 3] = res[0];
                    if (xr >= 0 && xr < g.wp) rb[(size_t)xr * 4 + 3] = res[0];
                } else {
                    const uint32_t wv = res[0] | ((uint32_t)res[1 % SC] << 8) | ((uint32_t)res[2 % SC] << 16);
                    row[plane + g.rh + x] = wv;
                    if (xl >= 0) row[plane + xl] = wv;
                    if (xr >= 0 && xr < g.wp) row[plane + xr] = wv;
                }
            }
        }
    };
    // software-pipelined by one row like pass A: between two barriers a warp scans row y and finishes row y-1
    const int n_steps = 2 * r + (y1 - y0);
    for (int t = 0; t < n_steps; ++t) {
        add(qin[0], 1.0f);
#pragma unroll
        for (int i = 0; i + 1 < PF; ++i) qin[i] = qin[i + 1];
        qin[PF - 1] = fetch(reflect(y0 - r + t + PF, g.h));
        if (t < 2 * r) continue;
        const int y = y0 + t - 2 * r;
        float *P = fbuf + ((y - y0) & 1) * (Q * NX);
        {
            float pre[C];
            float run = 0.0f;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                run += V[c];
                pre[c] = run;
            }
            float incl = run;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const float tt = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += tt;
            }
            // exclusive prefix by a shift, NOT incl - run: the chunk right of the last needed column may hold
            // unwritten coefficients (inf/NaN garbage), and inf - inf would poison this lane's own prefix
            float excl = __shfl_up_sync(0xffffffffu, incl, 1);
            if (lane == 0) excl = 0.0f;
            float *dstp = P + pl * NX + lane * C;
#pragma unroll
            for (int c = 0; c < C; c += 4)
                *reinterpret_cast<float4 *>(dstp + c) =
                    make_float4(pre[c] + excl, pre[c + 1] + excl, pre[c + 2] + excl, pre[c + 3] + excl);
        }
        add(qout[0], -1.0f);
#pragma unroll
        for (int i = 0; i + 1 < PF; ++i) qout[i] = qout[i + 1];
        qout[PF - 1] = fetch(reflect(y - r + PF, g.h));
        if (y > y0) math_row(y - 1, fbuf + ((y - 1 - y0) & 1) * (Q * NX));
        __syncthreads();
    }
    math_row(y1 - 1, fbuf + ((y1 - 1 - y0) & 1) * (Q * NX));
}

// ---- host -------------------------------------------------------------------------------------------
struct Plan {
    int C, strips, twe, rh, wp, segs, seg_rows;
};

static int halo(int r) { return (r + 1 + 3) & ~3; }
// + 16: the chunk (<= 16 columns) that holds the right-most needed column must lie inside the row
static int pitch(int w, int r) { return (w + 2 * halo(r) + 16 + 3) & ~3; }

// tuning overrides (development only): RF_GF2_SEGS_A / RF_GF2_SEGS_B force the number of row segments
static int env_int(const char *name)
{
    const char *e = getenv(name);
    return e ? atoi(e) : 0;
}

static Plan make_plan(int h, int w, int r)
{
    Plan best{};
    long best_cost = -1;
    const int rh = halo(r);
    for (int C : {8, 12, 16}) {
        const int tw = 32 * C - 2 * rh;
        if (tw < 32) continue;
        const int strips = (w + tw - 1) / tw;
        int twe = ((w + strips - 1) / strips + 3) & ~3;
        if (twe > tw) twe = tw & ~3;
        const long cost = (long)strips * 32 * C;  // columns touched per image row
        if (best_cost < 0 || cost < best_cost) {
            best_cost = cost;
            best = Plan{C, strips, twe, rh, pitch(w, r), 1, h};
        }
    }
    return best;
}

template <int SC, int C>
static int launch(Args a, const Plan &p, int iterations, cudaStream_t st)
{
    constexpr int QA_ = 9 + 4 * SC, QS_ = 4 * SC, QB_ = 4 * SC, NX = 32 * C, NTA = 32 * n_groups<SC>();
    const size_t smem_a = (size_t)2 * QA_ * NX * sizeof(uint32_t);
    const size_t smem_s = (size_t)2 * QS_ * NX * sizeof(uint32_t);
    const size_t smem_b = (size_t)2 * QB_ * NX * sizeof(float);
    static bool configured[64] = {};
    static int occ_a = 1, occ_s = 1, occ_b = 1;
    int dev = 0;
    RF_CUDA_TRY(cudaGetDevice(&dev));
    if (!configured[dev & 63]) {
        RF_CUDA_TRY(cudaFuncSetAttribute(pass_a_kernel<SC, C, FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        RF_CUDA_TRY(cudaFuncSetAttribute(pass_a_kernel<SC, C, FULL_STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        RF_CUDA_TRY(cudaFuncSetAttribute(pass_a_kernel<SC, C, SRC_ONLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        RF_CUDA_TRY(cudaFuncSetAttribute(pass_b_kernel<SC, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        RF_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_a, pass_a_kernel<SC, C, FULL_STORE>, NTA, smem_a));
        RF_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_s, pass_a_kernel<SC, C, SRC_ONLY>, NTA, smem_s));
        RF_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_b, pass_b_kernel<SC, C>, 32 * 4 * SC, smem_b));
        if (occ_a < 1) occ_a = 1;
        if (occ_s < 1) occ_s = 1;
        if (occ_b < 1) occ_b = 1;
        configured[dev & 63] = true;
    }
    dim3 pgrid((a.wp + 255) / 256, (a.h + PACK_ROWS - 1) / PACK_ROWS, a.n);
    pack_kernel<SC><<<pgrid, 256, 0, st>>>(a);
    RF_LAUNCH_CHECK("gf2::pack_kernel");
    // Row segments: the vertical sums make a column strictly sequential, so small batches are split into
    // row segments (each pays 2r warm-up rows) until the grid fills ONE wave of resident CTAs -- measured:
    // more than one wave loses to tail effects, fewer leaves SMs idle (profiles/r01_gf2_segments.txt).
    static const int force_a = env_int("RF_GF2_SEGS_A"), force_b = env_int("RF_GF2_SEGS_B");
    const long ctas = (long)p.strips * a.n;
    // shortest segment: 64 rows for pass A, 48 for pass B, whose warm-up rows are cheap (measured at 64 x 512x384:
    // 6 segments 1.50 ms for three iterations, 8 segments 1.45 ms, 10 segments 1.67 ms)
    auto pick = [&](int occ, int forced, int min_rows) {
        if (forced > 0) return forced;
        const int max_segs = a.h / min_rows > 1 ? a.h / min_rows : 1;
        long s = (long)sm_count() * occ / ctas;
        if (s < 1) s = 1;
        if (s > max_segs) s = max_segs;
        return (int)s;
    };
    const int sa = pick(occ_a, force_a, 64), ss = pick(occ_s, force_a, 64), sb = pick(occ_b, force_b, 48);
    auto grid_for = [&](int segs) {
        a.seg_rows = (a.h + segs - 1) / segs;
        return dim3(p.strips, (a.h + a.seg_rows - 1) / a.seg_rows, a.n);
    };
    for (int it = 0; it < iterations; ++it) {
        const bool last = it == iterations - 1;
        if (it == 0) {
            const dim3 grid = grid_for(sa);
            if (iterations > 1)
                pass_a_kernel<SC, C, FULL_STORE><<<grid, NTA, smem_a, st>>>(a);
            else
                pass_a_kernel<SC, C, FULL><<<grid, NTA, smem_a, st>>>(a);
        } else {
            const dim3 grid = grid_for(ss);
            pass_a_kernel<SC, C, SRC_ONLY><<<grid, NTA, smem_s, st>>>(a);
        }
        RF_LAUNCH_CHECK("gf2::pass_a_kernel");
        a.store_dst = last ? 1 : 0;
        a.store_packed = last ? 0 : 1;
        const dim3 grid = grid_for(sb);
        pass_b_kernel<SC, C><<<grid, 32 * 4 * SC, smem_b, st>>>(a);
        RF_LAUNCH_CHECK("gf2::pass_b_kernel");
    }
    return RF_OK;
}

// single reflection must cover the halo, the exact-integer FP32 sums need r <= 64
// Small windows go to the generic kernels (exact float(double(S) / k^2) means): with a handful of samples per
// window the covariance is often near-singular and the 1-ulp mean of this path gets amplified (measured on 'flat'
// guides, r = 1, eps <= 0.05: 0.5 % of bytes move by 1 LSB and single bytes by 2; the generic path is bit-equal to
// the oracle there).  From r = 8 up the deviation is below 1e-3 of the bytes, +-1 LSB.
constexpr int MIN_RADIUS = 8;
bool supported(int r, int h, int w) { return r >= MIN_RADIUS && r <= MAX_RADIUS && w >= halo(r) && h <= 65535; }

// packed planes + coefficient planes (+ nine guide-statistics planes when the filter is iterated)
size_t workspace_per_image(int sc, int h, int w, int r, int iterations)
{
    const size_t plane = (size_t)h * pitch(w, r);
    return plane * 4 * (sc == 1 ? 1 : 2) + plane * 4 * 4 * sc + (iterations > 1 ? plane * 4 * 9 : 0);
}

int run(const uint8_t *guide, const uint8_t *src, int sc, uint8_t *dst, void *ws, int n, int h, int w, int r,
        double eps, int iterations, cudaStream_t st)
{
    Args a;
    a.guide = guide;
    a.src = src;
    a.dst = dst;
    a.n = n;
    a.h = h;
    a.w = w;
    a.r = r;
    a.eps = (float)eps;
    const int k = 2 * r + 1;
    a.inv_area = (float)(1.0 / ((double)k * k));
    const Plan p = make_plan(h, w, r);
    a.rh = p.rh;
    a.wp = p.wp;
    a.twe = p.twe;
    a.seg_rows = p.seg_rows;
    a.store_dst = 1;
    a.store_packed = 0;
    const size_t plane = (size_t)h * p.wp;
    a.packed = (uint32_t *)ws;
    a.ab = (float *)((uint32_t *)ws + (size_t)n * (sc == 1 ? 1 : 2) * plane);
    a.gstat = iterations > 1 ? a.ab + (size_t)n * sc * 4 * plane : nullptr;
#define RF_GF2_LAUNCH(SC_)                                            \
    switch (p.C) {                                                    \
        case 8: return launch<SC_, 8>(a, p, iterations, st);          \
        case 12: return launch<SC_, 12>(a, p, iterations, st);        \
        default: return launch<SC_, 16>(a, p, iterations, st);        \
    }
    if (sc == 1) { RF_GF2_LAUNCH(1) }
    RF_GF2_LAUNCH(3)
#undef RF_GF2_LAUNCH
}

}  // namespace gf2
}  // namespace rf
