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
//   * pass A: every lane owns C consecutive columns of a 32*C wide strip and keeps the vertical window sums of
//     its quantities in FP32 registers -- integers below 2^23 ((2r+1)*255^2 for r <= 64), so the FFMA
//     updates (add the entering row, subtract the leaving one) are exact;
//   * the horizontal direction is a per-lane serial prefix over its C columns plus ONE 5-step warp scan
//     of the lane totals per quantity (uint32, wrap-around safe): 5/C shuffles per column;
//   * the 9 + 4*SC quantities are split over the warps of the CTA (each warp scans the whole strip for
//     its 3-4 quantities); prefixes go to a double-buffered shared-memory row, and after ONE
//     __syncthreads per row every thread turns P[x+r] - P[x-r-1] of ONE PIXEL PAIR into means and solves the
//     two 3x3 systems as one packed (f32x2) instruction stream, in the oracle's non-contracted multiply/add
//     order;
//   * pass A does not store the coefficients (a0, a1, a2, b) themselves but their VERTICAL PREFIX SUM down the
//     rows of its row segment (a thread owns the same columns in every row, so this is one packed add per
//     coefficient).  That makes pass B row-independent: the vertical window sum of a coefficient plane is a
//     signed sum of 2-4 prefix rows (row_terms() below: PV[y+r] - PV[y-r-1], plus the totals of segments the
//     window crosses and the BORDER_REFLECT mirror parts), so pass B has no sliding state, no warm-up rows and
//     no leaving-row stream, and every prefix row is read from DRAM once (its second use, 2r+1 rows later, hits
//     L2 because the grid walks the rows in order);
//   * pass B: one warp per (output row, coefficient plane); the prefix rows arrive in shared memory through
//     TMA (cp.async.bulk.tensor on a 3-D map of the planes, two slots per warp, mbarrier completion; columns
//     beyond the padded row are zero-filled by the hardware), the warp sums its terms, runs the same
//     per-lane prefix + warp scan over the row in FP32, and after one __syncthreads the threads of the row
//     combine the four box means with the guide of one pixel pair each (packed arithmetic) and store uint8;
//   * no FP64: the box mean is float(S) * float(1/k^2) (<= 1 ulp from OpenCV's float(double(S)/k^2);
//     measured effect ~2.5e-5 of the output bytes move by 1 LSB);
//   * iterated filtering with one guide (createGuidedFilter(guide, r, eps) reused for several filter() calls,
//     SURVEY 8f-4; the "3 x GF" configuration): the first pass A also stores mean(I) and the inverse
//     covariance per pixel, later iterations only accumulate the 4*SC source quantities and read those nine
//     floats back; pass B writes its uint8 output straight into the packed planes, so later iterations need
//     no pack kernel.  The arithmetic per pixel is the same function in both modes: results are byte-identical
//     to repeated single calls.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <mutex>

#include <cstdlib>
#include <type_traits>

#include "common.cuh"

namespace rf {
namespace gf2 {

constexpr int MAX_RADIUS = 64;  // (2r+1) * 65025 < 2^23
constexpr int QMAX = 21;

constexpr int MAXT = 12;                  // most prefix rows one window sum needs (row_terms)
constexpr int PLAN_STRIDE = 2 * MAXT + 4;  // ints per row of the row plan: count, pad, pad, pad, rows, weights

struct Args {
    const uint8_t *guide;  // [n][h][w][3]
    const uint8_t *src;    // [n][h][w][SC]
    uint32_t *packed;      // [n][NP][h][wp]  NP = 1 (SC == 1: B,G,R,p) or 2 (SC == 3: B,G,R,0 / p0,p1,p2,0)
    float *ab;             // [n][SC][4][h][wp]  vertical prefix sums (restarted every seg_rows rows) of
                           // (a0, a1, a2, b), columns padded like `packed`
    float *gstat;          // [n][9][h][wg]  guide statistics kept for iterated filtering: mean I (3) and the
                           // inverse of cov(I) + eps*Id (00, 01, 02, 11, 12, 22); NULL when not wanted
    const int *plan;       // [h][PLAN_STRIDE]  row plan of pass B (plan_kernel)
    uint8_t *dst;          // [n][h][w][SC]
    int store_dst;         // pass B writes dst (last iteration)
    int store_packed;      // pass B writes its output into the source bytes of `packed` (input of the next iteration)
    int n, h, w, r;
    int rh;          // halo columns on each side: round_up(r + 1, 4)
    int wp;          // padded row pitch in pixels: round_up(w + 2 * rh + 16, 4)
    int wg;          // row pitch of the statistics planes (no halo): round_up(w, 4)
    int twe;         // pass A: output columns per strip (multiple of 4)
    int seg_rows;    // pass A: output rows per CTA = rows per prefix segment
    int twe_b;       // pass B: output columns per strip (multiple of 4)
    float eps;
    float inv_area;  // 1 / (2r+1)^2
};

// ---- packed FP32 pairs (one 64-bit register = two adjacent pixels) ---------------------------------
typedef unsigned long long f2;
__device__ __forceinline__ f2 mul2(f2 a, f2 b)
{
    f2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b)
{
    f2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f2 sub2(f2 a, f2 b)
{
    f2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f2 dup2(float v) { return pack2(v, v); }

// ---- row plan of pass B ------------------------------------------------------------------------------
// Pass A stores PV[y] = sum of the coefficient rows seg_start(y) .. y (the prefix restarts at every multiple of
// seg_rows).  The vertical window sum of output row y -- rows y-r .. y+r under BORDER_REFLECT
// (fedcba|abcdefgh|hgfedcb) -- is then a signed sum of a few prefix rows.
struct RowTerms {
    int cnt;  // > MAXT: does not fit (never for h > r and seg_rows >= 64)
    int row[MAXT];
    float wgt[MAXT];
};

__host__ __device__ inline void push_term(RowTerms &t, int row, float w)
{
    for (int i = 0; i < t.cnt && i < MAXT; ++i)
        if (t.row[i] == row) {
            t.wgt[i] += w;
            return;
        }
    if (t.cnt < MAXT) {
        t.row[t.cnt] = row;
        t.wgt[t.cnt] = w;
    }
    ++t.cnt;
}

// image rows lo .. hi (0 <= lo <= hi < h)
__host__ __device__ inline void add_range(RowTerms &t, int lo, int hi, int seg_rows)
{
    push_term(t, hi, 1.0f);
    int s = (hi / seg_rows) * seg_rows;  // first row of hi's segment
    while (s > lo) {                     // the range reaches into the previous segment: add that segment's total
        push_term(t, s - 1, 1.0f);
        s -= seg_rows;
    }
    if (lo > s) push_term(t, lo - 1, -1.0f);
}

__host__ __device__ inline void row_terms(RowTerms &t, int y, int h, int r, int seg_rows)
{
    t.cnt = 0;
    int p = y - r;
    const int end = y + r;
    while (p <= end) {
        const int m = p >= 0 ? p / h : -((-p + h - 1) / h);  // which reflected copy of the image p falls into
        const int q = p - m * h;
        const int rest = end - p + 1;
        const int len = h - q < rest ? h - q : rest;
        if ((m & 1) == 0)
            add_range(t, q, q + len - 1, seg_rows);  // upright copy
        else
            add_range(t, h - q - len, h - 1 - q, seg_rows);  // mirrored copy
        p += len;
    }
}

__global__ void plan_kernel(int *plan, int h, int r, int seg_rows)
{
    const int y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= h) return;
    RowTerms t;
    row_terms(t, y, h, r, seg_rows);
    int *o = plan + (size_t)y * PLAN_STRIDE;
    o[0] = t.cnt < MAXT ? t.cnt : MAXT;
    for (int i = 0; i < MAXT; ++i) {
        o[4 + i] = i < t.cnt ? t.row[i] : 0;
        o[4 + MAXT + i] = __float_as_int(i < t.cnt ? t.wgt[i] : 0.0f);
    }
}

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
        for (int c = 0; c < C; ++c) {
            // V = 2^23 + (window sum): its bit pattern is 0x4B000000 + the integer, no conversion needed
            run += __float_as_uint(V[q][c]);
            pre[c] = run;
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

// window sums of the pixel pair (i, i+1) -> box means.  Sums of single channels stay below 2^23: exact
// integer->float on the FMA pipe.  (Four 32-bit loads: which of the two index pairs is 8-byte aligned depends on
// the parity of the radius.)
template <bool R_ODD>
__device__ __forceinline__ f2 box_mean2(const uint32_t *Pq, int i, int r, uint32_t bias, bool linear, f2 inv_area2)
{
    // i is even: P[i-r-1], P[i-r] (odd radius) or P[i+r], P[i+r+1] (even radius) is one aligned 64-bit load
    uint32_t h0, h1, l0, l1;
    if (R_ODD) {
        const uint2 lo = *reinterpret_cast<const uint2 *>(Pq + i - r - 1);
        l0 = lo.x, l1 = lo.y, h0 = Pq[i + r], h1 = Pq[i + r + 1];
    } else {
        const uint2 hi = *reinterpret_cast<const uint2 *>(Pq + i + r);
        h0 = hi.x, h1 = hi.y, l0 = Pq[i - r - 1], l1 = Pq[i - r];
    }
    const uint32_t d0 = h0 - l0 - bias, d1 = h1 - l1 - bias;
    f2 sf;
    if (linear)
        sf = add2(pack2(__uint_as_float(d0 | 0x4B000000u), __uint_as_float(d1 | 0x4B000000u)), dup2(-8388608.0f));
    else
        sf = pack2((float)d0, (float)d1);
    return mul2(sf, inv_area2);
}

// inverse of cov(I) + eps*Id from the nine guide means m[0..8]; out: inv 00, 01, 02, 11, 12, 22.  Two pixels per
// instruction; every step is the separately rounded multiply / add / subtract of the oracle.
__device__ __forceinline__ void guide_inverse2(const f2 *m, float eps, f2 *inv)
{
    const f2 e2 = dup2(eps);
    // cov(I) + eps*Id, symmetric storage 0:(0,0) 1:(0,1) 2:(0,2) 3:(1,1) 4:(1,2) 5:(2,2)
    const f2 c00 = add2(sub2(m[3], mul2(m[0], m[0])), e2);
    const f2 c01 = sub2(m[4], mul2(m[0], m[1]));
    const f2 c02 = sub2(m[5], mul2(m[0], m[2]));
    const f2 c11 = add2(sub2(m[6], mul2(m[1], m[1])), e2);
    const f2 c12 = sub2(m[7], mul2(m[1], m[2]));
    const f2 c22 = add2(sub2(m[8], mul2(m[2], m[2])), e2);
    const f2 f00 = sub2(mul2(c11, c22), mul2(c12, c12));
    const f2 f01 = sub2(mul2(c12, c02), mul2(c01, c22));
    const f2 f02 = sub2(mul2(c01, c12), mul2(c11, c02));
    const f2 f11 = sub2(mul2(c22, c00), mul2(c02, c02));
    const f2 f12 = sub2(mul2(c02, c01), mul2(c12, c00));
    const f2 f22 = sub2(mul2(c00, c11), mul2(c01, c01));
    f2 det = mul2(c00, f00);
    det = add2(det, mul2(c01, f01));
    det = add2(det, mul2(c02, f02));
    float d0, d1;
    unpack2(det, d0, d1);
    if (eps < 1e-2f) {
        if (fabsf(d0) < 1e-6f) d0 = 1e-6f;
        if (fabsf(d1) < 1e-6f) d1 = 1e-6f;
    }
    // one correctly rounded reciprocal instead of six divisions: each entry is within 1 ulp of cof/det
    const f2 rdet = pack2(__frcp_rn(d0), __frcp_rn(d1));
    inv[0] = mul2(f00, rdet);
    inv[1] = mul2(f01, rdet);
    inv[2] = mul2(f02, rdet);
    inv[3] = mul2(f11, rdet);
    inv[4] = mul2(f12, rdet);
    inv[5] = mul2(f22, rdet);
}

// a = inv * (mean(I p) - mean(I) mean(p)),  b = mean(p) - a . mean(I);  ms = (mean p, mean p*I0, p*I1, p*I2)
__device__ __forceinline__ void solve_source2(const f2 *ms, const f2 *mi, const f2 *inv, f2 *v)
{
    const f2 mp = ms[0];
    const f2 k0 = sub2(ms[1], mul2(mp, mi[0]));
    const f2 k1 = sub2(ms[2], mul2(mp, mi[1]));
    const f2 k2 = sub2(ms[3], mul2(mp, mi[2]));
    f2 a0 = mul2(inv[0], k0);
    a0 = add2(a0, mul2(inv[1], k1));
    a0 = add2(a0, mul2(inv[2], k2));
    f2 a1 = mul2(inv[1], k0);
    a1 = add2(a1, mul2(inv[3], k1));
    a1 = add2(a1, mul2(inv[4], k2));
    f2 a2 = mul2(inv[2], k0);
    a2 = add2(a2, mul2(inv[4], k1));
    a2 = add2(a2, mul2(inv[5], k2));
    f2 b = sub2(mp, mul2(a0, mi[0]));
    b = sub2(b, mul2(a1, mi[1]));
    b = sub2(b, mul2(a2, mi[2]));
    v[0] = a0;
    v[1] = a1;
    v[2] = a2;
    v[3] = b;
}

// stores a pixel pair (8-byte aligned address), or only its first pixel when the second lies outside the image
__device__ __forceinline__ void st_pair(float *p, f2 v, bool both)
{
    if (both) {
        *reinterpret_cast<f2 *>(p) = v;
    } else {
        float lo, hi;
        unpack2(v, lo, hi);
        *p = lo;
    }
}

// ---- mbarrier / TMA helpers ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    while (!done)
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
}

// one lane of a converged warp
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *tm, int c0, int c1, int c2, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}

// ---- pass A ---------------------------------------------------------------------------------------
// MODE FULL: all 9 + 4*SC quantities.  FULL_STORE: the same, and the guide statistics go to g.gstat.
// SRC_ONLY: only the 4*SC source quantities; mean(I) and the inverse covariance come from g.gstat.
//
// The packed rows arrive through TMA: an elected lane of warp 0 requests, RING-1 steps ahead, the row that enters
// the vertical window and (once rows are emitted) the row that leaves it, as one box of NX pixels per plane
// (3-D map of the packed planes; columns beyond the padded row are zero-filled and only feed prefixes right of
// every window).  Every warp reads its lanes' chunks from the ring with conflict-free 128-bit loads and releases
// the slot through an mbarrier, so DRAM latency is hidden by the ring, not by resident warps, and no row lives
// in registers across steps.
// ring depth in steps: the statistics of SRC_ONLY make its slots four times larger, three of them keep four CTAs per SM
template <int MODE>
__host__ __device__ constexpr int ring_depth() { return MODE == 2 ? 3 : 4; }

template <int SC, int C>
__device__ __forceinline__ RowRaw<SC, C> read_row(const uint32_t *slot, int lane)
{
    constexpr int NX = 32 * C;
    RowRaw<SC, C> r;
    const uint4 *p = reinterpret_cast<const uint4 *>(slot + lane * C);
#pragma unroll
    for (int c = 0; c < C; c += 4) {
        const uint4 t = p[c / 4];
        r.w0[c] = t.x, r.w0[c + 1] = t.y, r.w0[c + 2] = t.z, r.w0[c + 3] = t.w;
    }
    if (SC == 3) {
        const uint4 *q = reinterpret_cast<const uint4 *>(slot + NX + lane * C);
#pragma unroll
        for (int c = 0; c < (SC == 1 ? 0 : C); c += 4) {
            const uint4 t = q[c / 4];
            r.w1[c] = t.x, r.w1[c + 1] = t.y, r.w1[c + 2] = t.z, r.w1[c + 3] = t.w;
        }
    } else {
        r.w1[0] = 0u;
    }
    return r;
}

// bytes of one ring slot: entering row + leaving row (NP planes of NX pixels each), and in SRC_ONLY mode the nine
// cached statistics planes of the 2 * threads pixels the CTA solves per row
template <int SC, int C, int MODE>
__host__ __device__ constexpr size_t pass_a_slot_bytes()
{
    return (size_t)2 * (SC == 1 ? 1 : 2) * 32 * C * 4 + (MODE == 2 ? (size_t)9 * 2 * 32 * (SC == 1 ? 4 : 6) * 4 : 0);
}
template <int SC, int C, int MODE>
__host__ __device__ constexpr size_t pass_a_smem()
{
    // prefix rows [2][q][NX], ring of RING slots, 2 * RING mbarriers
    return (size_t)2 * (MODE == 2 ? 4 * SC : 9 + 4 * SC) * 32 * C * 4 + ring_depth<MODE>() * pass_a_slot_bytes<SC, C, MODE>() +
           2 * ring_depth<MODE>() * 8;
}

template <int SC, int C, int MODE>
__global__ void __launch_bounds__(32 * n_groups<SC>(), SC == 1 ? 4 : 1)
    pass_a_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_gs, const Args g)
{
    constexpr int QB = MODE == SRC_ONLY ? 9 : 0;  // first quantity this kernel accumulates
    constexpr int Q = 9 + 4 * SC - QB, NX = 32 * C, NP = SC == 1 ? 1 : 2, NG = n_groups<SC>(), NT = 32 * NG;
    constexpr int RING = ring_depth<MODE>();
    constexpr uint32_t ROW_BYTES = NP * NX * 4, GS_BYTES = 9 * 2 * NT * 4;
    constexpr uint32_t SLOT_WORDS = (uint32_t)(pass_a_slot_bytes<SC, C, MODE>() / 4);
    // [2][Q][NX] prefixes | ring [RING] x { in [NP][NX], out [NP][NX], (SRC_ONLY) stats [9][2 * NT] } | full[RING], empty[RING]
    extern __shared__ __align__(128) uint32_t pbuf[];
    uint32_t *ring = pbuf + 2 * Q * NX;
    const uint32_t bars = smem_u32(ring + RING * SLOT_WORDS);
    const int tid = threadIdx.x, lane = tid & 31;
    const int group = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int img = blockIdx.z;
    const int sx0 = blockIdx.x * g.twe;  // first output column of the strip (image coordinates) = padded
                                         // coordinate of the strip's first column (origin sx0 - rh, + rh padding)
    const int y0 = blockIdx.y * g.seg_rows;
    const int y1 = min(g.h, y0 + g.seg_rows);
    const size_t plane = (size_t)g.h * g.wp;
    const int r = g.r;
    const int n_out = min(g.twe, g.w - sx0);  // <= 2 * NT: every thread solves at most one pixel pair per row
    const uint32_t bias = (uint32_t)(2 * r + 1) * 0x4B000000u;  // what the biased elements add to a window
    const size_t gplane = (size_t)g.h * g.wg;
    float *GS = MODE == FULL ? nullptr : g.gstat + (size_t)img * 9 * gplane;
    const f2 ia2 = dup2(g.inv_area);

    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < RING; ++k) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bars + 8 * k) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bars + 8 * (RING + k)), "r"(NG) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // vertical window sums, carried as 2^23 + sum (exact integers below 2^24): the scan reads their bit patterns
    float V[NQG][C];
#pragma unroll
    for (int q = 0; q < NQG; ++q)
#pragma unroll
        for (int c = 0; c < C; ++c) V[q][c] = 8388608.0f;

    // this thread's pixel pair: output columns sx0 + idx, sx0 + idx + 1 of every row of the segment
    const int idx = 2 * tid;
    const bool active = idx < n_out;
    const bool both = idx + 1 < n_out;  // false: the image ends between the two (odd width)
    const int x = sx0 + idx;
    const bool edge = x < g.rh || x + 1 >= g.w - g.rh;  // has mirrored copies in the halo columns
    float *const ab0 = g.ab + (size_t)img * SC * 4 * plane;
    // vertical prefix sums of the coefficients down the rows of this segment (what pass B consumes)
    f2 pv[SC][4];
#pragma unroll
    for (int c = 0; c < SC; ++c)
#pragma unroll
        for (int k = 0; k < 4; ++k) pv[c][k] = 0ull;

    // the per-pixel solve of one output row from its prefixes P (and, in SRC_ONLY mode, the statistics gs)
    auto math_row = [&](auto r_odd, const int y, const uint32_t *P, const f2 *gs) {
        constexpr bool R_ODD = decltype(r_odd)::value;
        if (!active) return;
        const int i = g.rh + idx;
        const size_t row_off = (size_t)y * g.wp;
        f2 mi[3], inv[6];
        if (MODE == SRC_ONLY) {
#pragma unroll
            for (int k = 0; k < 3; ++k) mi[k] = gs[k];
#pragma unroll
            for (int k = 0; k < 6; ++k) inv[k] = gs[3 + k];
        } else {
            f2 m[9];
#pragma unroll
            for (int q = 0; q < 9; ++q) m[q] = box_mean2<R_ODD>(P + q * NX, i, r, bias, q_is_linear(q), ia2);
            guide_inverse2(m, g.eps, inv);
            mi[0] = m[0];
            mi[1] = m[1];
            mi[2] = m[2];
            if (MODE == FULL_STORE) {
#pragma unroll
                for (int k = 0; k < 3; ++k) st_pair(GS + k * gplane + (size_t)y * g.wg + x, mi[k], both);
#pragma unroll
                for (int k = 0; k < 6; ++k) st_pair(GS + (3 + k) * gplane + (size_t)y * g.wg + x, inv[k], both);
            }
        }
#pragma unroll
        for (int c = 0; c < SC; ++c) {
            f2 ms[4], v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int q = 9 + 4 * c + k;
                ms[k] = box_mean2<R_ODD>(P + (q - QB) * NX, i, r, bias, k == 0, ia2);
            }
            solve_source2(ms, mi, inv, v);
            float *o = ab0 + (size_t)c * 4 * plane + row_off;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                pv[c][k] = add2(pv[c][k], v[k]);
                st_pair(o + k * plane + g.rh + x, pv[c][k], both);
            }
            if (edge) {
                // mirrored copies for the halo columns pass B will read (BORDER_REFLECT: -1-j <-> j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int xe = x + e;
                    if (e == 1 && !both) break;
                    const int xl = xe < g.rh ? g.rh - 1 - xe : -1;                    // padded column of the left mirror
                    const int xr = xe >= g.w - g.rh ? g.rh + 2 * g.w - 1 - xe : -1;  // padded column of the right mirror
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        float lo, hi;
                        unpack2(pv[c][k], lo, hi);
                        const float val = e == 0 ? lo : hi;
                        if (xl >= 0) o[k * plane + xl] = val;
                        if (xr >= 0 && xr < g.wp) o[k * plane + xr] = val;
                    }
                }
            }
        }
    };
    // SRC_ONLY: the statistics of this thread's pixel pair, from the slot of step t
    auto load_stats = [&](const int t, f2 *gs) {
        const f2 *p = reinterpret_cast<const f2 *>(ring + (t % RING) * SLOT_WORDS + 2 * NP * NX) + tid;
#pragma unroll
        for (int j = 0; j < 9; ++j) gs[j] = p[j * NT];
    };

    // Step t: the row y0 - r + t enters the window; from t = 2r on, output row y = y0 + t - 2r is emitted and the
    // row y - r leaves afterwards; from t = 2r + 1 on, row y - 1 is solved (SRC_ONLY: with its cached statistics,
    // which travel in the same slot; the last row is solved in a step of its own, t = n_steps).  Warp 0 requests
    // the data of step t (warp-uniform code, elected lane).
    const int n_steps = 2 * r + (y1 - y0);
    const int n_req = MODE == SRC_ONLY ? n_steps + 1 : n_steps;
    auto request = [&](const int t) {
        const int s = t % RING;
        if (t >= RING) mbar_wait(bars + 8 * (RING + s), (uint32_t)(t / RING - 1) & 1u);  // every warp released use t - RING
        const bool has_in = t < n_steps;
        const bool has_out = t >= 2 * r && has_in;
        const bool has_gs = MODE == SRC_ONLY && t > 2 * r;
        const int yin = reflect(y0 - r + t, g.h);
        const int yout = has_out ? reflect(y0 + t - 3 * r, g.h) : 0;
        if (elect_one()) {
            const uint32_t full = bars + 8 * s;
            const uint32_t dst = smem_u32(ring + s * SLOT_WORDS);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full),
                         "r"((has_in ? ROW_BYTES : 0u) + (has_out ? ROW_BYTES : 0u) + (has_gs ? GS_BYTES : 0u))
                         : "memory");
            if (has_in) tma_load_3d(dst, &tmap, sx0 / 2, yin, img * NP, full);
            if (has_out) tma_load_3d(dst + ROW_BYTES, &tmap, sx0 / 2, yout, img * NP, full);
            if (has_gs) tma_load_3d(dst + 2 * ROW_BYTES, &tmap_gs, sx0 / 2, y0 + t - 2 * r - 1, img * 9, full);
        }
    };
    auto release = [&](const int t) {
        __syncwarp();
        if (elect_one())
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bars + 8 * (RING + t % RING)) : "memory");
    };
    if (group == 0)
        for (int t = 0; t < RING - 1 && t < n_req; ++t) request(t);

    // Software-pipelined by one row: between two barriers a warp scans row y AND solves row y-1 (whose prefixes
    // all warps stored before the previous barrier), so the shuffle-latency-bound scan overlaps the arithmetic
    // of the solve.
    const bool r_is_odd = (r & 1) != 0;
    f2 gs[MODE == SRC_ONLY ? 9 : 1];
    for (int t = 0; t < n_steps; ++t) {
        if (group == 0 && t + RING - 1 < n_req) request(t + RING - 1);
        const uint32_t *slot = ring + (t % RING) * SLOT_WORDS;
        mbar_wait(bars + 8 * (t % RING), (uint32_t)(t / RING) & 1u);
        {
            const RowRaw<SC, C> row_in = read_row<SC, C>(slot, lane);
            if (t < 2 * r) release(t);
            if (MODE == SRC_ONLY) {
                RF_GF2_DISPATCH_SRC(accumulate, V, row_in, 1.0f)
            } else {
                RF_GF2_DISPATCH(accumulate, V, row_in, 1.0f)
            }
        }
        if (t < 2 * r) continue;
        const int y = y0 + t - 2 * r;
        uint32_t *P = pbuf + ((y - y0) & 1) * (Q * NX);
        if (MODE == SRC_ONLY) {
            RF_GF2_DISPATCH_SRC(scan_store, V, P, NX, lane)
        } else {
            RF_GF2_DISPATCH(scan_store, V, P, NX, lane)
        }
        {
            const RowRaw<SC, C> row_out = read_row<SC, C>(slot + NP * NX, lane);
            if (MODE != SRC_ONLY) release(t);
            if (MODE == SRC_ONLY) {
                RF_GF2_DISPATCH_SRC(accumulate, V, row_out, -1.0f)
            } else {
                RF_GF2_DISPATCH(accumulate, V, row_out, -1.0f)
            }
        }
        if (y > y0) {
            const uint32_t *Pm = pbuf + ((y - 1 - y0) & 1) * (Q * NX);
            if (MODE == SRC_ONLY) load_stats(t, gs);
            if (r_is_odd)
                math_row(std::true_type{}, y - 1, Pm, gs);
            else
                math_row(std::false_type{}, y - 1, Pm, gs);
        }
        if (MODE == SRC_ONLY) release(t);
        // One barrier per row: the prefixes of row y are visible to everyone after it, and everyone has
        // finished reading the other buffer (row y-1), which the next step overwrites.
        __syncthreads();
    }
    if (MODE == SRC_ONLY) {
        mbar_wait(bars + 8 * (n_steps % RING), (uint32_t)(n_steps / RING) & 1u);
        load_stats(n_steps, gs);
    }
    {
        const uint32_t *Pm = pbuf + ((y1 - 1 - y0) & 1) * (Q * NX);
        if (r_is_odd)
            math_row(std::true_type{}, y1 - 1, Pm, gs);
        else
            math_row(std::false_type{}, y1 - 1, Pm, gs);
    }
}

// ---- pass B ---------------------------------------------------------------------------------------
// One warp per (output row, coefficient plane), RB rows per CTA.  No state is carried from row to row.
template <int SC, int RB>
__host__ __device__ constexpr int pass_b_warps() { return 4 * SC * RB; }

// q = mean(a) . I + mean(b) for the pixel pairs first_idx, first_idx + idx_step, ... of output row y of a strip, from the
// horizontal prefixes of the four vertical window sums (plane k at Prow + k * pstride); writes uint8 to dst and / or into
// the packed planes (the input of the next iteration, mirrored halo columns included).
template <int SC, bool R_ODD>
__device__ __forceinline__ void solve_row(const Args &g, const float *Prow, int pstride, int img, int y, int sx0,
                                          int first_idx, int idx_step)
{
    constexpr int NP = SC == 1 ? 1 : 2;
    const int r = g.r;
    const size_t plane = (size_t)g.h * g.wp;
    const size_t img_px = (size_t)g.h * g.w;
    const uint32_t *PK = g.packed + (size_t)img * NP * plane + (size_t)y * g.wp;
    const int n_out = min(g.twe_b, g.w - sx0);
    const f2 ia2 = dup2(g.inv_area);
    const bool w_even = (g.w & 1) == 0;
for (int idx = first_idx; idx < n_out; idx += idx_step) {
        const bool both = idx + 1 < n_out;
        const int i = g.rh + idx;
        const int x = sx0 + idx;
        const uint2 gw = *reinterpret_cast<const uint2 *>(PK + g.rh + x);
        const f2 i0 = b2f2(gw.x, gw.y, 0, false), i1 = b2f2(gw.x, gw.y, 1, false), i2 = b2f2(gw.x, gw.y, 2, false);
        uint32_t res0[SC], res1[SC];
#pragma unroll
        for (int c = 0; c < SC; ++c) {
            f2 m[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float *Pq = Prow + (4 * c + k) * pstride;
                // (P[i+r] - P[i-r-1], P[i+r+1] - P[i-r]); i is even, so one of the two pairs is an aligned 64-bit load
                f2 hi, lo;
                if (R_ODD) {
                    hi = pack2(Pq[i + r], Pq[i + r + 1]);
                    lo = *reinterpret_cast<const f2 *>(Pq + i - r - 1);
                } else {
                    hi = *reinterpret_cast<const f2 *>(Pq + i + r);
                    lo = pack2(Pq[i - r - 1], Pq[i - r]);
                }
                m[k] = mul2(sub2(hi, lo), ia2);
            }
            f2 v = m[3];
            v = add2(v, mul2(m[0], i0));
            v = add2(v, mul2(m[1], i1));
            v = add2(v, mul2(m[2], i2));
            float v0, v1;
            unpack2(v, v0, v1);
            res0[c] = sat_u8(v0);
            res1[c] = sat_u8(v1);
        }
        if (g.store_dst) {
            uint8_t *o = g.dst + (img * img_px + (size_t)y * g.w + x) * SC;
            if (SC == 1) {
                if (both && w_even)
                    *reinterpret_cast<uint16_t *>(o) = (uint16_t)(res0[0] | (res1[0] << 8));
                else {
                    o[0] = (uint8_t)res0[0];
                    if (both) o[1] = (uint8_t)res1[0];
                }
            } else {
                if (both && w_even) {
                    uint16_t *o2 = reinterpret_cast<uint16_t *>(o);
                    o2[0] = (uint16_t)(res0[0] | (res0[1 % SC] << 8));
                    o2[1] = (uint16_t)(res0[2 % SC] | (res1[0] << 8));
                    o2[2] = (uint16_t)(res1[1 % SC] | (res1[2 % SC] << 8));
                } else {
#pragma unroll
                    for (int c = 0; c < SC; ++c) o[c] = (uint8_t)res0[c];
                    if (both)
#pragma unroll
                        for (int c = 0; c < SC; ++c) o[SC + c] = (uint8_t)res1[c];
                }
            }
        }
        if (g.store_packed) {
            // the next iteration filters this output: put it where pack_kernel would have put it, mirrored
            // halo columns included.  Other CTAs of this launch read only the guide bytes of these words.
            uint32_t *row = g.packed + (size_t)img * NP * plane + (size_t)y * g.wp + (SC == 1 ? 0 : plane);
            const uint32_t w0 = SC == 1 ? (gw.x & 0x00FFFFFFu) | (res0[0] << 24)
                                        : res0[0] | (res0[1 % SC] << 8) | (res0[2 % SC] << 16);
            const uint32_t w1 = SC == 1 ? (gw.y & 0x00FFFFFFu) | (res1[0] << 24)
                                        : res1[0] | (res1[1 % SC] << 8) | (res1[2 % SC] << 16);
            if (both)
                *reinterpret_cast<uint2 *>(row + g.rh + x) = make_uint2(w0, w1);
            else
                row[g.rh + x] = w0;
            if (x < g.rh || x + 1 >= g.w - g.rh) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    if (e == 1 && !both) break;
                    const int xe = x + e;
                    const uint32_t wv = e == 0 ? w0 : w1;
                    const int xl = xe < g.rh ? g.rh - 1 - xe : -1;
                    const int xr = xe >= g.w - g.rh ? g.rh + 2 * g.w - 1 - xe : -1;
                    if (xl >= 0) row[xl] = wv;
                    if (xr >= 0 && xr < g.wp) row[xr] = wv;
                }
            }
        }
    }
}

// The whole issue path of pass B is warp-uniform: the warp index, the output row and the row plan are broadcast
// values, and the loads are issued by an elected lane.  (Issued from `if (lane == 0)` the compiler wrapped every
// cp.async.bulk.tensor in per-operand R2UR + vote loops: 60 of the 75 instructions of the term loop,
// profiles/r02_gf_v0_blocks.txt.)
template <int SC, int C, int RB, int NS>
__global__ void __launch_bounds__(32 * pass_b_warps<SC, RB>(), SC == 1 ? 8 / RB : 1)
    pass_b_kernel(const __grid_constant__ CUtensorMap tmap, const Args g)
{
    constexpr int Q = 4 * SC, NX = 32 * C, NW = Q * RB;
    constexpr uint32_t ROW_BYTES = NX * 4;  // one term: the warp's row chunk of one prefix plane
    extern __shared__ __align__(128) float slots[];  // [NW][NS][NX], then NW * NS mbarriers
    uint64_t *bars = reinterpret_cast<uint64_t *>(slots + NW * NS * NX);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int pl = warp % Q, rr = warp / Q;
    const int img = blockIdx.z;
    const int sx0 = blockIdx.y * g.twe_b;
    const int y = blockIdx.x * RB + rr;
    const int r = g.r;
    float *slot0 = slots + (warp * NS) * NX;

    if (y < g.h) {
        const uint32_t bar0 = smem_u32(bars + warp * NS);
        // this row's plan, one entry per lane: [0] = term count, [4 + t] = prefix row, [4 + MAXT + t] = weight
        const int mine = lane < PLAN_STRIDE ? __ldg(g.plan + (size_t)y * PLAN_STRIDE + lane) : 0;
        const int cnt = __shfl_sync(0xffffffffu, mine, 0);
        const int c0 = sx0 / 2, c2 = img * Q + pl;
        // two boxes per term: box = half a row chunk (NX/4 8-byte elements)
        auto issue = [&](int t) {
            const int row = __shfl_sync(0xffffffffu, mine, 4 + t);
            const uint32_t bar = bar0 + 8 * (t % NS);
            const uint32_t dst = smem_u32(slot0 + (t % NS) * NX);
            if (elect_one()) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(ROW_BYTES) : "memory");
                asm volatile(
                    "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                    ::"r"(dst), "l"(&tmap), "r"(c0), "r"(row), "r"(c2), "r"(bar)
                    : "memory");
                asm volatile(
                    "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                    ::"r"(dst + ROW_BYTES / 2), "l"(&tmap), "r"(c0 + NX / 4), "r"(row), "r"(c2), "r"(bar)
                    : "memory");
            }
        };
        if (elect_one()) {
#pragma unroll
            for (int k = 0; k < NS; ++k) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8 * k) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < NS; ++k)
            if (k < cnt) issue(k);
        float acc[C];
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = 0.0f;
        for (int t = 0; t < cnt; ++t) {
            const int s = t % NS;
            const float wgt = __int_as_float(__shfl_sync(0xffffffffu, mine, 4 + MAXT + t));
            mbar_wait(bar0 + 8 * s, (uint32_t)(t / NS) & 1u);
            const float4 *sp = reinterpret_cast<const float4 *>(slot0 + s * NX + lane * C);
            const f2 w2 = dup2(wgt);
#pragma unroll
            for (int c = 0; c < C; c += 4) {
                const float4 v = sp[c / 4];
                f2 a01 = pack2(acc[c], acc[c + 1]), a23 = pack2(acc[c + 2], acc[c + 3]);
                ffma2(a01, w2, pack2(v.x, v.y));  // weights are +-1 (+-2): exact products
                ffma2(a23, w2, pack2(v.z, v.w));
                unpack2(a01, acc[c], acc[c + 1]);
                unpack2(a23, acc[c + 2], acc[c + 3]);
            }
            __syncwarp();  // every lane has read the slot before the next term may overwrite it
            if (t + NS < cnt) issue(t + NS);
        }
        // horizontal prefix over the strip: serial over the lane's C columns, then one warp scan of the lane totals
        float run = 0.0f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            run += acc[c];
            acc[c] = run;
        }
        float incl = run;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const float tt = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += tt;
        }
        float excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 0.0f;
        // the prefixes replace the first slot (all of this warp's loads have landed and been consumed)
        float4 *dp = reinterpret_cast<float4 *>(slot0 + lane * C);
        const f2 e2 = dup2(excl);
#pragma unroll
        for (int c = 0; c < C; c += 4) {
            float p0, p1, p2, p3;
            unpack2(add2(pack2(acc[c], acc[c + 1]), e2), p0, p1);
            unpack2(add2(pack2(acc[c + 2], acc[c + 3]), e2), p2, p3);
            dp[c / 4] = make_float4(p0, p1, p2, p3);
        }
    }
    __syncthreads();
    if (y >= g.h) return;

    // the Q warps of a row share its pixels: one pixel pair per thread and step
    const float *Prow = slots + (rr * Q * NS) * NX;  // plane k of this row: Prow + k * NS * NX
    if (r & 1)
        solve_row<SC, true>(g, Prow, NS * NX, img, y, sx0, 2 * (pl * 32 + lane), 2 * 32 * Q);
    else
        solve_row<SC, false>(g, Prow, NS * NX, img, y, sx0, 2 * (pl * 32 + lane), 2 * 32 * Q);
}

// ---- host -------------------------------------------------------------------------------------------
struct Plan {
    int C, strips, twe;        // pass A strips
    int CB, strips_b, twe_b;   // pass B strips
    int rh, wp;
};

static int halo(int r) { return (r + 1 + 3) & ~3; }
// + 16: the chunk (<= 16 columns) that holds the right-most needed column must lie inside the row
static int pitch(int w, int r) { return (w + 2 * halo(r) + 16 + 3) & ~3; }

// tuning override (development only): RF_GF2_SEGS_A forces the number of row segments of pass A
static int env_int(const char *name)
{
    const char *e = getenv(name);
    return e ? atoi(e) : 0;
}

static Plan make_plan(int sc, int h, int w, int r)
{
    Plan best{};
    const int rh = halo(r);
    best.rh = rh;
    best.wp = pitch(w, r);
    long best_cost = -1;
    const int max_pairs = 32 * (sc == 1 ? 4 : 6);  // pass A solves one pixel pair per thread and row
    for (int C : {8, 12}) {  // wider chunks never pay: a strip emits at most 2 * threads columns
        int tw = 32 * C - 2 * rh;
        if (tw > 2 * max_pairs) tw = 2 * max_pairs;
        if (tw < 32) continue;
        const int strips = (w + tw - 1) / tw;
        int twe = ((w + strips - 1) / strips + 3) & ~3;
        if (twe > tw) twe = tw & ~3;
        const long cost = (long)strips * 32 * C;  // columns touched per image row
        if (best_cost < 0 || cost < best_cost) {
            best_cost = cost;
            best.C = C;
            best.strips = strips;
            best.twe = twe;
        }
    }
    best_cost = -1;
    for (int C : {12, 20}) {  // 48- and 80-byte lane chunks: conflict-free 128-bit shared-memory accesses
        const int tw = 32 * C - 2 * rh;
        if (tw < 32) continue;
        const int strips = (w + tw - 1) / tw;
        int twe = ((w + strips - 1) / strips + 3) & ~3;
        if (twe > tw) twe = tw & ~3;
        const long cost = (long)strips * 32 * C;
        if (best_cost < 0 || cost < best_cost) {
            best_cost = cost;
            best.CB = C;
            best.strips_b = strips;
            best.twe_b = twe;
        }
    }
    return best;
}

// does every row's window fit into MAXT prefix rows for this segmentation?  (always for h > r and segments of at
// least 64 rows; checked because the segment count depends on the batch size)
static bool plan_fits(int h, int r, int seg_rows)
{
    static std::mutex mu;
    static int last[4] = {0, 0, 0, 0};  // h, r, seg_rows, result
    std::lock_guard<std::mutex> lock(mu);
    if (last[0] == h && last[1] == r && last[2] == seg_rows) return last[3] != 0;
    bool ok = true;
    RowTerms t;
    for (int y = 0; y < h && ok; ++y) {
        row_terms(t, y, h, r, seg_rows);
        ok = t.cnt <= MAXT;
    }
    last[0] = h;
    last[1] = r;
    last[2] = seg_rows;
    last[3] = ok ? 1 : 0;
    return ok;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
static PFN_cuTensorMapEncodeTiled tensor_map_encoder()
{
    static PFN_cuTensorMapEncodeTiled fn = [] {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return (PFN_cuTensorMapEncodeTiled)f;
    }();
    return fn;
}

// 3-D map of 32-bit planes [planes][h][wp] as 8-byte elements: (wp / 2, h, planes); box = (box_x, 1, box_planes)
static int make_tensor_map(CUtensorMap *tm, void *base, int wp, int h, size_t planes, int box_x, int box_planes)
{
    PFN_cuTensorMapEncodeTiled enc = tensor_map_encoder();
    if (!enc) return fail(RF_ECUDA, "gf2: cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[3] = {(cuuint64_t)wp / 2, (cuuint64_t)h, (cuuint64_t)planes};
    const cuuint64_t strides[2] = {(cuuint64_t)wp * 4, (cuuint64_t)h * wp * 4};
    const cuuint32_t box[3] = {(cuuint32_t)box_x, 1, (cuuint32_t)box_planes};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult rc = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, base, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return fail(RF_ECUDA, "gf2: cuTensorMapEncodeTiled failed (CUresult %d)", (int)rc);
    return RF_OK;
}

template <int SC, int CB, int RB, int NS>
static int launch_b(const Args &a, const Plan &p, cudaStream_t st)
{
    constexpr int NW = pass_b_warps<SC, RB>(), NX = 32 * CB;
    constexpr size_t smem = (size_t)NW * NS * NX * sizeof(float) + (size_t)NW * NS * sizeof(uint64_t);
    static DeviceOnce once;
    int dev = 0;
    RF_CUDA_TRY(cudaGetDevice(&dev));
    {
        std::lock_guard<std::mutex> lock(once.mu);
        if (!once.done[dev & 63]) {
            RF_CUDA_TRY(cudaFuncSetAttribute(pass_b_kernel<SC, CB, RB, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            once.done[dev & 63] = true;
        }
    }
    CUtensorMap tm;
    const int rc = make_tensor_map(&tm, a.ab, a.wp, a.h, (size_t)a.n * 4 * SC, NX / 4, 1);
    if (rc != RF_OK) return rc;
    const dim3 grid((a.h + RB - 1) / RB, p.strips_b, a.n);
    pass_b_kernel<SC, CB, RB, NS><<<grid, 32 * NW, smem, st>>>(tm, a);
    RF_LAUNCH_CHECK("gf2::pass_b_kernel");
    return RF_OK;
}

template <int SC, int C>
static int launch(Args a, const Plan &p, int iterations, cudaStream_t st)
{
    constexpr int NX = 32 * C, NTA = 32 * n_groups<SC>();
    constexpr int NP = SC == 1 ? 1 : 2;
    const size_t smem_a = pass_a_smem<SC, C, FULL>();
    const size_t smem_s = pass_a_smem<SC, C, SRC_ONLY>();
    static DeviceOnce once;
    int dev = 0;
    RF_CUDA_TRY(cudaGetDevice(&dev));
    {
        std::lock_guard<std::mutex> lock(once.mu);
        if (!once.done[dev & 63]) {
            RF_CUDA_TRY(cudaFuncSetAttribute(pass_a_kernel<SC, C, FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            RF_CUDA_TRY(cudaFuncSetAttribute(pass_a_kernel<SC, C, FULL_STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            RF_CUDA_TRY(cudaFuncSetAttribute(pass_a_kernel<SC, C, SRC_ONLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            once.done[dev & 63] = true;
        }
    }
    // packed rows as TMA boxes: one strip chunk (NX pixels) of every plane of an image
    CUtensorMap tm_pk;
    {
        const int rc = make_tensor_map(&tm_pk, a.packed, a.wp, a.h, (size_t)a.n * NP, NX / 2, NP);
        if (rc != RF_OK) return rc;
    }
    // cached statistics of the pixel pairs a CTA solves per row: nine planes x 2 * threads pixels
    CUtensorMap tm_gs = tm_pk;
    if (iterations > 1) {
        const int rc = make_tensor_map(&tm_gs, a.gstat, a.wg, a.h, (size_t)a.n * 9, NTA, 9);
        if (rc != RF_OK) return rc;
    }
    dim3 pgrid((a.wp + 255) / 256, (a.h + PACK_ROWS - 1) / PACK_ROWS, a.n);
    pack_kernel<SC><<<pgrid, 256, 0, st>>>(a);
    RF_LAUNCH_CHECK("gf2::pack_kernel");
    // Row segments: the vertical sums make a column strictly sequential, so the rows are split into segments of
    // about 96 rows (each pays 2r warm-up rows; measured optimum for 64 x 512x384 and within 4 % of the best for
    // 8 x 4K, profiles/r02_gf_segments.txt).  The segmentation depends on the image height ONLY: the prefix rows
    // pass A writes restart at segment boundaries and their FP32 rounding depends on where they restart, so a
    // segmentation chosen from the batch size would make the bytes of an image depend on how a batch is sharded
    // or chunked.  One segmentation for every iteration: the row plan of pass B is built for exactly it.
    static const int force_a = env_int("RF_GF2_SEGS_A");
    int segs = force_a > 0 ? force_a : (a.h + 48) / 96;
    if (segs < 1) segs = 1;
    a.seg_rows = (a.h + segs - 1) / segs;
    if (!plan_fits(a.h, a.r, a.seg_rows)) a.seg_rows = a.h;  // one segment: at most two terms per reflected copy
    const dim3 grid_a(p.strips, (a.h + a.seg_rows - 1) / a.seg_rows, a.n);
    plan_kernel<<<(a.h + 127) / 128, 128, 0, st>>>(const_cast<int *>(a.plan), a.h, a.r, a.seg_rows);
    RF_LAUNCH_CHECK("gf2::plan_kernel");
    for (int it = 0; it < iterations; ++it) {
        const bool last = it == iterations - 1;
        if (it == 0) {
            if (iterations > 1)
                pass_a_kernel<SC, C, FULL_STORE><<<grid_a, NTA, smem_a, st>>>(tm_pk, tm_gs, a);
            else
                pass_a_kernel<SC, C, FULL><<<grid_a, NTA, smem_a, st>>>(tm_pk, tm_gs, a);
        } else {
            pass_a_kernel<SC, C, SRC_ONLY><<<grid_a, NTA, smem_s, st>>>(tm_pk, tm_gs, a);
        }
        RF_LAUNCH_CHECK("gf2::pass_a_kernel");
        a.store_dst = last ? 1 : 0;
        a.store_packed = last ? 0 : 1;
        // one output row per CTA, two slots per warp: measured best of RB in {1, 2, 4} x NS in {2, 3, 4} (89 us against
        // 106 / 138 us for 2 / 4 rows per CTA and 95 / 114 us for 3 / 4 slots at 64 x 512x384: resident CTAs hide the
        // plan -> TMA -> scan -> solve latency chain, not deeper prefetch within one.  A persistent, warp-specialised
        // form -- one producer warp keeping a ring of prefix rows full for four consumer warps per CTA -- was built and
        // measured in round 2: byte-identical, but 111 / 134 / 187 us with 3 / 4 / 8 ring slots (16 / 12 / 8 consumer
        // warps per SM): the scan + solve of a row need the 36 resident warps more than the loads need a deeper ring)
        const int rc = p.CB == 12 ? launch_b<SC, 12, 1, 2>(a, p, st) : launch_b<SC, 20, 1, 2>(a, p, st);
        if (rc != RF_OK) return rc;
    }
    return RF_OK;
}

// single reflection must cover the halo, the exact-integer FP32 sums need r <= 64
// Small windows go to the generic kernels (exact float(double(S) / k^2) means): with a handful of samples per
// window the covariance is often near-singular and the 1-ulp mean of this path gets amplified (measured on 'flat'
// guides, r = 1, eps <= 0.05: 0.5 % of bytes move by 1 LSB and single bytes by 2; the generic path is bit-equal to
// the oracle there).  From r = 8 up the deviation is below 1e-3 of the bytes, +-1 LSB.
constexpr int MIN_RADIUS = 8;
bool supported(int r, int h, int w) { return r >= MIN_RADIUS && r <= MAX_RADIUS && w >= halo(r) && h > r && h <= 65535; }

static size_t plan_bytes(int h) { return ((size_t)h * PLAN_STRIDE * sizeof(int) + 15) & ~(size_t)15; }

// packed planes + prefix planes (+ nine guide-statistics planes when the filter is iterated) + the row plan (needed
// once per call; counted per image so that any chunk of the batch has room for it)
size_t workspace_per_image(int sc, int h, int w, int r, int iterations)
{
    const size_t plane = (size_t)h * pitch(w, r);
    const size_t gplane = (size_t)h * ((w + 3) & ~3);  // statistics planes carry no halo
    return plane * 4 * (sc == 1 ? 1 : 2) + plane * 4 * 4 * sc + (iterations > 1 ? gplane * 4 * 9 : 0) + plan_bytes(h);
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
    const Plan p = make_plan(sc, h, w, r);
    a.rh = p.rh;
    a.wp = p.wp;
    a.wg = (w + 3) & ~3;
    a.twe = p.twe;
    a.twe_b = p.twe_b;
    a.seg_rows = h;
    a.store_dst = 1;
    a.store_packed = 0;
    const size_t plane = (size_t)h * p.wp;
    a.packed = (uint32_t *)ws;
    a.ab = (float *)((uint32_t *)ws + (size_t)n * (sc == 1 ? 1 : 2) * plane);
    a.gstat = iterations > 1 ? a.ab + (size_t)n * sc * 4 * plane : nullptr;
    a.plan = (const int *)(a.ab + (size_t)n * sc * 4 * plane + (iterations > 1 ? (size_t)n * 9 * h * a.wg : 0));
#define RF_GF2_LAUNCH(SC_)                                            \
    switch (p.C) {                                                    \
        case 8: return launch<SC_, 8>(a, p, iterations, st);          \
        default: return launch<SC_, 12>(a, p, iterations, st);        \
    }
    if (sc == 1) { RF_GF2_LAUNCH(1) }
    RF_GF2_LAUNCH(3)
#undef RF_GF2_LAUNCH
}

}  // namespace gf2
}  // namespace rf

// row plan of pass B for one output row (diagnostic entry point: lets the CPU-side tests check the window
// decomposition against a brute-force sum).  Returns the number of terms (> 12 = does not fit).
extern "C" int rf_guided_row_terms(int h, int radius, int seg_rows, int y, int *rows, float *weights)
{
    if (h < 1 || radius < 0 || seg_rows < 1 || y < 0 || y >= h || !rows || !weights) return -1;
    rf::gf2::RowTerms t;
    rf::gf2::row_terms(t, y, h, radius, seg_rows);
    for (int i = 0; i < t.cnt && i < rf::gf2::MAXT; ++i) {
        rows[i] = t.row[i];
        weights[i] = t.wgt[i];
    }
    return t.cnt;
}
