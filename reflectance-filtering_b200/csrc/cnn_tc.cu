// Per-pixel MLP forward on the 5th-generation tensor cores (tcgen05 + TMEM), 3xTF32 split operands.
//
// Replaces caffe.Net(...).forward() at /root/reference/decompose_with_trained_CNN.py:82-95 for the
// shipped graph family (uniform hidden width 32; network_definition.prototxt:17-165).  All convolutions
// are 1x1, so each hidden layer is D[128 px x 32] = A[128 px x 32] * W^T[32 x 32] per tile of 128 pixels.
//
//   * every layer runs as tcgen05.mma kind::tf32 (M=128, N=32, K=8) with the accumulator in TMEM: conv0 as
//     K = 8 over (r, g, b, 1, 0...) from the exact sRGB->linear table, conv1..conv4 as K = 40 = 32 activations +
//     one K-step whose A block is the constant (1, 0 x 7) and whose weight row is the bias -- so the
//     accumulator already holds Caffe's dot + bias.  Only the 160 -> 1 fusing layer stays on the CUDA cores,
//     as a running packed (FFMA2) dot product in every epilogue;
//   * single TF32 misses the 1e-3 tolerance (7e-3), so operands are split x = hi + lo with
//     hi = x truncated to TF32 (what the tensor core reads anyway) and lo = x - hi, and each layer issues
//     A_hi*W_hi + A_hi*W_lo + A_lo*W_hi: 3 MMAs for conv0, 5 + 5 + 4 per hidden layer, 59 per tile; measured
//     max error vs the FP32 oracle ~8e-6;
//   * activations never leave the SM and never touch shared memory: epilogue = tcgen05.ld (thread t of
//     warp w owns TMEM lane 32w+t = pixel t of the tile, all 32 outputs) -> ReLU, fuse FFMA2, split (one LOP3 +
//     half an FADD2 per value) -> tcgen05.st back into tensor memory as the A operand of the next layer's MMAs
//     (A-from-TMEM form); only the weight planes are read from shared memory (canonical K-major no-swizzle
//     UMMA layout);
//   * the MMA issue path is warp-uniform (broadcast values + elect.sync): 14 back-to-back UTCHMMA per layer.
//     A first version issued from `if (thread == 0)` and spent ~16 instructions of per-operand R2UR
//     broadcasts per MMA while 127 threads waited (1.02 -> 0.76 ms);
//   * what bounds it: the per-layer sync chain (st -> barrier -> issue -> commit -> mbarrier -> ld) is latency
//     that only OTHER tile pipelines hide (measured: 2 pipelines per SM 0.97 ms, 4: 0.69 ms, 5: 0.60 ms; -61 % MMAs
//     = -10 % time, -32 % epilogue instructions = -3 %, a quarter of the accumulator read = no change).  Tensor
//     memory (512 columns) bounds the pipelines: 96 columns each (D 32, A_hi 32, A_lo 32; conv0's A overlays
//     A_hi) + ONE constant block shared by all = five warpgroup pipelines in one persistent CTA per SM.
// SASS: UTCHMMA / LDTM / STTM / UTCBAR (tcgen05.mma / .ld / .st / .commit).  The exact-FP32 kernel in cnn.cu
// remains for other widths and as the in-library cross-check (RF_CNN_FP32=1).
#include <cstdlib>

#include "common.cuh"

namespace rf {
namespace cnntc {

constexpr int TILE_M = 128;
constexpr int CW = 32;  // hidden width
constexpr int WGS = 5;            // warpgroups per CTA (one CTA per SM), each runs its own tile pipeline
constexpr int THREADS = 128 * WGS;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// canonical K-major, no swizzle: 8-row x 16-byte core matrices; rows of a core matrix 16 B apart,
// core matrices 128 B apart along M/N (SBO) and `lbo` bytes apart along K
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
    return d;                // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}

// instruction descriptor: D=F32, A=B=TF32, both K-major, N=32, M=128
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((CW >> 3) << 17) | ((TILE_M >> 4) << 24);

// A operand from tensor memory (lanes = rows of the tile, one 32-bit column per TF32 element)
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(IDESC), "r"(accumulate), "r"(0u)
        : "memory");
}

// this thread's TMEM lane, 32 consecutive columns
__device__ __forceinline__ void tmem_store32(uint32_t taddr, const float (&v)[32])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]),
        "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]), "f"(v[16]), "f"(v[17]), "f"(v[18]),
        "f"(v[19]), "f"(v[20]), "f"(v[21]), "f"(v[22]), "f"(v[23]), "f"(v[24]), "f"(v[25]), "f"(v[26]), "f"(v[27]),
        "f"(v[28]), "f"(v[29]), "f"(v[30]), "f"(v[31])
        : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    unsigned long long spins = 0;
    while (!done) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (++spins > (1ull << 26)) __trap();  // never hang the GPU on a lost completion
    }
}

__device__ __forceinline__ float tf32_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// packed two-lane FP32 add on a 64-bit register pair (sm_100 FADD2)
__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b)
{
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// -(x truncated to TF32): one LOP3 ((x & mask) ^ sign)
__device__ __forceinline__ float neg_tf32_hi(float x)
{
    return __uint_as_float((__float_as_uint(x) & 0xFFFFE000u) ^ 0x80000000u);
}
// one lane of a converged warp
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// this thread's TMEM lane, 8 consecutive columns
__device__ __forceinline__ void tmem_store8(uint32_t taddr, float v0, float v1, float v2, float v3, float v4, float v5,
                                            float v6, float v7)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "f"(v0),
                 "f"(v1), "f"(v2), "f"(v3), "f"(v4), "f"(v5), "f"(v6), "f"(v7)
                 : "memory");
}

constexpr int K0 = 8;                     // conv0 as an MMA: K = (r, g, b, 1, 0, 0, 0, 0), the 1 carries the bias
constexpr int KH = 40;                    // hidden layers: 32 activations, the constant 1 (bias row), 7 x padding
constexpr int B0_BYTES = CW * K0 * 4;     // 1 KB per conv0 weight plane
constexpr int BH_BYTES = CW * KH * 4;     // 5 KB per hidden weight plane

// Tensor-memory columns.  A layer's sync chain (st -> barrier -> MMA issue -> commit -> mbarrier -> ld) is latency
// that only other pipelines can hide (measured: 2 pipelines per SM 0.97 ms, 4 pipelines 0.69 ms), and tensor
// memory (512 columns) bounds their number: 96 columns per pipeline + ONE constant block shared by all.
constexpr int T_D = 0;       // [0, 32)    accumulator
constexpr int T_AH = 32;     // [32, 64)   A_hi: activations (the tensor core truncates them to TF32)
constexpr int T_AL = 64;     // [64, 96)   A_lo: what the truncation drops
constexpr int T_XH = 32;     // [32, 40)   conv0 A_hi: (r, g, b, 1, 0 x 4) linear RGB   (before the first activations land)
constexpr int T_XL = 40;     // [40, 48)   conv0 A_lo
constexpr int T_COLS = 96;
constexpr int T_ONE = WGS * T_COLS;  // [480, 488) the K-step that carries the bias row: (1, 0 x 7) in every lane

// shared memory map (bytes)
struct Smem {
    int b0, b, fw, lut_hi, lut_lo, bar, slot, total;
};
__host__ __device__ inline Smem smem_map(int n_hidden)
{
    Smem s;
    s.b0 = 0;                                          // conv0: hi plane, lo plane
    s.b = s.b0 + 2 * B0_BYTES;                         // per hidden MMA layer: hi plane, lo plane
    s.fw = s.b + (n_hidden - 1) * 2 * BH_BYTES;        // fusing weights n_hidden x 32, then the fuse bias
    s.lut_hi = s.fw + (n_hidden * CW + 4) * 4;         // sRGB -> linear table and its TF32 remainder
    s.lut_lo = s.lut_hi + 256 * 4;
    s.bar = s.lut_lo + 256 * 4;                        // one 8-byte mbarrier per warpgroup
    s.slot = s.bar + 8 * WGS;
    s.total = s.slot + 8;
    return s;
}

// byte offset of element (n, k) in a canonical K-major no-swizzle plane with 32 rows (N)
__device__ __forceinline__ int plane_off(int n, int k) { return (k >> 2) * 512 + (n >> 3) * 128 + (n & 7) * 16 + (k & 3) * 4; }

__global__ void __launch_bounds__(THREADS) mlp_tc_kernel(const float *__restrict__ params, int n_hidden,
                                                         const float *__restrict__ lut_g,
                                                         const uint8_t *__restrict__ bgr, size_t n_px,
                                                         float *__restrict__ out_f32, uint8_t *__restrict__ out_u8)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    const Smem sm = smem_map(n_hidden);
    const int tid = threadIdx.x, warp = tid >> 5, wg = tid >> 7, wtid = tid & 127;
    float *fw = reinterpret_cast<float *>(smem + sm.fw);
    float *lut_hi = reinterpret_cast<float *>(smem + sm.lut_hi);
    float *lut_lo = reinterpret_cast<float *>(smem + sm.lut_lo);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + sm.bar) + wg;
    uint32_t *slot = reinterpret_cast<uint32_t *>(smem + sm.slot);

    // ---- stage the model: parameter block is [W0 | b0 | W1 | b1 | ... | fuse_w | fuse_b] ----------------
    // Every layer becomes a pair of K-major planes W[n][k] (the word itself: the tensor core reads its TF32
    // truncation; and the remainder), with the bias as the extra row k = 3 (conv0) / k = 32 (hidden).
    for (int i = tid; i < CW * K0; i += THREADS) {
        const int n = i / K0, k = i % K0;
        const float wv = k < 3 ? params[3 * n + k] : (k == 3 ? params[96 + n] : 0.0f);
        *reinterpret_cast<float *>(smem + sm.b0 + plane_off(n, k)) = wv;
        *reinterpret_cast<float *>(smem + sm.b0 + B0_BYTES + plane_off(n, k)) = tf32_lo(wv);
    }
    for (int i = tid; i < 256; i += THREADS) {
        const float v = lut_g[i];
        lut_hi[i] = v;
        lut_lo[i] = tf32_lo(v);
    }
    {
        const float *q = params + 96 + CW;
        for (int l = 1; l < n_hidden; ++l) {
            uint8_t *hi = smem + sm.b + (l - 1) * 2 * BH_BYTES, *lo = hi + BH_BYTES;
            for (int i = tid; i < CW * KH; i += THREADS) {
                const int n = i / KH, k = i % KH;
                const float wv = k < CW ? q[n * CW + k] : (k == CW ? q[CW * CW + n] : 0.0f);
                *reinterpret_cast<float *>(hi + plane_off(n, k)) = wv;
                *reinterpret_cast<float *>(lo + plane_off(n, k)) = tf32_lo(wv);
            }
            q += CW * CW + CW;
        }
        for (int i = tid; i < n_hidden * CW + 1; i += THREADS) fw[i] = q[i];
    }
    if (wtid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // weight planes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // warp-uniform copies (broadcast from lane 0) so that the MMA issue path stays on the uniform datapath
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *slot, 0);
    const int wg_u = __shfl_sync(0xffffffffu, tid >> 7, 0);
    const bool issuer_warp = __shfl_sync(0xffffffffu, warp & 3, 0) == 0;
    const uint32_t tmem = tmem_base + wg_u * T_COLS;
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;  // a warp may only touch its own 32 lanes
    const uint32_t bar_s = smem_u32(reinterpret_cast<uint64_t *>(smem + sm.bar) + wg_u);
    // shared-memory matrix descriptors: LBO 512 (next 4 k), SBO 128 (next 8 n), version 1; + (address >> 4)
    const uint64_t desc_base = make_desc(0, 512, 128);
    const uint32_t b0_s = smem_u32(smem + sm.b0), b_s = smem_u32(smem + sm.b);
    const float fb = fw[n_hidden * CW];
    // the constant 1 (bias row) and its padding: written once by warpgroup 0 for all 128 lanes, read by every
    // pipeline's MMAs
    if (wg_u == 0) {
        tmem_store8(tmem_base + T_ONE + lane_sel, 1.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    uint32_t parity = 0;
    const size_t n_tiles = (n_px + TILE_M - 1) / TILE_M;
    const size_t tile0 = (size_t)blockIdx.x * WGS + wg, tile_step = (size_t)gridDim.x * WGS;
    auto load_px = [&](size_t tile, uint32_t &c0, uint32_t &c1, uint32_t &c2) {
        const size_t p = tile * TILE_M + wtid;
        const uint8_t *px = bgr + 3 * (p < n_px ? p : n_px - 1);
        c0 = px[0];
        c1 = px[1];
        c2 = px[2];
    };
    uint32_t cb = 0, cg = 0, cr = 0;
    if (tile0 < n_tiles) load_px(tile0, cb, cg, cr);
    for (size_t tile = tile0; tile < n_tiles; tile += tile_step) {
        const size_t p = tile * TILE_M + wtid;
        const bool valid = p < n_px;
        // BGR -> RGB, sRGB -> linear (exact table), split for the tensor core
        tmem_store8(tmem + T_XH + lane_sel, lut_hi[cr], lut_hi[cg], lut_hi[cb], 1.0f, 0.0f, 0.0f, 0.0f, 0.0f);
        tmem_store8(tmem + T_XL + lane_sel, lut_lo[cr], lut_lo[cg], lut_lo[cb], 0.0f, 0.0f, 0.0f, 0.0f, 0.0f);
        // the next tile's pixel arrives while this one runs through the layers
        if (tile + tile_step < n_tiles) load_px(tile + tile_step, cb, cg, cr);

        // the fusing dot product runs as two interleaved partial sums (even / odd channels) in one 64-bit pair
        unsigned long long z2 = pack2(0.0f, 0.0f);
        for (int l = 0; l < n_hidden; ++l) {
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("bar.sync %0, 128;" ::"r"(1 + wg_u) : "memory");  // this warpgroup only
            if (issuer_warp) {
                if (elect_one()) {
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (l == 0) {
                        // D = X_hi W0_hi + X_hi W0_lo + X_lo W0_hi   (K = 8: one instruction each)
                        const uint64_t dh = desc_base | (uint64_t)((b0_s >> 4) & 0x3FFF);
                        const uint64_t dl = desc_base | (uint64_t)(((b0_s + B0_BYTES) >> 4) & 0x3FFF);
                        mma_tf32_ts(tmem + T_D, tmem + T_XH, dh, 0);
                        mma_tf32_ts(tmem + T_D, tmem + T_XH, dl, 1);
                        mma_tf32_ts(tmem + T_D, tmem + T_XL, dh, 1);
                    } else {
                        // D = A_hi W_hi + A_hi W_lo (K = 40: activations and the bias row) + A_lo W_hi (K = 32)
                        const uint32_t bh = b_s + (l - 1) * 2 * BH_BYTES, bl = bh + BH_BYTES;
                        const uint64_t dh = desc_base | (uint64_t)((bh >> 4) & 0x3FFF);
                        const uint64_t dl = desc_base | (uint64_t)((bl >> 4) & 0x3FFF);
#pragma unroll
                        for (int j = 0; j < KH / 8; ++j)  // next K step: two core-matrix columns = 1024 B = 64 units
                            mma_tf32_ts(tmem + T_D, j < CW / 8 ? tmem + T_AH + 8 * j : tmem_base + T_ONE,
                                        dh + (uint64_t)(j * 64), j != 0);
#pragma unroll
                        for (int j = 0; j < KH / 8; ++j)
                            mma_tf32_ts(tmem + T_D, j < CW / 8 ? tmem + T_AH + 8 * j : tmem_base + T_ONE,
                                        dl + (uint64_t)(j * 64), 1);
#pragma unroll
                        for (int j = 0; j < CW / 8; ++j)
                            mma_tf32_ts(tmem + T_D, tmem + T_AL + 8 * j, dh + (uint64_t)(j * 64), 1);
                    }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_s)
                                 : "memory");
                }
                __syncwarp();
            }
            mbar_wait(bar_s, parity);
            parity ^= 1;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t acc[CW];
            const uint32_t taddr = tmem + T_D + lane_sel;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(acc[0]), "=r"(acc[1]), "=r"(acc[2]), "=r"(acc[3]), "=r"(acc[4]), "=r"(acc[5]), "=r"(acc[6]),
                  "=r"(acc[7]), "=r"(acc[8]), "=r"(acc[9]), "=r"(acc[10]), "=r"(acc[11]), "=r"(acc[12]), "=r"(acc[13]),
                  "=r"(acc[14]), "=r"(acc[15]), "=r"(acc[16]), "=r"(acc[17]), "=r"(acc[18]), "=r"(acc[19]),
                  "=r"(acc[20]), "=r"(acc[21]), "=r"(acc[22]), "=r"(acc[23]), "=r"(acc[24]), "=r"(acc[25]),
                  "=r"(acc[26]), "=r"(acc[27]), "=r"(acc[28]), "=r"(acc[29]), "=r"(acc[30]), "=r"(acc[31])
                : "r"(taddr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            // the accumulator already holds dot + bias (Caffe order); ReLU, fuse FMA (packed)
            float h[CW];
            const float4 *fl = reinterpret_cast<const float4 *>(fw + l * CW);
#pragma unroll
            for (int o = 0; o < CW; o += 4) {
                const float4 f4 = fl[o >> 2];
                h[o] = fmaxf(__uint_as_float(acc[o]), 0.0f);
                h[o + 1] = fmaxf(__uint_as_float(acc[o + 1]), 0.0f);
                h[o + 2] = fmaxf(__uint_as_float(acc[o + 2]), 0.0f);
                h[o + 3] = fmaxf(__uint_as_float(acc[o + 3]), 0.0f);
                ffma2(z2, pack2(f4.x, f4.y), pack2(h[o], h[o + 1]));
                ffma2(z2, pack2(f4.z, f4.w), pack2(h[o + 2], h[o + 3]));
            }
            if (l + 1 < n_hidden) {
                // activations -> tensor memory as the next layer's A: A_hi = the word itself, A_lo = what the
                // TF32 truncation drops (packed: h + (-(h & mask)) on both halves of a register pair)
                float lo[CW];
#pragma unroll
                for (int o = 0; o < CW; o += 2) {
                    const unsigned long long l2 =
                        fadd2(pack2(h[o], h[o + 1]), pack2(neg_tf32_hi(h[o]), neg_tf32_hi(h[o + 1])));
                    unpack2(l2, lo[o], lo[o + 1]);
                }
                tmem_store32(tmem + T_AH + lane_sel, h);
                tmem_store32(tmem + T_AL + lane_sel, lo);
            }
        }
        float z_even, z_odd;
        unpack2(z2, z_even, z_odd);
        const float z = z_even + z_odd;
        const float r = __fdiv_rn(1.0f, 1.0f + expf(-(z + fb)));
        if (valid) {
            if (out_f32) out_f32[p] = r;
            if (out_u8) out_u8[p] = (uint8_t)__float2int_rz(__fmul_rn(r, 255.0f));
        }
        // the next tile's A writes and MMAs are ordered behind this tile's TMEM reads by the
        // before_thread_sync fence + warpgroup barrier at the top of its first MMA layer
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

bool supported(int width, int n_hidden) { return width == CW && n_hidden >= 2 && n_hidden <= 8; }

bool disabled_by_env()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("RF_CNN_FP32");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

int launch(const float *d_params, int n_hidden, const float *d_lut, const uint8_t *bgr, size_t n_px, float *out_f32,
           uint8_t *out_u8, cudaStream_t st)
{
    const Smem sm = smem_map(n_hidden);
    const size_t smem = (size_t)sm.total + 1024;  // slack for the 1024-byte alignment of the base
    static DeviceOnce once;
    int dev = 0;
    RF_CUDA_TRY(cudaGetDevice(&dev));
    {
        std::lock_guard<std::mutex> lock(once.mu);
        if (!once.done[dev & 63]) {
            RF_CUDA_TRY(cudaFuncSetAttribute(mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            once.done[dev & 63] = true;
        }
    }
    const size_t n_tiles = (n_px + TILE_M - 1) / TILE_M;
    // one CTA per SM: it allocates all 512 tensor-memory columns for its WGS pipelines
    size_t blocks = (size_t)sm_count();
    if (blocks > (n_tiles + WGS - 1) / WGS) blocks = (n_tiles + WGS - 1) / WGS;
    mlp_tc_kernel<<<(unsigned)blocks, THREADS, smem, st>>>(d_params, n_hidden, d_lut, bgr, n_px, out_f32, out_u8);
    RF_LAUNCH_CHECK("mlp_tc_kernel");
    return RF_OK;
}

}  // namespace cnntc
}  // namespace rf
