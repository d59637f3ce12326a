// Per-pixel MLP ("CNN" of 1x1 convolutions) forward for sm_100a, exact-FP32 CUDA-core version.
//
// Replaces caffe.Net(...).forward() at /root/reference/decompose_with_trained_CNN.py:82-95 on the
// graph /root/reference/network_definition.prototxt:17-165, with the input transform
// (decompose...py:57-69, image_utils.py:32-39: /255, BGR->RGB, sRGB->linear) fused in as an exact
// 256-entry table and the optional uint8 output (image_utils.py:68, truncation) fused out.
//
// Each thread carries two adjacent pixels through all layers in registers; activations of the two
// pixels are the two halves of 64-bit registers so every MAC pair is one packed FFMA2 with the
// (warp-uniform) weight broadcast from shared memory by LDS.128.  Summation order is Caffe's:
// dot product over input channels in ascending order, then + bias, then ReLU; the fusing layer is
// accumulated layer by layer in concat order.  See DESIGN.md "K2" for the roofline.
#include <cmath>
#include <vector>

#include "common.cuh"

struct rf_cnn {
    int device;
    int n_hidden;
    int width;       // uniform hidden width C
    int n_params;    // floats in d_params
    float *d_params; // device copy of the caller's parameter block
    float *d_lut;    // 256 floats
};

namespace rf {
namespace cnn {

constexpr int THREADS = 128;
constexpr int MAX_HIDDEN = 8;

template <int C>
__device__ __forceinline__ void layer_first(const float *W, const float *b, const unsigned long long x[3],
                                            unsigned long long h[C])
{
#pragma unroll
    for (int o = 0; o < C; ++o) {
        unsigned long long acc = 0ull;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float w = W[o * 3 + k];
            ffma2(acc, pack2(w, w), x[k]);
        }
        float lo, hi;
        unpack2(acc, lo, hi);
        const float bo = b[o];
        h[o] = pack2(fmaxf(lo + bo, 0.0f), fmaxf(hi + bo, 0.0f));
    }
}

template <int C>
__device__ __forceinline__ void layer_hidden(const float *W, const float *b, const unsigned long long a[C],
                                             unsigned long long h[C])
{
#pragma unroll
    for (int o = 0; o < C; ++o) {
        unsigned long long acc = 0ull;
        const float4 *Wr = reinterpret_cast<const float4 *>(W + o * C);
#pragma unroll
        for (int k4 = 0; k4 < C / 4; ++k4) {
            const float4 w = Wr[k4];
            ffma2(acc, pack2(w.x, w.x), a[4 * k4 + 0]);
            ffma2(acc, pack2(w.y, w.y), a[4 * k4 + 1]);
            ffma2(acc, pack2(w.z, w.z), a[4 * k4 + 2]);
            ffma2(acc, pack2(w.w, w.w), a[4 * k4 + 3]);
        }
        float lo, hi;
        unpack2(acc, lo, hi);
        const float bo = b[o];
        h[o] = pack2(fmaxf(lo + bo, 0.0f), fmaxf(hi + bo, 0.0f));
    }
}

template <int C>
__device__ __forceinline__ void fuse_accumulate(const float *fw, const unsigned long long h[C],
                                                unsigned long long &z)
{
#pragma unroll
    for (int o = 0; o < C; ++o) {
        const float w = fw[o];
        ffma2(z, pack2(w, w), h[o]);
    }
}

__device__ __forceinline__ float sigmoid_caffe(float z) { return __fdiv_rn(1.0f, 1.0f + expf(-z)); }

// params layout in shared memory = the caller's block: [W0 | b0 | W1 | b1 | ... | fuse_w | fuse_b].
// The first layer's W0 (C x 3) is not 16-byte friendly; it is read as scalars.
template <int C>
__global__ void __launch_bounds__(THREADS) mlp_kernel(const float *__restrict__ params, int n_params,
                                                      int n_hidden, const float *__restrict__ lut,
                                                      const uint8_t *__restrict__ bgr, size_t n_px,
                                                      float *__restrict__ out_f32, uint8_t *__restrict__ out_u8)
{
    extern __shared__ __align__(16) float sm[];
    float *lut_s = sm;          // 256
    float *par_s = sm + 256;    // n_params (16-byte aligned: 256 floats precede it)
    for (int i = threadIdx.x; i < 256; i += THREADS) lut_s[i] = lut[i];
    for (int i = threadIdx.x; i < n_params; i += THREADS) par_s[i] = params[i];
    __syncthreads();

    const size_t n_pairs = (n_px + 1) / 2;
    const float *fuse_w = par_s + (3 * C + C) + (size_t)(n_hidden - 1) * (C * C + C);
    for (size_t pair = (size_t)blockIdx.x * THREADS + threadIdx.x; pair < n_pairs;
         pair += (size_t)gridDim.x * THREADS) {
        const size_t p0 = 2 * pair;
        const bool has1 = p0 + 1 < n_px;
        const uint8_t *px = bgr + 3 * p0;
        uint8_t c[6];
#pragma unroll
        for (int i = 0; i < 3; ++i) c[i] = px[i];
#pragma unroll
        for (int i = 3; i < 6; ++i) c[i] = has1 ? px[i] : px[i - 3];
        // BGR -> RGB and sRGB -> linear through the table
        unsigned long long x[3];
        x[0] = pack2(lut_s[c[2]], lut_s[c[5]]);
        x[1] = pack2(lut_s[c[1]], lut_s[c[4]]);
        x[2] = pack2(lut_s[c[0]], lut_s[c[3]]);

        unsigned long long a[C], h[C];
        unsigned long long z = 0ull;
        layer_first<C>(par_s, par_s + 3 * C, x, a);
        fuse_accumulate<C>(fuse_w, a, z);
        const float *q = par_s + 3 * C + C;
        for (int l = 1; l < n_hidden; ++l) {
            layer_hidden<C>(q, q + C * C, a, h);
            fuse_accumulate<C>(fuse_w + l * C, h, z);
#pragma unroll
            for (int o = 0; o < C; ++o) a[o] = h[o];
            q += C * C + C;
        }
        float z0, z1;
        unpack2(z, z0, z1);
        const float fb = fuse_w[n_hidden * C];
        const float r0 = sigmoid_caffe(z0 + fb), r1 = sigmoid_caffe(z1 + fb);
        if (out_f32) {
            out_f32[p0] = r0;
            if (has1) out_f32[p0 + 1] = r1;
        }
        if (out_u8) {
            // image_utils.py:68: (image * 255).astype(np.uint8) -- float32 product, truncation
            out_u8[p0] = (uint8_t)__float2int_rz(__fmul_rn(r0, 255.0f));
            if (has1) out_u8[p0 + 1] = (uint8_t)__float2int_rz(__fmul_rn(r1, 255.0f));
        }
    }
}

template <int C>
static int launch(const rf_cnn *net, const uint8_t *bgr, size_t n_px, float *out_f32, uint8_t *out_u8,
                  cudaStream_t st)
{
    const size_t smem = (256 + (size_t)net->n_params) * sizeof(float);
    static DeviceOnce once;
    {
        std::lock_guard<std::mutex> lock(once.mu);
        if (!once.done[net->device & 63]) {
            RF_CUDA_TRY(cudaFuncSetAttribute(mlp_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            once.done[net->device & 63] = true;
        }
    }
    const size_t n_pairs = (n_px + 1) / 2;
    size_t blocks = (n_pairs + THREADS - 1) / THREADS;
    const size_t cap = (size_t)sm_count() * 8;  // persistent-ish: each CTA loads the weights once
    if (blocks > cap) blocks = cap;
    mlp_kernel<C><<<(unsigned)blocks, THREADS, smem, st>>>(net->d_params, net->n_params, net->n_hidden, net->d_lut,
                                                          bgr, n_px, out_f32, out_u8);
    RF_LAUNCH_CHECK("mlp_kernel");
    return RF_OK;
}

}  // namespace cnn
}  // namespace rf

namespace rf {
namespace cnntc {  // cnn_tc.cu: tcgen05 / TMEM kernel for hidden width 32
bool supported(int width, int n_hidden);
bool disabled_by_env();
int launch(const float *d_params, int n_hidden, const float *d_lut, const uint8_t *bgr, size_t n_px, float *out_f32,
           uint8_t *out_u8, cudaStream_t st);
}  // namespace cnntc
}  // namespace rf

using namespace rf;

extern "C" int rf_cnn_create(const float *params, const int *dims, int n_hidden, const float *srgb_lut256,
                             rf_cnn **out)
{
    if (!params || !dims || !out) return fail(RF_EINVAL, "rf_cnn_create: NULL argument");
    if (n_hidden < 1 || n_hidden > cnn::MAX_HIDDEN)
        return fail(RF_EUNSUPPORTED, "rf_cnn_create: n_hidden %d outside 1..%d", n_hidden, cnn::MAX_HIDDEN);
    if (dims[0] != 3) return fail(RF_EINVAL, "rf_cnn_create: the network input must have 3 channels, got %d", dims[0]);
    const int C = dims[1];
    for (int i = 1; i <= n_hidden; ++i)
        if (dims[i] != C)
            return fail(RF_EUNSUPPORTED, "rf_cnn_create: hidden widths must be uniform (layer %d has %d, layer 1 has %d)",
                        i, dims[i], C);
    if (!(C == 8 || C == 16 || C == 32 || C == 64))
        return fail(RF_EUNSUPPORTED, "rf_cnn_create: hidden width %d not in {8,16,32,64}", C);
    int n_params = 3 * C + C + (n_hidden - 1) * (C * C + C) + n_hidden * C + 1;
    std::vector<float> lut(256);
    if (srgb_lut256) {
        for (int v = 0; v < 256; ++v) lut[v] = srgb_lut256[v];
    } else {
        for (int v = 0; v < 256; ++v) {  // image_utils.py:32-39 in double, stored as float32
            const double s = v / 255.0;
            lut[v] = (float)(s <= 0.04045 ? s / 12.92 : std::pow((s + 0.055) / 1.055, 2.4));
        }
    }
    rf_cnn *net = new (std::nothrow) rf_cnn();
    if (!net) return fail(RF_ENOMEM, "rf_cnn_create: out of host memory");
    net->n_hidden = n_hidden;
    net->width = C;
    net->n_params = n_params;
    net->d_params = nullptr;
    net->d_lut = nullptr;
    cudaError_t e = cudaGetDevice(&net->device);
    if (e == cudaSuccess) e = cudaMalloc(&net->d_params, n_params * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&net->d_lut, 256 * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(net->d_params, params, n_params * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(net->d_lut, lut.data(), 256 * sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaFree(net->d_params);
        cudaFree(net->d_lut);
        delete net;
        return fail(RF_ECUDA, "rf_cnn_create: %s", cudaGetErrorString(e));
    }
    *out = net;
    return RF_OK;
}

extern "C" void rf_cnn_destroy(rf_cnn *net)
{
    if (!net) return;
    cudaFree(net->d_params);
    cudaFree(net->d_lut);
    delete net;
}

extern "C" int rf_cnn_forward_u8(const rf_cnn *net, const uint8_t *bgr, int n, int h, int w, float *out_f32,
                                 uint8_t *out_u8, void *stream)
{
    if (!net || !bgr) return fail(RF_EINVAL, "rf_cnn_forward_u8: NULL argument");
    if (!out_f32 && !out_u8) return fail(RF_EINVAL, "rf_cnn_forward_u8: no output requested");
    if (n < 0 || h < 1 || w < 1) return fail(RF_EINVAL, "rf_cnn_forward_u8: bad shape n=%d h=%d w=%d", n, h, w);
    if (n == 0) return RF_OK;
    int dev = -1;
    RF_CUDA_TRY(cudaGetDevice(&dev));
    if (dev != net->device)
        return fail(RF_EINVAL, "rf_cnn_forward_u8: model lives on device %d, current device is %d", net->device, dev);
    const size_t n_px = (size_t)n * h * w;
    cudaStream_t st = (cudaStream_t)stream;
    if (cnntc::supported(net->width, net->n_hidden) && !cnntc::disabled_by_env())
        return cnntc::launch(net->d_params, net->n_hidden, net->d_lut, bgr, n_px, out_f32, out_u8, st);
    switch (net->width) {
        case 8: return cnn::launch<8>(net, bgr, n_px, out_f32, out_u8, st);
        case 16: return cnn::launch<16>(net, bgr, n_px, out_f32, out_u8, st);
        case 32: return cnn::launch<32>(net, bgr, n_px, out_f32, out_u8, st);
        case 64: return cnn::launch<64>(net, bgr, n_px, out_f32, out_u8, st);
    }
    return fail(RF_EINVAL, "rf_cnn_forward_u8: corrupt handle");
}
