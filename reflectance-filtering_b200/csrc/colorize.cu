// Colorized side outputs of the decomposition on the device.
//
// Replaces, for a batch of images, what /root/reference/decompose_with_trained_CNN.py:122-128 does on the
// host with numpy:  colorize (image_utils.py:76-81)  ->  imwrite(..., sRGB=True) = normalize by the 99.9th
// percentile with 'lower' selection when max > 1 (image_utils.py:84-92), the reference's rgb_to_srgb
// (1.055 inside the power, image_utils.py:42-49) and truncation to uint8 (image_utils.py:68).
// The reference computes all of this in float64 (np.mean of uint8 -> float64, float64 / float32), so the
// kernels do too: it is a few hundred FP64 operations per pixel once per image, not a hot loop.
//
//   values kernel : shading = (b+g+r)/3 / r_intensity ; reflectance_c = c / max(shading, eps)  (float64 planes
//                   in the workspace) + per-image maxima
//   select        : exact k-th smallest (k = floor((n-1) * 0.999), supplied by the host with numpy's own
//                   arithmetic) by a 6-pass most-significant-digit radix select over the IEEE bit patterns
//                   (all values are >= 0, so unsigned order == numeric order); 2048-bin histograms,
//                   privatised in shared memory
//   write kernel  : x / percentile, clip, quirky gamma, (uint8)(x * 255)
#include "common.cuh"

namespace rf {
namespace colorize {

constexpr int BINS = 2048;  // 11-bit digits
constexpr int PASSES = 6;   // 11 * 6 = 66 >= 64

struct Problem {          // one per (image, output): output 0 = reflectance (3*h*w values), 1 = shading (h*w)
    unsigned long long prefix;   // bits decided so far (MSB side)
    unsigned long long k;        // rank still to find inside the current prefix bucket
    unsigned long long maxbits;  // bit pattern of the maximum value
    double percentile;           // result
};

__device__ __forceinline__ int digit_shift(int pass) { return 64 - 11 * (pass + 1) > 0 ? 64 - 11 * (pass + 1) : 0; }
__device__ __forceinline__ int digit_bits(int pass) { return pass < 5 ? 11 : 9; }

__global__ void values_kernel(const uint8_t *__restrict__ bgr, const float *__restrict__ intensity, int n, size_t hw,
                              double eps, double *__restrict__ refl, double *__restrict__ shad, Problem *prob)
{
    const int img = blockIdx.y;
    unsigned long long mr = 0, ms = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (size_t)gridDim.x * blockDim.x) {
        const uint8_t *px = bgr + (img * hw + i) * 3;
        const double b = px[0], g = px[1], r = px[2];
        const double norm_input = (b + g + r) / 3.0;              // np.mean(image, axis=2)
        const double s = norm_input / (double)intensity[img * hw + i];
        const double den = fmax(s, eps);
        const double rb = b / den, rg = g / den, rr = r / den;
        shad[img * hw + i] = s;
        double *o = refl + (img * hw + i) * 3;
        o[0] = rb;
        o[1] = rg;
        o[2] = rr;
        ms = max(ms, (unsigned long long)__double_as_longlong(s));
        mr = max(mr, (unsigned long long)__double_as_longlong(fmax(rb, fmax(rg, rr))));
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        mr = max(mr, __shfl_down_sync(0xffffffffu, mr, d));
        ms = max(ms, __shfl_down_sync(0xffffffffu, ms, d));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&prob[2 * img].maxbits, mr);
        atomicMax(&prob[2 * img + 1].maxbits, ms);
    }
}

// histogram of the current digit over the elements that match the decided prefix
__global__ void hist_kernel(const double *__restrict__ refl, const double *__restrict__ shad, size_t hw, int pass,
                            const Problem *__restrict__ prob, unsigned int *__restrict__ hist)
{
    __shared__ unsigned int sh[BINS];
    const int p = blockIdx.y;  // problem index
    const int img = p >> 1;
    const bool is_shading = p & 1;
    const unsigned long long *v = reinterpret_cast<const unsigned long long *>(is_shading ? shad + img * hw : refl + img * hw * 3);
    const size_t len = is_shading ? hw : hw * 3;
    for (int i = threadIdx.x; i < BINS; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const int shift = digit_shift(pass), bits = digit_bits(pass);
    const unsigned long long prefix = prob[p].prefix;
    const int decided = 11 * pass;  // number of MSBs fixed so far
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned long long x = v[i];
        const bool match = decided == 0 || (x >> (64 - decided)) == (prefix >> (64 - decided));
        if (match) atomicAdd(&sh[(unsigned int)(x >> shift) & ((1u << bits) - 1)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < BINS; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[(size_t)p * BINS + i], sh[i]);
}

// one block per problem: pick the bin that holds rank k, extend the prefix, clear the histogram
__global__ void pick_kernel(int pass, Problem *prob, unsigned int *hist)
{
    const int p = blockIdx.x;
    unsigned int *h = hist + (size_t)p * BINS;
    if (threadIdx.x == 0) {
        unsigned long long k = prob[p].k, cum = 0;
        int bin = 0;
        const int nb = 1 << digit_bits(pass);
        for (; bin < nb; ++bin) {
            if (cum + h[bin] > k) break;
            cum += h[bin];
        }
        if (bin == nb) bin = nb - 1;
        prob[p].k = k - cum;
        prob[p].prefix |= (unsigned long long)bin << digit_shift(pass);
        if (pass == PASSES - 1) prob[p].percentile = __longlong_as_double((long long)prob[p].prefix);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < BINS; i += blockDim.x) h[i] = 0;
}

__device__ __forceinline__ uint8_t finish(double x, bool scale, double pct)
{
    if (scale) {
        x = x / pct;                       // img /= np.percentile(img, 99.9, method='lower')
        x = fmin(fmax(x, 0.0), 1.0);       // np.clip(img, 0, 1)
    }
    // reference rgb_to_srgb: 1.055 inside the power (image_utils.py:48)
    const double y = x <= 0.0031308 ? x * 12.92 : pow(1.055 * x, 1.0 / 2.4) - 0.055;
    return (uint8_t)(int)(y * 255.0);       // (image * 255).astype(np.uint8): truncation
}

__global__ void write_kernel(const double *__restrict__ refl, const double *__restrict__ shad, size_t hw,
                             const Problem *__restrict__ prob, uint8_t *__restrict__ out_refl,
                             uint8_t *__restrict__ out_shad)
{
    const int img = blockIdx.y;
    const Problem pr = prob[2 * img], ps = prob[2 * img + 1];
    const bool scale_r = __longlong_as_double((long long)pr.maxbits) > 1.0;  // normalize only if max > 1
    const bool scale_s = __longlong_as_double((long long)ps.maxbits) > 1.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (size_t)gridDim.x * blockDim.x) {
        const double *r = refl + (img * hw + i) * 3;
        uint8_t *o = out_refl + (img * hw + i) * 3;
        o[0] = finish(r[0], scale_r, pr.percentile);
        o[1] = finish(r[1], scale_r, pr.percentile);
        o[2] = finish(r[2], scale_r, pr.percentile);
        out_shad[img * hw + i] = finish(shad[img * hw + i], scale_s, ps.percentile);
    }
}

__global__ void init_kernel(Problem *prob, int n, unsigned long long k_refl, unsigned long long k_shad, unsigned int *hist)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 2 * n) {
        prob[i].prefix = 0;
        prob[i].k = (i & 1) ? k_shad : k_refl;
        prob[i].maxbits = 0;
        prob[i].percentile = 0.0;
    }
    for (size_t j = i; j < (size_t)2 * n * BINS; j += (size_t)gridDim.x * blockDim.x) hist[j] = 0;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace colorize
}  // namespace rf

using namespace rf;

extern "C" size_t rf_colorize_workspace_bytes(int n, int h, int w)
{
    if (n < 1 || h < 1 || w < 1) return 0;
    const size_t hw = (size_t)h * w;
    return colorize::align_up(n * hw * 4 * sizeof(double), 256) + colorize::align_up(2 * n * sizeof(colorize::Problem), 256) +
           (size_t)2 * n * colorize::BINS * sizeof(unsigned int);
}

extern "C" int rf_colorize_u8(const uint8_t *bgr, const float *intensity, int n, int h, int w, double eps,
                              unsigned long long k_reflectance, unsigned long long k_shading, uint8_t *out_reflectance,
                              uint8_t *out_shading, void *ws, size_t ws_bytes, void *stream)
{
    if (!bgr || !intensity || !out_reflectance || !out_shading || !ws) return fail(RF_EINVAL, "rf_colorize_u8: NULL pointer");
    if (n < 0 || h < 1 || w < 1) return fail(RF_EINVAL, "rf_colorize_u8: bad shape n=%d h=%d w=%d", n, h, w);
    if (n == 0) return RF_OK;
    if (n > 32767) return fail(RF_EUNSUPPORTED, "rf_colorize_u8: more than 32767 images per call");
    if (ws_bytes < rf_colorize_workspace_bytes(n, h, w)) return fail(RF_EINVAL, "rf_colorize_u8: workspace too small");
    if ((uintptr_t)ws % 16) return fail(RF_EINVAL, "rf_colorize_u8: workspace must be 16-byte aligned");
    const size_t hw = (size_t)h * w;
    if (k_reflectance >= 3 * hw || k_shading >= hw) return fail(RF_EINVAL, "rf_colorize_u8: rank outside the image");
    cudaStream_t st = (cudaStream_t)stream;
    double *refl = (double *)ws;
    double *shad = refl + (size_t)n * hw * 3;
    colorize::Problem *prob = (colorize::Problem *)((char *)ws + colorize::align_up(n * hw * 4 * sizeof(double), 256));
    unsigned int *hist = (unsigned int *)((char *)prob + colorize::align_up(2 * n * sizeof(colorize::Problem), 256));

    const int sms = sm_count();
    colorize::init_kernel<<<(2 * n * colorize::BINS + 255) / 256 > 4 * sms ? 4 * sms : (2 * n * colorize::BINS + 255) / 256, 256, 0, st>>>(
        prob, n, k_reflectance, k_shading, hist);
    RF_LAUNCH_CHECK("colorize::init_kernel");
    unsigned bx = (unsigned)((hw + 255) / 256);
    if (bx > (unsigned)(8 * sms)) bx = 8 * sms;
    colorize::values_kernel<<<dim3(bx, n), 256, 0, st>>>(bgr, intensity, n, hw, eps, refl, shad, prob);
    RF_LAUNCH_CHECK("colorize::values_kernel");
    for (int pass = 0; pass < colorize::PASSES; ++pass) {
        colorize::hist_kernel<<<dim3(bx, 2 * n), 256, 0, st>>>(refl, shad, hw, pass, prob, hist);
        RF_LAUNCH_CHECK("colorize::hist_kernel");
        colorize::pick_kernel<<<2 * n, 256, 0, st>>>(pass, prob, hist);
        RF_LAUNCH_CHECK("colorize::pick_kernel");
    }
    colorize::write_kernel<<<dim3(bx, n), 256, 0, st>>>(refl, shad, hw, prob, out_reflectance, out_shading);
    RF_LAUNCH_CHECK("colorize::write_kernel");
    return RF_OK;
}
