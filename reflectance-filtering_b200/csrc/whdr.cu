// WHDR (weighted human disagreement rate, Bell et al. 2014) of reflectance images on the device.
//
// Replaces the Python loop of /root/reference/training/layers/whdr_layer.py:253-287 (whdr), with
// _lightness (:182-198) and the comparison blob of createNumpyArrayWithComparisonsForIIW.py:616-649 /
// _extract_valid_comparisons_with_actual_size (whdr_layer.py:240-251):
//
//   blob [n][max+1][6] float64: row c < count = (x1, y1, x2, y2, darker, weight) with x, y in [0,1) relative
//   coordinates, darker in {0 = 'E', 1, 2}; the last row holds (count, file name, 0); unused rows are NaN.
//   pixel coordinates are int(x * width), int(y * height)  (float64 product, truncation)
//   lightness L = max(float32 eps, mean over channels) in float32 (np.mean of a float32 vector: ((r0+r1)+r2)/3)
//   judgement: l2/l1 > 1+delta -> 1 ; l1/l2 > 1+delta -> 2 ; else 0          (float32 ratios)
//   per image: error_sum / weight_sum over the comparisons (0 when there are none)
//
// One CTA per image; every thread takes a strided subset of the comparisons and the float64 partial sums are
// combined by a fixed-shape tree, so results are run-to-run deterministic (the sequential Python sum can
// differ from the tree in the last bits of the double).
#include "common.cuh"

namespace rf {
namespace whdr {

constexpr int THREADS = 256;

__global__ void __launch_bounds__(THREADS)
whdr_kernel(const float *__restrict__ refl, int c, int h, int w, const double *__restrict__ blob, int max_cmp,
            float one_plus_delta, bool pixel_coords, double *__restrict__ out, int *__restrict__ bad)
{
    const int img = blockIdx.x;
    const double *cmp = blob + (size_t)img * (max_cmp + 1) * 6;
    const float *r = refl + (size_t)img * c * h * w;
    const size_t plane = (size_t)h * w;
    const double cnt_d = cmp[(size_t)max_cmp * 6];
    int count = cnt_d >= 0.0 && cnt_d <= (double)max_cmp ? (int)cnt_d : -1;
    if (count < 0) {  // NaN or out of range meta row
        if (threadIdx.x == 0) {
            atomicOr(bad, 1);
            out[2 * img] = 0.0;
            out[2 * img + 1] = 0.0;
        }
        return;
    }
    const float eps = 1.1920928955078125e-07f;  // np.finfo(np.float32).eps
    double err = 0.0, wsum = 0.0;
    for (int i = threadIdx.x; i < count; i += THREADS) {
        const double *q = cmp + (size_t)i * 6;
        // res[:, [0, 2]] = (res[:, [0, 2]] * width).astype(int): float64 product, truncation
        const double sx = pixel_coords ? 1.0 : (double)w, sy = pixel_coords ? 1.0 : (double)h;
        const double fx1 = q[0] * sx, fy1 = q[1] * sy, fx2 = q[2] * sx, fy2 = q[3] * sy;
        const bool finite = fabs(fx1) < 2e9 && fabs(fy1) < 2e9 && fabs(fx2) < 2e9 && fabs(fy2) < 2e9;  // false for NaN
        const int x1 = finite ? (int)fx1 : -1, y1 = finite ? (int)fy1 : -1;
        const int x2 = finite ? (int)fx2 : -1, y2 = finite ? (int)fy2 : -1;
        const int darker = (int)q[4];
        const double weight = q[5];
        if ((unsigned)x1 >= (unsigned)w || (unsigned)x2 >= (unsigned)w || (unsigned)y1 >= (unsigned)h ||
            (unsigned)y2 >= (unsigned)h) {
            atomicOr(bad, 2);  // numpy would raise IndexError
            continue;
        }
        const size_t o1 = (size_t)y1 * w + x1, o2 = (size_t)y2 * w + x2;
        float l1, l2;
        if (c == 3) {
            l1 = __fdiv_rn(__fadd_rn(__fadd_rn(r[o1], r[plane + o1]), r[2 * plane + o1]), 3.0f);
            l2 = __fdiv_rn(__fadd_rn(__fadd_rn(r[o2], r[plane + o2]), r[2 * plane + o2]), 3.0f);
        } else {
            l1 = r[o1];
            l2 = r[o2];
        }
        l1 = fmaxf(eps, l1);
        l2 = fmaxf(eps, l2);
        int alg = 0;
        if (__fdiv_rn(l2, l1) > one_plus_delta)
            alg = 1;
        else if (__fdiv_rn(l1, l2) > one_plus_delta)
            alg = 2;
        if (alg != darker) err += weight;
        wsum += weight;
    }
    __shared__ double s_err[THREADS], s_w[THREADS];
    s_err[threadIdx.x] = err;
    s_w[threadIdx.x] = wsum;
    __syncthreads();
    for (int d = THREADS / 2; d > 0; d >>= 1) {
        if (threadIdx.x < d) {
            s_err[threadIdx.x] += s_err[threadIdx.x + d];
            s_w[threadIdx.x] += s_w[threadIdx.x + d];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[2 * img] = s_err[0];
        out[2 * img + 1] = s_w[0];
    }
}

}  // namespace whdr
}  // namespace rf

using namespace rf;

extern "C" int rf_whdr_f32(const float *reflectance, int c, int n, int h, int w, const double *comparisons,
                           int max_comparisons, double delta, unsigned flags, double *out_sums, int *bad_flag,
                           void *stream)
{
    if (!reflectance || !comparisons || !out_sums || !bad_flag) return fail(RF_EINVAL, "rf_whdr_f32: NULL pointer");
    if (c != 1 && c != 3) return fail(RF_EINVAL, "rf_whdr_f32: expecting 1 or 3 channels to compute lightness, got %d", c);
    if (n < 0 || h < 1 || w < 1 || max_comparisons < 0) return fail(RF_EINVAL, "rf_whdr_f32: bad shape");
    if (!(delta >= 0.0)) return fail(RF_EINVAL, "rf_whdr_f32: delta must be >= 0");
    if (n == 0) return RF_OK;
    whdr::whdr_kernel<<<n, whdr::THREADS, 0, (cudaStream_t)stream>>>(reflectance, c, h, w, comparisons, max_comparisons,
                                                                    (float)(1.0 + delta), (flags & RF_WHDR_PIXEL_COORDS) != 0,
                                                                    out_sums, bad_flag);
    RF_LAUNCH_CHECK("whdr::whdr_kernel");
    return RF_OK;
}
