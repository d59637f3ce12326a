/*
 * rf_oracle.c -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the checker for the CUDA path: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  Nothing under
 * reflectance-filtering_b200/ links, imports or calls it.
 *
 * What it restates (file:line into /root/reference, SURVEY.md Appendix A for the
 * un-vendored third parties whose arithmetic the reference calls):
 *   orc_srgb_lut            image_utils.py:32-39 (srgb_to_rgb) on the 256 codes v/255
 *                           that decompose_with_trained_CNN.py:57-69 can produce
 *   orc_mlp_forward_*       net.forward() at decompose_with_trained_CNN.py:90 on the graph
 *                           network_definition.prototxt:17-165 (Caffe CPU layers, A.1)
 *   orc_joint_bilateral_u8  cv2.ximgproc.jointBilateralFilter call at
 *                           filter_reflectance.py:60-64 (OpenCV-contrib 3.1.0
 *                           joint_bilateral_filter.cpp, jointBilateralFilter_8u, A.2)
 *   orc_joint_bilateral_u8_border / orc_joint_bilateral_f32   the rest of that function's surface (borderType
 *                           argument, CV_32F images: SURVEY 8f-4), not reachable from the reference CLI
 *   orc_box_mean_reflect    cv::boxFilter(CV_32F, normalize, BORDER_REFLECT) as used by
 *                           ximgproc's guided filter (A.3 step 2)
 *   orc_guided_u8           cv2.ximgproc.guidedFilter call at filter_reflectance.py:67-70
 *                           (guided_filter.cpp, colour guide, A.3; and its 1-channel-guide case)
 *
 * Pinning (see oracle/README.md and tests/test_oracle_pins.py): the MLP is checked against
 * cv2.dnn reading the reference's own prototxt+caffemodel, the bilateral filter is
 * bit-identical to cv2.bilateralFilter when joint == src, the box mean is checked against
 * cv2.boxFilter.  The guided filter as a whole is "parity unpinned" against a real
 * ximgproc binary (none is installable offline); its building block is pinned.
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_OK 0
#define ORC_EINVAL 1
#define ORC_ENOMEM 2

void orc_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---- borders (OpenCV borderInterpolate) --------------------------------- */
/* REFLECT_101: gfedcb|abcdefgh|gfedcba ; REFLECT: fedcba|abcdefgh|hgfedcb */
static int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) {
        if (p < 0) p = -p;
        else p = 2 * len - 2 - p;
    }
    return p;
}

static int reflect(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) {
        if (p < 0) p = -p - 1;
        else p = 2 * len - 1 - p;
    }
    return p;
}

/* cv::borderInterpolate for the border types of copyMakeBorder: 0 CONSTANT (returns -1: the padded value is 0),
 * 1 REPLICATE aaaaaa|abcdefgh|hhhhhhh, 2 REFLECT fedcba|abcdefgh|hgfedcb, 3 WRAP cdefgh|abcdefgh|abcdefg,
 * 4 REFLECT_101 gfedcb|abcdefgh|gfedcba (BORDER_DEFAULT)                                                     */
static int border_index(int p, int len, int border_type)
{
    if (p >= 0 && p < len) return p;
    switch (border_type) {
        case 0: return -1;
        case 1: return p < 0 ? 0 : len - 1;
        case 2: return reflect(p, len);
        case 3: {
            int q = p % len;
            return q < 0 ? q + len : q;
        }
        default: return reflect101(p, len);
    }
}

static uint8_t sat_u8(float v)
{
    /* saturate_cast<uchar>(float) = cvRound (round-half-even) then clamp */
    long r = lrintf(v);
    return (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
}

/* ---- sRGB -> linear LUT -------------------------------------------------- */
void orc_srgb_lut(float *lut256)
{
    for (int v = 0; v < 256; ++v) {
        double s = v / 255.0;
        double lin = (s <= 0.04045) ? s / 12.92 : pow((s + 0.055) / 1.055, 2.4);
        lut256[v] = (float)lin;
    }
}

/* ---- per-pixel MLP -------------------------------------------------------- */
/* params: for each hidden layer W[out][in] then b[out]; then fuse w[sum(out)], fuse b.
 * dims: n_hidden+1 entries (dims[0] == 3).  bgr: interleaved uint8 BGR pixels.
 * Caffe semantics: out = W*in (sgemm) ; out += b ; relu in place ; concat of the post-ReLU
 * maps in layer order ; fuse conv ; sigmoid 1/(1+exp(-z)).                                  */
#define ORC_MAXW 256

int orc_mlp_forward_f32(const uint8_t *bgr, long n_px, const float *params, const int *dims,
                        int n_hidden, const float *lut256, float *out)
{
    if (n_hidden < 1 || n_hidden > 16 || dims[0] != 3) return ORC_EINVAL;
    for (int l = 1; l <= n_hidden; ++l)
        if (dims[l] < 1 || dims[l] > ORC_MAXW) return ORC_EINVAL;
#pragma omp parallel for schedule(static)
    for (long p = 0; p < n_px; ++p) {
        float a[ORC_MAXW], h[ORC_MAXW];
        /* BGR -> RGB, /255 -> linear via the exact 256-entry table */
        a[0] = lut256[bgr[3 * p + 2]];
        a[1] = lut256[bgr[3 * p + 1]];
        a[2] = lut256[bgr[3 * p + 0]];
        const float *q = params;
        const float *fw = params;
        for (int l = 0; l < n_hidden; ++l) fw += dims[l] * dims[l + 1] + dims[l + 1];
        float z = 0.0f;
        for (int l = 0; l < n_hidden; ++l) {
            const int cin = dims[l], cout = dims[l + 1];
            const float *W = q, *b = q + cin * cout;
            for (int o = 0; o < cout; ++o) {
                float s = 0.0f;
                for (int k = 0; k < cin; ++k) s += W[o * cin + k] * a[k];
                s += b[o];
                h[o] = s > 0.0f ? s : 0.0f;
            }
            for (int o = 0; o < cout; ++o) {
                z += fw[o] * h[o];
                a[o] = h[o];
            }
            fw += cout;
            q += cin * cout + cout;
        }
        z += fw[0];
        out[p] = 1.0f / (1.0f + expf(-z));
    }
    return ORC_OK;
}

int orc_mlp_forward_f64(const uint8_t *bgr, long n_px, const float *params, const int *dims,
                        int n_hidden, double *out)
{
    if (n_hidden < 1 || n_hidden > 16 || dims[0] != 3) return ORC_EINVAL;
    for (int l = 1; l <= n_hidden; ++l)
        if (dims[l] < 1 || dims[l] > ORC_MAXW) return ORC_EINVAL;
#pragma omp parallel for schedule(static)
    for (long p = 0; p < n_px; ++p) {
        double a[ORC_MAXW], h[ORC_MAXW];
        for (int c = 0; c < 3; ++c) {
            double s = bgr[3 * p + 2 - c] / 255.0;
            a[c] = (s <= 0.04045) ? s / 12.92 : pow((s + 0.055) / 1.055, 2.4);
            a[c] = (double)(float)a[c]; /* the blob is float32 (decompose...py:88) */
        }
        const float *q = params;
        const float *fw = params;
        for (int l = 0; l < n_hidden; ++l) fw += dims[l] * dims[l + 1] + dims[l + 1];
        double z = 0.0;
        for (int l = 0; l < n_hidden; ++l) {
            const int cin = dims[l], cout = dims[l + 1];
            const float *W = q, *b = q + cin * cout;
            for (int o = 0; o < cout; ++o) {
                double s = 0.0;
                for (int k = 0; k < cin; ++k) s += (double)W[o * cin + k] * a[k];
                s += b[o];
                h[o] = s > 0.0 ? s : 0.0;
            }
            for (int o = 0; o < cout; ++o) {
                z += (double)fw[o] * h[o];
                a[o] = h[o];
            }
            fw += cout;
            q += cin * cout + cout;
        }
        z += fw[0];
        out[p] = 1.0 / (1.0 + exp(-z));
    }
    return ORC_OK;
}

/* image_utils.py:68 -- (image*255).astype(np.uint8): product in the array's dtype
 * (float32 here), truncation toward zero.                                           */
void orc_quantize_trunc(const float *x, long n, uint8_t *out)
{
    for (long i = 0; i < n; ++i) out[i] = (uint8_t)(x[i] * 255.0f);
}

/* ---- joint bilateral filter, 8-bit --------------------------------------- */
int orc_joint_bilateral_u8_border(const uint8_t *joint, int jc, const uint8_t *src, int sc, uint8_t *dst,
                                  int h, int w, double sigma_color, double sigma_space, int d, int border_type)
{
    if (!(jc == 1 || jc == 3) || !(sc == 1 || sc == 3) || h < 1 || w < 1) return ORC_EINVAL;
    if (border_type < 0 || border_type > 4) return ORC_EINVAL;
    if (sigma_color <= 0) sigma_color = 1;
    if (sigma_space <= 0) sigma_space = 1;
    int radius = d <= 0 ? (int)lrint(sigma_space * 1.5) : d / 2;
    if (radius < 1) radius = 1;
    const double gcc = -0.5 / (sigma_color * sigma_color);
    const double gsc = -0.5 / (sigma_space * sigma_space);

    float *cw = (float *)malloc(sizeof(float) * 256 * jc);
    const int side = 2 * radius + 1;
    float *sw = (float *)malloc(sizeof(float) * side * side);
    int *oi = (int *)malloc(sizeof(int) * side * side);
    int *oj = (int *)malloc(sizeof(int) * side * side);
    const int ph = h + 2 * radius, pw = w + 2 * radius;
    uint8_t *pj = (uint8_t *)malloc((size_t)ph * pw * jc);
    uint8_t *ps = (uint8_t *)malloc((size_t)ph * pw * sc);
    if (!cw || !sw || !oi || !oj || !pj || !ps) {
        free(cw); free(sw); free(oi); free(oj); free(pj); free(ps);
        return ORC_ENOMEM;
    }
    for (int i = 0; i < 256 * jc; ++i) cw[i] = (float)exp((double)i * i * gcc);
    int maxk = 0;
    for (int i = -radius; i <= radius; ++i)
        for (int j = -radius; j <= radius; ++j) {
            double r = sqrt((double)i * i + (double)j * j);
            if (r > radius) continue;
            sw[maxk] = (float)exp(r * r * gsc);
            oi[maxk] = i;
            oj[maxk] = j;
            ++maxk;
        }
    /* copyMakeBorder(..., borderType) of both images (the call site uses the default, BORDER_REFLECT_101;
     * BORDER_CONSTANT pads with 0) */
    for (int y = 0; y < ph; ++y) {
        const int sy = border_index(y - radius, h, border_type);
        for (int x = 0; x < pw; ++x) {
            const int sx = border_index(x - radius, w, border_type);
            const int inside = sy >= 0 && sx >= 0;
            for (int c = 0; c < jc; ++c)
                pj[((size_t)y * pw + x) * jc + c] = inside ? joint[((size_t)sy * w + sx) * jc + c] : 0;
            for (int c = 0; c < sc; ++c)
                ps[((size_t)y * pw + x) * sc + c] = inside ? src[((size_t)sy * w + sx) * sc + c] : 0;
        }
    }
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < h; ++y) {
        for (int x = 0; x < w; ++x) {
            const uint8_t *j0 = pj + ((size_t)(y + radius) * pw + (x + radius)) * jc;
            float wsum = 0.0f, s0 = 0.0f, s1 = 0.0f, s2 = 0.0f;
            for (int k = 0; k < maxk; ++k) {
                const size_t q = (size_t)(y + radius + oi[k]) * pw + (x + radius + oj[k]);
                const uint8_t *jk = pj + q * jc;
                int alpha = 0;
                for (int c = 0; c < jc; ++c) alpha += abs((int)j0[c] - (int)jk[c]);
                const float wt = sw[k] * cw[alpha];
                const uint8_t *sk = ps + q * sc;
                s0 += wt * sk[0];
                if (sc == 3) {
                    s1 += wt * sk[1];
                    s2 += wt * sk[2];
                }
                wsum += wt;
            }
            uint8_t *o = dst + ((size_t)y * w + x) * sc;
            o[0] = sat_u8(s0 / wsum);
            if (sc == 3) {
                o[1] = sat_u8(s1 / wsum);
                o[2] = sat_u8(s2 / wsum);
            }
        }
    }
    free(cw); free(sw); free(oi); free(oj); free(pj); free(ps);
    return ORC_OK;
}

int orc_joint_bilateral_u8(const uint8_t *joint, int jc, const uint8_t *src, int sc, uint8_t *dst,
                           int h, int w, double sigma_color, double sigma_space, int d)
{
    return orc_joint_bilateral_u8_border(joint, jc, src, sc, dst, h, w, sigma_color, sigma_space, d, 4);
}

/* ---- joint bilateral filter, CV_32F (not reachable from the reference CLI; SURVEY A.2 last bullet, 8f-4) ------
 * jointBilateralFilter_32f of joint_bilateral_filter.cpp [upstream, recollection], the structure of imgproc's
 * bilateralFilter_32f, against which the joint == src case is pinned (tests/test_oracle_pins.py):
 *   colorRange = (max - min over all joint values) * jc;  kExpNumBins = 4096 * jc;  scaleIndex = kExpNumBins / colorRange
 *   expLUT[i] = (float)exp((i / scaleIndex)^2 * (-0.5 / sigmaColor^2)),  i = 0 .. kExpNumBins + 1
 *   per tap: alpha = sum_c |J0_c - Jk_c| * scaleIndex; idx = (int)alpha; alpha -= idx;
 *            w = spaceWeight[k] * (expLUT[idx] + alpha * (expLUT[idx+1] - expLUT[idx]));  dst = sum(w src) / sum(w)
 * A constant joint image (max - min < FLT_EPSILON) makes every range weight 1: the spatial Gaussian over the disc
 * (upstream switches to a square GaussianBlur there [recollection]; imgproc copies the source -- both unverifiable
 * offline, so the degenerate case is defined by the formula itself and stated in DESIGN.md).                    */
int orc_joint_bilateral_f32(const float *joint, int jc, const float *src, int sc, float *dst, int h, int w,
                            double sigma_color, double sigma_space, int d, int border_type)
{
    if (!(jc == 1 || jc == 3) || !(sc == 1 || sc == 3) || h < 1 || w < 1) return ORC_EINVAL;
    if (border_type < 0 || border_type > 4) return ORC_EINVAL;
    if (sigma_color <= 0) sigma_color = 1;
    if (sigma_space <= 0) sigma_space = 1;
    int radius = d <= 0 ? (int)lrint(sigma_space * 1.5) : d / 2;
    if (radius < 1) radius = 1;
    const double gcc = -0.5 / (sigma_color * sigma_color);
    const double gsc = -0.5 / (sigma_space * sigma_space);
    float mn = joint[0], mx = joint[0];
    for (size_t i = 0; i < (size_t)h * w * jc; ++i) {
        if (joint[i] < mn) mn = joint[i];
        if (joint[i] > mx) mx = joint[i];
    }
    const int flat = fabs((double)mx - (double)mn) < 1.1920928955078125e-7;
    const int nbins = 4096 * jc;
    const float color_range = (float)((double)mx - (double)mn) * jc;
    const float scale_index = flat ? 0.0f : nbins / color_range;
    const int side = 2 * radius + 1;
    float *lut = (float *)malloc(sizeof(float) * (nbins + 2));
    float *sw = (float *)malloc(sizeof(float) * side * side);
    int *oi = (int *)malloc(sizeof(int) * side * side);
    int *oj = (int *)malloc(sizeof(int) * side * side);
    if (!lut || !sw || !oi || !oj) {
        free(lut); free(sw); free(oi); free(oj);
        return ORC_ENOMEM;
    }
    for (int i = 0; i < nbins + 2; ++i) {
        const double val = flat ? 0.0 : i / (double)scale_index;
        lut[i] = (float)exp(val * val * gcc);
    }
    int maxk = 0;
    for (int i = -radius; i <= radius; ++i)
        for (int j = -radius; j <= radius; ++j) {
            double r = sqrt((double)i * i + (double)j * j);
            if (r > radius) continue;
            sw[maxk] = (float)exp(r * r * gsc);
            oi[maxk] = i;
            oj[maxk] = j;
            ++maxk;
        }
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < h; ++y) {
        for (int x = 0; x < w; ++x) {
            const float *j0 = joint + ((size_t)y * w + x) * jc;
            float wsum = 0.0f, s0 = 0.0f, s1 = 0.0f, s2 = 0.0f;
            for (int k = 0; k < maxk; ++k) {
                const int yy = border_index(y + oi[k], h, border_type), xx = border_index(x + oj[k], w, border_type);
                const int inside = yy >= 0 && xx >= 0;
                const float zero[3] = {0.0f, 0.0f, 0.0f};
                const float *jk = inside ? joint + ((size_t)yy * w + xx) * jc : zero;
                const float *sk = inside ? src + ((size_t)yy * w + xx) * sc : zero;
                float alpha = 0.0f;
                for (int c = 0; c < jc; ++c) alpha += fabsf(j0[c] - jk[c]);
                alpha *= scale_index;
                int idx = (int)alpha;
                if (idx > nbins) idx = nbins;  /* a zero-padded border can exceed the image's own range */
                alpha -= (float)idx;
                const float wt = sw[k] * (lut[idx] + alpha * (lut[idx + 1] - lut[idx]));
                s0 += wt * sk[0];
                if (sc == 3) {
                    s1 += wt * sk[1];
                    s2 += wt * sk[2];
                }
                wsum += wt;
            }
            float *o = dst + ((size_t)y * w + x) * sc;
            o[0] = s0 / wsum;
            if (sc == 3) {
                o[1] = s1 / wsum;
                o[2] = s2 / wsum;
            }
        }
    }
    free(lut); free(sw); free(oi); free(oj);
    return ORC_OK;
}

/* ---- box mean with double running sums, BORDER_REFLECT ------------------- */
/* cv::boxFilter for CV_32F: RowSum<float,double> (sliding add/sub along the row)
 * then ColumnSum<double,float> (per-column running sum over 2r+1 rows), output
 * (float)(sum * scale), scale = 1.0/((2r+1)^2).                                      */
int orc_box_mean_reflect(const float *src, float *dst, int h, int w, int r)
{
    if (h < 1 || w < 1 || r < 0) return ORC_EINVAL;
    const int k = 2 * r + 1;
    const double scale = 1.0 / ((double)k * k);
    const int pw = w + 2 * r;
    double *rows = (double *)malloc(sizeof(double) * (size_t)h * w); /* row sums of source rows */
    if (!rows) return ORC_ENOMEM;
#pragma omp parallel
    {
        float *pad = (float *)malloc(sizeof(float) * pw);
#pragma omp for schedule(static)
        for (int y = 0; y < h; ++y) {
            const float *s = src + (size_t)y * w;
            for (int x = 0; x < pw; ++x) pad[x] = s[reflect(x - r, w)];
            double acc = 0.0;
            for (int i = 0; i < k; ++i) acc += (double)pad[i];
            double *D = rows + (size_t)y * w;
            D[0] = acc;
            for (int x = 0; x < w - 1; ++x) {
                acc += (double)pad[x + k] - (double)pad[x];
                D[x + 1] = acc;
            }
        }
        free(pad);
    }
    /* column pass: SUM[x] = sum of the first k-1 padded rows, then add/emit/subtract */
#pragma omp parallel
    {
#pragma omp for schedule(static)
        for (int xb = 0; xb < w; xb += 64) {
            const int xe = xb + 64 < w ? xb + 64 : w;
            double sum[64];
            for (int x = xb; x < xe; ++x) sum[x - xb] = 0.0;
            for (int i = 0; i < k - 1; ++i) {
                const double *R = rows + (size_t)reflect(i - r, h) * w;
                for (int x = xb; x < xe; ++x) sum[x - xb] += R[x];
            }
            for (int y = 0; y < h; ++y) {
                const double *Rn = rows + (size_t)reflect(y + r, h) * w;
                const double *Ro = rows + (size_t)reflect(y - r, h) * w;
                float *D = dst + (size_t)y * w;
                for (int x = xb; x < xe; ++x) {
                    const double s0 = sum[x - xb] + Rn[x];
                    D[x] = (float)(s0 * scale);
                    sum[x - xb] = s0 - Ro[x];
                }
            }
        }
    }
    free(rows);
    return ORC_OK;
}

/* ---- guided filter, colour (3-channel) guide, 8-bit in / 8-bit out -------- */
/* convertTo(CV_32F) of one channel, without scaling; f32 != 0: the image is CV_32F already */
static void plane_from(const void *img, int f32, int c, int cn, long n, float *out)
{
    if (f32)
        for (long i = 0; i < n; ++i) out[i] = ((const float *)img)[i * cn + c];
    else
        for (long i = 0; i < n; ++i) out[i] = (float)((const uint8_t *)img)[i * cn + c];
}

/* convertTo(depth of src) of the float result: uint8 rounds half-even and saturates, CV_32F is copied */
static void store_result(const float *acc, long count, int f32, void *dst)
{
    if (f32)
        memcpy(dst, acc, sizeof(float) * count);
    else
        for (long i = 0; i < count; ++i) ((uint8_t *)dst)[i] = sat_u8(acc[i]);
}

/* 1-channel guide (guided_filter.cpp, gCnNum == 1; SURVEY A.3 "analogous special cases", [upstream-recollection]):
 * the 1x1 "matrix" cov(I) + eps is inverted as a reciprocal, alpha = cov(I, p) * inv, beta = mean(p) - alpha * mean(I),
 * q = mean(alpha) * I + mean(beta).  Not reachable from the reference CLI (cv2.imread hands it 3 channels); part of
 * the cv2.ximgproc.guidedFilter surface behind apply_filter (filter_reflectance.py:67-70). */
static int guided_gray_guide(const void *guide, int gf32, const void *src, int sf32, int sc, void *dst, int h, int w,
                             int radius, float eps)
{
    const long n = (long)h * w;
    enum { NPL = 8 };
    float *buf = (float *)malloc(sizeof(float) * n * NPL);
    float *acc = (float *)malloc(sizeof(float) * n * sc);
    if (!buf || !acc) { free(buf); free(acc); return ORC_ENOMEM; }
    float *I = buf, *mI = buf + n, *inv = buf + 2 * n, *tmp = buf + 3 * n, *p = buf + 4 * n, *mp = buf + 5 * n,
          *a = buf + 6 * n, *b = buf + 7 * n;
    int rc = ORC_OK;
    plane_from(guide, gf32, 0, 1, n, I);
    rc |= orc_box_mean_reflect(I, mI, h, w, radius);
    for (long i = 0; i < n; ++i) tmp[i] = I[i] * I[i];
    rc |= orc_box_mean_reflect(tmp, inv, h, w, radius);
    for (long i = 0; i < n; ++i) {
        float v = inv[i] - mI[i] * mI[i];
        v += eps;
        inv[i] = 1.0f / v;
    }
    for (int si = 0; si < sc && rc == ORC_OK; ++si) {
        plane_from(src, sf32, si, sc, n, p);
        rc |= orc_box_mean_reflect(p, mp, h, w, radius);
        for (long i = 0; i < n; ++i) tmp[i] = p[i] * I[i];
        rc |= orc_box_mean_reflect(tmp, a, h, w, radius);
        for (long i = 0; i < n; ++i) {
            float c = a[i] - mp[i] * mI[i];
            float al = c * inv[i];
            a[i] = al;
            tmp[i] = mp[i] - al * mI[i];
        }
        rc |= orc_box_mean_reflect(tmp, b, h, w, radius);
        rc |= orc_box_mean_reflect(a, tmp, h, w, radius);
        for (long i = 0; i < n; ++i) acc[i * sc + si] = b[i] + tmp[i] * I[i];
    }
    store_result(acc, n * sc, sf32, dst);
    free(buf); free(acc);
    return rc;
}

/* guide / src: uint8 or (gf32 / sf32 != 0) CV_32F; dst has the depth of src (dDepth = -1).  The CV_32F depths are the
 * rest of cv2.ximgproc.guidedFilter's surface, not reachable from the reference CLI. */
int orc_guided_any(const void *guide, int gf32, int gc, const void *src, int sf32, int sc, void *dst,
                   int h, int w, int radius, double eps_d)
{
    if (!(gc == 1 || gc == 3) || !(sc == 1 || sc == 3) || h < 1 || w < 1 || radius < 0) return ORC_EINVAL;
    if (gc == 1) return guided_gray_guide(guide, gf32, src, sf32, sc, dst, h, w, radius, (float)eps_d);
    const long n = (long)h * w;
    const float eps = (float)eps_d;
    /* planes: I[3], mI[3], cov[6] -> inv[6], tmp, p, mp, c[3], a[3], b */
    enum { NPL = 3 + 3 + 6 + 1 + 1 + 1 + 3 + 3 + 1 };
    float *buf = (float *)malloc(sizeof(float) * n * NPL);
    float *acc = (float *)malloc(sizeof(float) * n * sc);
    if (!buf || !acc) { free(buf); free(acc); return ORC_ENOMEM; }
    float *I[3], *mI[3], *cv[6], *tmp, *p, *mp, *c[3], *a[3], *b;
    {
        float *q = buf;
        for (int i = 0; i < 3; ++i) { I[i] = q; q += n; }
        for (int i = 0; i < 3; ++i) { mI[i] = q; q += n; }
        for (int i = 0; i < 6; ++i) { cv[i] = q; q += n; }
        tmp = q; q += n; p = q; q += n; mp = q; q += n;
        for (int i = 0; i < 3; ++i) { c[i] = q; q += n; }
        for (int i = 0; i < 3; ++i) { a[i] = q; q += n; }
        b = q;
    }
    int rc = ORC_OK;
    for (int i = 0; i < 3; ++i) {
        plane_from(guide, gf32, i, 3, n, I[i]);
        rc |= orc_box_mean_reflect(I[i], mI[i], h, w, radius);
    }
    /* symmetric 3x3: index (k,l), k<=l -> 0:(0,0) 1:(0,1) 2:(0,2) 3:(1,1) 4:(1,2) 5:(2,2) */
    static const int ka[6] = {0, 0, 0, 1, 1, 2}, kb[6] = {0, 1, 2, 1, 2, 2};
    for (int m = 0; m < 6; ++m) {
        for (long i = 0; i < n; ++i) tmp[i] = I[ka[m]][i] * I[kb[m]][i];
        rc |= orc_box_mean_reflect(tmp, cv[m], h, w, radius);
        for (long i = 0; i < n; ++i) {
            float v = cv[m][i] - mI[ka[m]][i] * mI[kb[m]][i];
            if (ka[m] == kb[m]) v += eps;
            cv[m][i] = v;
        }
    }
    /* inverse by cofactors; cv[] becomes inv[] */
#define S(k, l) s[(k) <= (l) ? sym[(k)][(l)] : sym[(l)][(k)]]
    static const int sym[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) {
        float s[6], cof[6];
        for (int m = 0; m < 6; ++m) s[m] = cv[m][i];
        for (int m = 0; m < 6; ++m) {
            const int k = ka[m], l = kb[m];
            const int k1 = (k + 1) % 3, k2 = (k + 2) % 3, l1 = (l + 1) % 3, l2 = (l + 2) % 3;
            cof[m] = S(k1, l1) * S(k2, l2) - S(k1, l2) * S(k2, l1);
        }
        float det = s[0] * cof[0];
        det += s[1] * cof[1];
        det += s[2] * cof[2];
        if (eps < 1e-2f && fabsf(det) < 1e-6f) det = 1e-6f;
        for (int m = 0; m < 6; ++m) cv[m][i] = cof[m] / det;
    }
#undef S
    for (int si = 0; si < sc && rc == ORC_OK; ++si) {
        plane_from(src, sf32, si, sc, n, p);
        rc |= orc_box_mean_reflect(p, mp, h, w, radius);
        for (int g = 0; g < 3; ++g) {
            for (long i = 0; i < n; ++i) tmp[i] = p[i] * I[g][i];
            rc |= orc_box_mean_reflect(tmp, c[g], h, w, radius);
            for (long i = 0; i < n; ++i) c[g][i] -= mp[i] * mI[g][i];
        }
        for (long i = 0; i < n; ++i) {
            float al[3];
            for (int g = 0; g < 3; ++g) {
                float v = cv[sym[g][0]][i] * c[0][i];
                v += cv[sym[g][1]][i] * c[1][i];
                v += cv[sym[g][2]][i] * c[2][i];
                al[g] = v;
            }
            float be = mp[i];
            for (int g = 0; g < 3; ++g) be -= al[g] * mI[g][i];
            a[0][i] = al[0]; a[1][i] = al[1]; a[2][i] = al[2];
            tmp[i] = be;
        }
        rc |= orc_box_mean_reflect(tmp, b, h, w, radius);
        for (int g = 0; g < 3; ++g) {
            rc |= orc_box_mean_reflect(a[g], tmp, h, w, radius);
            memcpy(a[g], tmp, sizeof(float) * n);
        }
        for (long i = 0; i < n; ++i) {
            float v = b[i];
            for (int g = 0; g < 3; ++g) v += a[g][i] * I[g][i];
            acc[i * sc + si] = v;
        }
    }
    store_result(acc, n * sc, sf32, dst);
    free(buf); free(acc);
    return rc;
}

int orc_guided_u8(const uint8_t *guide, int gc, const uint8_t *src, int sc, uint8_t *dst,
                  int h, int w, int radius, double eps_d)
{
    return orc_guided_any(guide, 0, gc, src, 0, sc, dst, h, w, radius, eps_d);
}
