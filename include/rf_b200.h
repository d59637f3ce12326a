/*
 * rf_b200.h -- C ABI of librf_b200.so: the CNN -> joint bilateral / guided filter hot path
 * of tnestmeyer/reflectance-filtering as hand-written sm_100a kernels.
 *
 * The reference has no native interface of its own: its hot path calls three third-party
 * native entry points from Python.  Each function below replaces one of those call sites
 * (file:line into the reference):
 *
 *   rf_cnn_create / rf_cnn_forward_u8   caffe.Net(prototxt, caffe.TEST, weights=...) and
 *                                       net.forward()      decompose_with_trained_CNN.py:104-106, :82-95
 *                                       (input transform :57-69 + image_utils.py:32-39 is fused in;
 *                                        the optional u8 output is image_utils.py:68 truncation)
 *   rf_joint_bilateral_u8               cv2.ximgproc.jointBilateralFilter(joint, src, d, sigmaColor,
 *                                       sigmaSpace)        filter_reflectance.py:60-64
 *   rf_guided_u8                        cv2.ximgproc.guidedFilter(guide, src, radius, eps)
 *                                                          filter_reflectance.py:67-70
 *   rf_whdr_f32                         whdr(reflectance, comparisons, delta)
 *                                                          training/layers/whdr_layer.py:253-287
 *   rf_colorize_u8                      image_utils.colorize / normalize / rgb_to_srgb / imwrite
 *                                       quantisation       image_utils.py:42-49,60-92
 *
 * Conventions
 *   - Every image pointer is a DEVICE pointer owned by the caller (e.g. a torch allocation) on the
 *     device selected with rf_set_device(); images are dense, interleaved HWC uint8, batch-major
 *     [n][h][w][c].  No torch types cross this boundary.
 *   - All calls are asynchronous on `stream` (a cudaStream_t passed as void*; NULL = default
 *     stream).  The hot calls do not allocate or synchronise, except the first use of a new
 *     (sigma_color, sigma_space, d) triple in rf_joint_bilateral_u8, which builds and caches a
 *     small weight table on the device.
 *   - Every function returns RF_OK (0) or an RF_E* status and never throws; rf_last_error()
 *     returns a thread-local message for the last failure.
 *   - There is no CPU fallback: without a CUDA device every compute call returns RF_ECUDA.
 */
#ifndef RF_B200_H
#define RF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RF_OK 0
#define RF_EINVAL 1      /* bad argument (shape, channel count, NULL pointer) */
#define RF_EUNSUPPORTED 2 /* valid for the reference but outside this build's limits */
#define RF_ECUDA 3       /* CUDA runtime error; message in rf_last_error() */
#define RF_ENOMEM 4

/* flags for rf_joint_bilateral_u8 */
#define RF_BF_GRAY_REPLICATED 1u /* joint, src and dst are 1-channel planes that stand for three equal
                                    channels (what cv2.imread makes of the CNN's gray PNG): the range
                                    distance is 3*|dJ| and the three output channels are equal */

typedef struct rf_cnn rf_cnn; /* opaque model handle (device-resident weights + sRGB table) */

int rf_version(void);
const char *rf_last_error(void);
int rf_set_device(int device);
int rf_device_info(int *sm_count, int *cc_major, int *cc_minor, size_t *total_mem);
/* number of kernels this library has launched since load (all threads) */
unsigned long long rf_launch_count(void);

/* ---- CNN (per-pixel MLP with skip concat) --------------------------------------------------
 * params: for each hidden layer W[out][in] row-major then b[out]; then the fusing weights
 *         w[sum(out_i)] (concat order = layer order) and its bias.
 * dims:   n_hidden+1 ints, dims[0] == 3, all hidden widths equal and one of 8, 16, 32, 64 (the family
 *         create_convStaticSkipLayers of the reference's training code produces; the shipped net is 5 x 32),
 *         1 <= n_hidden <= 8; anything else returns RF_EUNSUPPORTED.  Width 32 runs on the tensor cores
 *         (tcgen05, 3xTF32 split operands, max abs error ~1e-5 on the reflectance); other widths run the
 *         exact-FP32 CUDA-core kernel.
 * srgb_lut256: the exact table float32(srgb_to_rgb(v/255.0)), v = 0..255, computed by the host in
 *         float64 as the reference does; NULL lets the library compute it with pow() in double. */
int rf_cnn_create(const float *params, const int *dims, int n_hidden, const float *srgb_lut256,
                  rf_cnn **out);
void rf_cnn_destroy(rf_cnn *net);
/* bgr: uint8 [n][h][w][3].  out_f32 (float [n][h][w], linear reflectance intensity in (0,1)) and
 * out_u8 (uint8 [n][h][w] = trunc(r * 255)) may each be NULL, not both. */
int rf_cnn_forward_u8(const rf_cnn *net, const uint8_t *bgr, int n, int h, int w, float *out_f32,
                      uint8_t *out_u8, void *stream);

/* ---- joint bilateral filter, 8-bit ---------------------------------------------------------
 * joint [n][h][w][jc], src / dst [n][h][w][sc], jc, sc in {1, 3}.  Semantics of OpenCV-contrib
 * 3.1.0 jointBilateralFilter_8u: radius = d <= 0 ? cvRound(1.5 * sigma_space) : d / 2 (at least 1),
 * taps on the disc i*i + j*j <= radius^2, BORDER_REFLECT_101, weights exp(-k^2 / 2 sigma_space^2) *
 * exp(-alpha^2 / 2 sigma_color^2) with alpha = sum_c |J0_c - Jk_c|, output round-half-even.
 * sigma <= 0 is replaced by 1 as OpenCV does (the Python operator rejects it earlier).
 * joint may alias src (self-guided); dst must not alias either.
 * Radii up to rf_joint_bilateral_fast_max_radius() run on the shared-memory-tiled kernels; larger ones (up to
 * rf_joint_bilateral_max_radius(), else RF_EUNSUPPORTED) on the generic kernel. */
int rf_joint_bilateral_u8(const uint8_t *joint, int jc, const uint8_t *src, int sc, uint8_t *dst,
                          int n, int h, int w, double sigma_color, double sigma_space, int d,
                          unsigned flags, void *stream);

/* The rest of cv2.ximgproc.jointBilateralFilter(joint, src, d, sigmaColor, sigmaSpace[, dst[, borderType]])
 * (the call site filter_reflectance.py:60-64 passes whatever ndarrays it is given):
 * border_type = OpenCV's cv::BorderTypes value as copyMakeBorder applies it to joint and src (CONSTANT pads with 0). */
#define RF_BORDER_CONSTANT 0
#define RF_BORDER_REPLICATE 1
#define RF_BORDER_REFLECT 2
#define RF_BORDER_WRAP 3
#define RF_BORDER_REFLECT_101 4 /* cv2.BORDER_DEFAULT, what rf_joint_bilateral_u8 uses */
int rf_joint_bilateral_u8_border(const uint8_t *joint, int jc, const uint8_t *src, int sc, uint8_t *dst,
                                 int n, int h, int w, double sigma_color, double sigma_space, int d,
                                 unsigned flags, int border_type, void *stream);

/* CV_32F images (jointBilateralFilter_32f): float joint [n][h][w][jc], src / dst [n][h][w][sc].  The range weight is
 * interpolated in an exp table of 4096 * jc bins scaled to each joint image's own value range; the spatial part is
 * as in the 8-bit version; no rounding of the output.  `workspace`: device scratch of at least
 * rf_joint_bilateral_f32_workspace_bytes(n, jc) bytes (the per-image tables). */
size_t rf_joint_bilateral_f32_workspace_bytes(int n, int jc);
int rf_joint_bilateral_f32(const float *joint, int jc, const float *src, int sc, float *dst, int n, int h, int w,
                           double sigma_color, double sigma_space, int d, int border_type, void *workspace,
                           size_t workspace_bytes, void *stream);
int rf_joint_bilateral_geometry(double sigma_space, int d, int *radius, int *taps);
int rf_joint_bilateral_max_radius(void);
int rf_joint_bilateral_fast_max_radius(void);

/* ---- guided filter, 8-bit -------------------------------------------------------------------
 * guide [n][h][w][gc], gc in {3, 1}; src / dst [n][h][w][sc], sc in {1, 3}.  Semantics of ximgproc
 * guidedFilter (He et al.) on the 0..255 scale: (2r+1)^2 box means with BORDER_REFLECT, cov(I) + eps * Id,
 * per-pixel 3x3 inverse (gc = 1: reciprocal of var(I) + eps), q = mean(a) . I + mean(b), output
 * round-half-even saturated.  The reference CLI always passes a 3-channel guide (cv2.imread); 1-channel
 * guides are the rest of the cv2.ximgproc.guidedFilter surface behind apply_filter.
 * A 1-channel src whose three channels would be equal (the CNN reflectance) is filtered once.
 * ws: caller-owned scratch of at least rf_guided_workspace_bytes(...) bytes. */
size_t rf_guided_workspace_bytes(int sc, int n, int h, int w, int radius);
int rf_guided_u8(const uint8_t *guide, int gc, const uint8_t *src, int sc, uint8_t *dst, int n,
                 int h, int w, int radius, double eps, void *ws, size_t ws_bytes, void *stream);
int rf_guided_max_radius(void);

/* CV_32F guide and source (guidedFilter converts every depth to float without scaling; the result has the depth of
 * src): float planes in, float out, no rounding.  Runs on the generic (any radius) kernels; `ws` of at least
 * rf_guided_f32_workspace_bytes(...) bytes.  A uint8 / float mix is handled by converting the uint8 image to float
 * first (what the Python mirror does). */
size_t rf_guided_f32_workspace_bytes(int sc, int n, int h, int w);
int rf_guided_f32(const float *guide, int gc, const float *src, int sc, float *dst, int n, int h, int w,
                  int radius, double eps, void *ws, size_t ws_bytes, void *stream);
/* The same guide applied `iterations` times: iteration k filters the uint8 output of iteration k-1
 * (createGuidedFilter(guide, radius, eps) reused for several ->filter() calls; the reference's "3 x GF"
 * setting re-runs filter_reflectance.py on its own output).  Byte-identical to `iterations` calls of
 * rf_guided_u8; the guide statistics (mean I, inverse of cov(I) + eps*Id) are computed once and kept in
 * the workspace. */
size_t rf_guided_iterated_workspace_bytes(int sc, int n, int h, int w, int radius, int iterations);
int rf_guided_iterated_u8(const uint8_t *guide, int gc, const uint8_t *src, int sc, uint8_t *dst, int n,
                          int h, int w, int radius, double eps, int iterations, void *ws, size_t ws_bytes,
                          void *stream);

/* Diagnostic: the row plan of the guided filter's second pass for output row y.  The first pass stores, per
 * coefficient plane, prefix sums down the image rows that restart every seg_rows rows; the (2*radius+1)-row window
 * sum of row y under BORDER_REFLECT is sum_i weights[i] * prefix_row[rows[i]].  rows / weights hold at least 12
 * entries; returns the number of terms (more than 12: this segmentation is not used), -1 on bad arguments.
 * Host-only, no CUDA call (lets CPU tests check the decomposition against a brute-force window sum). */
int rf_guided_row_terms(int h, int radius, int seg_rows, int y, int *rows, float *weights);

/* ---- layout helpers ------------------------------------------------------------------------ */
/* gray [n_px] -> bgr [n_px][3] with three equal channels (what cv2.imread returns for the CNN PNG) */
int rf_replicate_gray_u8(const uint8_t *gray, uint8_t *bgr, size_t n_px, void *stream);
/* bgr [n_px][3] -> first channel [n_px]; *all_equal_flag (device int, may be NULL) is cleared if any
 * pixel has unequal channels; it is never set by this call */
int rf_extract_gray_u8(const uint8_t *bgr, uint8_t *gray, size_t n_px, int *all_equal_flag,
                       void *stream);

/* ---- colorized side outputs of the decomposition (decompose_with_trained_CNN.py:122-128) ----------------
 * For every image: colorize (image_utils.py:76-81) on the raw 0..255 BGR image and the float32 reflectance
 * intensity, then what imwrite(..., sRGB=True) does to each result (image_utils.py:60-92): divide by the
 * 99.9th percentile ('lower' selection) and clip if max > 1, the reference's rgb_to_srgb, truncate to uint8.
 * All arithmetic in float64 as numpy does.  k_reflectance / k_shading: 0-based rank of the percentile
 * element among the 3*h*w reflectance / h*w shading values = floor((count - 1) * (99.9 / 100)) evaluated by
 * the caller in float64 (numpy's own formula).  out_reflectance [n][h][w][3] (BGR order), out_shading [n][h][w]. */
size_t rf_colorize_workspace_bytes(int n, int h, int w);
int rf_colorize_u8(const uint8_t *bgr, const float *intensity, int n, int h, int w, double eps,
                   unsigned long long k_reflectance, unsigned long long k_shading, uint8_t *out_reflectance,
                   uint8_t *out_shading, void *ws, size_t ws_bytes, void *stream);

/* ---- WHDR of reflectance images (training/layers/whdr_layer.py:253-287) ---------------------
 * reflectance: float [n][c][h][w] (planar, c in {1, 3}; the CNN's out_f32 is the c = 1 case).
 * comparisons: float64 [n][max_comparisons + 1][6], the blob of createNumpyArrayWithComparisonsForIIW.py:616-649:
 *   row i < count = (x1, y1, x2, y2, darker, weight), darker 0 = equal / 1 / 2; last row = (count, file name, 0).
 *   Coordinates are relative ([0,1), scaled as int(x * w), int(y * h) like whdr_layer.py:240-251) unless
 *   RF_WHDR_PIXEL_COORDS is set (already-scaled values, what whdr() itself receives).
 * out_sums: device float64 [n][2] = (error_sum, weight_sum) per image; WHDR = error_sum / weight_sum, 0 if the
 *   weight sum is 0.  bad_flag: device int, OR-ed with 1 for an invalid count row, 2 for a coordinate outside the
 *   image (numpy raises IndexError there; such comparisons are skipped); never cleared by this call. */
#define RF_WHDR_PIXEL_COORDS 1u
int rf_whdr_f32(const float *reflectance, int c, int n, int h, int w, const double *comparisons,
                int max_comparisons, double delta, unsigned flags, double *out_sums, int *bad_flag, void *stream);

/* ---- aggregate statistics (the one value a multi-GPU run may all-reduce) --------------------
 * stats[0] += n_px ; [1] += sum(out) ; [2] += sum(out^2) ; [3] += sum|out - in| ; device doubles */
int rf_accumulate_stats_u8(const uint8_t *in, const uint8_t *out, size_t n_bytes, double *stats4,
                           void *stream);

#ifdef __cplusplus
}
#endif
#endif /* RF_B200_H */
