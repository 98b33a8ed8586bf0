/*
 * vsb.h -- C-ABI of libvsb_b200.so: the B200 (sm_100a) implementation of the per-layer hot path
 * of aaron1138/Voxel-Stack-Blender (Z-blend -> LUT -> XY filters).
 *
 * The reference has no FFI (it is pure Python over OpenCV/NumPy/SciPy/Numba wheels); the entry
 * points below replace the third-party call sites of its hot path one for one.  Each declaration
 * cites the reference lines it replaces (paths relative to the reference repository root).  The
 * binding a maintainer adds on the reference side is a ctypes stub, shown in INTEGRATION.md and
 * implemented in voxel-stack-blender_b200/vsb_native.py.
 *
 * Conventions
 *   - all image pointers are DEVICE pointers, row-major uint8 [H][pitch], pitch in bytes;
 *   - "bits" are bit-packed binary masks: uint32 words, `wpr` words per row (>= ceil(W/32),
 *     a multiple of 4), pixel x of a row lives in word x>>5, bit x&31; bits at x >= W are 0;
 *   - `stream` is a cudaStream_t passed as void* (0 = default stream);
 *   - functions are asynchronous on `stream`, own no memory, are re-entrant per stream, and
 *     return 0 on success or a negative vsb_status; vsb_last_error_string() describes the last
 *     failure of the calling thread;
 *   - there is no CPU fallback anywhere behind this interface.
 */
#ifndef VSB_H_
#define VSB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VSB_MAX_PRIORS 16 /* prior masks one fused launch can take (more are pre-OR'ed with vsb_bits_or) */
#define VSB_MAX_XY_OPS 16
#define VSB_FAST_REACH 63 /* largest chamfer reach (ceil(F)-1) of the shared-memory tile path */

typedef enum {
    VSB_OK = 0,
    VSB_ERR_ARG = -1,     /* bad argument (null pointer, size, alignment, unsupported parameter) */
    VSB_ERR_CUDA = -2,    /* CUDA runtime error; see vsb_last_error_string() */
    VSB_ERR_UNSUPPORTED = -3
} vsb_status;

typedef void *vsb_stream_t;

const char *vsb_last_error_string(void);
int vsb_version(void);
/* number of kernel launches issued through this library by the calling process (bench.py's
 * gpu_launches claim) */
unsigned long long vsb_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * K1  threshold + bit-pack.      replaces cv2.threshold(img,127,255,THRESH_BINARY)
 *     processing_core.py:44 (load_image).  bit = gray > thresh.
 * ------------------------------------------------------------------------------------------- */
int vsb_threshold_pack(const uint8_t *gray, size_t pitch, int W, int H, int thresh,
                       uint32_t *bits, int wpr, vsb_stream_t stream);

/* bits -> 0/255 bytes (the `binary_image` load_image returns).  processing_core.py:44-45 */
int vsb_bits_unpack(const uint32_t *bits, int wpr, int W, int H, uint8_t *out, size_t pitch,
                    vsb_stream_t stream);

/* OR of n packed masks.          replaces the copy + cv2.bitwise_or loop
 *     processing_core.py:61-65 (find_prior_combined_white_mask).  srcs = host array of device ptrs */
int vsb_bits_or(const uint32_t *const *srcs, int n, int wpr, int H, uint32_t *dst,
                vsb_stream_t stream);

/* rec = OR(priors) & ~cur, optional population count (cv2.countNonZero) into *count_out (device
 * uint64, must be zeroed by the caller).   processing_core.py:119,122,259-260,350-351 */
int vsb_maskdiff(const uint32_t *cur, const uint32_t *const *priors, int n, int wpr, int H,
                 uint32_t *rec_out, unsigned long long *count_out, vsb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Chamfer table  T[(R+1)*(R+1)] float32 (host helper, no device work):  T[M][m] = the 5x5
 * chamfer value (a=1, b=1.4, c=2.1969) of displacement (M,m) in canonical float32 summation
 * order.  cv2.distanceTransform(src, DIST_L2, 5) == min over zero pixels of T[max|d|][min|d|].
 * ------------------------------------------------------------------------------------------- */
int vsb_chamfer_table(float *T_host, int R);

/* K2  dense chamfer distance map, clipped: dist = min(D, clip) where D is the distance to the
 *     nearest set bit of `seed_bits`, exact for D < R+1; pixels with no seed within Chebyshev
 *     reach R get `clip`.     replaces cv2.distanceTransform(~cur, DIST_L2, 5)
 *     processing_core.py:129,250,379.   T = device copy of vsb_chamfer_table(R), R <= 63. */
int vsb_dt_chamfer5(const uint32_t *seed_bits, int wpr, int W, int H, int R, const float *T_dev,
                    float clip, float *dist, size_t dist_pitch_elems, vsb_stream_t stream);

/* same windowed tile kernel with the exact Euclidean metric: dist = min(sqrt(d2), clip), exact for
 * d < R+1 (R <= 63). */
int vsb_dt_exact_windowed(const uint32_t *seed_bits, int wpr, int W, int H, int R, float clip,
                          float *dist, size_t dist_pitch_elems, vsb_stream_t stream);

/* K2' exact Euclidean distance transform, unbounded reach (cv2.DIST_MASK_PRECISE analogue), csrc/vsb_edt.cu:
 *     separable lower-envelope formulation -- column pass (distance to the nearest set bit of the same column, uint16:
 *     32x32 bit squares transposed with warp shuffles, block carries) then, along every row, the lower envelope of the
 *     parabolas (x - x')^2 + g(x',y)^2, evaluated by bisection on the monotone column of the ruling parabola (one CTA
 *     per row, the row of g staged with one bulk copy).  O(log W) per pixel at worst, whatever the distances are.
 *     d2_out (optional) = exact integer squared distance (INT32_MAX if the image holds no set bit);
 *     dist_out (optional) = sqrtf(d2) (3e38 there).  W, H <= 32767; wpr a multiple of 4; scratch: device buffer of
 *     vsb_edt_exact_scratch_bytes(W, H) bytes, 128-byte aligned. */
size_t vsb_edt_exact_scratch_bytes(int W, int H);
int vsb_edt_exact(const uint32_t *seed_bits, int wpr, int W, int H, void *scratch, int32_t *d2_out, float *dist_out,
                  size_t pitch_elems, vsb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Unbounded reach (csrc/vsb_unbounded.cu): column distances + pruned row search.
 *   vsb_coldist      : g[y][x] = vertical distance to the nearest set bit of column x (uint16, 0xFFFF = none) and
 *                      gmin64[y][b] = min of g over columns 64b..64b+63 (uint16 [H][ceil(W/64)]).
 *   vsb_dt_unbounded : dense distance map, metric chamfer5 (exact integer chamfer cost in 2^-23 units, correctly
 *                      rounded to float32) or exact (d2 = integer squared distance, dist = sqrtf); 3e38 / INT32_MAX
 *                      where the image has no set bit.     replaces cv2.distanceTransform(src, DIST_L2, 5 | PRECISE)
 *   vsb_zblend_unbounded : fixed fade with any F (norm 0) or the global min-max normalisation (norm 1,
 *                      processing_core.py:141-145: minMaxLoc over the receding pixels, (d-min)/(max-min), zeros if
 *                      max <= min), merged with gray, leading LUT.  Scratch: g, gmin64, dist float [H][pitch],
 *                      minmax2 = 2 x uint32.
 * ------------------------------------------------------------------------------------------- */
/* number of uint16 elements the `gmin64` scratch of the three entry points below must hold (block minima plus the
 * per-band first/last-white-row records of the column pass) */
size_t vsb_coldist_scratch_elems(int W, int H, int wpr);
int vsb_coldist(const uint32_t *seed_bits, int wpr, int W, int H, uint16_t *g, size_t g_pitch_elems,
                uint16_t *gmin64, vsb_stream_t stream);
int vsb_dt_unbounded(const uint32_t *seed_bits, int wpr, int W, int H, int metric, uint16_t *g,
                     size_t g_pitch_elems, uint16_t *gmin64, float *dist_out, int32_t *d2_out,
                     size_t pitch_elems, vsb_stream_t stream);
int vsb_zblend_unbounded(const uint8_t *gray, size_t gray_pitch, int W, int H, const uint32_t *cur_bits,
                         const uint32_t *const *prior_bits, int n_priors, int wpr, int metric, int norm,
                         float fade, float den, uint16_t *g, size_t g_pitch_elems, uint16_t *gmin64,
                         float *dist, size_t dist_pitch_elems, uint32_t *minmax2, const uint8_t *lut256_dev,
                         uint8_t *out, size_t out_pitch, vsb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K3  fused Z-blend: rec mask, windowed chamfer distance, fade gradient, merge with the gray
 *     layer, composed LUT gather.  One launch per layer.
 *     replaces processing_core.py:107-153 (fixed fade), :222-284 (weighted stack), :442-460
 *     (merge_to_output) and lut_manager.py:57-63 (apply_z_lut) for leading apply_lut ops.
 * ------------------------------------------------------------------------------------------- */
typedef enum {
    VSB_Z_FIXED = 0,    /* g = trunc(255*(1 - min(d,F)/den)) on rec               :138-152 */
    VSB_Z_WEIGHTED = 1, /* g = trunc(255*(sum_i rec_i ? (1-min(d,F_i)/den_i)*w_i : 0)/W)  :252-282 */
    VSB_Z_CONST = 2,    /* g = const_val on rec (degenerate F <= 0)                 :138-140 */
    VSB_Z_DIST = 3,     /* no image output: write min(d,F) at rec pixels into dist_out (Enhanced EDT
                           stage 1, :379-381) */
    VSB_Z_ENHANCED = 4, /* stack level only: VSB_Z_DIST + vsb_ccl8_max + vsb_enhanced_fade  :338-404 */
    VSB_Z_NONE = 5,     /* vsb_layer_fused only: no Z-blend, the input already is the stage-1 image */
    VSB_Z_MINMAX = 6,   /* stack level only: vsb_zblend_unbounded norm 1 (global min-max, :141-145) */
    VSB_Z_FIXED_FAR = 7, /* stack level only: vsb_zblend_unbounded norm 0 (fixed fade, reach > VSB_FAST_REACH) */
    VSB_Z_WEIGHTED_FAR = 8, /* stack level only: vsb_weighted_unbounded (a fade distance beyond VSB_FAST_REACH) */
    VSB_Z_ENHANCED_FAR = 9  /* stack level only: vsb_rec_dist_unbounded + vsb_ccl8_max + vsb_enhanced_fade */
} vsb_zmode;

typedef enum {
    VSB_METRIC_CHAMFER5 = 0, /* the reference's metric: cv2.distanceTransform(.., DIST_L2, 5) */
    VSB_METRIC_EXACT = 1     /* exact Euclidean (sqrt of the integer squared distance)       */
} vsb_metric;

typedef struct {
    int mode;      /* vsb_zmode */
    int metric;    /* vsb_metric */
    int n_priors;  /* <= VSB_MAX_PRIORS, newest first */
    int reach;     /* R = ceil(max fade)-1, <= VSB_FAST_REACH */
    int const_val; /* VSB_Z_CONST */
    float fade;    /* F as float32 */
    float den;     /* F > 0 ? F : 1 */
    float total_weight;
    float weights[VSB_MAX_PRIORS];
    float fades[VSB_MAX_PRIORS];
    float dens[VSB_MAX_PRIORS];
    int enhanced_arith;  /* VSB_Z_ENHANCED: 0 = float32 (SciPy flavour), 1 = float64 (Numba flavour) */
    double fade_limit;
    /* Enhanced EDT, anisotropic correction (processing_core.py:357-377): size of the nearest-resized mask the
     * distance transform runs on (0 = disabled) and the search reach there (covers fade_limit + 3) */
    int aniso_w, aniso_h, aniso_reach;   /* VSB_Z_ENHANCED: F as the Python float the reference passes on */
} vsb_zparams;

int vsb_zblend_fused(const uint8_t *gray, size_t gray_pitch, int W, int H, const uint32_t *cur_bits,
                     const uint32_t *const *prior_bits /* host array of device ptrs */, int wpr,
                     const vsb_zparams *params /* host */, const float *T_dev,
                     const uint8_t *lut256_dev /* may be NULL = identity */, uint8_t *out,
                     size_t out_pitch, float *dist_out /* VSB_Z_DIST only */,
                     size_t dist_pitch_elems, vsb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K5  Enhanced EDT: 8-connected components of rec + per-component max of the clipped distance,
 *     then the ROI-normalised fade.   replaces cv2.connectedComponents (processing_core.py:383),
 *     ndimage.maximum / numba pass 1 (:291, :313-320) and the two fades (:295-302 float32,
 *     :323-334 float64), merge (:458) and the leading LUT.
 *     labels: int32 [H*W] scratch; at rec pixels it ends up holding the linear index y*W+x of the
 *     component's root pixel (other entries are never touched);
 *     compmax: uint32 [H*W] scratch holding the float bits of the maximum, indexed by root.
 * ------------------------------------------------------------------------------------------- */
int vsb_ccl8_max(const uint32_t *rec_bits, int wpr, int W, int H, const float *dist,
                 size_t dist_pitch_elems, int32_t *labels, uint32_t *compmax, vsb_stream_t stream);

int vsb_enhanced_fade(const uint8_t *gray, size_t gray_pitch, int W, int H, const uint32_t *rec_bits,
                      int wpr, const float *dist, size_t dist_pitch_elems, const int32_t *labels,
                      const uint32_t *compmax, double fade_limit, int arith /* 0 = float32 (SciPy), 1 = float64 (Numba) */,
                      const uint8_t *lut256_dev, uint8_t *out, size_t out_pitch, vsb_stream_t stream);

/* Enhanced EDT stage 1 in one launch: the packed receding mask (= vsb_maskdiff) and min(d, fade) at its pixels
 * (= vsb_zblend_fused in VSB_Z_DIST mode).   processing_core.py:350, :379-381 */
int vsb_rec_dist(const uint32_t *cur_bits, const uint32_t *const *prior_bits, int wpr, int W, int H,
                 const vsb_zparams *params, const float *T_dev, uint32_t *rec_bits_out, float *dist_out,
                 size_t dist_pitch_elems, vsb_stream_t stream);

/* the two per-label fades exactly as the reference exposes them to its tests (dense, contiguous
 * arrays; labels 1..num_labels-1 from any labelling, 0 = background):
 *     arith 0 = _calculate_receding_gradient_field_enhanced_edt_scipy  processing_core.py:287-303
 *     arith 1 = _calculate_receding_gradient_field_enhanced_edt_numba  processing_core.py:305-336
 * cmax: uint32 [num_labels] scratch. */
int vsb_enhanced_from_labels(const float *dist, const int32_t *labels, int W, int H, int num_labels,
                             double fade_limit, int arith, uint32_t *cmax, uint8_t *out,
                             vsb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K4  XY operations.
 * ------------------------------------------------------------------------------------------- */
/* lut gather                         lut_manager.py:57-63 */
int vsb_lut_apply(const uint8_t *in, size_t in_pitch, int W, int H, const uint8_t *lut256_dev,
                  uint8_t *out, size_t out_pitch, vsb_stream_t stream);

/* max(a,b)                           processing_core.py:458 (merge_to_output) */
int vsb_merge_max(const uint8_t *a, size_t a_pitch, const uint8_t *b, size_t b_pitch, int W, int H,
                  uint8_t *out, size_t out_pitch, vsb_stream_t stream);

/* cv2.GaussianBlur on uint8: Q8.8 taps (host arrays, odd lengths <= 63, each summing to 256),
 * (sum + 2^15) >> 16, BORDER_REFLECT_101.        xy_blend_processor.py:27-33 */
int vsb_gaussian_q88(const uint8_t *in, size_t in_pitch, int W, int H, const int32_t *kx, int nkx,
                     const int32_t *ky, int nky, uint8_t *out, size_t out_pitch, vsb_stream_t stream);

/* cv2.bilateralFilter on uint8: taps (dy[k],dx[k]) in row-major order with float32 space weights
 * sw[k], colour weights cw[256] (host arrays), float32 accumulate with FMA, round-half-even,
 * BORDER_REFLECT_101.                              xy_blend_processor.py:35-40 */
int vsb_bilateral(const uint8_t *in, size_t in_pitch, int W, int H, int radius, int ntaps,
                  const int8_t *dy, const int8_t *dx, const float *sw, const float *cw256,
                  uint8_t *out, size_t out_pitch, vsb_stream_t stream);

/* cv2.medianBlur on uint8: exact k x k median, BORDER_REPLICATE.   xy_blend_processor.py:42-46 */
int vsb_median(const uint8_t *in, size_t in_pitch, int W, int H, int ksize, uint8_t *out,
               size_t out_pitch, vsb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * The fused per-layer path (two launches per layer).
 *   vsb_pack_flags : K1 plus one uint16 per 128x32 tile: the tile's single gray value, or 0x100 = mixed; bit 0x200 is
 *                    set when every byte of the tile is 0 or 255 (the input the reference asks for, README.md:25).
 *                    flags: uint16 [ceil(H/32)][ceil(W/128)];  wpr must be 4*ceil(W/128).
 *                    With threshold 127 and 16-byte aligned rows this is the persistent TMA-fed kernel of
 *                    csrc/vsb_pack_tma.cu (cp.async.bulk.tensor loads into per-warp mbarrier rings, bulk tensor
 *                    stores for lut[gray]); other thresholds / alignments take the plain-load kernel.
 *   vsb_layer_fused: Z-blend (FIXED / WEIGHTED / CONST, or NONE) -> lut1 -> [bilateral] -> [Gaussian] -> lut2
 *                    in ONE kernel, intermediates in shared memory.  Same arithmetic, bit for bit, as the
 *                    separate entry points above.  cur_flags / prior_flags (from vsb_pack_flags) are optional: they
 *                    select the binary fast path (see general_tiles) and let uniform neighbourhoods leave early.
 *     replaces processing_pipeline.py:98-107 (process_z_blending, merge_to_output, process_xy_pipeline)
 *     for XY chains of the shape [apply_lut*] [bilateral_filter d<=9] [gaussian_blur k<=7] [apply_lut*].
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int has_bilateral, radius, ntaps; /* radius <= 4, ntaps <= 96 */
    int8_t tap_dy[96], tap_dx[96];
    float tap_sw[96];
    float cw[256];
    int nkx, nky;                     /* 0 = no Gaussian; odd, <= 7 */
    int32_t kx[7], ky[7];
    /* bilateral shortcut of the binary fast path (vsb_bilateral_two_valued_min_same): a window that holds only
     * lut1[0] / lut1[255] with at least this many taps equal to the centre rounds to the centre; 0 = never assume it */
    int two_valued_min_same;
} vsb_xy_fused;

/* Smallest number m of taps (centre included) equal to the centre value for which cv2.bilateralFilter, on a window
 * holding only the two values v0 and v1, provably returns the centre: worst case over tap placements of
 * |v1 - v0| * Wo * cw[|v1-v0|] / (Ws + Wo * cw[|v1-v0|]) < 0.49 (float32 accumulation error is < 1e-3).  Returns 0 if no
 * m <= ntaps gives that.  Host helper, no device work.     xy_blend_processor.py:35-40 */
int vsb_bilateral_two_valued_min_same(const float *tap_sw, int ntaps, const float *cw256, int v0, int v1);

int vsb_pack_flags(const uint8_t *gray, size_t pitch, int W, int H, int thresh, uint32_t *bits, int wpr,
                   uint16_t *flags, vsb_stream_t stream);

/* Enhanced EDT, second half fused with the XY chain: the ROI-normalised fade (vsb_enhanced_fade's arithmetic, `arith` =
 * params->enhanced_arith, limit = params->fade_limit) evaluated on the receding pixels from the distances, component
 * roots and per-root maxima that vsb_zblend_fused(VSB_Z_DIST) + vsb_ccl8_max produced, then merge -> lut1 -> [bilateral]
 * -> [Gaussian] -> lut2 exactly like vsb_layer_fused.   processing_core.py:287-336, :404, :458 + the XY chain. */
int vsb_layer_fused_enhanced(const uint8_t *in, size_t in_pitch, int W, int H, const uint32_t *cur_bits,
                             const uint32_t *const *prior_bits, int wpr, const vsb_zparams *params, const float *dist,
                             size_t dist_pitch_elems, const int32_t *labels, const uint32_t *compmax,
                             const uint8_t *lut1_dev, const vsb_xy_fused *xy /* host, may be NULL */,
                             const uint8_t *lut2_dev, uint8_t *out, size_t out_pitch,
                             const uint16_t *cur_flags /* NULL, or the layer's tile flags */,
                             int *general_tiles /* NULL, or int[3 * (1 + tiles)] scratch: see vsb_layer_fused */,
                             vsb_stream_t stream);

/* LUT-only pipelines (no XY filter): one pass over the gray layer writes the packed mask AND lut[gray] (the final
 * value of every pixel that does not recede); vsb_layer_patch then computes the fade on the receding pixels alone and
 * rewrites them with lut[max(gray, gradient)].  Same results as vsb_layer_fused without XY stage, the gray layer is
 * read from HBM once and untouched tiles cost one look at their mask words. */
int vsb_pack_flags_lut(const uint8_t *gray, size_t pitch, int W, int H, int thresh, uint32_t *bits, int wpr,
                       uint16_t *flags, const uint8_t *lut256_dev /* NULL = identity */, uint8_t *out, size_t out_pitch,
                       const uint32_t *const *prior_bits,
                       const uint16_t *const *prior_flags /* NULL, or the priors' tile flags (vsb_pack_flags*): tiles
                                         where every prior is uniformly black skip the prior-mask reads */,
                       int n_priors,
                       int *rec_tiles /* NULL, or int[1 + tiles]: [0] receives the number of 128x32 tiles holding a
                                         receding pixel w.r.t. the given priors, [1..] their indices */,
                       vsb_stream_t stream);
int vsb_layer_patch(const uint8_t *in, size_t in_pitch, int W, int H, const uint32_t *cur_bits,
                    const uint32_t *const *prior_bits, int wpr, const vsb_zparams *params, const float *T_dev,
                    const uint8_t *lut1_dev, const int *rec_tiles /* from vsb_pack_flags_lut, or NULL = look at every tile */,
                    uint8_t *out /* pre-filled by vsb_pack_flags_lut */, size_t out_pitch, vsb_stream_t stream);

/* work counters of vsb_layer_fused (debugging / DESIGN.md numbers): op 1 = enable + clear, 2 = read 8 counters
 * (uniform tiles, worked tiles, tiles with receding pixels, receding pixels, bilateral pixels, Gaussian words,
 * 0, 0) and clear, 0 = disable */
int vsb_layer_stats(int op, unsigned long long *out8);

/* test hook of vsb_layer_fused's mask-only fast path: size of the per-tile list arena (entries; 0 = default 7424) and of the
 * receding-pixel list (0 = default 2560).  Smaller values send ordinary tiles down the routes that only unusually full tiles
 * take (general kernel for tiles fuller than the arena or -- at image borders -- than the list); results must not change. */
int vsb_layer_fast_path_limits(int arena_entries, int receding_entries);

int vsb_layer_fused(const uint8_t *in, size_t in_pitch, int W, int H, const uint32_t *cur_bits,
                    const uint32_t *const *prior_bits, const uint16_t *cur_flags,
                    const uint16_t *const *prior_flags, int wpr, const vsb_zparams *params, const float *T_dev,
                    const uint8_t *lut1_dev, const vsb_xy_fused *xy /* host, may be NULL */,
                    const uint8_t *lut2_dev, uint8_t *out, size_t out_pitch,
                    int *general_tiles /* NULL, or int[3 * (1 + tiles)] device scratch.  With it and cur_flags, tiles whose
                                          3x3 tile neighbourhood is strictly binary (flag bit 0x200) run in k_layer2, which
                                          derives everything from the packed masks and never reads `in`; the other
                                          tiles are listed here and taken by the general kernel */,
                    vsb_stream_t stream);

/* vsb_zblend_unbounded's two halves, for the other modes with fade distances beyond VSB_FAST_REACH:
 * distances at the receding pixels only (dist is untouched elsewhere; minmax2 receives their min / max float bits) */
int vsb_rec_dist_unbounded(const uint32_t *cur_bits, const uint32_t *const *prior_bits, int n_priors, int wpr, int W, int H,
                           int metric, uint16_t *g, size_t g_pitch_elems, uint16_t *gmin64, float *dist,
                           size_t dist_pitch_elems, uint32_t *minmax2, vsb_stream_t stream);
/* weighted stack (processing_core.py:222-284) with unbounded distances; params->weights / fades / dens /
 * total_weight / n_priors / metric as for VSB_Z_WEIGHTED */
int vsb_weighted_unbounded(const uint8_t *gray, size_t gray_pitch, int W, int H, const uint32_t *cur_bits,
                           const uint32_t *const *prior_bits, int wpr, const vsb_zparams *params, uint16_t *g,
                           size_t g_pitch_elems, uint16_t *gmin64, float *dist, size_t dist_pitch_elems, uint32_t *minmax2,
                           const uint8_t *lut256_dev, uint8_t *out, size_t out_pitch, vsb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * unsharp mask, resize, and the two resizes of the anisotropic
 * Enhanced-EDT branch.
 * ------------------------------------------------------------------------------------------- */
/* cv2.addWeighted(img, 1+amount, blurred, -amount, 0) followed by the absdiff / THRESH_BINARY_INV select of
 * apply_unsharp_mask (xy_blend_processor.py:58-67); `blurred` = vsb_gaussian_q88 of img with a k x k kernel.
 * float32: saturate(rint(fma(img, alpha, blurred * beta))); threshold <= 0 keeps the sharpened value everywhere. */
int vsb_unsharp_combine(const uint8_t *img, size_t pitch, const uint8_t *blurred, size_t bpitch, int W, int H,
                        float amount, int threshold, uint8_t *out, size_t opitch, vsb_stream_t stream);

/* cv2.resize(image, (width, height), interpolation=...) of apply_resize (xy_blend_processor.py:84-90).  A plan holds
 * the coordinate / coefficient tables of one (source size, target size, interpolation) triple on the current
 * device; it is immutable and may be shared by streams. */
enum vsb_interp {
    VSB_INTER_NEAREST = 0,   /* "NEAREST"  */
    VSB_INTER_LINEAR = 1,    /* "BILINEAR": Q11 fixed point */
    VSB_INTER_CUBIC = 2,     /* "BICUBIC":  OpenCV's own kernel (cv2 built with IPP differs by +-1 LSB) */
    VSB_INTER_LANCZOS4 = 3,  /* "LANCZOS4" and the reference's fallback for unknown names */
    VSB_INTER_AREA = 4,      /* "AREA": box sums / float area weights / 2-tap when enlarging */
    VSB_INTER_LINEAR_F32 = 5 /* float32 INTER_LINEAR (distance map of processing_core.py:373) */
};
typedef struct vsb_resize_plan vsb_resize_plan;
int vsb_resize_plan_create(int src_W, int src_H, int dst_W, int dst_H, int interp, vsb_resize_plan **out_plan);
int vsb_resize_plan_destroy(vsb_resize_plan *plan);
int vsb_resize_plan_dims(const vsb_resize_plan *plan, int *src_W, int *src_H, int *dst_W, int *dst_H);
int vsb_resize_u8(const vsb_resize_plan *plan, const uint8_t *in, size_t in_pitch, uint8_t *out, size_t out_pitch,
                  vsb_stream_t stream);
/* cv2.resize(resized_dist_map, (W, H), INTER_LINEAR) on float32 (processing_core.py:373); pitches in elements */
int vsb_resize_f32_linear(const vsb_resize_plan *plan, const float *in, size_t in_pitch_elems, float *out,
                          size_t out_pitch_elems, vsb_stream_t stream);
/* cv2.resize(distance_transform_src, (new_w, new_h), INTER_NEAREST) (processing_core.py:369) on the bit-packed mask */
int vsb_resize_bits_nearest(const vsb_resize_plan *plan, const uint32_t *bits, int wpr, uint32_t *out_bits, int out_wpr,
                            vsb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * ROI_FADE.
 * ------------------------------------------------------------------------------------------- */
/* cv2.connectedComponentsWithStats(binary_image, 8, CV_32S) of identify_rois (processing_core.py:69-105) on the
 * bit-packed mask.  labels: int32 [H*W] scratch (root pixel index at foreground pixels); cid: int32 [H*W], on return
 * component id + 1 at foreground pixels and 0 elsewhere (ids are compact, in no particular order); stats_dev[id] /
 * *count_dev: per-component statistics and the number of components (may exceed max_components: rerun with a larger
 * table).  first_block = smallest (y/2)*ceil(W/2) + x/2 over the component: sorting by it reproduces the library's
 * label order, which the tracker's tie-breaking depends on. */
typedef struct {
    int root;                 /* linear index of the first pixel (raster order) */
    int area;
    int x0, y0, x1, y1;       /* inclusive bounding box */
    int first_block;
    int pad;
    unsigned long long sum_x, sum_y; /* centroid = sum / area */
} vsb_roi_stat;
int vsb_roi_components(const uint32_t *bits, int wpr, int W, int H, int32_t *labels, int32_t *cid,
                       vsb_roi_stat *stats_dev, int *count_dev, int max_components, vsb_stream_t stream);

/* _calculate_receding_gradient_field_roi_fade (processing_core.py:155-220) + merge_to_output + leading apply_lut.
 * eligible_dev[id] != 0: component id is an ROI (area >= min_size) that is not skipped as raft / support.
 * dilate_radius = min(51, 2*int(F)+1) / 2; reach: search reach in pixels (table T_dev = vsb_chamfer_table(reach));
 * use_fixed = config.use_fixed_fade_receding (0: per-ROI min-max normalisation, needs minmax_scratch uint32[2*n]). */
int vsb_roi_fade(const uint8_t *gray, size_t gray_pitch, int W, int H, const uint32_t *cur_bits,
                 const uint32_t *const *prior_bits, int n_priors, int wpr, const int32_t *cid,
                 const uint8_t *eligible_dev, int n_components, int dilate_radius, int reach, const float *T_dev,
                 int use_fixed, float fade, float den, uint32_t *minmax_scratch, const uint8_t *lut256_dev,
                 uint8_t *out, size_t out_pitch, vsb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Host-buffer entry points (the plugin path measured as `e2e`): pinned or pageable HOST arrays
 * in, HOST array out; device staging buffers are owned by an opaque stack context.
 *     replaces ProcessingPipelineThread._process_single_image_task + the producer loop
 *     processing_pipeline.py:74-113,188-223 for FIXED_FADE / WEIGHTED_STACK.
 * ------------------------------------------------------------------------------------------- */
typedef struct vsb_stack vsb_stack;

typedef struct {
    int type;   /* 0 none, 1 lut, 2 gaussian, 3 bilateral, 4 median, 5 unsharp mask (Gaussian taps in kx/ky), 6 resize */
    uint8_t lut[256];
    int32_t kx[63], ky[63];
    int nkx, nky;
    int radius, ntaps;
    int8_t tap_dy[1024], tap_dx[1024];
    float tap_sw[1024];
    float cw[256];
    int median_ksize;
    float unsharp_amount;             /* type 5: apply_unsharp_mask, xy_blend_processor.py:48-67 */
    int unsharp_threshold;
    int resize_w, resize_h, interp;   /* type 6: resolved target size (apply_resize's rule, :71-82) and vsb_interp */
} vsb_xy_op;

int vsb_stack_create(int W, int H, int n_priors, const vsb_zparams *zp, const vsb_xy_op *ops,
                     int n_ops, vsb_stack **out_stack);
int vsb_stack_destroy(vsb_stack *s);
/* size of the layers the stack emits (differs from W x H when the XY chain holds a resize) */
int vsb_stack_out_size(vsb_stack *s, int *out_W, int *out_H);
/* forget the sliding window of prior masks (start of a new stack / Z-shard) */
int vsb_stack_reset(vsb_stack *s);
/* feed a halo layer: thresholded and pushed into the window, no output */
int vsb_stack_push_halo(vsb_stack *s, const uint8_t *gray_host, size_t pitch);
/* the same for a halo layer that is already resident on the device (asynchronous on the stack's stream): what a Z-shard
 * whose slab lives in HBM calls for the receding_layers layers below its first one */
int vsb_stack_push_halo_device(vsb_stack *s, const uint8_t *gray_dev, size_t pitch);
/* one layer, host in -> host out (H2D, kernels, D2H; synchronous on return) */
int vsb_stack_process_host(vsb_stack *s, const uint8_t *gray_host, size_t in_pitch,
                           uint8_t *out_host, size_t out_pitch);
/* pipelined variant: H2D, kernels and D2H of consecutive layers overlap on three streams.
 * submit returns a ticket (>= 0); the output is complete after vsb_stack_wait(ticket).  At most
 * vsb_stack_slots() layers may be in flight; host buffers should be pinned. */
int vsb_stack_submit_host(vsb_stack *s, const uint8_t *gray_host, size_t in_pitch,
                          uint8_t *out_host, size_t out_pitch);
int vsb_stack_wait(vsb_stack *s, int ticket);
int vsb_stack_slots(vsb_stack *s);
/* one layer, device in -> device out on the stack's stream (asynchronous) */
int vsb_stack_process_device(vsb_stack *s, const uint8_t *gray_dev, size_t in_pitch,
                             uint8_t *out_dev, size_t out_pitch);
/* n device-resident layers in one call: layer i is read at gray_dev + i*in_layer_stride (the stride may
 * be negative to walk a slab downwards) and written to out_dev + (i % out_ring)*out_layer_stride */
int vsb_stack_process_device_batch(vsb_stack *s, const uint8_t *gray_dev, size_t in_pitch,
                                   long long in_layer_stride, uint8_t *out_dev, size_t out_pitch,
                                   long long out_layer_stride, int out_ring, int n_layers);
int vsb_stack_sync(vsb_stack *s);
/* per-stage CUDA-event timing of the layer kernels on the stack's compute stream (measurement aid):
 * enable(1) starts recording an event pair around every stage; read() synchronises, returns the number of
 * stages and fills the summed milliseconds, kernel-launch counts and 32-byte stage names, then clears. */
int vsb_stack_profile_enable(vsb_stack *s, int on);
int vsb_stack_profile_read(vsb_stack *s, int max_stages, float *ms_sum, int *launches, char *names);
vsb_stream_t vsb_stack_stream(vsb_stack *s);

/* ---------------------------------------------------------------------------------------------
 * Synthetic stack rasteriser (measurement utility): layer z of a primitive
 * table (x, y, r0, slope, z_start, z_end) float32[6] per disk, plus an optional raft rectangle.
 * Pixel = 255 iff inside any primitive.
 * ------------------------------------------------------------------------------------------- */
int vsb_synth_layer(const float *prims_dev, int n_prims, const float *raft /* host: x0,y0,x1,y1,corner_r,z_end or NULL */,
                    int z, int W, int H, uint8_t *out, size_t pitch, vsb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VSB_H_ */
