/*
 * vsb_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement (plain C) of the
 * arithmetic that the reference's per-layer hot path delegates to third-party
 * wheels (OpenCV / SciPy / Numba).  Nothing under voxel-stack-blender_b200/ may
 * link, load or call this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg use it, and only as the checker.
 *
 * Third-party algorithms restated here (their sources are NOT in /root/reference;
 * the reference pins opencv-python==4.12.0.88, scipy==1.16.1, numba==0.61.2 in
 * requirements.txt:9,10,20,25):
 *   - cv2.distanceTransform(src, DIST_L2, 5)      processing_core.py:129,250,379
 *       5x5 chamfer, a=1, b=1.4, c=2.1969 (Borgefors two-pass).  Restated twice:
 *       orc_chamfer5_twopass  = the sequential raster algorithm itself,
 *       orc_chamfer5          = its closed form  min_q T[max|d|][min|d|]  with a
 *       canonical float32 summation order (the definition the CUDA kernels match
 *       bit-for-bit).  Pinned against cv2 by tests/golden (max |diff| < 1e-4 px).
 *   - cv2.connectedComponents (8-connectivity)      processing_core.py:383
 *   - scipy.ndimage.maximum / numba pass 1          processing_core.py:291,313-320
 *   - cv2.GaussianBlur on uint8 (Q8.8 fixed point)  xy_blend_processor.py:33
 *   - cv2.bilateralFilter on uint8                  xy_blend_processor.py:40
 *   - cv2.medianBlur on uint8 (replicate border)    xy_blend_processor.py:46
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off; FMA only via fmaf()).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

static const float CH_A = 1.0f;
static const float CH_B = 1.4f;
static const float CH_C = 2.1969f;

/* ------------------------------------------------------------------------- */
/* Chamfer table: T[M*(R+1)+m], 0<=m<=M<=R, float32 sums in canonical order:
 * axis steps first (exact integer), then diagonal steps one by one, then
 * knight steps one by one.  Entries with m>M mirror (m,M).                     */
ORC_API void orc_chamfer5_table(float *T, int R)
{
    int n = R + 1;
    for (int M = 0; M <= R; ++M) {
        for (int m = 0; m <= M; ++m) {
            volatile float v;
            if (2 * m <= M) {
                v = (float)(M - 2 * m) * CH_A;
                for (int k = 0; k < m; ++k) v = v + CH_C;
            } else {
                v = 0.0f;
                for (int k = 0; k < 2 * m - M; ++k) v = v + CH_B;
                for (int k = 0; k < M - m; ++k) v = v + CH_C;
            }
            T[M * n + m] = v;
            T[m * n + M] = v;
        }
    }
}

/* Brute force definition: dist(p) = min over seeds q with Chebyshev |p-q|<=R of
 * T[max][min]; +inf if none.  O(W*H*(2R+1)^2): tiny images only.               */
ORC_API void orc_chamfer5_brute(const uint8_t *seed, int W, int H, int R,
                                const float *T, float *dist)
{
    int n = R + 1;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            float best = INFINITY;
            for (int dy = -R; dy <= R; ++dy) {
                int yy = y + dy;
                if (yy < 0 || yy >= H) continue;
                for (int dx = -R; dx <= R; ++dx) {
                    int xx = x + dx;
                    if (xx < 0 || xx >= W) continue;
                    if (!seed[(size_t)yy * W + xx]) continue;
                    int ax = abs(dx), ay = abs(dy);
                    float t = T[ax * n + ay];
                    if (t < best) best = t;
                }
            }
            dist[(size_t)y * W + x] = best;
        }
}

/* Same result as the brute force, separable: column pass gives the vertical
 * distance g to the nearest seed in each column (capped at R+1), row pass takes
 * min_k T[max(|k|,g)][min(|k|,g)] with the bound T>=max(|k|,g) pruning k.       */
ORC_API int orc_chamfer5(const uint8_t *seed, int W, int H, int R,
                         const float *T, float *dist)
{
    int n = R + 1;
    int cap = R + 1;
    uint16_t *g = (uint16_t *)malloc((size_t)W * H * sizeof(uint16_t));
    if (!g) return -1;
    for (int x = 0; x < W; ++x) {
        int run = cap;
        for (int y = 0; y < H; ++y) {
            run = seed[(size_t)y * W + x] ? 0 : (run < cap ? run + 1 : cap);
            g[(size_t)y * W + x] = (uint16_t)run;
        }
        run = cap;
        for (int y = H - 1; y >= 0; --y) {
            run = seed[(size_t)y * W + x] ? 0 : (run < cap ? run + 1 : cap);
            if (run < g[(size_t)y * W + x]) g[(size_t)y * W + x] = (uint16_t)run;
        }
    }
#pragma omp parallel for schedule(dynamic, 16)
    for (int y = 0; y < H; ++y) {
        const uint16_t *gr = g + (size_t)y * W;
        for (int x = 0; x < W; ++x) {
            float best = INFINITY;
            for (int k = 0; k <= R && (float)k < best; ++k) {
                int xl = x - k, xr = x + k;
                if (xl >= 0 && gr[xl] <= R) {
                    float t = T[k * n + gr[xl]];
                    if (t < best) best = t;
                }
                if (k && xr < W && gr[xr] <= R) {
                    float t = T[k * n + gr[xr]];
                    if (t < best) best = t;
                }
            }
            dist[(size_t)y * W + x] = best;
        }
    }
    free(g);
    return 0;
}

/* The raster algorithm cv2.distanceTransform(.., DIST_L2, 5) runs (Borgefors):
 * forward pass over the 8 causal neighbours of the 5x5 mask, backward pass over
 * the 8 anti-causal ones, float32 adds, out-of-image = +inf.  Unbounded reach.
 * Used to show the closed form above equals the library algorithm up to float
 * summation order (tests/test_oracle.py).                                     */
ORC_API void orc_chamfer5_twopass(const uint8_t *seed, int W, int H, float *dist)
{
    static const int fdx[8] = {-1, 1, -2, -1, 0, 1, 2, -1};
    static const int fdy[8] = {-2, -2, -1, -1, -1, -1, -1, 0};
    const float fc[8] = {CH_C, CH_C, CH_C, CH_B, CH_A, CH_B, CH_C, CH_A};
    for (size_t i = 0; i < (size_t)W * H; ++i) dist[i] = seed[i] ? 0.0f : INFINITY;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            float v = dist[(size_t)y * W + x];
            if (v == 0.0f) continue;
            for (int k = 0; k < 8; ++k) {
                int xx = x + fdx[k], yy = y + fdy[k];
                if (xx < 0 || xx >= W || yy < 0 || yy >= H) continue;
                float t = dist[(size_t)yy * W + xx] + fc[k];
                if (t < v) v = t;
            }
            dist[(size_t)y * W + x] = v;
        }
    for (int y = H - 1; y >= 0; --y)
        for (int x = W - 1; x >= 0; --x) {
            float v = dist[(size_t)y * W + x];
            if (v == 0.0f) continue;
            for (int k = 0; k < 8; ++k) {
                int xx = x - fdx[k], yy = y - fdy[k];
                if (xx < 0 || xx >= W || yy < 0 || yy >= H) continue;
                float t = dist[(size_t)yy * W + xx] + fc[k];
                if (t < v) v = t;
            }
            dist[(size_t)y * W + x] = v;
        }
}

/* Exact squared Euclidean distance to the nearest seed (int32), separable
 * column scan + Felzenszwalb-Huttenlocher lower envelope per row.  Oracle for
 * the `exact` metric (cv2.DIST_MASK_PRECISE equivalent); -1 where no seed.     */
ORC_API int orc_edt_exact_sq(const uint8_t *seed, int W, int H, int32_t *d2)
{
    const int64_t INF = (int64_t)1 << 40;
    int64_t *g = (int64_t *)malloc((size_t)W * H * sizeof(int64_t));
    if (!g) return -1;
    for (int x = 0; x < W; ++x) {
        int64_t run = INF;
        for (int y = 0; y < H; ++y) {
            run = seed[(size_t)y * W + x] ? 0 : (run >= INF ? INF : run + 1);
            g[(size_t)y * W + x] = run;
        }
        run = INF;
        for (int y = H - 1; y >= 0; --y) {
            run = seed[(size_t)y * W + x] ? 0 : (run >= INF ? INF : run + 1);
            if (run < g[(size_t)y * W + x]) g[(size_t)y * W + x] = run;
        }
    }
    int any = 0;
    for (size_t i = 0; i < (size_t)W * H && !any; ++i) any = seed[i] != 0;
#pragma omp parallel
    {
        int *v = (int *)malloc((size_t)W * sizeof(int));
        double *z = (double *)malloc(((size_t)W + 1) * sizeof(double));
#pragma omp for schedule(dynamic, 16)
        for (int y = 0; y < H; ++y) {
            const int64_t *gr = g + (size_t)y * W;
            int32_t *out = d2 + (size_t)y * W;
            if (!any) { for (int x = 0; x < W; ++x) out[x] = -1; continue; }
            int k = -1;
            for (int q = 0; q < W; ++q) {
                if (gr[q] >= INF) continue;
                double fq = (double)gr[q] * (double)gr[q];
                if (k < 0) { k = 0; v[0] = q; z[0] = -INFINITY; z[1] = INFINITY; continue; }
                double s;
                for (;;) {
                    int p = v[k];
                    double fp = (double)gr[p] * (double)gr[p];
                    s = ((fq + (double)q * q) - (fp + (double)p * p)) / (2.0 * (q - p));
                    if (s <= z[k] && k > 0) { --k; continue; }
                    break;
                }
                ++k; v[k] = q; z[k] = s; z[k + 1] = INFINITY;
            }
            if (k < 0) { for (int x = 0; x < W; ++x) out[x] = -1; continue; }
            int j = 0;
            for (int x = 0; x < W; ++x) {
                while (z[j + 1] < (double)x) ++j;
                /* guard against boundary ties: check neighbours of j as well */
                int64_t best = INF * 4;
                for (int t = (j > 0 ? j - 1 : 0); t <= (j < k ? j + 1 : k); ++t) {
                    int p = v[t];
                    int64_t c = (int64_t)(x - p) * (x - p) + gr[p] * gr[p];
                    if (c < best) best = c;
                }
                out[x] = best > 0x7fffffff ? 0x7fffffff : (int32_t)best;
            }
        }
        free(v); free(z);
    }
    free(g);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* 8-connected components of mask!=0, labels 1..n-1 in raster order of first
 * pixel, background 0.  Returns num_labels (= components + 1), like
 * cv2.connectedComponents (processing_core.py:383).                           */
static int32_t uf_find(int32_t *p, int32_t a)
{
    while (p[a] != a) { p[a] = p[p[a]]; a = p[a]; }
    return a;
}
ORC_API int orc_ccl8(const uint8_t *mask, int W, int H, int32_t *labels)
{
    size_t n = (size_t)W * H;
    int32_t *parent = (int32_t *)malloc((n / 2 + 2 + W) * sizeof(int32_t));
    if (!parent) return -1;
    int32_t next = 1;
    parent[0] = 0;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            size_t i = (size_t)y * W + x;
            if (!mask[i]) { labels[i] = 0; continue; }
            int32_t l = 0;
            int32_t nb[4];
            int c = 0;
            if (x > 0 && labels[i - 1]) nb[c++] = labels[i - 1];
            if (y > 0) {
                if (x > 0 && labels[i - W - 1]) nb[c++] = labels[i - W - 1];
                if (labels[i - W]) nb[c++] = labels[i - W];
                if (x + 1 < W && labels[i - W + 1]) nb[c++] = labels[i - W + 1];
            }
            if (!c) { l = next; parent[next] = next; ++next; }
            else {
                l = uf_find(parent, nb[0]);
                for (int k = 1; k < c; ++k) {
                    int32_t r = uf_find(parent, nb[k]);
                    if (r < l) { parent[l] = r; l = r; }
                    else if (r > l) parent[r] = l;
                }
            }
            labels[i] = l;
        }
    int32_t *remap = (int32_t *)calloc((size_t)next, sizeof(int32_t));
    int32_t cnt = 1;
    for (size_t i = 0; i < n; ++i) {
        if (!labels[i]) continue;
        int32_t r = uf_find(parent, labels[i]);
        if (!remap[r]) remap[r] = cnt++;
        labels[i] = remap[r];
    }
    free(remap); free(parent);
    return cnt;
}

/* per-label maximum of d over label>0 (float32 compare), mx[0]=0.
 * processing_core.py:291 (ndimage.maximum) and :313-320 (numba pass 1).       */
ORC_API void orc_label_max(const float *d, const int32_t *labels, size_t n,
                           int num_labels, float *mx)
{
    for (int i = 0; i < num_labels; ++i) mx[i] = 0.0f;
    for (size_t i = 0; i < n; ++i) {
        int32_t l = labels[i];
        if (l > 0 && d[i] > mx[l]) mx[l] = d[i];
    }
}

/* ------------------------------------------------------------------------- */
static inline int reflect101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        if (i >= n) i = 2 * (n - 1) - i;
    }
    return i;
}
static inline int clampi(int i, int n) { return i < 0 ? 0 : (i >= n ? n - 1 : i); }

/* cv2.GaussianBlur on uint8: separable Q8.8 integer kernels kx (len nkx), ky
 * (len nky), each summing to 256; out = (sum_ij kx_i*ky_j*p + 2^15) >> 16,
 * BORDER_REFLECT_101.  xy_blend_processor.py:27-33.                            */
ORC_API void orc_gauss_q88(const uint8_t *src, int W, int H, const int32_t *kx, int nkx,
                           const int32_t *ky, int nky, uint8_t *dst)
{
    int rx = nkx / 2, ry = nky / 2;
    uint32_t *tmp = (uint32_t *)malloc((size_t)W * H * sizeof(uint32_t));
#pragma omp parallel for
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            uint32_t s = 0;
            for (int i = 0; i < nkx; ++i)
                s += (uint32_t)kx[i] * src[(size_t)y * W + reflect101(x + i - rx, W)];
            tmp[(size_t)y * W + x] = s;
        }
#pragma omp parallel for
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            uint32_t s = 0;
            for (int j = 0; j < nky; ++j)
                s += (uint32_t)ky[j] * tmp[(size_t)reflect101(y + j - ry, H) * W + x];
            uint32_t v = (s + 32768u) >> 16;
            dst[(size_t)y * W + x] = (uint8_t)(v > 255 ? 255 : v);
        }
    free(tmp);
}

/* cv2.bilateralFilter on single-channel uint8 (xy_blend_processor.py:35-40):
 * circular tap set, float32 tables computed in double, taps accumulated in
 * row-major order with one float32 FMA per tap for sum(v*w) and a float32 add
 * for sum(w); result = round-half-even(sum/wsum); BORDER_REFLECT_101.
 * mode 0: round (the generic OpenCV path);  mode 1: truncate (the IPP path the
 * pip wheel takes for d in {3,5}, SURVEY finding 6).                           */
ORC_API void orc_bilateral(const uint8_t *src, int W, int H, int d, double sigma_color,
                           double sigma_space, int mode, uint8_t *dst)
{
    if (sigma_color <= 0) sigma_color = 1;
    if (sigma_space <= 0) sigma_space = 1;
    double gc = -0.5 / (sigma_color * sigma_color);
    double gs = -0.5 / (sigma_space * sigma_space);
    int radius = d <= 0 ? (int)lrint(sigma_space * 1.5) : d / 2;
    if (radius < 1) radius = 1;
    int dd = 2 * radius + 1;
    float cw[256];
    for (int i = 0; i < 256; ++i) cw[i] = (float)exp((double)i * i * gc);
    float *sw = (float *)malloc((size_t)dd * dd * sizeof(float));
    int *oi = (int *)malloc((size_t)dd * dd * sizeof(int));
    int *oj = (int *)malloc((size_t)dd * dd * sizeof(int));
    int maxk = 0;
    for (int i = -radius; i <= radius; ++i)
        for (int j = -radius; j <= radius; ++j) {
            double r = sqrt((double)i * i + (double)j * j);
            if (r > radius) continue;
            sw[maxk] = (float)exp(r * r * gs);
            oi[maxk] = i; oj[maxk] = j; ++maxk;
        }
#pragma omp parallel for schedule(dynamic, 8)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            int v0 = src[(size_t)y * W + x];
            float s = 0.0f, ws = 0.0f;
            for (int k = 0; k < maxk; ++k) {
                int v = src[(size_t)reflect101(y + oi[k], H) * W + reflect101(x + oj[k], W)];
                float w = sw[k] * cw[abs(v - v0)];
                s = fmaf((float)v, w, s);
                ws = ws + w;
            }
            float q = s / ws;
            int o = mode ? (int)q : (int)lrintf(q);
            dst[(size_t)y * W + x] = (uint8_t)(o < 0 ? 0 : (o > 255 ? 255 : o));
        }
    free(sw); free(oi); free(oj);
}

/* cv2.medianBlur on uint8: exact k*k median (the (k*k/2)-th smallest, 0-based),
 * BORDER_REPLICATE.  xy_blend_processor.py:42-46.                              */
ORC_API void orc_median(const uint8_t *src, int W, int H, int k, uint8_t *dst)
{
    int r = k / 2, need = (k * k) / 2;
#pragma omp parallel for schedule(dynamic, 8)
    for (int y = 0; y < H; ++y) {
        int hist[256];
        for (int x = 0; x < W; ++x) {
            memset(hist, 0, sizeof hist);
            for (int dy = -r; dy <= r; ++dy)
                for (int dx = -r; dx <= r; ++dx)
                    ++hist[src[(size_t)clampi(y + dy, H) * W + clampi(x + dx, W)]];
            int acc = 0, v = 0;
            for (; v < 256; ++v) { acc += hist[v]; if (acc > need) break; }
            dst[(size_t)y * W + x] = (uint8_t)v;
        }
    }
}

/* Chamfer 5x5 with unbounded reach, exact integer arithmetic: costs in units of 2^-23 (a = 2^23, b = 1.4f and
 * c = 2.1969f as exact binary fractions), result = the correctly rounded float32 of the exact minimum.  This is the
 * definition the unbounded device path (csrc/vsb_unbounded.cu) matches bit for bit; it differs from the tabulated
 * sequential-sum definition above by at most a few float32 ulps.  +inf where the image has no seed.
 * Used for the global min-max normalisation (processing_core.py:141-145) and fade distances > 64 px.          */
ORC_API int orc_chamfer5_unbounded(const uint8_t *seed, int W, int H, float *dist)
{
    const int64_t A = 8388608, B = 11744051, C = 18428932;
    const int32_t NONE = 0x7fffffff;
    int32_t *g = (int32_t *)malloc((size_t)W * H * sizeof(int32_t));
    if (!g) return -1;
    for (int x = 0; x < W; ++x) {
        int32_t run = NONE;
        for (int y = 0; y < H; ++y) {
            run = seed[(size_t)y * W + x] ? 0 : (run == NONE ? NONE : run + 1);
            g[(size_t)y * W + x] = run;
        }
        run = NONE;
        for (int y = H - 1; y >= 0; --y) {
            run = seed[(size_t)y * W + x] ? 0 : (run == NONE ? NONE : run + 1);
            if (run < g[(size_t)y * W + x]) g[(size_t)y * W + x] = run;
        }
    }
#pragma omp parallel for schedule(dynamic, 16)
    for (int y = 0; y < H; ++y) {
        const int32_t *gr = g + (size_t)y * W;
        for (int x = 0; x < W; ++x) {
            int64_t best = INT64_MAX;
            for (int side = -1; side <= 1; side += 2)
                for (int k = (side < 0 ? 0 : 1);; ++k) {
                    int xx = x + side * k;
                    if (xx < 0 || xx >= W || (int64_t)k * A >= best) break;
                    if (gr[xx] == NONE) continue;
                    int M = k > gr[xx] ? k : gr[xx], m = k > gr[xx] ? gr[xx] : k;
                    int64_t c = (2 * m <= M) ? (int64_t)(M - 2 * m) * A + (int64_t)m * C
                                             : (int64_t)(2 * m - M) * B + (int64_t)(M - m) * C;
                    if (c < best) best = c;
                }
            dist[(size_t)y * W + x] = best == INT64_MAX ? INFINITY : (float)((double)best / 8388608.0);
        }
    }
    free(g);
    return 0;
}
