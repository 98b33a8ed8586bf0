// vsb_roi.cu -- ROI_FADE gradient.
//
// Replaces _calculate_receding_gradient_field_roi_fade (processing_core.py:155-220) + merge_to_output (:442-460) +
// the leading apply_lut: for every ROI that takes part, a fade of the distance to *that ROI's* pixels on the receding
// pixels inside the ROI's box dilation, maximum over ROIs.
//
// The reference loops over ROIs with one full-image dilate + distance transform each.  Here the loop is turned inside
// out: a receding pixel looks at the labelled current mask around it once.  Every labelled seed pixel within the
// search reach updates the running minimum of its component; a component "covers" the pixel when one of its pixels
// lies within Chebyshev distance K (the k x k box dilation, k = min(51, 2*int(F)+1), K = k/2).  Fixed fade: the
// result is the fade of the smallest distance among covering components (the fade is monotone, so max over ROIs of
// the fade = fade of the min).  Min-max fade: pass 1 reduces per-component min / max over the component's zone,
// pass 2 normalises per component and takes the maximum.
#include "vsb_common.cuh"

#define ROI_SLOTS 8 // components tracked per sweep of the window; more are handled by further sweeps

struct RoiParams {
    const uint8_t *gray; size_t gpitch;
    const uint32_t *cur; const uint32_t *prior[VSB_MAX_PRIORS]; int n_priors; int wpr;
    const int32_t *cid; const uint8_t *elig;
    const float *T; int R, K;
    int W, H;
    int fixed; float fade, den;
    uint32_t *mm; // min-max mode: [2 * n_components] float bits (min initialised to +inf bits, max to 0)
    const uint8_t *lut;
    uint8_t *out; size_t opitch;
};

// One sweep over the window of pixel (x, y): collects up to ROI_SLOTS eligible components with id > lo (smallest ids
// first is not needed -- any ROI_SLOTS of them; the caller advances `lo` to the largest id seen in the sweep only when
// the table overflowed, see below).  Returns the number of components found; `over` reports whether some component
// did not fit.
struct RoiHit { int id; float d; int near; };

__device__ __forceinline__ int roi_sweep(const RoiParams &p, int x, int y, int lo, int hi, RoiHit *hit, bool &over)
{
    int n = 0;
    over = false;
    const int R = p.R, n1 = p.R + 1;
    int last = -1, last_slot = -1;
    for (int dy = -R; dy <= R; ++dy) {
        const int yy = y + dy;
        if (yy < 0 || yy >= p.H) continue;
        const int xa = max(x - R, 0), xb = min(x + R, p.W - 1);
        const int ady = abs(dy);
        for (int wi = xa >> 5; wi <= (xb >> 5); ++wi) {
            uint32_t m = p.cur[(size_t)yy * p.wpr + wi];
            if (!m) continue;
            const int w0 = wi << 5;
            if (w0 < xa) m &= 0xFFFFFFFFu << (xa - w0);
            if (w0 + 31 > xb) m &= 0xFFFFFFFFu >> (w0 + 31 - xb);
            while (m) {
                const int b = __ffs(m) - 1;
                m &= m - 1;
                const int qx = w0 + b;
                const int id = p.cid[(size_t)yy * p.W + qx] - 1;
                if (id <= lo || id >= hi) continue;
                int slot;
                if (id == last) slot = last_slot;
                else {
                    if (!p.elig[id]) continue;
                    slot = -1;
                    for (int k = 0; k < n; ++k) if (hit[k].id == id) { slot = k; break; }
                    if (slot < 0) {
                        if (n == ROI_SLOTS) {
                            // keep the ROI_SLOTS smallest ids of this sweep: replace the largest kept id if this one is smaller
                            int big = 0;
                            for (int k = 1; k < n; ++k) if (hit[k].id > hit[big].id) big = k;
                            over = true;
                            if (id > hit[big].id) { last = -1; continue; }
                            slot = big;
                        } else slot = n++;
                        hit[slot].id = id; hit[slot].d = 3.0e38f; hit[slot].near = 0;
                        // a replaced slot restarts its minimum: rows already swept are revisited because the caller
                        // re-sweeps with hi = (largest kept id + 1) when `over` is set
                    }
                    last = id; last_slot = slot;
                }
                const int adx = abs(qx - x);
                const int M = max(adx, ady), mn = min(adx, ady);
                hit[slot].d = fminf(hit[slot].d, p.T[M * n1 + mn]);
                if (M <= p.K) hit[slot].near = 1;
            }
        }
    }
    return n;
}

__device__ __forceinline__ bool roi_is_rec(const RoiParams &p, int x, int y)
{
    const size_t o = (size_t)y * p.wpr + (x >> 5);
    uint32_t acc = 0;
    for (int j = 0; j < p.n_priors; ++j) acc |= p.prior[j][o];
    return ((acc & ~p.cur[o]) >> (x & 31)) & 1u;
}

// Enumerates every covering component of a receding pixel exactly once, ROI_SLOTS at a time in increasing id order:
// a sweep keeps the smallest ids in (lo, +inf); when it overflowed, it is repeated restricted to (lo, max kept id]
// so that the kept minima are complete, then lo advances.
template <typename F>
__device__ __forceinline__ void roi_for_each(const RoiParams &p, int x, int y, F &&fn)
{
    RoiHit hit[ROI_SLOTS];
    int lo = -1;
    for (;;) {
        bool over;
        int n = roi_sweep(p, x, y, lo, 0x7fffffff, hit, over);
        int top = -1;
        for (int k = 0; k < n; ++k) top = max(top, hit[k].id);
        if (over) { bool o2; n = roi_sweep(p, x, y, lo, top + 1, hit, o2); }
        for (int k = 0; k < n; ++k) if (hit[k].near) fn(hit[k].id, hit[k].d);
        if (!over) break;
        lo = top;
    }
}

__global__ void __launch_bounds__(256) k_roi_minmax(const __grid_constant__ RoiParams p)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= p.W || y >= p.H || !roi_is_rec(p, x, y)) return;
    roi_for_each(p, x, y, [&](int id, float d) {
        atomicMin(&p.mm[2 * id], __float_as_uint(d));
        atomicMax(&p.mm[2 * id + 1], __float_as_uint(d));
    });
}

__global__ void __launch_bounds__(256) k_roi_fade(const __grid_constant__ RoiParams p)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= p.W || y >= p.H) return;
    int g = 0;
    if (p.K > 0 && roi_is_rec(p, x, y)) {
        if (p.fixed) {
            float best = 3.0e38f;
            bool any = false;
            roi_for_each(p, x, y, [&](int, float d) { best = fminf(best, d); any = true; });
            if (any) {
                const float n = __fdiv_rn(fminf(best, p.fade), p.den);
                g = (int)__fmul_rn(255.0f, __fsub_rn(1.0f, n)) & 255;
            }
        } else {
            roi_for_each(p, x, y, [&](int id, float d) {
                const float mn = __uint_as_float(p.mm[2 * id]), mx = __uint_as_float(p.mm[2 * id + 1]);
                if (mx <= mn) return; // (max_val <= min_val: the ROI is skipped, processing_core.py:197-198)
                // (d - float32(min)) / float32(double(max) - double(min)), then 255 * (1 - n), truncated
                const float den = (float)((double)mx - (double)mn);
                const float n = __fdiv_rn(__fsub_rn(d, mn), den);
                const int v = (int)__fmul_rn(255.0f, __fsub_rn(1.0f, n)) & 255;
                g = max(g, v);
            });
        }
    }
    const int raw = p.gray ? (int)p.gray[(size_t)y * p.gpitch + x] : 0;
    const int m = max(raw, g);
    p.out[(size_t)y * p.opitch + x] = p.lut ? p.lut[m] : (uint8_t)m;
}

__global__ void k_roi_mm_init(uint32_t *mm, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { mm[2 * i] = 0x7f800000u; mm[2 * i + 1] = 0u; }
}

extern "C" int vsb_roi_fade(const uint8_t *gray, size_t gray_pitch, int W, int H, const uint32_t *cur_bits,
                            const uint32_t *const *prior_bits, int n_priors, int wpr, const int32_t *cid,
                            const uint8_t *eligible_dev, int n_components, int dilate_radius, int reach, const float *T_dev,
                            int use_fixed, float fade, float den, uint32_t *minmax_scratch, const uint8_t *lut256_dev,
                            uint8_t *out, size_t out_pitch, vsb_stream_t stream)
{
    VSB_REQUIRE(cur_bits && cid && out && W > 0 && H > 0 && wpr >= (W + 31) / 32, "vsb_roi_fade: null pointer / geometry");
    VSB_REQUIRE(n_priors >= 0 && n_priors <= VSB_MAX_PRIORS && (n_priors == 0 || prior_bits), "vsb_roi_fade: priors");
    VSB_REQUIRE(n_components >= 0 && (n_components == 0 || eligible_dev), "vsb_roi_fade: eligibility table missing");
    VSB_REQUIRE(dilate_radius >= 0 && dilate_radius <= 25 && reach >= 0 && reach <= VSB_FAST_REACH && (T_dev || dilate_radius == 0),
                "vsb_roi_fade: dilation radius %d / reach %d", dilate_radius, reach);
    VSB_REQUIRE(use_fixed || minmax_scratch || n_components == 0, "vsb_roi_fade: min-max scratch missing");
    VSB_REQUIRE(!gray || gray_pitch >= (size_t)W, "vsb_roi_fade: gray pitch < width");
    VSB_REQUIRE(out_pitch >= (size_t)W, "vsb_roi_fade: out pitch < width");
    RoiParams p;
    memset(&p, 0, sizeof p);
    p.gray = gray; p.gpitch = gray_pitch; p.cur = cur_bits; p.n_priors = n_priors; p.wpr = wpr;
    for (int i = 0; i < n_priors; ++i) p.prior[i] = prior_bits[i];
    p.cid = cid; p.elig = eligible_dev; p.T = T_dev; p.R = reach; p.K = (n_components > 0 && n_priors > 0) ? dilate_radius : 0;
    p.W = W; p.H = H; p.fixed = use_fixed ? 1 : 0; p.fade = fade; p.den = den; p.mm = minmax_scratch;
    p.lut = lut256_dev; p.out = out; p.opitch = out_pitch;
    cudaStream_t s = (cudaStream_t)stream;
    dim3 grid((W + 31) / 32, (H + 7) / 8);
    if (!use_fixed && p.K > 0) {
        k_roi_mm_init<<<(n_components + 255) / 256, 256, 0, s>>>(minmax_scratch, n_components);
        int rc = vsb_check_launch("k_roi_mm_init");
        if (rc) return rc;
        k_roi_minmax<<<grid, 256, 0, s>>>(p);
        rc = vsb_check_launch("k_roi_minmax");
        if (rc) return rc;
    }
    k_roi_fade<<<grid, 256, 0, s>>>(p);
    return vsb_check_launch("k_roi_fade");
}
