// vsb_geom.cu -- unsharp mask, image resize (five interpolation modes) and the float32 / bit-mask resizes of the
// anisotropic Enhanced-EDT branch.
//
//   apply_unsharp_mask   xy_blend_processor.py:48-67   vsb_unsharp_combine  (the Gaussian is vsb_gaussian_q88)
//   apply_resize         xy_blend_processor.py:69-90   vsb_resize_plan_create / vsb_resize_u8
//   anisotropic branch   processing_core.py:357-377    vsb_resize_bits_nearest, vsb_resize_f32_linear
//
// The coordinate / coefficient tables are built on the host exactly the way the library call behind the reference
// builds them (double scale factors, float32 fractions, Q11 short coefficients, float32 area weights) and live in a
// plan object; the kernels are gathers -- one thread per output pixel, every arithmetic step in the precision and
// order the library call uses (the CPU restatement the tests compare against is pinned to the reference outputs).
#include "vsb_common.cuh"

#include <cmath>
#include <vector>

// ---------------------------------------------------------------------------------------------------------------
// unsharp: dst = saturate(rint(fma(img, 1 + amount, blurred * -amount))), optionally keeping img where the change
// is at most `threshold` grey levels
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t unsharp_px(uint32_t a, uint32_t b, float alpha, float beta, int thr)
{
    const float t = __fmaf_rn((float)a, alpha, __fmul_rn((float)b, beta));
    const int s = min(max(__float2int_rn(t), 0), 255);
    if (thr > 0 && abs((int)a - s) <= thr) return a;
    return (uint32_t)s;
}

__global__ void __launch_bounds__(256) k_unsharp(const uint8_t *__restrict__ img, size_t pitch, const uint8_t *__restrict__ blr,
                                                 size_t bpitch, int W, int H, float alpha, float beta, int thr,
                                                 uint8_t *__restrict__ out, size_t opitch, int vec)
{
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (y >= H || x0 >= W) return;
    const uint8_t *pa = img + (size_t)y * pitch + x0, *pb = blr + (size_t)y * bpitch + x0;
    uint8_t *po = out + (size_t)y * opitch + x0;
    if (vec && x0 + 16 <= W) {
        const uint4 va = vsb_ld16_stream(pa), vb = vsb_ld16_stream(pb);
        const uint32_t wa[4] = {va.x, va.y, va.z, va.w}, wb[4] = {vb.x, vb.y, vb.z, vb.w};
        uint32_t wo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t o = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                o |= unsharp_px((wa[j] >> (8 * k)) & 255u, (wb[j] >> (8 * k)) & 255u, alpha, beta, thr) << (8 * k);
            wo[j] = o;
        }
        vsb_st16_stream(po, make_uint4(wo[0], wo[1], wo[2], wo[3]));
    } else {
        for (int k = 0; k < 16 && x0 + k < W; ++k) po[k] = (uint8_t)unsharp_px(pa[k], pb[k], alpha, beta, thr);
    }
}

extern "C" int vsb_unsharp_combine(const uint8_t *img, size_t pitch, const uint8_t *blurred, size_t bpitch, int W, int H,
                                   float amount, int threshold, uint8_t *out, size_t opitch, vsb_stream_t stream)
{
    VSB_REQUIRE(img && blurred && out && W > 0 && H > 0, "vsb_unsharp_combine: null pointer / empty image");
    VSB_REQUIRE(pitch >= (size_t)W && bpitch >= (size_t)W && opitch >= (size_t)W, "vsb_unsharp_combine: pitch < width");
    // cv2.addWeighted takes double scalars and narrows them to float32 (alpha = 1 + amount computed in double)
    const float alpha = (float)(1.0 + (double)amount), beta = (float)(-(double)amount);
    const int vec = vsb_aligned16(img, pitch) && vsb_aligned16(blurred, bpitch) && vsb_aligned16(out, opitch);
    dim3 block(32, 8), grid(((W + 15) / 16 + 31) / 32, (H + 7) / 8);
    k_unsharp<<<grid, block, 0, (cudaStream_t)stream>>>(img, pitch, blurred, bpitch, W, H, alpha, beta, threshold, out, opitch, vec);
    return vsb_check_launch("k_unsharp");
}

// ---------------------------------------------------------------------------------------------------------------
// resize plan: host-built tables in one device allocation
// ---------------------------------------------------------------------------------------------------------------
enum { RK_NEAREST = 0, RK_LINEAR = 1, RK_TAPS = 2, RK_AREA_BOX = 3, RK_AREA_TAB = 4 };

struct vsb_resize_plan {
    int sW, sH, dW, dH, interp;
    int kind;          // RK_*
    int ks;            // taps per axis (RK_LINEAR: 2, RK_TAPS: 4 or 8)
    int cubic;         // RK_TAPS: float vertical pass on the first dW - dW % 8 columns
    int bx, by;        // RK_AREA_BOX factors
    int *xi, *yi;      // source indices  [d * ks + k]   (RK_NEAREST: ks = 1; RK_AREA_TAB: CSR columns)
    int *xc, *yc;      // Q11 coefficients [d * ks + k]
    float *xf, *yf;    // float32 coefficients (RK_AREA_TAB weights, float32 linear)
    int *xs, *ys;      // RK_AREA_TAB: CSR row starts [d + 1]
    void *dev;         // the single device allocation all of the above point into
    int dev_id;
};

namespace {

struct Axis {
    std::vector<int> idx, coef, start;
    std::vector<float> fco;
};

inline void scales(int ssz, int dsz, double &inv, double &scale)
{
    inv = (double)dsz / (double)ssz;
    scale = 1.0 / inv;
}

inline int q11(float c)
{
    long v = lrintf(c * 2048.0f); // saturate_cast<short>(float): round half to even
    return (int)(v < -32768 ? -32768 : (v > 32767 ? 32767 : v));
}

void frac(int ssz, int dsz, bool area_up, std::vector<int> &s, std::vector<float> &f)
{
    double inv, scale;
    scales(ssz, dsz, inv, scale);
    s.resize(dsz); f.resize(dsz);
    for (int d = 0; d < dsz; ++d) {
        if (!area_up) {
            float fx = (float)((d + 0.5) * scale - 0.5);
            const int sx = (int)floorf(fx);
            fx -= (float)sx;
            s[d] = sx; f[d] = fx;
        } else {
            const int sx = (int)floor(d * scale);
            float fx = (float)((d + 1) - (sx + 1) * inv);
            fx = fx <= 0.0f ? 0.0f : fx - floorf(fx);
            s[d] = sx; f[d] = fx;
        }
    }
}

inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

void axis_nearest(int ssz, int dsz, Axis &a)
{
    double inv, scale;
    scales(ssz, dsz, inv, scale);
    a.idx.resize(dsz);
    for (int d = 0; d < dsz; ++d) {
        const int s = (int)floor(d * scale);
        a.idx[d] = s < ssz - 1 ? s : ssz - 1;
    }
}

// two taps; the horizontal table clamps the fraction at the image edges, the vertical one clamps the rows
void axis_linear(int ssz, int dsz, bool is_x, bool area_up, Axis &a)
{
    std::vector<int> s; std::vector<float> f;
    frac(ssz, dsz, area_up, s, f);
    a.idx.resize(2 * (size_t)dsz); a.coef.resize(2 * (size_t)dsz); a.fco.resize(2 * (size_t)dsz);
    for (int d = 0; d < dsz; ++d) {
        int sx = s[d]; float fx = f[d];
        if (is_x) {
            if (sx < 0) { fx = 0.0f; sx = 0; }
            if (sx >= ssz - 1) { fx = 0.0f; sx = ssz - 1; }
        }
        a.idx[2 * d] = clampi(sx, 0, ssz - 1); a.idx[2 * d + 1] = clampi(sx + 1, 0, ssz - 1);
        const float w0 = 1.0f - fx;
        a.fco[2 * d] = w0; a.fco[2 * d + 1] = fx;
        a.coef[2 * d] = q11(w0); a.coef[2 * d + 1] = q11(fx);
    }
}

void cubic_coeffs(float x, float *c)
{
    const float A = -0.75f;
    volatile float c0 = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A; // volatile: no excess precision games
    volatile float c1 = ((A + 2) * x - (A + 3)) * x * x + 1;
    volatile float c2 = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
    c[0] = c0; c[1] = c1; c[2] = c2;
    c[3] = 1.0f - c[0] - c[1] - c[2];
}

void lanczos4_coeffs(float x, float *c)
{
    static const double s45 = 0.70710678118654752440084436210485;
    static const double cs[8][2] = {{1, 0}, {-s45, -s45}, {0, 1}, {s45, -s45}, {-1, 0}, {s45, s45}, {0, -1}, {-s45, s45}};
    float sum = 0.0f;
    const double y0 = -(double)(x + 3.0f) * M_PI * 0.25, s0 = sin(y0), c0 = cos(y0);
    for (int i = 0; i < 8; ++i) {
        const float y0_ = (x + 3.0f - (float)i);
        if (fabsf(y0_) >= 1e-6f) {
            const double y = -(double)y0_ * M_PI * 0.25;
            c[i] = (float)((cs[i][0] * s0 + cs[i][1] * c0) / (y * y));
        } else {
            c[i] = 1e30f; // the tap that sits on a sample: after normalisation it is 1 and the others vanish
        }
        sum += c[i];
    }
    sum = 1.0f / sum;
    for (int i = 0; i < 8; ++i) c[i] *= sum;
}

void axis_taps(int ssz, int dsz, int ks, Axis &a)
{
    std::vector<int> s; std::vector<float> f;
    frac(ssz, dsz, false, s, f);
    a.idx.resize((size_t)ks * dsz); a.coef.resize((size_t)ks * dsz);
    float c[8];
    for (int d = 0; d < dsz; ++d) {
        if (ks == 4) cubic_coeffs(f[d], c); else lanczos4_coeffs(f[d], c);
        for (int k = 0; k < ks; ++k) {
            a.idx[(size_t)d * ks + k] = clampi(s[d] - (ks / 2 - 1) + k, 0, ssz - 1);
            a.coef[(size_t)d * ks + k] = q11(c[k]);
        }
    }
}

void axis_area(int ssz, int dsz, Axis &a)
{
    double inv, scale;
    scales(ssz, dsz, inv, scale);
    a.start.assign(1, 0);
    for (int dx = 0; dx < dsz; ++dx) {
        const double fsx1 = dx * scale, fsx2 = fsx1 + scale;
        const double cw = scale < ssz - fsx1 ? scale : ssz - fsx1;
        int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
        sx2 = sx2 < ssz - 1 ? sx2 : ssz - 1;
        sx1 = sx1 < sx2 ? sx1 : sx2;
        if (sx1 - fsx1 > 1e-3) { a.idx.push_back(sx1 - 1); a.fco.push_back((float)((sx1 - fsx1) / cw)); }
        for (int sx = sx1; sx < sx2; ++sx) { a.idx.push_back(sx); a.fco.push_back((float)(1.0 / cw)); }
        if (fsx2 - sx2 > 1e-3) {
            double w = fsx2 - sx2;
            w = w < 1.0 ? w : 1.0;
            w = w < cw ? w : cw;
            a.idx.push_back(sx2); a.fco.push_back((float)(w / cw));
        }
        a.start.push_back((int)a.idx.size());
    }
}

size_t up16(size_t v) { return (v + 15) & ~(size_t)15; }

} // namespace

extern "C" int vsb_resize_plan_create(int sW, int sH, int dW, int dH, int interp, vsb_resize_plan **out_plan)
{
    VSB_REQUIRE(out_plan, "vsb_resize_plan_create: null out pointer");
    *out_plan = nullptr;
    VSB_REQUIRE(sW > 0 && sH > 0 && dW > 0 && dH > 0, "vsb_resize_plan_create: empty image %dx%d -> %dx%d", sW, sH, dW, dH);
    VSB_REQUIRE(interp >= VSB_INTER_NEAREST && interp <= VSB_INTER_LINEAR_F32, "vsb_resize_plan_create: interpolation %d", interp);
    vsb_resize_plan *p = new vsb_resize_plan();
    p->sW = sW; p->sH = sH; p->dW = dW; p->dH = dH; p->interp = interp;
    Axis ax, ay;
    switch (interp) {
    case VSB_INTER_NEAREST: p->kind = RK_NEAREST; p->ks = 1; axis_nearest(sW, dW, ax); axis_nearest(sH, dH, ay); break;
    case VSB_INTER_LINEAR:
    case VSB_INTER_LINEAR_F32: p->kind = RK_LINEAR; p->ks = 2; axis_linear(sW, dW, true, false, ax); axis_linear(sH, dH, false, false, ay); break;
    case VSB_INTER_CUBIC: p->kind = RK_TAPS; p->ks = 4; p->cubic = 1; axis_taps(sW, dW, 4, ax); axis_taps(sH, dH, 4, ay); break;
    case VSB_INTER_LANCZOS4: p->kind = RK_TAPS; p->ks = 8; axis_taps(sW, dW, 8, ax); axis_taps(sH, dH, 8, ay); break;
    default: // AREA: three code paths of the library call
        if (dW > sW || dH > sH) { p->kind = RK_LINEAR; p->ks = 2; axis_linear(sW, dW, true, true, ax); axis_linear(sH, dH, false, true, ay); }
        else if (sW % dW == 0 && sH % dH == 0) { p->kind = RK_AREA_BOX; p->bx = sW / dW; p->by = sH / dH; }
        else { p->kind = RK_AREA_TAB; axis_area(sW, dW, ax); axis_area(sH, dH, ay); }
    }
    // one allocation: [xi | yi | xc | yc | xf | yf | xs | ys]
    const size_t sz[8] = {ax.idx.size() * 4, ay.idx.size() * 4, ax.coef.size() * 4, ay.coef.size() * 4,
                          ax.fco.size() * 4, ay.fco.size() * 4, ax.start.size() * 4, ay.start.size() * 4};
    const void *src[8] = {ax.idx.data(), ay.idx.data(), ax.coef.data(), ay.coef.data(),
                          ax.fco.data(), ay.fco.data(), ax.start.data(), ay.start.data()};
    size_t off[8], total = 0;
    for (int i = 0; i < 8; ++i) { off[i] = total; total += up16(sz[i]); }
    if (total == 0) total = 16;
    std::vector<unsigned char> host(total, 0);
    for (int i = 0; i < 8; ++i) if (sz[i]) memcpy(host.data() + off[i], src[i], sz[i]);
    cudaError_t e = cudaGetDevice(&p->dev_id);
    if (e == cudaSuccess) e = cudaMalloc(&p->dev, total);
    if (e == cudaSuccess) e = cudaMemcpy(p->dev, host.data(), total, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        vsb_set_error("vsb_resize_plan_create: %s", cudaGetErrorString(e));
        if (p->dev) cudaFree(p->dev);
        delete p;
        return VSB_ERR_CUDA;
    }
    unsigned char *b = (unsigned char *)p->dev;
    p->xi = (int *)(b + off[0]); p->yi = (int *)(b + off[1]); p->xc = (int *)(b + off[2]); p->yc = (int *)(b + off[3]);
    p->xf = (float *)(b + off[4]); p->yf = (float *)(b + off[5]); p->xs = (int *)(b + off[6]); p->ys = (int *)(b + off[7]);
    *out_plan = p;
    return VSB_OK;
}

extern "C" int vsb_resize_plan_destroy(vsb_resize_plan *p)
{
    if (!p) return VSB_OK;
    if (p->dev) cudaFree(p->dev);
    delete p;
    return VSB_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint8_t sat_u8(int v) { return (uint8_t)min(max(v, 0), 255); }

__global__ void __launch_bounds__(256) k_rs_nearest(const uint8_t *__restrict__ in, size_t pitch, uint8_t *__restrict__ out,
                                                    size_t opitch, int dW, int dH, const int *__restrict__ xi,
                                                    const int *__restrict__ yi)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dW || y >= dH) return;
    out[(size_t)y * opitch + x] = in[(size_t)yi[y] * pitch + xi[x]];
}

// 2 taps, Q11: rows are S = a0*p0 + a1*p1 (int); column ((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16), (+2)>>2
__global__ void __launch_bounds__(256) k_rs_linear(const uint8_t *__restrict__ in, size_t pitch, uint8_t *__restrict__ out,
                                                   size_t opitch, int dW, int dH, const int *__restrict__ xi,
                                                   const int *__restrict__ xc, const int *__restrict__ yi,
                                                   const int *__restrict__ yc)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dW || y >= dH) return;
    const int x0 = xi[2 * x], x1 = xi[2 * x + 1], a0 = xc[2 * x], a1 = xc[2 * x + 1];
    const uint8_t *r0 = in + (size_t)yi[2 * y] * pitch, *r1 = in + (size_t)yi[2 * y + 1] * pitch;
    const int S0 = r0[x0] * a0 + r0[x1] * a1, S1 = r1[x0] * a0 + r1[x1] * a1;
    const int v = (((yc[2 * y] * (S0 >> 4)) >> 16) + ((yc[2 * y + 1] * (S1 >> 4)) >> 16) + 2) >> 2;
    out[(size_t)y * opitch + x] = sat_u8(v);
}

// KS taps per axis, Q11 x Q11: fixed point (sum + 2^21) >> 22 with the int wrap-around of the C code; cubic runs the
// column pass in float32 on the first `nfloat` columns (beta / 2^22, nested multiply-adds, round half even)
template <int KS>
__global__ void __launch_bounds__(256) k_rs_taps(const uint8_t *__restrict__ in, size_t pitch, uint8_t *__restrict__ out,
                                                 size_t opitch, int dW, int dH, const int *__restrict__ xi,
                                                 const int *__restrict__ xc, const int *__restrict__ yi,
                                                 const int *__restrict__ yc, int nfloat)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dW || y >= dH) return;
    int xs[KS], xa[KS];
#pragma unroll
    for (int k = 0; k < KS; ++k) { xs[k] = xi[x * KS + k]; xa[k] = xc[x * KS + k]; }
    int S[KS];
#pragma unroll
    for (int j = 0; j < KS; ++j) {
        const uint8_t *row = in + (size_t)yi[y * KS + j] * pitch;
        int h = 0;
#pragma unroll
        for (int k = 0; k < KS; ++k) h += row[xs[k]] * xa[k];
        S[j] = h;
    }
    if (KS == 4 && x < nfloat) {
        const float scale = 1.0f / (2048.0f * 2048.0f);
        float t = __fmul_rn((float)S[3], __fmul_rn((float)yc[y * 4 + 3], scale));
#pragma unroll
        for (int k = 2; k >= 0; --k) t = __fadd_rn(__fmul_rn((float)S[k], __fmul_rn((float)yc[y * 4 + k], scale)), t);
        out[(size_t)y * opitch + x] = sat_u8(__float2int_rn(t));
    } else {
        uint32_t acc = 0;
#pragma unroll
        for (int k = 0; k < KS; ++k) acc += (uint32_t)(S[k] * yc[y * KS + k]);
        out[(size_t)y * opitch + x] = sat_u8((int)(acc + (1u << 21)) >> 22);
    }
}

__global__ void __launch_bounds__(256) k_rs_box(const uint8_t *__restrict__ in, size_t pitch, uint8_t *__restrict__ out,
                                                size_t opitch, int dW, int dH, int bx, int by)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dW || y >= dH) return;
    int sum = 0;
    for (int j = 0; j < by; ++j) {
        const uint8_t *row = in + (size_t)(y * by + j) * pitch + (size_t)x * bx;
        for (int i = 0; i < bx; ++i) sum += row[i];
    }
    int v;
    if (bx == 2 && by == 2) v = (sum + 2) >> 2;
    else v = __float2int_rn(__fmul_rn((float)sum, __fdiv_rn(1.0f, (float)(bx * by))));
    out[(size_t)y * opitch + x] = sat_u8(v);
}

// general area: float32 weights; per source row a fresh horizontal sum in table order, rows accumulated in order
__global__ void __launch_bounds__(256) k_rs_area(const uint8_t *__restrict__ in, size_t pitch, uint8_t *__restrict__ out,
                                                 size_t opitch, int dW, int dH, const int *__restrict__ xs,
                                                 const int *__restrict__ xi, const float *__restrict__ xf,
                                                 const int *__restrict__ ys, const int *__restrict__ yi,
                                                 const float *__restrict__ yf)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dW || y >= dH) return;
    const int k0 = xs[x], k1 = xs[x + 1], j0 = ys[y], j1 = ys[y + 1];
    float acc = 0.0f;
    for (int j = j0; j < j1; ++j) {
        const uint8_t *row = in + (size_t)yi[j] * pitch;
        float buf = 0.0f;
        for (int k = k0; k < k1; ++k) buf = __fadd_rn(buf, __fmul_rn((float)row[xi[k]], xf[k]));
        const float term = __fmul_rn(yf[j], buf);
        acc = (j == j0) ? term : __fadd_rn(acc, term);
    }
    out[(size_t)y * opitch + x] = sat_u8(__float2int_rn(acc));
}

__global__ void __launch_bounds__(256) k_rs_linear_f32(const float *__restrict__ in, size_t pitch, float *__restrict__ out,
                                                       size_t opitch, int dW, int dH, const int *__restrict__ xi,
                                                       const float *__restrict__ xf, const int *__restrict__ yi,
                                                       const float *__restrict__ yf)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dW || y >= dH) return;
    const int x0 = xi[2 * x], x1 = xi[2 * x + 1];
    const float a0 = xf[2 * x], a1 = xf[2 * x + 1];
    const float *r0 = in + (size_t)yi[2 * y] * pitch, *r1 = in + (size_t)yi[2 * y + 1] * pitch;
    const float S0 = __fadd_rn(__fmul_rn(r0[x0], a0), __fmul_rn(r0[x1], a1));
    const float S1 = __fadd_rn(__fmul_rn(r1[x0], a0), __fmul_rn(r1[x1], a1));
    out[(size_t)y * opitch + x] = __fadd_rn(__fmul_rn(S0, yf[2 * y]), __fmul_rn(S1, yf[2 * y + 1]));
}

// nearest-neighbour resize of a bit-packed mask: one thread per output word, 32 gathered source bits
__global__ void __launch_bounds__(256) k_rs_bits(const uint32_t *__restrict__ bits, int wpr, uint32_t *__restrict__ out,
                                                 int owpr, int dW, int dH, const int *__restrict__ xi,
                                                 const int *__restrict__ yi)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (w >= owpr || y >= dH) return;
    const uint32_t *row = bits + (size_t)yi[y] * wpr;
    uint32_t o = 0;
    for (int b = 0; b < 32; ++b) {
        const int x = w * 32 + b;
        if (x < dW) { const int sx = xi[x]; o |= ((row[sx >> 5] >> (sx & 31)) & 1u) << b; }
    }
    out[(size_t)y * owpr + w] = o;
}

static int plan_check(const vsb_resize_plan *p, const char *who)
{
    VSB_REQUIRE(p, "%s: null plan", who);
    int dev = -1;
    VSB_CUDA(cudaGetDevice(&dev));
    VSB_REQUIRE(dev == p->dev_id, "%s: plan was created on device %d, current device is %d", who, p->dev_id, dev);
    return VSB_OK;
}

extern "C" int vsb_resize_u8(const vsb_resize_plan *p, const uint8_t *in, size_t pitch, uint8_t *out, size_t opitch,
                             vsb_stream_t stream)
{
    int rc = plan_check(p, "vsb_resize_u8");
    if (rc) return rc;
    VSB_REQUIRE(in && out && in != out, "vsb_resize_u8: null pointer / in-place");
    VSB_REQUIRE(pitch >= (size_t)p->sW && opitch >= (size_t)p->dW, "vsb_resize_u8: pitch < width");
    VSB_REQUIRE(p->interp != VSB_INTER_LINEAR_F32, "vsb_resize_u8: the plan is a float32 plan");
    cudaStream_t s = (cudaStream_t)stream;
    dim3 block(32, 8), grid((p->dW + 31) / 32, (p->dH + 7) / 8);
    switch (p->kind) {
    case RK_NEAREST: k_rs_nearest<<<grid, block, 0, s>>>(in, pitch, out, opitch, p->dW, p->dH, p->xi, p->yi); break;
    case RK_LINEAR: k_rs_linear<<<grid, block, 0, s>>>(in, pitch, out, opitch, p->dW, p->dH, p->xi, p->xc, p->yi, p->yc); break;
    case RK_TAPS:
        if (p->ks == 4) k_rs_taps<4><<<grid, block, 0, s>>>(in, pitch, out, opitch, p->dW, p->dH, p->xi, p->xc, p->yi, p->yc, p->dW - p->dW % 8);
        else k_rs_taps<8><<<grid, block, 0, s>>>(in, pitch, out, opitch, p->dW, p->dH, p->xi, p->xc, p->yi, p->yc, 0);
        break;
    case RK_AREA_BOX: k_rs_box<<<grid, block, 0, s>>>(in, pitch, out, opitch, p->dW, p->dH, p->bx, p->by); break;
    default: k_rs_area<<<grid, block, 0, s>>>(in, pitch, out, opitch, p->dW, p->dH, p->xs, p->xi, p->xf, p->ys, p->yi, p->yf);
    }
    return vsb_check_launch("k_resize");
}

extern "C" int vsb_resize_f32_linear(const vsb_resize_plan *p, const float *in, size_t pitch_elems, float *out,
                                     size_t opitch_elems, vsb_stream_t stream)
{
    int rc = plan_check(p, "vsb_resize_f32_linear");
    if (rc) return rc;
    VSB_REQUIRE(in && out && in != out, "vsb_resize_f32_linear: null pointer / in-place");
    VSB_REQUIRE(p->interp == VSB_INTER_LINEAR_F32, "vsb_resize_f32_linear: the plan is not a float32 linear plan");
    VSB_REQUIRE(pitch_elems >= (size_t)p->sW && opitch_elems >= (size_t)p->dW, "vsb_resize_f32_linear: pitch < width");
    dim3 block(32, 8), grid((p->dW + 31) / 32, (p->dH + 7) / 8);
    k_rs_linear_f32<<<grid, block, 0, (cudaStream_t)stream>>>(in, pitch_elems, out, opitch_elems, p->dW, p->dH, p->xi, p->xf, p->yi, p->yf);
    return vsb_check_launch("k_rs_linear_f32");
}

extern "C" int vsb_resize_bits_nearest(const vsb_resize_plan *p, const uint32_t *bits, int wpr, uint32_t *out_bits,
                                       int out_wpr, vsb_stream_t stream)
{
    int rc = plan_check(p, "vsb_resize_bits_nearest");
    if (rc) return rc;
    VSB_REQUIRE(bits && out_bits && bits != out_bits, "vsb_resize_bits_nearest: null pointer / in-place");
    VSB_REQUIRE(p->interp == VSB_INTER_NEAREST, "vsb_resize_bits_nearest: the plan is not a nearest-neighbour plan");
    VSB_REQUIRE(wpr * 32 >= p->sW && out_wpr * 32 >= p->dW, "vsb_resize_bits_nearest: words per row too small");
    dim3 block(32, 8), grid((out_wpr + 31) / 32, (p->dH + 7) / 8);
    k_rs_bits<<<grid, block, 0, (cudaStream_t)stream>>>(bits, wpr, out_bits, out_wpr, p->dW, p->dH, p->xi, p->yi);
    return vsb_check_launch("k_rs_bits");
}

extern "C" int vsb_resize_plan_dims(const vsb_resize_plan *p, int *sW, int *sH, int *dW, int *dH)
{
    VSB_REQUIRE(p, "vsb_resize_plan_dims: null plan");
    if (sW) *sW = p->sW;
    if (sH) *sH = p->sH;
    if (dW) *dW = p->dW;
    if (dH) *dH = p->dH;
    return VSB_OK;
}
