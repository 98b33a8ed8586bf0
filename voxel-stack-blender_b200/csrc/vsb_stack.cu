// vsb_stack.cu -- the per-stack runtime behind the host-buffer entry points: sliding window of packed
// prior masks, device staging slots, three-stream H2D / compute / D2H pipeline, XY op chain.
//
// Replaces the producer loop + per-layer task of ProcessingPipelineThread
// (processing_pipeline.py:74-113, :163, :188-223): deque(maxlen=receding_layers) of thresholded masks,
// newest-first snapshot per layer, Z-blend -> merge -> XY chain.  PNG decode/encode stays with the caller.
#include <vector>
#include <new>
#include <cstring>
#include <cstdlib>
#include "vsb_common.cuh"

#define STACK_SLOTS 4
#define STACK_MAX_WINDOW 128 // receding_layers the window can hold (the GUI offers 0..100, ui_components.py:168)

struct DevOp {
    int type; // 1 lut, 2 gaussian, 3 bilateral, 4 median, 5 unsharp mask, 6 resize
    uint8_t *lut_dev;
    vsb_xy_op op; // host copy of the parameters
    int iw, ih, ow, oh;      // geometry before / after the op (only a resize changes it)
    vsb_resize_plan *plan;   // type 6
};

struct vsb_stack {
    int W, H, wpr, n_priors;
    size_t pitch; // device pitch of the input staging slots (multiple of 128)
    int oW, oH;   // size of the emitted layers (a resize in the XY chain changes it)
    size_t opitch, tpitch; // pitch of the output slots / of the chain temporaries (widest intermediate)
    size_t tsize;          // bytes of one chain temporary
    uint8_t *tmp_blur;     // unsharp mask: the blurred image
    // anisotropic Enhanced EDT (allocated on first use)
    vsb_resize_plan *an_near, *an_lin;
    uint32_t *an_bits; float *an_dist, *an_T; int an_wpr;
    vsb_zparams zp;
    std::vector<DevOp> ops;      // XY chain after LUT composition
    uint8_t *lead_lut_dev;       // leading apply_lut ops composed, or nullptr
    float *T_dev;
    cudaStream_t s_h2d, s_comp, s_d2h;
    // Consecutive layers of a device batch alternate between s_comp (lane 0) and s_aux (lane 1): the bandwidth-bound
    // pack pass of layer z+1 runs beside the issue- / latency-bound layer kernel of layer z.  The window ring holds
    // n_priors + 2 masks so that the pack of z+1 never overwrites a mask layer z still reads.
    cudaStream_t s_aux;
    cudaEvent_t ev_fork, ev_join, ev_pack[2];
    bool overlap_ok;             // the configured path keeps no per-layer scratch besides the per-lane one
    uint32_t *or_scratch[2];     // more than VSB_MAX_PRIORS receding layers: OR of the window, per lane
    cudaEvent_t ev_h2d[STACK_SLOTS], ev_comp[STACK_SLOTS], ev_d2h[STACK_SLOTS];
    bool slot_used[STACK_SLOTS];
    uint8_t *gray_dev[STACK_SLOTS], *out_dev[STACK_SLOTS];
    uint8_t *tmp[2];
    std::vector<uint32_t *> ring; // n_priors + 1 packed masks
    int head, count;
    long long submitted;
    // Enhanced-EDT scratch (allocated on first use)
    uint32_t *rec_bits;
    float *dist;
    int32_t *labels;
    uint32_t *cmax;
    // unbounded-reach scratch (min-max normalisation, fade distances beyond the tile reach; allocated on first use)
    uint16_t *ub_g, *ub_gmin;
    float *ub_dist;
    uint32_t *ub_mm;
    // fused per-layer path (vsb_layer.cu): taken when the XY chain has the shape [bilateral][gaussian][lut]
    bool fused;
    vsb_xy_fused fx;
    uint8_t *lut2_dev;
    std::vector<uint16_t *> flag_ring;
    int *rec_tiles[2];     // LUT-only chains: work list of the patch kernel (1 + tiles ints), per lane
    int *gen_tiles[2];     // XY chains: tiles k_layer2 leaves to the general kernel / to its dense launch (3 x (1 + tiles) ints: two lists and the tile classes), per lane
    // environment switches, read once at creation
    bool env_no_patch, env_no_fused_enh, env_no_overlap;
    // optional per-stage CUDA-event profiling (bench.py roofline leg)
    bool prof_on;
    struct ProfRec { int stage; int launches; cudaEvent_t a, b; };
    std::vector<ProfRec> prof;
};

enum { ST_PACK = 0, ST_ZBLEND, ST_MASKDIFF, ST_ZDIST, ST_CCL, ST_EFADE, ST_PACKF, ST_LAYER, ST_OP0 };
static const char *kStageNames[] = {"threshold_pack", "zblend_fused", "maskdiff", "zblend_dist", "ccl8_max", "enhanced_fade",
                                    "pack_flags", "layer_fused"};
static const char *kOpNames[] = {"none", "lut_apply", "gaussian_q88", "bilateral", "median", "unsharp_mask", "resize"};

struct ProfScope {
    vsb_stack *s; size_t idx; bool on;
    cudaStream_t st_;
    ProfScope(vsb_stack *st, int stage, int launches, cudaStream_t stream) : s(st), idx(0), on(st->prof_on), st_(stream)
    {
        if (!on) return;
        vsb_stack::ProfRec r; r.stage = stage; r.launches = launches; r.a = r.b = nullptr;
        if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) { on = false; return; }
        cudaEventRecord(r.a, st_);
        idx = s->prof.size();
        s->prof.push_back(r);
    }
    ~ProfScope() { if (on) cudaEventRecord(s->prof[idx].b, st_); }
};

static void compose(uint8_t *acc, const uint8_t *next)
{
    uint8_t t[256];
    for (int i = 0; i < 256; ++i) t[i] = next[acc[i]]; // final = cur[final], pyside_xy_blend_tab.py:314-321
    memcpy(acc, t, 256);
}

extern "C" int vsb_stack_destroy(vsb_stack *s)
{
    if (!s) return VSB_OK;
    if (s->s_h2d) cudaStreamSynchronize(s->s_h2d);
    if (s->s_comp) cudaStreamSynchronize(s->s_comp);
    if (s->s_aux) cudaStreamSynchronize(s->s_aux);
    if (s->s_d2h) cudaStreamSynchronize(s->s_d2h);
    for (auto &o : s->ops) { if (o.lut_dev) cudaFree(o.lut_dev); if (o.plan) vsb_resize_plan_destroy(o.plan); }
    if (s->tmp_blur) cudaFree(s->tmp_blur);
    for (int i = 0; i < 2; ++i) {
        if (s->rec_tiles[i]) cudaFree(s->rec_tiles[i]);
        if (s->gen_tiles[i]) cudaFree(s->gen_tiles[i]);
        if (s->or_scratch[i]) cudaFree(s->or_scratch[i]);
        if (s->ev_pack[i]) cudaEventDestroy(s->ev_pack[i]);
    }
    if (s->ev_fork) cudaEventDestroy(s->ev_fork);
    if (s->ev_join) cudaEventDestroy(s->ev_join);
    if (s->an_near) vsb_resize_plan_destroy(s->an_near);
    if (s->an_lin) vsb_resize_plan_destroy(s->an_lin);
    if (s->an_bits) cudaFree(s->an_bits);
    if (s->an_dist) cudaFree(s->an_dist);
    if (s->an_T) cudaFree(s->an_T);
    if (s->lead_lut_dev) cudaFree(s->lead_lut_dev);
    if (s->T_dev) cudaFree(s->T_dev);
    for (int i = 0; i < STACK_SLOTS; ++i) {
        if (s->gray_dev[i]) cudaFree(s->gray_dev[i]);
        if (s->out_dev[i]) cudaFree(s->out_dev[i]);
        if (s->ev_h2d[i]) cudaEventDestroy(s->ev_h2d[i]);
        if (s->ev_comp[i]) cudaEventDestroy(s->ev_comp[i]);
        if (s->ev_d2h[i]) cudaEventDestroy(s->ev_d2h[i]);
    }
    for (int i = 0; i < 2; ++i) if (s->tmp[i]) cudaFree(s->tmp[i]);
    for (auto p : s->ring) if (p) cudaFree(p);
    for (auto p : s->flag_ring) if (p) cudaFree(p);
    if (s->lut2_dev) cudaFree(s->lut2_dev);
    for (auto &r : s->prof) { if (r.a) cudaEventDestroy(r.a); if (r.b) cudaEventDestroy(r.b); }
    if (s->rec_bits) cudaFree(s->rec_bits);
    if (s->dist) cudaFree(s->dist);
    if (s->labels) cudaFree(s->labels);
    if (s->cmax) cudaFree(s->cmax);
    if (s->ub_g) cudaFree(s->ub_g);
    if (s->ub_gmin) cudaFree(s->ub_gmin);
    if (s->ub_dist) cudaFree(s->ub_dist);
    if (s->ub_mm) cudaFree(s->ub_mm);
    if (s->s_h2d) cudaStreamDestroy(s->s_h2d);
    if (s->s_comp) cudaStreamDestroy(s->s_comp);
    if (s->s_aux) cudaStreamDestroy(s->s_aux);
    if (s->s_d2h) cudaStreamDestroy(s->s_d2h);
    delete s;
    return VSB_OK;
}

#define STACK_CUDA(call)                                                                                       \
    do {                                                                                                       \
        cudaError_t e__ = (call);                                                                              \
        if (e__ != cudaSuccess) {                                                                              \
            vsb_set_error("%s failed: %s", #call, cudaGetErrorString(e__));                                    \
            vsb_stack_destroy(s);                                                                              \
            return VSB_ERR_CUDA;                                                                               \
        }                                                                                                      \
    } while (0)

extern "C" int vsb_stack_create(int W, int H, int n_priors, const vsb_zparams *zp, const vsb_xy_op *ops, int n_ops,
                                vsb_stack **out_stack)
{
    VSB_REQUIRE(out_stack && zp, "vsb_stack_create: null pointer");
    VSB_REQUIRE(W > 0 && H > 0 && n_priors >= 0 && n_priors <= STACK_MAX_WINDOW,
                "vsb_stack_create: bad geometry or more than %d receding layers", STACK_MAX_WINDOW);
    VSB_REQUIRE(n_ops >= 0 && n_ops <= VSB_MAX_XY_OPS && (n_ops == 0 || ops), "vsb_stack_create: bad op list");
    const bool far_mode = zp->mode == VSB_Z_MINMAX || zp->mode == VSB_Z_FIXED_FAR || zp->mode == VSB_Z_WEIGHTED_FAR ||
                          zp->mode == VSB_Z_ENHANCED_FAR;
    VSB_REQUIRE(far_mode || (zp->reach >= 0 && zp->reach <= VSB_FAST_REACH), "vsb_stack_create: reach %d > %d", zp->reach,
                VSB_FAST_REACH);
    vsb_stack *s = new (std::nothrow) vsb_stack();
    VSB_REQUIRE(s, "vsb_stack_create: out of memory");
    s->W = W; s->H = H; s->wpr = ((W + 127) / 128) * 4; s->n_priors = n_priors;
    s->pitch = ((size_t)W + 127) & ~(size_t)127;
    s->zp = *zp;
    s->lead_lut_dev = nullptr; s->T_dev = nullptr;
    s->s_h2d = s->s_comp = s->s_d2h = s->s_aux = nullptr;
    s->ev_fork = s->ev_join = nullptr;
    for (int i = 0; i < 2; ++i) { s->ev_pack[i] = nullptr; s->rec_tiles[i] = nullptr; s->gen_tiles[i] = nullptr; s->or_scratch[i] = nullptr; }
    s->overlap_ok = false;
    s->env_no_patch = getenv("VSB_NO_PATCH") != nullptr;
    s->env_no_fused_enh = getenv("VSB_NO_FUSED_ENHANCED") != nullptr;
    s->env_no_overlap = getenv("VSB_NO_OVERLAP") != nullptr;
    for (int i = 0; i < STACK_SLOTS; ++i) {
        s->ev_h2d[i] = s->ev_comp[i] = s->ev_d2h[i] = nullptr;
        s->slot_used[i] = false; s->gray_dev[i] = s->out_dev[i] = nullptr;
    }
    s->tmp[0] = s->tmp[1] = nullptr; s->tmp_blur = nullptr;
    s->an_near = s->an_lin = nullptr; s->an_bits = nullptr; s->an_dist = s->an_T = nullptr; s->an_wpr = 0;
    s->oW = W; s->oH = H; s->opitch = s->tpitch = s->pitch; s->tsize = 0;
    s->head = 0; s->count = 0; s->submitted = 0;
    s->rec_bits = nullptr; s->dist = nullptr; s->labels = nullptr; s->cmax = nullptr;
    s->ub_g = s->ub_gmin = nullptr; s->ub_dist = nullptr; s->ub_mm = nullptr;
    s->prof_on = false;
    s->fused = false; s->lut2_dev = nullptr;
    memset(&s->fx, 0, sizeof s->fx);

    // XY chain: compose runs of apply_lut; a leading run is folded into the Z-blend epilogue
    uint8_t lead[256];
    bool have_lead = false, leading = true;
    for (int i = 0; i < n_ops; ++i) {
        const vsb_xy_op &op = ops[i];
        if (op.type == 0) continue;
        if (op.type == 1) {
            if (leading) {
                if (!have_lead) { memcpy(lead, op.lut, 256); have_lead = true; }
                else compose(lead, op.lut);
            } else if (!s->ops.empty() && s->ops.back().type == 1) {
                compose(s->ops.back().op.lut, op.lut);
            } else {
                DevOp d; d.type = 1; d.lut_dev = nullptr; d.plan = nullptr; d.op = op; s->ops.push_back(d);
            }
            continue;
        }
        leading = false;
        if (op.type == 4 && op.median_ksize <= 1) continue; // apply_median_blur no-op, xy_blend_processor.py:45
        if (op.type < 1 || op.type > 6) { vsb_set_error("vsb_stack_create: unknown op type %d", op.type); delete s; return VSB_ERR_ARG; }
        DevOp d; d.type = op.type; d.lut_dev = nullptr; d.plan = nullptr; d.op = op; s->ops.push_back(d);
    }
    {   // geometry along the chain
        int cw = W, ch = H;
        size_t maxw = (size_t)W, maxh = (size_t)H;
        bool bad = false;
        for (auto &o : s->ops) {
            o.iw = cw; o.ih = ch;
            if (o.type == 6) {
                if (o.op.resize_w <= 0 || o.op.resize_h <= 0 || o.op.interp < VSB_INTER_NEAREST || o.op.interp > VSB_INTER_AREA) { bad = true; break; }
                cw = o.op.resize_w; ch = o.op.resize_h;
            }
            o.ow = cw; o.oh = ch;
            if ((size_t)cw > maxw) maxw = (size_t)cw;
            if ((size_t)ch > maxh) maxh = (size_t)ch;
        }
        if (bad) { vsb_set_error("vsb_stack_create: resize op needs a positive target size and a vsb_interp mode"); delete s; return VSB_ERR_ARG; }
        s->oW = cw; s->oH = ch;
        s->opitch = ((size_t)cw + 127) & ~(size_t)127;
        s->tpitch = (maxw + 127) & ~(size_t)127;
        s->tsize = s->tpitch * maxh;
    }
    {   // does the composed chain have the fusable shape [bilateral r<=4][gaussian k<=7][lut] ?
        size_t i = 0;
        const size_t n = s->ops.size();
        bool ok = getenv("VSB_NO_FUSED") == nullptr && W >= 8 && H >= 8 &&
                  (zp->mode == VSB_Z_FIXED || zp->mode == VSB_Z_WEIGHTED || zp->mode == VSB_Z_CONST || zp->mode == VSB_Z_ENHANCED || far_mode);
        const uint8_t *lut2 = nullptr;
        if (ok && i < n && s->ops[i].type == 3) {
            const vsb_xy_op &o = s->ops[i].op;
            if (o.radius >= 1 && o.radius <= 4 && o.ntaps <= 96) {
                s->fx.has_bilateral = 1; s->fx.radius = o.radius; s->fx.ntaps = o.ntaps;
                memcpy(s->fx.tap_dy, o.tap_dy, o.ntaps); memcpy(s->fx.tap_dx, o.tap_dx, o.ntaps);
                memcpy(s->fx.tap_sw, o.tap_sw, o.ntaps * sizeof(float)); memcpy(s->fx.cw, o.cw, sizeof s->fx.cw);
                ++i;
            }
        }
        if (ok && i < n && s->ops[i].type == 2 && s->ops[i].op.nkx <= 7 && s->ops[i].op.nky <= 7) {
            const vsb_xy_op &o = s->ops[i].op;
            s->fx.nkx = o.nkx; s->fx.nky = o.nky;
            memcpy(s->fx.kx, o.kx, o.nkx * sizeof(int32_t)); memcpy(s->fx.ky, o.ky, o.nky * sizeof(int32_t));
            ++i;
        }
        if (ok && i < n && s->ops[i].type == 1) { lut2 = s->ops[i].op.lut; ++i; }
        s->fused = ok && i == n;
        if (s->fused && lut2) {
            STACK_CUDA(cudaMalloc((void **)&s->lut2_dev, 256));
            STACK_CUDA(cudaMemcpy(s->lut2_dev, lut2, 256, cudaMemcpyHostToDevice));
        }
    }
    STACK_CUDA(cudaStreamCreateWithFlags(&s->s_h2d, cudaStreamNonBlocking));
    STACK_CUDA(cudaStreamCreateWithFlags(&s->s_comp, cudaStreamNonBlocking));
    STACK_CUDA(cudaStreamCreateWithFlags(&s->s_d2h, cudaStreamNonBlocking));
    STACK_CUDA(cudaStreamCreateWithFlags(&s->s_aux, cudaStreamNonBlocking));
    STACK_CUDA(cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
    STACK_CUDA(cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) STACK_CUDA(cudaEventCreateWithFlags(&s->ev_pack[i], cudaEventDisableTiming));
    for (int i = 0; i < STACK_SLOTS; ++i) {
        STACK_CUDA(cudaEventCreateWithFlags(&s->ev_h2d[i], cudaEventDisableTiming));
        STACK_CUDA(cudaEventCreateWithFlags(&s->ev_comp[i], cudaEventDisableTiming));
        STACK_CUDA(cudaEventCreateWithFlags(&s->ev_d2h[i], cudaEventDisableTiming));
        STACK_CUDA(cudaMalloc((void **)&s->gray_dev[i], s->pitch * H));
        STACK_CUDA(cudaMalloc((void **)&s->out_dev[i], s->opitch * s->oH));
    }
    if (!s->ops.empty()) {
        STACK_CUDA(cudaMalloc((void **)&s->tmp[0], s->tsize));
        if (s->ops.size() > 1) STACK_CUDA(cudaMalloc((void **)&s->tmp[1], s->tsize));
    }
    for (auto &o : s->ops) {
        if (o.type == 5 && !s->tmp_blur) STACK_CUDA(cudaMalloc((void **)&s->tmp_blur, s->tsize));
        if (o.type == 6 && vsb_resize_plan_create(o.iw, o.ih, o.ow, o.oh, o.op.interp, &o.plan) != VSB_OK) { vsb_stack_destroy(s); return VSB_ERR_CUDA; }
    }
    if (have_lead) {
        STACK_CUDA(cudaMalloc((void **)&s->lead_lut_dev, 256));
        STACK_CUDA(cudaMemcpy(s->lead_lut_dev, lead, 256, cudaMemcpyHostToDevice));
    }
    for (auto &o : s->ops)
        if (o.type == 1) {
            STACK_CUDA(cudaMalloc((void **)&o.lut_dev, 256));
            STACK_CUDA(cudaMemcpy(o.lut_dev, o.op.lut, 256, cudaMemcpyHostToDevice));
        }
    {
        const int treach = far_mode ? 0 : zp->reach;
        const int n1 = treach + 1;
        std::vector<float> T((size_t)n1 * n1);
        vsb_chamfer_table(T.data(), treach);
        STACK_CUDA(cudaMalloc((void **)&s->T_dev, T.size() * sizeof(float)));
        STACK_CUDA(cudaMemcpy(s->T_dev, T.data(), T.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    s->ring.assign((size_t)n_priors + 2, nullptr);
    for (auto &p : s->ring) STACK_CUDA(cudaMalloc((void **)&p, (size_t)s->wpr * H * sizeof(uint32_t)));
    if (n_priors > VSB_MAX_PRIORS)
        for (int i = 0; i < 2; ++i) STACK_CUDA(cudaMalloc((void **)&s->or_scratch[i], (size_t)s->wpr * H * sizeof(uint32_t)));
    if (s->fused) {
        const size_t ntiles = (size_t)((W + VSB_TW - 1) / VSB_TW) * ((H + VSB_TH - 1) / VSB_TH);
        s->flag_ring.assign((size_t)n_priors + 2, nullptr);
        for (auto &p : s->flag_ring) STACK_CUDA(cudaMalloc((void **)&p, ntiles * sizeof(uint16_t)));
        for (int i = 0; i < 2; ++i) {
            STACK_CUDA(cudaMalloc((void **)&s->rec_tiles[i], (ntiles + 1) * sizeof(int)));
            STACK_CUDA(cudaMalloc((void **)&s->gen_tiles[i], 3 * (ntiles + 1) * sizeof(int)));
        }
        if (s->fx.has_bilateral) { // may a two-valued window be taken as the identity?  (values after the leading LUT)
            const int v0 = have_lead ? lead[0] : 0, v1 = have_lead ? lead[255] : 255;
            s->fx.two_valued_min_same = vsb_bilateral_two_valued_min_same(s->fx.tap_sw, s->fx.ntaps, s->fx.cw, v0, v1);
        }
    }
    // two-lane overlap: only the paths whose per-layer scratch is per lane (the fused layer kernel and the LUT-only
    // patch path); Enhanced EDT / unbounded / unfused chains share scratch images and stay on one stream
    s->overlap_ok = !s->env_no_overlap && s->fused &&
                    (zp->mode == VSB_Z_FIXED || zp->mode == VSB_Z_WEIGHTED || zp->mode == VSB_Z_CONST);
    *out_stack = s;
    return VSB_OK;
}

extern "C" int vsb_stack_reset(vsb_stack *s)
{
    VSB_REQUIRE(s, "vsb_stack_reset: null stack");
    s->head = 0; s->count = 0;
    return VSB_OK;
}

extern "C" int vsb_stack_slots(vsb_stack *s) { return s ? STACK_SLOTS : 0; }

extern "C" int vsb_stack_out_size(vsb_stack *s, int *out_W, int *out_H)
{
    VSB_REQUIRE(s, "vsb_stack_out_size: null stack");
    if (out_W) *out_W = s->oW;
    if (out_H) *out_H = s->oH;
    return VSB_OK;
}
extern "C" vsb_stream_t vsb_stack_stream(vsb_stack *s) { return s ? (vsb_stream_t)s->s_comp : nullptr; }

static int ensure_unbounded_scratch(vsb_stack *s)
{
    if (s->ub_g) return VSB_OK;
    const size_t n = (size_t)s->W * s->H;
    VSB_CUDA(cudaMalloc((void **)&s->ub_g, n * sizeof(uint16_t)));
    VSB_CUDA(cudaMalloc((void **)&s->ub_gmin, vsb_coldist_scratch_elems(s->W, s->H, s->wpr) * sizeof(uint16_t)));
    VSB_CUDA(cudaMalloc((void **)&s->ub_dist, n * sizeof(float)));
    VSB_CUDA(cudaMalloc((void **)&s->ub_mm, 2 * sizeof(uint32_t)));
    return VSB_OK;
}

static int ensure_enhanced_scratch(vsb_stack *s)
{
    if (s->dist) return VSB_OK;
    const size_t n = (size_t)s->W * s->H;
    VSB_CUDA(cudaMalloc((void **)&s->rec_bits, (size_t)s->wpr * s->H * sizeof(uint32_t)));
    VSB_CUDA(cudaMalloc((void **)&s->dist, n * sizeof(float)));
    VSB_CUDA(cudaMalloc((void **)&s->labels, n * sizeof(int32_t)));
    VSB_CUDA(cudaMalloc((void **)&s->cmax, n * sizeof(uint32_t)));
    return VSB_OK;
}

static int ensure_aniso_scratch(vsb_stack *s, const vsb_zparams &zp)
{
    if (s->an_near) return VSB_OK;
    const int aw = zp.aniso_w, ah = zp.aniso_h;
    VSB_REQUIRE(aw > 0 && ah > 0 && zp.aniso_reach >= 0 && zp.aniso_reach <= VSB_FAST_REACH,
                "anisotropic Enhanced EDT: resized mask %dx%d / reach %d not supported", aw, ah, zp.aniso_reach);
    int rc = vsb_resize_plan_create(s->W, s->H, aw, ah, VSB_INTER_NEAREST, &s->an_near);
    if (rc) return rc;
    rc = vsb_resize_plan_create(aw, ah, s->W, s->H, VSB_INTER_LINEAR_F32, &s->an_lin);
    if (rc) return rc;
    s->an_wpr = ((aw + 127) / 128) * 4;
    VSB_CUDA(cudaMalloc((void **)&s->an_bits, (size_t)s->an_wpr * ah * sizeof(uint32_t)));
    VSB_CUDA(cudaMalloc((void **)&s->an_dist, (size_t)aw * ah * sizeof(float)));
    const int n1 = zp.aniso_reach + 1;
    std::vector<float> T((size_t)n1 * n1);
    vsb_chamfer_table(T.data(), zp.aniso_reach);
    VSB_CUDA(cudaMalloc((void **)&s->an_T, T.size() * sizeof(float)));
    VSB_CUDA(cudaMemcpy(s->an_T, T.data(), T.size() * sizeof(float), cudaMemcpyHostToDevice));
    return VSB_OK;
}

// The window's prior masks for this layer, newest first.  More than VSB_MAX_PRIORS of them (the GUI allows 100): every
// mode but the weighted stack only uses their OR (processing_core.py:61-65, :346), so they are OR-ed into the lane's
// scratch mask in chunks and handed on as one prior; the weighted stack only ever reads the first len(weights) <= 16.
static int window_priors(vsb_stack *s, int lane, vsb_stream_t st, const uint32_t **pri, int *np_out)
{
    const int nring = s->n_priors + 2;
    const int np = s->count < s->n_priors ? s->count : s->n_priors;
    const bool weighted = s->zp.mode == VSB_Z_WEIGHTED || s->zp.mode == VSB_Z_WEIGHTED_FAR;
    if (np <= VSB_MAX_PRIORS || weighted) {
        const int n = np < VSB_MAX_PRIORS ? np : VSB_MAX_PRIORS;
        for (int i = 0; i < n; ++i) pri[i] = s->ring[(s->head - 1 - i + 2 * nring) % nring];
        *np_out = n;
        return VSB_OK;
    }
    uint32_t *acc = s->or_scratch[lane];
    for (int done = 0; done < np;) {
        const uint32_t *src[VSB_MAX_PRIORS];
        int n = 0;
        if (done > 0) src[n++] = acc;
        while (n < VSB_MAX_PRIORS && done < np) { src[n++] = s->ring[(s->head - 1 - done + 2 * nring) % nring]; ++done; }
        const int rc = vsb_bits_or(src, n, s->wpr, s->H, acc, st);
        if (rc) return rc;
    }
    pri[0] = acc;
    *np_out = 1;
    return VSB_OK;
}

// all kernels of one layer on the lane's stream (lane 0 = s_comp, lane 1 = s_aux); gray/out are device pointers.
// `signal_pack`: record ev_pack[lane] once the layer's mask (and tile flags) are written -- what the next layer,
// running on the other lane, waits for.
static int run_layer(vsb_stack *s, const uint8_t *gray, size_t gpitch, uint8_t *out, size_t opitch, bool emit, int lane = 0,
                     bool signal_pack = false)
{
    cudaStream_t cst = lane ? s->s_aux : s->s_comp;
    vsb_stream_t st = (vsb_stream_t)cst;
    const int nring = s->n_priors + 2;
    uint32_t *cur = s->ring[s->head];
    int rc;
    const uint32_t *pri[VSB_MAX_PRIORS];
    int np = 0;
    rc = window_priors(s, lane, st, pri, &np);
    if (rc) return rc;
    // LUT-only chain with a tile-reach fade: the pack pass also writes lut[gray], the layer kernel only patches the
    // receding pixels (the gray layer is read once)
    const bool patch_path = s->fused && emit && s->ops.empty() && !s->env_no_patch &&
                            (s->zp.mode == VSB_Z_FIXED || s->zp.mode == VSB_Z_WEIGHTED || s->zp.mode == VSB_Z_CONST);
    if (patch_path) {
        ProfScope ps(s, ST_PACKF, 1, cst);
        const uint16_t *pfl[VSB_MAX_PRIORS];
        const bool have_flags = s->n_priors <= VSB_MAX_PRIORS; // an OR-ed window has no tile flags
        for (int i = 0; i < np; ++i) pfl[i] = have_flags ? s->flag_ring[(s->head - 1 - i + 2 * nring) % nring] : nullptr;
        rc = vsb_pack_flags_lut(gray, gpitch, s->W, s->H, 127, cur, s->wpr, s->flag_ring[s->head], s->lead_lut_dev, out, opitch,
                                pri, have_flags ? pfl : nullptr, np, s->rec_tiles[lane], st);
    }
    else if (s->fused) { ProfScope ps(s, ST_PACKF, 1, cst); rc = vsb_pack_flags(gray, gpitch, s->W, s->H, 127, cur, s->wpr, s->flag_ring[s->head], st); }
    else { ProfScope ps(s, ST_PACK, 1, cst); rc = vsb_threshold_pack(gray, gpitch, s->W, s->H, 127, cur, s->wpr, st); }
    if (rc) return rc;
    if (signal_pack) VSB_CUDA(cudaEventRecord(s->ev_pack[lane], cst));
    if (emit) {
        vsb_zparams zp = s->zp;
        if (zp.mode == VSB_Z_WEIGHTED || zp.mode == VSB_Z_WEIGHTED_FAR) {
            // k = min(len(priors), len(weights)); total weight is over the k used weights (processing_core.py:231-238)
            const int k = np < zp.n_priors ? np : zp.n_priors;
            float tot = 0.0f;
            double totd = 0.0;
            for (int i = 0; i < k; ++i) totd += (double)zp.weights[i];
            tot = (float)totd;
            zp.n_priors = (totd > 0.0) ? k : 0;
            zp.total_weight = tot;
        } else {
            zp.n_priors = np;
        }
        const bool far_mode = zp.mode == VSB_Z_MINMAX || zp.mode == VSB_Z_FIXED_FAR || zp.mode == VSB_Z_WEIGHTED_FAR;
        const bool enhanced = zp.mode == VSB_Z_ENHANCED || zp.mode == VSB_Z_ENHANCED_FAR;
        if (s->fused && !enhanced && !far_mode) {
            const uint16_t *pfl[VSB_MAX_PRIORS];
            const bool window_flags = s->n_priors <= VSB_MAX_PRIORS; // an OR-ed window has no tile flags
            for (int i = 0; i < np; ++i) pfl[i] = window_flags ? s->flag_ring[(s->head - 1 - i + 2 * nring) % nring] : nullptr;
            if (patch_path) {
                ProfScope ps(s, ST_LAYER, 1, cst);
                rc = vsb_layer_patch(gray, gpitch, s->W, s->H, cur, pri, s->wpr, &zp, s->T_dev, s->lead_lut_dev, s->rec_tiles[lane], out, opitch, st);
            } else {
                ProfScope ps(s, ST_LAYER, 1, cst);
                // the tile-flag shortcut saves the input read of quiet tiles; it only pays once the kernel is HBM bound
                rc = vsb_layer_fused(gray, gpitch, s->W, s->H, cur, pri, s->flag_ring[s->head], pfl, s->wpr, &zp, s->T_dev,
                                     s->lead_lut_dev, &s->fx, s->lut2_dev, out, opitch, s->gen_tiles[lane], st);
            }
            if (rc) return rc;
            s->head = (s->head + 1) % nring;
            if (s->count < s->n_priors) ++s->count;
            return VSB_OK;
        }
        const size_t nops = s->ops.size();
        uint8_t *zdst = nops ? s->tmp[0] : out;
        const size_t zpitch = nops ? s->tpitch : opitch;
        if (far_mode) {
            rc = ensure_unbounded_scratch(s);
            if (rc) return rc;
            ProfScope ps(s, ST_ZBLEND, 5, cst);
            if (zp.mode == VSB_Z_WEIGHTED_FAR)
                rc = vsb_weighted_unbounded(gray, gpitch, s->W, s->H, cur, pri, s->wpr, &zp, s->ub_g, (size_t)s->W, s->ub_gmin,
                                            s->ub_dist, (size_t)s->W, s->ub_mm, s->lead_lut_dev, zdst, zpitch, st);
            else
                rc = vsb_zblend_unbounded(gray, gpitch, s->W, s->H, cur, pri, np, s->wpr, zp.metric, zp.mode == VSB_Z_MINMAX ? 1 : 0,
                                          zp.fade, zp.den, s->ub_g, (size_t)s->W, s->ub_gmin, s->ub_dist, (size_t)s->W, s->ub_mm,
                                          s->lead_lut_dev, zdst, zpitch, st);
            if (rc) return rc;
        } else if (enhanced) {
            rc = ensure_enhanced_scratch(s);
            if (rc) return rc;
            const bool plain_enh = zp.mode == VSB_Z_ENHANCED && zp.aniso_w == 0; // receding mask + distances in one launch
            if (!plain_enh) {
                ProfScope ps(s, ST_MASKDIFF, 1, cst);
                rc = vsb_maskdiff(cur, pri, np, s->wpr, s->H, s->rec_bits, nullptr, st);
                if (rc) return rc;
            }
            if (zp.mode == VSB_Z_ENHANCED_FAR) {
                rc = ensure_unbounded_scratch(s);
                if (rc) return rc;
                ProfScope ps(s, ST_ZDIST, 4, cst);
                rc = vsb_rec_dist_unbounded(cur, pri, np, s->wpr, s->W, s->H, zp.metric, s->ub_g, (size_t)s->W, s->ub_gmin, s->dist,
                                            (size_t)s->W, s->ub_mm, st);
            } else if (zp.aniso_w > 0) {
                // processing_core.py:357-377: nearest-resize the mask, transform there, bilinear-resize the map back.
                // Distances stay in resized-pixel units; pixels farther than the reach get a stand-in >= fade_limit,
                // which fades to 0 exactly like the true distance (bilinear taps are adjacent pixels, reach covers +3)
                rc = ensure_aniso_scratch(s, zp);
                if (rc) return rc;
                ProfScope ps(s, ST_ZDIST, 3, cst);
                const float far = (float)((zp.fade_limit > 0.0 ? zp.fade_limit : 0.0) + 4.0);
                rc = vsb_resize_bits_nearest(s->an_near, cur, s->wpr, s->an_bits, s->an_wpr, st);
                if (!rc) rc = vsb_dt_chamfer5(s->an_bits, s->an_wpr, zp.aniso_w, zp.aniso_h, zp.aniso_reach, s->an_T, far, s->an_dist,
                                              (size_t)zp.aniso_w, st);
                if (!rc) rc = vsb_resize_f32_linear(s->an_lin, s->an_dist, (size_t)zp.aniso_w, s->dist, (size_t)s->W, st);
            } else {
                ProfScope ps(s, ST_ZDIST, 1, cst);
                rc = vsb_rec_dist(cur, pri, s->wpr, s->W, s->H, &zp, s->T_dev, s->rec_bits, s->dist, (size_t)s->W, st);
            }
            if (rc) return rc;
            { ProfScope ps(s, ST_CCL, 3, cst); rc = vsb_ccl8_max(s->rec_bits, s->wpr, s->W, s->H, s->dist, (size_t)s->W, s->labels, s->cmax, st); }
            if (rc) return rc;
            if (s->fused && !s->env_no_fused_enh) {
                // the fade itself runs inside the layer kernel, straight into the XY chain (no intermediate image)
                ProfScope ps(s, ST_LAYER, 1, cst);
                rc = vsb_layer_fused_enhanced(gray, gpitch, s->W, s->H, cur, pri, s->wpr, &zp, s->dist, (size_t)s->W, s->labels,
                                              s->cmax, s->lead_lut_dev, &s->fx, s->lut2_dev, out, opitch, s->flag_ring[s->head],
                                              s->gen_tiles[lane], st);
                if (rc) return rc;
                s->head = (s->head + 1) % nring;
                if (s->count < s->n_priors) ++s->count;
                return VSB_OK;
            }
            { ProfScope ps(s, ST_EFADE, 1, cst);
              rc = vsb_enhanced_fade(gray, gpitch, s->W, s->H, s->rec_bits, s->wpr, s->dist, (size_t)s->W, s->labels,
                                     s->cmax, zp.fade_limit, zp.enhanced_arith, s->lead_lut_dev, zdst, zpitch, st); }
            if (rc) return rc;
        } else {
            { ProfScope ps(s, ST_ZBLEND, 1, cst);
              rc = vsb_zblend_fused(gray, gpitch, s->W, s->H, cur, pri, s->wpr, &zp, s->T_dev, s->lead_lut_dev, zdst, zpitch,
                                    nullptr, 0, st); }
            if (rc) return rc;
        }
        const uint8_t *src = zdst;
        size_t spitch = zpitch;
        if (s->fused && nops) { // Enhanced EDT: its own stage 1, then the fused XY chain
            vsb_zparams zn = zp; zn.mode = VSB_Z_NONE; zn.n_priors = 0;
            ProfScope ps(s, ST_LAYER, 1, cst);
            rc = vsb_layer_fused(src, spitch, s->W, s->H, nullptr, nullptr, nullptr, nullptr, s->wpr, &zn, nullptr, nullptr,
                                 &s->fx, s->lut2_dev, out, opitch, nullptr, st);
            if (rc) return rc;
            s->head = (s->head + 1) % nring;
            if (s->count < s->n_priors) ++s->count;
            return VSB_OK;
        }
        for (size_t i = 0; i < nops; ++i) {
            const bool last = (i + 1 == nops);
            uint8_t *dst = last ? out : s->tmp[(i + 1) & 1];
            const size_t dpitch = last ? opitch : s->tpitch;
            const DevOp &o = s->ops[i];
            ProfScope ps(s, ST_OP0 + (int)i, o.type == 5 ? 3 : 1, cst);
            switch (o.type) {
            case 1: rc = vsb_lut_apply(src, spitch, o.iw, o.ih, o.lut_dev, dst, dpitch, st); break;
            case 2: rc = vsb_gaussian_q88(src, spitch, o.iw, o.ih, o.op.kx, o.op.nkx, o.op.ky, o.op.nky, dst, dpitch, st); break;
            case 3: rc = vsb_bilateral(src, spitch, o.iw, o.ih, o.op.radius, o.op.ntaps, o.op.tap_dy, o.op.tap_dx,
                                       o.op.tap_sw, o.op.cw, dst, dpitch, st); break;
            case 4: rc = vsb_median(src, spitch, o.iw, o.ih, o.op.median_ksize, dst, dpitch, st); break;
            case 5:
                rc = vsb_gaussian_q88(src, spitch, o.iw, o.ih, o.op.kx, o.op.nkx, o.op.ky, o.op.nky, s->tmp_blur, s->tpitch, st);
                if (!rc) rc = vsb_unsharp_combine(src, spitch, s->tmp_blur, s->tpitch, o.iw, o.ih, o.op.unsharp_amount,
                                                  o.op.unsharp_threshold, dst, dpitch, st);
                break;
            case 6: rc = vsb_resize_u8(o.plan, src, spitch, dst, dpitch, st); break;
            default: rc = VSB_ERR_ARG;
            }
            if (rc) return rc;
            src = dst; spitch = dpitch;
        }
    }
    s->head = (s->head + 1) % nring;
    if (s->count < s->n_priors) ++s->count;
    return VSB_OK;
}

extern "C" int vsb_stack_process_device(vsb_stack *s, const uint8_t *gray_dev, size_t in_pitch, uint8_t *out_dev,
                                        size_t out_pitch)
{
    VSB_REQUIRE(s && gray_dev && out_dev, "vsb_stack_process_device: null pointer");
    return run_layer(s, gray_dev, in_pitch, out_dev, out_pitch, true);
}

extern "C" int vsb_stack_process_device_batch(vsb_stack *s, const uint8_t *gray_dev, size_t in_pitch,
                                              long long in_layer_stride, uint8_t *out_dev, size_t out_pitch,
                                              long long out_layer_stride, int out_ring, int n_layers)
{
    VSB_REQUIRE(s && gray_dev && out_dev && n_layers >= 0 && out_ring >= 1, "vsb_stack_process_device_batch: bad arguments");
    // two lanes: layer i runs on lane i & 1 and starts once the previous layer's mask is written; everything is forked
    // from and joined back to s_comp, so callers keep ordering their work on the one stream vsb_stack_stream() returns.
    // Needs two distinct output buffers in flight and no per-stage timing (the events of two lanes would interleave).
    const bool overlap = s->overlap_ok && !s->prof_on && n_layers > 1 && out_ring >= 2;
    if (overlap) {
        VSB_CUDA(cudaEventRecord(s->ev_fork, s->s_comp));
        VSB_CUDA(cudaStreamWaitEvent(s->s_aux, s->ev_fork, 0));
    }
    for (int i = 0; i < n_layers; ++i) {
        const int lane = overlap ? (i & 1) : 0;
        if (overlap && i > 0) VSB_CUDA(cudaStreamWaitEvent(lane ? s->s_aux : s->s_comp, s->ev_pack[lane ^ 1], 0));
        int rc = run_layer(s, gray_dev + (long long)i * in_layer_stride, in_pitch,
                           out_dev + (long long)(i % out_ring) * out_layer_stride, out_pitch, true, lane, overlap);
        if (rc) return rc;
    }
    if (overlap) {
        VSB_CUDA(cudaEventRecord(s->ev_join, s->s_aux));
        VSB_CUDA(cudaStreamWaitEvent(s->s_comp, s->ev_join, 0));
    }
    return VSB_OK;
}

extern "C" int vsb_stack_sync(vsb_stack *s)
{
    VSB_REQUIRE(s, "vsb_stack_sync: null stack");
    VSB_CUDA(cudaStreamSynchronize(s->s_h2d));
    VSB_CUDA(cudaStreamSynchronize(s->s_comp));
    VSB_CUDA(cudaStreamSynchronize(s->s_aux));
    VSB_CUDA(cudaStreamSynchronize(s->s_d2h));
    return VSB_OK;
}

static int submit(vsb_stack *s, const uint8_t *gray_host, size_t in_pitch, uint8_t *out_host, size_t out_pitch, bool emit)
{
    const int slot = (int)(s->submitted % STACK_SLOTS);
    if (s->slot_used[slot]) VSB_CUDA(cudaEventSynchronize(s->ev_d2h[slot])); // slot's previous layer fully drained
    // equal pitches: one linear copy up to the last valid byte (the padding columns travel along, the DMA engine
    // runs at full rate)
    if (in_pitch == s->pitch) VSB_CUDA(cudaMemcpyAsync(s->gray_dev[slot], gray_host, s->pitch * (size_t)(s->H - 1) + (size_t)s->W, cudaMemcpyHostToDevice, s->s_h2d));
    else VSB_CUDA(cudaMemcpy2DAsync(s->gray_dev[slot], s->pitch, gray_host, in_pitch, (size_t)s->W, (size_t)s->H,
                                    cudaMemcpyHostToDevice, s->s_h2d));
    VSB_CUDA(cudaEventRecord(s->ev_h2d[slot], s->s_h2d));
    VSB_CUDA(cudaStreamWaitEvent(s->s_comp, s->ev_h2d[slot], 0));
    int rc = run_layer(s, s->gray_dev[slot], s->pitch, s->out_dev[slot], s->opitch, emit);
    if (rc) return rc;
    VSB_CUDA(cudaEventRecord(s->ev_comp[slot], s->s_comp));
    VSB_CUDA(cudaStreamWaitEvent(s->s_d2h, s->ev_comp[slot], 0));
    if (emit) {
        if (out_pitch == s->opitch) VSB_CUDA(cudaMemcpyAsync(out_host, s->out_dev[slot], s->opitch * (size_t)(s->oH - 1) + (size_t)s->oW, cudaMemcpyDeviceToHost, s->s_d2h));
        else VSB_CUDA(cudaMemcpy2DAsync(out_host, out_pitch, s->out_dev[slot], s->opitch, (size_t)s->oW, (size_t)s->oH,
                                        cudaMemcpyDeviceToHost, s->s_d2h));
    }
    VSB_CUDA(cudaEventRecord(s->ev_d2h[slot], s->s_d2h));
    s->slot_used[slot] = true;
    const long long ticket = s->submitted++;
    return (int)(ticket & 0x3fffffff);
}

extern "C" int vsb_stack_submit_host(vsb_stack *s, const uint8_t *gray_host, size_t in_pitch, uint8_t *out_host,
                                     size_t out_pitch)
{
    VSB_REQUIRE(s && gray_host && out_host, "vsb_stack_submit_host: null pointer");
    VSB_REQUIRE(in_pitch >= (size_t)s->W && out_pitch >= (size_t)s->oW, "vsb_stack_submit_host: pitch < width");
    return submit(s, gray_host, in_pitch, out_host, out_pitch, true);
}

extern "C" int vsb_stack_wait(vsb_stack *s, int ticket)
{
    VSB_REQUIRE(s && ticket >= 0, "vsb_stack_wait: bad arguments");
    VSB_CUDA(cudaEventSynchronize(s->ev_d2h[ticket % STACK_SLOTS]));
    return VSB_OK;
}

extern "C" int vsb_stack_push_halo(vsb_stack *s, const uint8_t *gray_host, size_t pitch)
{
    VSB_REQUIRE(s && gray_host && pitch >= (size_t)s->W, "vsb_stack_push_halo: bad arguments");
    int t = submit(s, gray_host, pitch, nullptr, 0, false);
    if (t < 0) return t;
    return vsb_stack_wait(s, t);
}

extern "C" int vsb_stack_push_halo_device(vsb_stack *s, const uint8_t *gray_dev, size_t pitch)
{
    VSB_REQUIRE(s && gray_dev && pitch >= (size_t)s->W, "vsb_stack_push_halo_device: bad arguments");
    return run_layer(s, gray_dev, pitch, nullptr, 0, false);
}

extern "C" int vsb_stack_process_host(vsb_stack *s, const uint8_t *gray_host, size_t in_pitch, uint8_t *out_host,
                                      size_t out_pitch)
{
    int t = vsb_stack_submit_host(s, gray_host, in_pitch, out_host, out_pitch);
    if (t < 0) return t;
    return vsb_stack_wait(s, t);
}

// ---- per-stage profiling ---------------------------------------------------------------------------------
static void prof_clear(vsb_stack *s)
{
    for (auto &r : s->prof) { if (r.a) cudaEventDestroy(r.a); if (r.b) cudaEventDestroy(r.b); }
    s->prof.clear();
}

extern "C" int vsb_stack_profile_enable(vsb_stack *s, int on)
{
    VSB_REQUIRE(s, "vsb_stack_profile_enable: null stack");
    cudaStreamSynchronize(s->s_comp);
    prof_clear(s);
    s->prof_on = on != 0;
    return VSB_OK;
}

extern "C" int vsb_stack_profile_read(vsb_stack *s, int max_stages, float *ms_sum, int *launches, char *names /* max_stages x 32 */)
{
    VSB_REQUIRE(s && ms_sum && launches && names && max_stages > 0, "vsb_stack_profile_read: bad arguments");
    VSB_CUDA(cudaStreamSynchronize(s->s_comp));
    int nst = 0;
    for (int i = 0; i < max_stages; ++i) { ms_sum[i] = 0.0f; launches[i] = 0; names[i * 32] = 0; }
    for (auto &r : s->prof) {
        if (r.stage >= max_stages) continue;
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) continue;
        ms_sum[r.stage] += ms;
        launches[r.stage] += r.launches;
        if (r.stage + 1 > nst) nst = r.stage + 1;
    }
    for (int i = 0; i < nst; ++i) {
        const char *nm = i < ST_OP0 ? kStageNames[i]
                                    : ((size_t)(i - ST_OP0) < s->ops.size() ? kOpNames[s->ops[i - ST_OP0].type] : "?");
        snprintf(names + i * 32, 32, "%s", nm);
    }
    prof_clear(s);
    return nst;
}
