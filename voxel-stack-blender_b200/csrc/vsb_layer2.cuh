// vsb_layer2.cuh -- k_layer2: the fused per-layer kernel for STRICTLY BINARY regions (every input byte 0 or 255 -- what
// the reference asks of its input, README.md:25,68-70, and what k_pack_flags certifies per tile with FLAG_BINARY).
// Included by vsb_layer.cu; same arithmetic, bit for bit, as k_layer.
//
// In a binary region the stage-1 image (merge + leading LUT) is a function of the packed masks alone: LUT1[0] where the
// current mask is black, LUT1[255] where it is white, LUT1[gradient] on the receding pixels.  So this kernel never reads
// the gray layer.  Per 128x32 tile:
//   1. tile flags of the 3x3 neighbourhood: a non-binary neighbour hands the tile to the general kernel (gen_list);
//      nine equal uniform flags with nothing that can recede give the constant output at once;
//   2. the region's mask words (tile + XY halo, 6 words per row: current, and OR(priors) & ~current);
//   3. bit-plane morphology decides, per pixel, which filters can be skipped exactly:
//        - bilateral: a window without a receding pixel holds only the two values LUT1[0] / LUT1[255]; with at least three
//          taps equal to the centre (centre + two of its 4-neighbours, one LOP3 network) the host-checked bound says the
//          rounded result IS the centre (two_valued); everything else goes on the pixel list;
//        - Gaussian: words whose (bilateral + Gaussian)-radius window is one constant keep their value;
//   4. bits -> bytes for both stage buffers (4 pixels per PRMT-free multiply/select), receding pixels: row search, fade,
//      LUT1; then the listed bilateral pixels, the listed Gaussian words, trailing LUT, 128-bit stores.
#pragma once

#define L2W 6                        // mask words per region row: x0-32 .. x0+159
#define L2_PLANES 7                  // and_b, or_b, rec_b, and_h, or_h, rec_h, safe
#define L2_DENSE_MIN 2560            // receding pixels from which a tile takes the row-distance table
#define L2_HCOLS 142                 // table columns: region columns L2_HC0 .. L2_HC0 + 141 (receding pixels lie in 9 .. 150)
#define L2_HC0 9
#define EDT_FAR (1 << 20)
#define L2_HS 164                    // bytes per table column (the listed rows side by side; 41 words: conflict-free across columns)
#define L2_CLIFF_TILES 2048          // dense tiles in a layer from which the lower threshold applies
#define L2_DENSE_MIN_CLIFF 1024      // ... that threshold
#define L2_DV 10                     // dense tiles: region rows per lane in the table search
#define L2_DENSE_ROWS 164            // 142 x 164 bytes fit the space of the three lists (4480 + 1024 + 6656 entries)

// bits [lo, hi) of the 192-bit region span that fall into word k
__device__ __forceinline__ uint32_t span_mask(int k, int lo, int hi)
{
    const int a = max(lo - 32 * k, 0), b = min(hi - 32 * k, 32);
    if (a >= b) return 0u;
    const uint32_t upto_b = b >= 32 ? 0xFFFFFFFFu : ((1u << b) - 1u);
    return upto_b & ~((1u << a) - 1u);
}

// horizontal OR / AND of a bit row over |d| <= k (wl, w, wr = left neighbour word, the word, right neighbour word)
template <bool IS_AND>
__device__ __forceinline__ uint32_t hwin(uint32_t wl, uint32_t w, uint32_t wr, int k)
{
    uint32_t acc = w;
    for (int d = 1; d <= k; ++d) {
        const uint32_t a = __funnelshift_r(w, wr, d), b = __funnelshift_l(wl, w, d);
        acc = IS_AND ? (acc & a & b) : (acc | a | b);
    }
    return acc;
}

// 4 mask bits -> 4 bytes, K1 where the bit is set, K0 elsewhere (K0 / K1 replicated over the word)
__device__ __forceinline__ uint32_t expand4(uint32_t nib, uint32_t K0, uint32_t K1)
{
    const uint32_t m = ((nib * 0x00204081u) & 0x01010101u) * 0xFFu;
    return (m & K1) | (~m & K0);
}

// one tile.  DENSE: the instantiation of the second launch, which holds the row-distance-table path for the tiles the first one
// sets aside (interior tiles with thousands of receding pixels) -- kept out of the first launch's code
// SHAPE 1: the XY chain of Preset-Double-LUT (bilateral d = 7: radius 3, 29 taps; Gaussian 3 x 3) as compile-time constants, so the
// window loops unroll; SHAPE 2: the d = 7 bilateral alone (BASELINE config 3); SHAPE 0: any fused chain, from the parameters
template <int ZMODE, int METRIC, bool DENSE, int SHAPE>
__device__ __forceinline__ void layer2_tile(const LayerParams &p, const BilTapsL *__restrict__ taps, const int tx, const int ty)
{
    extern __shared__ __align__(16) uint8_t dsm[];
    float *s_T = (float *)(dsm + p.sm.T);
    uint8_t (*s1)[RG_W] = (uint8_t (*)[RG_W])(dsm + p.sm.s1);
    uint8_t (*s2buf)[RG_W] = (uint8_t (*)[RG_W])(dsm + p.sm.s2);
    uint32_t (*s_bits)[LBW] = (uint32_t (*)[LBW])(dsm + p.sm.bits);
    // one arena for the three lists: receding pixels from its start (at most p.rcap), the bilateral pixel list right behind
    // them (its room is what the receding pixels leave), the 1024 Gaussian words at its end (offset p.sm.glist)
    uint16_t *s_rlist = (uint16_t *)(dsm + p.sm.list);
    uint16_t *s_glist = s_rlist + p.sm.glist;
    uint8_t *s_h = (uint8_t *)s_rlist;                   // dense tiles: row-distance table [column][L2_HS], in the space of the lists
    __shared__ uint32_t s_c[RG_ROWS_MAX][8], s_r[RG_ROWS_MAX][8];      // region words: current mask, receding mask
    __shared__ uint32_t s_pl[L2_PLANES][RG_ROWS_MAX][8];
    __shared__ uint32_t s_rowany[(LB_ROWS_MAX + 31) / 32];
    __shared__ float s_cw2[512];
    __shared__ __align__(4) uint8_t s_lut1[256];
    __shared__ __align__(4) uint8_t s_lut2[256];
    __shared__ int s_cnt, s_gcnt, s_rcnt, s_nact;
    __shared__ int s_aside;
    __shared__ __align__(16) uint32_t s_wcode[VSB_THREADS / 32]; // per warp: receding / white / black seen in the region
    __shared__ __align__(4) uint8_t s_act[LB_ROWS_MAX + 4], s_lo[RG_ROWS_MAX], s_hi[RG_ROWS_MAX], s_mid[RG_ROWS_MAX]; // dense tiles: non-empty window rows, and per region row the ones in reach

    const int t = threadIdx.x, lane = t & 31;
    const int W = p.W, H = p.H, wpr = p.wpr;
    const int np = p.z.n_priors;
    const int r = t >> 3, c = t & 7;
    const int Hh = SHAPE == 1 ? 4 : (SHAPE == 2 ? 3 : p.halo), hb = SHAPE != 0 ? 3 : p.hb;
    const int rx = SHAPE == 1 ? 1 : (SHAPE == 2 ? 0 : p.nkx >> 1), ry = SHAPE == 1 ? 1 : (SHAPE == 2 ? 0 : p.nky >> 1), hg = max(rx, ry);
    const bool gauss = SHAPE == 1 ? true : (SHAPE == 2 ? false : p.nkx > 0);
    const int RH = VSB_TH + 2 * Hh;
    const int R = p.z.reach;
    const int TS = DENSE ? ((R + 1) | 1) : R + 1; // floats per row of the staged chamfer table
    const int x0 = tx * VSB_TW, y0 = ty * VSB_TH;
    const int oy = y0 + r, ox = x0 + c * 16;
    const int nvalid = (oy < H) ? max(0, min(16, W - ox)) : 0;

    // ---- 1. the tile's class (k_layer2_classify, from the tile flags of the 3x3 neighbourhood): constant tiles are written and
    //         left at once -- no table staging, no barrier -- and non-binary ones were listed for the general kernel
    const uint32_t cls = p.tile_class[ty * p.tiles_x + tx]; // the same word for every thread of the CTA
    if ((cls & 3u) == 1u) return;
    if ((cls & 3u) == 3u) { // listed for the dense launch by the pre-pass
        if (!DENSE) return;
    } else if (cls & 2u) {
        if (nvalid > 0) {
            const uint32_t cv = (cls >> 8) * 0x01010101u;
            uint8_t *dst = p.out + (size_t)oy * p.opitch + ox;
            const uint4 o = make_uint4(cv, cv, cv, cv);
            if (nvalid == 16 && p.aligned) vsb_st16_stream(dst, o);
            else vsb_st_bytes(dst, o, nvalid);
        }
        return;
    }
    // (no barrier here: the two LUTs, the counters and the region words all become visible at the barrier that ends step 2)
    if (t == 0) { s_cnt = 0; s_gcnt = 0; s_rcnt = 0; }
    if (t < 64 && p.lut1) ((uint32_t *)s_lut1)[t] = ((const uint32_t *)p.lut1)[t];
    if (t >= 64 && t < 128 && p.lut2) ((uint32_t *)s_lut2)[t - 64] = ((const uint32_t *)p.lut2)[t - 64];
    uint32_t k0b = 0u, k1b = 255u, K0 = 0u, K1 = 0xFFFFFFFFu; // LUT1[0], LUT1[255] (set after the barrier)
    auto constant_out = [&](uint32_t s1val) {
        if (nvalid > 0) {
            const uint32_t cv = (p.lut2 ? (uint32_t)s_lut2[s1val] : s1val) * 0x01010101u;
            uint8_t *dst = p.out + (size_t)oy * p.opitch + ox;
            const uint4 o = make_uint4(cv, cv, cv, cv);
            if (nvalid == 16 && p.aligned) vsb_st16_stream(dst, o);
            else vsb_st_bytes(dst, o, nvalid);
        }
    };

    // ---- 2. region mask words --------------------------------------------------------------------------------------------
    const int wx0 = (x0 >> 5) - 1;               // word column of region word 0
    const int img_lo = 32 - x0, img_hi = W - x0 + 32; // image columns in span coordinates
    const bool border = (x0 - Hh < 0) || (x0 + VSB_TW + Hh > W) || (y0 - Hh < 0) || (y0 + VSB_TH + Hh > H);
    const bool searching = ZL_SEARCHES(ZMODE);
    uint32_t code = 0, nrec = 0;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int i = t + u * VSB_THREADS;
        if (i < RH * L2W) {
            const int rr = i / L2W, k = i - rr * L2W;
            int y = y0 - Hh + rr;
            uint32_t rm; // the region's columns in this word
            bool word_in = true;
            if (!border) { // interior tile: the region is span bits 32 - Hh .. 159 + Hh, every row and word exists
                rm = k == 0 ? (Hh ? 0xFFFFFFFFu << (32 - Hh) : 0u) : (k == L2W - 1 ? (1u << Hh) - 1u : 0xFFFFFFFFu);
            } else {
                if (y < 0 || y >= H) y = vsb_reflect101(y, H); // BORDER_REFLECT_101 rows: the mirrored row's bits
                word_in = wx0 + k >= 0 && wx0 + k < wpr;
                rm = span_mask(k, max(32 - Hh, img_lo), min(160 + Hh, img_hi));
            }
            uint32_t cw = 0, rw = 0;
            if (word_in) {
                const size_t o = (size_t)y * wpr + (wx0 + k);
                cw = p.cur[o];
                uint32_t acc = 0;
                for (int j = 0; j < np; j += 4) { // four independent requests per trip
                    const uint32_t a0 = p.prior[j][o], a1 = j + 1 < np ? p.prior[j + 1][o] : 0u;
                    const uint32_t a2 = j + 2 < np ? p.prior[j + 2][o] : 0u, a3 = j + 3 < np ? p.prior[j + 3][o] : 0u;
                    acc |= a0 | a1 | a2 | a3;
                }
                rw = acc & ~cw & rm;
            }
            s_c[rr][k] = cw;
            s_r[rr][k] = rw;
            code |= (rw ? 1u : 0u) | ((cw & rm) ? 2u : 0u) | ((~cw & rm) ? 4u : 0u);
            if (!DENSE && searching) nrec += __popc(rw);
        }
    }
    code = __reduce_or_sync(0xffffffffu, code);
    if (!DENSE && searching) code |= __reduce_add_sync(0xffffffffu, nrec) << 3; // + the warp's receding pixels (at most 2048)
    if (lane == 0) s_wcode[t >> 5] = code;
    if (hb > 0) {
        const float cwt = taps->cw[t];
        s_cw2[255 + t] = cwt; s_cw2[255 - t] = cwt;
    }
    __syncthreads();
    if (p.lut1) { k0b = s_lut1[0]; k1b = s_lut1[255]; }
    K0 = k0b * 0x01010101u; K1 = k1b * 0x01010101u;
    const uint4 wc0 = *(const uint4 *)&s_wcode[0], wc1 = *(const uint4 *)&s_wcode[4];
    const int state = (int)((wc0.x | wc0.y | wc0.z | wc0.w | wc1.x | wc1.y | wc1.z | wc1.w) & 7u);
    const bool any_rec = (state & 1) != 0;
    if (!any_rec && (state & 6) != 6) { constant_out((state & 2) ? k1b : k0b); return; } // one value, nothing recedes
    // dense tiles (thousands of receding pixels: whole tiles right after a raft ends) are set aside for the second launch
    // before anything is staged for them
    if (!DENSE && searching && !border) {
        const int nrec_region = (int)((wc0.x >> 3) + (wc0.y >> 3) + (wc0.z >> 3) + (wc0.w >> 3) + (wc1.x >> 3) + (wc1.y >> 3) + (wc1.z >> 3) + (wc1.w >> 3));
        bool aside = nrec_region >= p.dense_from;
        // half-full tiles: only in a layer that sends thousands of tiles to the dense launch anyway (the pre-pass's count plus the
        // ones set aside so far -- right after a raft ended), where that launch takes them better than the lists do
        if (!aside && nrec_region >= L2_DENSE_MIN_CLIFF) { // (the count moves while the kernel runs: one thread reads it for the CTA)
            if (t == 0) s_aside = __ldcg(&p.dense_list[0]) >= L2_CLIFF_TILES;
            __syncthreads();
            aside = s_aside != 0;
        }
        if (aside) {
            if (t == 0) p.dense_list[1 + atomicAdd(&p.dense_list[0], 1)] = ty * p.tiles_x + tx;
            return;
        }
    }
    if (any_rec && searching && METRIC == VSB_METRIC_CHAMFER5) { // the chamfer table: only tiles with receding pixels read it
        const int nT = (R + 1) * (R + 1);                       // (first use is two barriers away)
        if (DENSE) { // rows an odd number of floats apart: the lanes of the table search read T[dx][k] with one k and 32 dx
            for (int i = t; i < nT; i += VSB_THREADS) { const int r = i / (R + 1); s_T[r * TS + (i - r * (R + 1))] = p.T[i]; }
        } else {
            const int nfull = ((((size_t)p.T) & 15) == 0) ? (nT >> 2) : 0;
            for (int i = t; i < nfull; i += VSB_THREADS) ((float4 *)s_T)[i] = ((const float4 *)p.T)[i];
            for (int i = 4 * nfull + t; i < nT; i += VSB_THREADS) s_T[i] = p.T[i];
        }
    }

    // out-of-image columns (left / right border tiles): mirrored bits, one thread per region row
    if (border && t < RH && ((x0 - Hh < 0) || (x0 + VSB_TW + Hh > W))) {
        auto getb = [&](const uint32_t *row, int sx) -> uint32_t { return (row[sx >> 5] >> (sx & 31)) & 1u; };
        auto setb = [&](uint32_t *row, int sx, uint32_t v) { row[sx >> 5] = (row[sx >> 5] & ~(1u << (sx & 31))) | (v << (sx & 31)); };
        for (int j = 1; j <= Hh; ++j) {
            if (x0 - Hh < 0) { // x = -j  <-  x = +j
                const int sd = img_lo - j, ss = img_lo + j;
                if (sd >= 0 && ss < 192) { setb(s_c[t], sd, getb(s_c[t], ss)); setb(s_r[t], sd, getb(s_r[t], ss)); }
            }
            if (x0 + VSB_TW + Hh > W) { // x = W-1+j  <-  x = W-1-j
                const int sd = img_hi - 1 + j, ss = img_hi - 1 - j;
                if (sd < 192 && ss >= 0) { setb(s_c[t], sd, getb(s_c[t], ss)); setb(s_r[t], sd, getb(s_r[t], ss)); }
            }
        }
    }
    if (border) __syncthreads();

    // ---- 3a. horizontal pass of the morphology + bits -> bytes + receding pixel list + search window -----------------------
    if (any_rec && searching) stage_mask_rows<LBW>(s_bits, p.cur, wpr, H, (x0 >> 5) - 3, y0 - Hh - R, RH + 2 * R);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int i = t + u * VSB_THREADS;
        if (i < RH * L2W) {
            const int rr = i / L2W, k = i - rr * L2W;
            const uint32_t cl = k > 0 ? s_c[rr][k - 1] : 0u, cc = s_c[rr][k], cr = k < L2W - 1 ? s_c[rr][k + 1] : 0u;
            const uint32_t rl = k > 0 ? s_r[rr][k - 1] : 0u, rc = s_r[rr][k], rq = k < L2W - 1 ? s_r[rr][k + 1] : 0u;
            s_pl[0][rr][k] = hwin<true>(cl, cc, cr, hb);
            s_pl[1][rr][k] = hwin<false>(cl, cc, cr, hb);
            s_pl[2][rr][k] = hwin<false>(rl, rc, rq, hb);
            s_pl[3][rr][k] = hwin<true>(cl, cc, cr, hb + rx);
            s_pl[4][rr][k] = hwin<false>(cl, cc, cr, hb + rx);
            s_pl[5][rr][k] = hwin<false>(rl, rc, rq, hb + rx);
            uint32_t safe = 0u;
            if (p.two_valued && rr > 0 && rr < RH - 1) { // centre + at least two of its 4-neighbours hold the same value
                const uint32_t sl = ~(cc ^ __funnelshift_l(cl, cc, 1)), sr = ~(cc ^ __funnelshift_r(cc, cr, 1));
                const uint32_t su = ~(cc ^ s_c[rr - 1][k]), sd = ~(cc ^ s_c[rr + 1][k]);
                safe = (sl & sr) | (su & sd) | ((sl | sr) & (su | sd));
            }
            s_pl[6][rr][k] = safe;
            if (rc) { // receding pixels of this word -> list (a tile holding more than the list does is handed on below)
                uint32_t rw = rc;
                int base = atomicAdd(&s_rcnt, __popc(rw));
                if (base + __popc(rw) > p.rcap) rw = 0u;
                while (rw) {
                    const int bq = __ffs(rw) - 1;
                    rw &= rw - 1;
                    s_rlist[base++] = (uint16_t)((rr << 8) | (32 * k + bq - 16));
                }
            }
        }
        if (i < RH * RG_CH) { // bytes of chunk (rr, cc): span bits 16*cc+16 .. +15
            const int rr = i / RG_CH, cc = i - rr * RG_CH;
            const uint32_t hw = (s_c[rr][(cc + 1) >> 1] >> (((cc + 1) & 1) * 16)) & 0xFFFFu;
            const uint4 v = make_uint4(expand4(hw & 15u, K0, K1), expand4((hw >> 4) & 15u, K0, K1),
                                       expand4((hw >> 8) & 15u, K0, K1), expand4(hw >> 12, K0, K1));
            *(uint4 *)&s1[rr][cc * 16] = v;
            *(uint4 *)&s2buf[rr][cc * 16] = v;
        }
    }
    __syncthreads();

    const int n_listed_rec = any_rec ? s_rcnt : 0;
    uint16_t *s_list = s_rlist + n_listed_rec;           // bilateral pixel list
    const int bcap = p.sm.glist - n_listed_rec;

    // ---- dense tiles (hundreds of receding pixels: wide rims, and whole tiles right after a raft ends): the per-pixel row walk
    //      probes the same window rows over and over, and most of them hold no white pixel.  List the non-empty window rows, take
    //      the nearest-white distance of every (column, listed row) once -- one byte each, rows of a column side by side -- and
    //      let a lane per column read them four rows at a time for a block of L2_DV pixels (step 4).  The table borrows the space
    //      of the three lists (the receding pixels are then taken from the mask words), so this runs BEFORE the filter lists are
    //      built.  Only the second launch (DENSE) holds this path; the first one set such tiles aside after its first barrier.
    bool use_h = false;
    if (any_rec && s_rcnt > p.rcap) { // more receding pixels than the list holds (border tiles, modes without search): general kernel
        if (t == 0) p.gen_list[1 + atomicAdd(&p.gen_list[0], 1)] = ty * p.tiles_x + tx;
        return;
    }
    if (DENSE && any_rec && searching && s_rcnt >= min(p.dense_from, L2_DENSE_MIN_CLIFF)) {
        const int brows = RH + 2 * R;
        stage_rowany_rows<LBW>(s_rowany, s_bits, brows);
        __syncthreads();
        if (t < 32) {
            int n = 0;
            for (int b0 = 0; b0 < brows; b0 += 32) {
                const uint32_t w = s_rowany[b0 >> 5];
                if ((w >> lane) & 1u) s_act[n + __popc(w & ((1u << lane) - 1u))] = (uint8_t)(b0 + lane);
                n += __popc(w);
            }
            if (lane < 4) s_act[min(n + lane, LB_ROWS_MAX + 3)] = 255; // padding of the last group of four: never in reach
            if (lane == 0) s_nact = n;
        }
        __syncthreads();
        const int nact = s_nact;
        use_h = nact <= L2_DENSE_ROWS && border == false;
        if (use_h) {
            if (t < RH) { // the listed rows within reach of region row t
                const int hr0 = t + R;
                int lo = 0;
                while (lo < nact && (int)s_act[lo] < hr0 - R) ++lo;
                int hi = lo;
                while (hi < nact && (int)s_act[hi] <= hr0 + R) ++hi;
                int mid = lo;
                while (mid < hi && (int)s_act[mid] <= hr0) ++mid;
                s_lo[t] = (uint8_t)lo; s_hi[t] = (uint8_t)hi; s_mid[t] = (uint8_t)mid; // [lo, mid): rows at or above, [mid, hi): below
            }
            const int nact4 = (nact + 3) & ~3;
            for (int i = t; i < nact4 * L2_HCOLS; i += VSB_THREADS) {
                const int a = i / L2_HCOLS, c = i - a * L2_HCOLS;
                s_h[c * L2_HS + a] = a < nact ? (uint8_t)nearest_dx(s_bits[s_act[a]], c + L2_HC0 + 80) : (uint8_t)255;
            }
            __syncthreads();
        }
    }
    // ---- 4. receding pixels: distance, fade, leading LUT -> both stage buffers ------------------------------------------------
    // (after 3b for ordinary tiles, from the list; before it for dense tiles, from the mask words and the table)
    const float Ff = p.z.fade;
    auto fade_pixel = [&](const int rr, const int rc, const float best) -> uint8_t {
        const int x = x0 - 16 + rc, y = y0 - Hh + rr;
        int g;
        if (ZMODE == VSB_Z_FIXED) {
            const float n = __fdiv_rn(fminf(best, Ff), p.z.den);
            g = (int)__fmul_rn(255.0f, __fsub_rn(1.0f, n));
        } else if (ZMODE == VSB_Z_WEIGHTED) {
            float acc = 0.0f;
            const size_t o = (size_t)y * wpr + (x >> 5);
            const uint32_t ncur = ~p.cur[o];
            for (int j = 0; j < np; ++j) {
                if (p.z.weights[j] > 0.0f && (((p.prior[j][o] & ncur) >> (x & 31)) & 1u)) {
                    const float n = __fsub_rn(1.0f, __fdiv_rn(fminf(best, p.z.fades[j]), p.z.dens[j]));
                    acc = __fadd_rn(acc, __fmul_rn(n, p.z.weights[j]));
                }
            }
            g = (int)__fmul_rn(255.0f, __fdiv_rn(acc, p.z.total_weight));
        } else if (ZMODE == VSB_Z_ENHANCED) {
            const float d = p.enh_dist[(size_t)y * p.enh_dpitch + x];
            const float mx = __uint_as_float(p.enh_cmax[p.enh_labels[(size_t)y * W + x]]);
            if (p.z.enhanced_arith == 0) {
                float den = (float)fmin((double)mx, p.z.fade_limit);
                if (den <= 0.0f) den = 1.0f;
                const float nrm = __fdiv_rn(fminf(fmaxf(d, 0.0f), den), den);
                g = (int)__fmul_rn(255.0f, __fsub_rn(1.0f, nrm));
            } else {
                const double den = fmin((double)mx, p.z.fade_limit);
                g = 0;
                if (den > 0.0) g = (int)(long long)__dmul_rn(__dsub_rn(1.0, __ddiv_rn(fmin((double)d, den), den)), 255.0);
            }
        } else {
            g = p.z.const_val;
        }
        const int m = g & 255; // merge with the raw gray value: a receding pixel is black, its gray value is 0
        const uint8_t v = p.lut1 ? s_lut1[m] : (uint8_t)m;
        s1[rr][rc] = v;
        s2buf[rr][rc] = v;
        return v;
    };
    if (DENSE && use_h) { // dense, interior tile: a warp per mask word and block of L2_DV region rows, a lane per pixel COLUMN
        // A lane owns the L2_DV pixels of its column in the block.  They share the column's table bytes (the nearest white
        // pixel of a listed row lies equally far in x from all of them) and differ only in |dy|, which is warp-uniform: a
        // listed row costs one byte compare for the block and, where it is wanted, one table read per pixel.  Every metric
        // here grows with |dy| for a fixed |dx| and is >= max(|dx|, |dy|), so a row can only improve one of the lane's
        // pixels if its white pixel is closer in x than the largest of their best distances so far, and than that of every
        // listed row that is nearer to ALL of them: the nearer rows on the same side of the block, and -- from L2_DV rows away
        // -- the rows inside the block.  Order: the rows inside the block, then upwards, then downwards, in groups of four
        // (one 32-bit read of the table each); a direction ends once no lane can do better than the group's smallest |dy|.
        const int nblk = (RH + L2_DV - 1) / L2_DV;
        for (int item = t >> 5; item < nblk * L2W; item += VSB_THREADS / 32) {
            const int blk = item / L2W, k = item - blk * L2W;
            const int rr0 = blk * L2_DV, nv = min(L2_DV, RH - rr0);
            uint32_t mine = 0u, anyrow = 0u; // bit j: pixel (rr0 + j, this lane's column) is receding
#pragma unroll
            for (int j = 0; j < L2_DV; ++j) {
                if (j < nv) {
                    const uint32_t rw = s_r[rr0 + j][k];
                    anyrow |= rw;
                    mine |= ((rw >> lane) & 1u) << j;
                }
            }
            if (anyrow == 0u) continue; // warp-uniform
            const int rc = 32 * k + lane - 16;
            const int hr0 = rr0 + R, hr1 = hr0 + nv - 1; // window rows of the block's first and last pixel row
            const uint8_t *hp = s_h + min(max(rc - L2_HC0, 0), L2_HCOLS - 1) * L2_HS;
            float best[L2_DV]; // chamfer: the distance; exact metric: its square (integers below 2^24, exact in float)
#pragma unroll
            for (int j = 0; j < L2_DV; ++j) best[j] = (mine >> j) & 1u ? 3.0e38f : 0.0f;
            const int lo = s_lo[rr0], hi = s_hi[rr0 + nv - 1];
            int i0 = s_mid[rr0];
            if (i0 > lo && (int)s_act[i0 - 1] == hr0) --i0; // [i0, i1): the listed rows inside the block
            const int i1 = max((int)s_mid[rr0 + nv - 1], i0);
            int dx_in = R + 1, dx_side = R + 1; // smallest |dx| among the rows inside the block / on the current side of it
            int thr = mine ? R : -1;            // dx <= thr can still improve one of this lane's pixels; < 0: nothing can
            uint32_t thrb = (uint32_t)R * 0x01010101u | 0x80808080u;
            auto bound = [&](const int dxdom) {
                float bm = best[0];
#pragma unroll
                for (int j = 1; j < L2_DV; ++j) bm = fmaxf(bm, best[j]);
                const int reach = METRIC == VSB_METRIC_CHAMFER5 ? (int)bm : (bm >= 1.0e30f ? R : (int)__fsqrt_ru(bm));
                thr = mine ? min(min(R, reach), dxdom - 1) : -1;
                thrb = (uint32_t)max(thr, 0) * 0x01010101u | 0x80808080u;
                return bm;
            };
            // rows g4 .. g4 + 3 of the list, those with index in [ia, ib); a4: their window rows.  Returns the smallest |dx| taken.
            auto group = [&](const int g4, const int ia, const int ib) {
                // table bytes are 0..64, or 255 for padding: per byte, bit 7 of (thr | 0x80) - (dx & 0x7f) is set iff dx <= thr
                const uint32_t w4 = *(const uint32_t *)(hp + g4);
                const uint32_t hit = thr >= 0 ? ((thrb - (w4 & 0x7f7f7f7fu)) & 0x80808080u) : 0u;
                const uint32_t any = __reduce_or_sync(0xffffffffu, hit);
                int dmin = R + 1;
                if (any == 0u) return dmin;
                const uint32_t a4 = *(const uint32_t *)(s_act + g4);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (!(any & (0x80u << (8 * q))) || g4 + q < ia || g4 + q >= ib) continue; // warp-uniform
                    const int d0 = (int)((a4 >> (8 * q)) & 255u) - hr0;                        // row - first pixel row
                    if (hit & (0x80u << (8 * q))) {
                        const int dx = (w4 >> (8 * q)) & 255;
                        dmin = min(dmin, dx);
                        if (METRIC == VSB_METRIC_CHAMFER5) {
                            const float *Tr = s_T + dx * TS; // the table is symmetric: T[dx][|dy|]
#pragma unroll
                            for (int j = 0; j < L2_DV; ++j) {
                                const int kk = abs(d0 - j);
                                if (kk <= R) best[j] = fminf(best[j], Tr[kk]);
                            }
                        } else {
                            const float dx2 = (float)(dx * dx);
#pragma unroll
                            for (int j = 0; j < L2_DV; ++j) {
                                const int kk = abs(d0 - j);
                                if (kk <= R) best[j] = fminf(best[j], dx2 + (float)(kk * kk));
                            }
                        }
                    }
                }
                return dmin;
            };
            for (int g4 = i0 & ~3; g4 < i1; g4 += 4) dx_in = min(dx_in, group(g4, i0, i1)); // inside: every row, no order
            float bm = bound(R + 1);
            for (int g4 = (i0 - 1) & ~3; g4 >= (lo & ~3) && i0 > lo; g4 -= 4) { // upwards: the group's nearest row is its last
                const int kmin = hr0 - (int)s_act[min(g4 + 3, i0 - 1)];
                if (__all_sync(0xffffffffu, thr < 0 || (METRIC == VSB_METRIC_CHAMFER5 ? bm <= (float)kmin : bm <= (float)(kmin * kmin)))) break;
                dx_side = min(dx_side, group(g4, lo, i0));
                bm = bound(kmin >= L2_DV ? min(dx_side, dx_in) : dx_side); // (the NEXT group is at least as far as this one)
            }
            dx_side = R + 1;
            bm = bound(R + 1);
            for (int g4 = i1 & ~3; g4 < hi; g4 += 4) {                          // downwards: its first
                const int kmin = (int)s_act[max(g4, i1)] - hr1;
                if (__all_sync(0xffffffffu, thr < 0 || (METRIC == VSB_METRIC_CHAMFER5 ? bm <= (float)kmin : bm <= (float)(kmin * kmin)))) break;
                dx_side = min(dx_side, group(g4, i1, hi));
                bm = bound(kmin >= L2_DV ? min(dx_side, dx_in) : dx_side);
            }
            // a receding pixel whose value came out as black's (beyond the fade distance) IS a black pixel for the filters:
            // only the others keep their mark, so the bypass planes below see a plain two-valued region there
#pragma unroll
            for (int j = 0; j < L2_DV; ++j) {
                if (j < nv) {
                    bool graded = false;
                    if ((mine >> j) & 1u) {
                        const float d = METRIC == VSB_METRIC_CHAMFER5 ? best[j] : (best[j] >= 1.0e30f ? 3.0e38f : __fsqrt_rn(best[j]));
                        graded = fade_pixel(rr0 + j, rc, d) != (uint8_t)k0b;
                    }
                    const uint32_t nw = __ballot_sync(0xffffffffu, graded);
                    if (lane == 0) s_r[rr0 + j][k] = nw;
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 2; ++u) { // the receding planes of 3a, again, from the marks that are left
            const int i = t + u * VSB_THREADS;
            if (i < RH * L2W) {
                const int rr = i / L2W, k = i - rr * L2W;
                const uint32_t rl = k > 0 ? s_r[rr][k - 1] : 0u, rc = s_r[rr][k], rq = k < L2W - 1 ? s_r[rr][k + 1] : 0u;
                s_pl[2][rr][k] = hwin<false>(rl, rc, rq, hb);
                s_pl[5][rr][k] = hwin<false>(rl, rc, rq, hb + rx);
            }
        }
        __syncthreads();
    }
    auto receding_listed = [&]() {
        const int total = s_rcnt;
        for (int i = t; i < total; i += VSB_THREADS) {
            const int codev = s_rlist[i];
            const int rr = codev >> 8, rc = codev & 255;
            const int x = x0 - 16 + rc, y = y0 - Hh + rr;
            if (y < 0 || y >= H || x < 0 || x >= W) continue; // mirrored positions are filled in below
            float best = 0.0f;
            if (ZL_SEARCHES(ZMODE)) best = tile_search<METRIC, LBW, false>(s_bits, s_rowany, s_T, R, rr + R, rc + 80, TS);
            fade_pixel(rr, rc, best);
        }
        __syncthreads();
        // REFLECT_101 of the gradient values into the out-of-image halo (border tiles): mirror the stage-1 bytes
        if (border) {
            for (int i = t; i < total; i += VSB_THREADS) {
                const int codev = s_rlist[i];
                const int rr = codev >> 8, rc = codev & 255;
                const int x = x0 - 16 + rc, y = y0 - Hh + rr;
                if (y >= 0 && y < H && x >= 0 && x < W) continue;
                const int ys = vsb_reflect101(y, H), xs = vsb_reflect101(x, W);
                const uint8_t v = s1[ys - (y0 - Hh)][xs - (x0 - 16)];
                s1[rr][rc] = v;
                s2buf[rr][rc] = v;
            }
            __syncthreads();
        }
    };

    // ---- 3b. vertical pass: which pixels need the bilateral, which words the Gaussian ---------------------------------------
    if (any_rec && searching && !use_h) stage_rowany_rows<LBW>(s_rowany, s_bits, RH + 2 * R);
    {
        const int nb_lo = Hh - hg, nb_hi = Hh + VSB_TH + hg; // region rows on which stage 2 is needed
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = t + u * VSB_THREADS;
            if (i < RH * L2W) {
                const int rr = i / L2W, k = i - rr * L2W;
                const int y = y0 - Hh + rr;
                const bool row_in = y >= 0 && y < H;
                if (hb > 0 && rr >= nb_lo && rr < nb_hi && row_in) {
                    uint32_t A = 0xFFFFFFFFu, O = 0u, Rr = 0u;
                    for (int d = -hb; d <= hb; ++d) { A &= s_pl[0][rr + d][k]; O |= s_pl[1][rr + d][k]; Rr |= s_pl[2][rr + d][k]; }
                    uint32_t nb = Rr | ((O & ~A) & ~s_pl[6][rr][k]);
                    nb &= span_mask(k, max(32 - hg, img_lo), min(160 + hg, img_hi));
                    if (nb) {
                        int base = atomicAdd(&s_cnt, __popc(nb));
                        if (base + __popc(nb) > bcap) nb = 0u; // arena full: the tile goes to the general kernel below
                        while (nb) {
                            const int bq = __ffs(nb) - 1;
                            nb &= nb - 1;
                            s_list[base++] = (uint16_t)((rr << 8) | (32 * k + bq - 16));
                        }
                    }
                }
                if (gauss && rr >= Hh && rr < Hh + VSB_TH && row_in) {
                    uint32_t A = 0xFFFFFFFFu, O = 0u, Rr = 0u;
                    const int vr = hb + ry;
                    for (int d = -vr; d <= vr; ++d) { A &= s_pl[3][rr + d][k]; O |= s_pl[4][rr + d][k]; Rr |= s_pl[5][rr + d][k]; }
                    uint32_t ng = (Rr | (O & ~A)) & span_mask(k, max(32, img_lo), min(160, img_hi));
                    ng = (ng | (ng >> 1) | (ng >> 2) | (ng >> 3)) & 0x11111111u; // one bit per 4-pixel word
                    if (ng) {
                        int base = atomicAdd(&s_gcnt, __popc(ng));
                        while (ng) {
                            const int bq = __ffs(ng) - 1;
                            ng &= ng - 1;
                            s_glist[base++] = (uint16_t)((rr << 8) | ((32 * k + bq - 16) >> 2));
                        }
                    }
                }
            }
        }
    }
    __syncthreads();
    if (hb > 0 && s_cnt > bcap) { // more bilateral pixels than the arena has room for: the general kernel takes the tile
        if (t == 0) p.gen_list[1 + atomicAdd(&p.gen_list[0], 1)] = ty * p.tiles_x + tx;
        return;
    }

    if (any_rec && !use_h) receding_listed();

    uint8_t (*S2)[RG_W] = hb > 0 ? s2buf : s1;                  // Gaussian input
    uint8_t (*OUT)[RG_W] = gauss ? (hb > 0 ? s1 : s2buf) : S2;  // what the store phase reads

    // ---- 5. bilateral on the listed pixels ------------------------------------------------------------------------------------
    if (hb > 0) {
        const int total = s_cnt;
        if (total > 0) {
            const int ntaps = SHAPE != 0 ? 29 : p.ntaps;
            for (int i = t; i < total; i += VSB_THREADS) {
                const int codev = s_list[i];
                const int rr = codev >> 8, rc = codev & 255;
                const uint8_t *ctr = &s1[rr][rc];
                const int v0 = *ctr;
                uint32_t cwa = (uint32_t)__cvta_generic_to_shared(s_cw2) + 4u * (uint32_t)(255 - v0);
                asm volatile("" : "+r"(cwa)); // opaque: keeps the per-tap address at one IMAD
                float sacc = 0.0f, ws = 0.0f;
#pragma unroll 4
                for (int k = 0; k < ntaps; ++k) {
                    const int v = ctr[p.tap_off[k]];
                    const float w = __fmul_rn(p.tap_sw[k], lds_f32(cwa + 4u * (uint32_t)v));
                    sacc = __fmaf_rn((float)v, w, sacc);
                    ws = __fadd_rn(ws, w);
                }
                const int o = __float2int_rn(__fdiv_rn(sacc, ws));
                s2buf[rr][rc] = (uint8_t)min(max(o, 0), 255);
            }
        }
        __syncthreads();
        if (border && hg > 0) { // stage-2 values of the out-of-image halo: mirror
            const int r_lo = Hh - hg, nr = VSB_TH + 2 * hg;
            for (int i = t; i < nr * RG_W; i += VSB_THREADS) {
                const int rr = r_lo + i / RG_W, rc = i % RG_W;
                const int y = y0 - Hh + rr, x = x0 - 16 + rc;
                if (x < -hg || x >= W + hg || y < -hg || y >= H + hg) continue;
                if (y < 0 || y >= H || x < 0 || x >= W) {
                    const int ys = vsb_reflect101(y, H), xs = vsb_reflect101(x, W);
                    s2buf[rr][rc] = s2buf[ys - (y0 - Hh)][xs - (x0 - 16)];
                }
            }
            __syncthreads();
        }
    }

    // ---- 6. Gaussian on the listed words -----------------------------------------------------------------------------------------
    if (gauss) {
        const int gtotal = s_gcnt;
        for (int i = t; i < gtotal; i += VSB_THREADS) {
            const int codev = s_glist[i];
            const int rr = codev >> 8, w = codev & 255;
            uint32_t o;
            switch (rx) {
            case 0: o = gauss_word_ry<0>(S2, rr, w, p.kx, p.ky, ry); break;
            case 1: o = gauss_word_ry<1>(S2, rr, w, p.kx, p.ky, ry); break;
            case 2: o = gauss_word_ry<2>(S2, rr, w, p.kx, p.ky, ry); break;
            default: o = gauss_word_ry<3>(S2, rr, w, p.kx, p.ky, ry); break;
            }
            *(uint32_t *)&OUT[rr][4 * w] = o;
        }
        __syncthreads();
    }

    // ---- 7. trailing LUT, store ------------------------------------------------------------------------------------------------------
    if (nvalid > 0) {
        VsbLut L2;
        L2.s = p.lut2 ? s_lut2 : nullptr;
        L2.l0 = (uint32_t)s_lut2[0] * 0x01010101u; L2.l255 = (uint32_t)s_lut2[255] * 0x01010101u;
        uint4 o = *(const uint4 *)&OUT[r + Hh][16 + 16 * c];
        o = L2.map16(o);
        uint8_t *dst = p.out + (size_t)oy * p.opitch + ox;
        if (nvalid == 16 && p.aligned) vsb_st16_stream(dst, o);
        else vsb_st_bytes(dst, o, nvalid);
    }
}

// One thread per tile: the decision every CTA of k_layer2 used to derive from 9 + 9 N tile flags behind a barrier.
//   class & 3: 0 = work, 1 = a neighbour is not strictly binary -> listed for the general kernel, 2 = constant output;
//   class >> 8: that constant (LUT2[LUT1[0 or 255]]).  A tile is constant when its 3x3 neighbourhood is one uniform value and
//   nothing can recede there (the value is white, or every prior's neighbourhood is uniformly black).
__global__ void __launch_bounds__(256) k_layer2_classify(const __grid_constant__ LayerParams p)
{
    const int tile = blockIdx.x * 256 + threadIdx.x;
    if (tile >= p.ntiles) return;
    const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
    const int np = p.z.n_priors;
    // 9 + 9 N independent requests: the flags are combined with bitwise operators so that no load waits for a branch
    int fi[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        const int ny = min(max(ty + q / 3 - 1, 0), p.tiles_y - 1), nx = min(max(tx + q % 3 - 1, 0), p.tiles_x - 1);
        fi[q] = ny * p.tiles_x + nx;
    }
    const uint32_t fc = p.cur_flags[tile];
    uint32_t bin_and = FLAG_BINARY, diff_or = 0u, pri_or = 0u;
    bool own_white = false; // some prior layer's tile is uniformly white
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        const uint32_t f = p.cur_flags[fi[q]];
        bin_and &= f;
        diff_or |= (f ^ fc) | (f & FLAG_MIXED);
    }
    for (int j = 0; j < np; ++j) {
        const uint16_t *pf = p.prior_flags[j];
        if (!pf) { pri_or |= FLAG_MIXED; continue; }
#pragma unroll
        for (int q = 0; q < 9; ++q) {
            const uint32_t g = pf[fi[q]];
            pri_or |= (g & FLAG_MIXED) | ((g & 0xFFu) > 127u ? 1u : 0u); // mixed, or a uniform white tile
            if (q == 4) own_white |= !(g & FLAG_MIXED) && (g & 0xFFu) > 127u;
        }
    }
    const bool all_bin = (bin_and & FLAG_BINARY) != 0u, same = diff_or == 0u, pri_black = pri_or == 0u;
    uint32_t cls = 0u;
    if (!all_bin) {
        cls = 1u;
        p.gen_list[1 + atomicAdd(&p.gen_list[0], 1)] = tile;
    } else if (same && ((fc & 0xFFu) > 127u || pri_black || np == 0)) {
        const uint32_t s1val = (fc & 0xFFu) > 127u ? (p.lut1 ? p.lut1[255] : 255u) : (p.lut1 ? p.lut1[0] : 0u);
        cls = 2u | ((p.lut2 ? (uint32_t)p.lut2[s1val] : s1val) << 8);
    } else if (p.route_dense && own_white && !(fc & FLAG_MIXED) && (fc & 0xFFu) <= 127u && tx > 0 && tx < p.tiles_x - 1 &&
               ty > 0 && ty < p.tiles_y - 1 && (tx + 2) * VSB_TW <= p.W && (ty + 2) * VSB_TH <= p.H) {
        // a black tile over a white one, away from the image border: every pixel recedes (right after a raft ends, most of
        // the plate) -- straight to the dense launch, k_layer2 does not stage it first to find that out
        cls = 3u;
        p.dense_list[1 + atomicAdd(&p.dense_list[0], 1)] = tile;
    }
    p.tile_class[tile] = (uint16_t)cls;
}

template <int ZMODE, int METRIC, int SHAPE>
__global__ void __launch_bounds__(VSB_THREADS, 4) k_layer2(const __grid_constant__ LayerParams p, const BilTapsL *__restrict__ taps)
{
    layer2_tile<ZMODE, METRIC, false, SHAPE>(p, taps, (int)blockIdx.x, (int)blockIdx.y);
}

// the tiles set aside by k_layer2: a few CTAs per SM stride over the list (empty on ordinary layers)
template <int ZMODE, int METRIC, int SHAPE>
__global__ void __launch_bounds__(VSB_THREADS, 3) k_layer2_dense(const __grid_constant__ LayerParams p, const BilTapsL *__restrict__ taps)
{
    const int n = p.dense_list[0];
    for (int li = (int)blockIdx.x; li < n; li += (int)gridDim.x) {
        const int tile = p.dense_list[1 + li];
        const int ty = tile / p.tiles_x;
        layer2_tile<ZMODE, METRIC, true, SHAPE>(p, taps, tile - ty * p.tiles_x, ty);
        __syncthreads(); // the next tile re-initialises the shared state
    }
}
