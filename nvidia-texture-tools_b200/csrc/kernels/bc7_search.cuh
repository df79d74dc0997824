// BC7 endpoint search as a per-thread state machine: one thread per *searcher*.
//
// The reference's refine() (src/bc7/avpcl_mode*.cpp: optimize_endpts -> optimize_one -> perturb_one / exhaustive) runs one
// independent sequential search per (candidate shape | rotation x index mode, region, lsb combination): a few hundred
// map_colors() trials each, every trial depending on the outcome of the previous ones.  A block has up to 432 such
// searchers over its eight modes.  Here each searcher is one thread, and the search is written as an explicit state
// machine whose only expensive step — map_colors of one trial endpoint pair over the region's texels — sits at a single
// point of the loop.  The 32 lanes of a warp are therefore always converged in the trial evaluation, whatever phase
// (first / alternating perturbation, +-3 exhaustive window, restart) each of them is in; only the few instructions
// that pick the next trial diverge.  The thread-per-candidate kernels (bc7.cuh) and the warp-per-candidate kernels
// (tests/simt_emu/bc7_coop.cuh, emulator only) execute the same search; all three give bit-identical blocks.
//
//   k_bc7_tiles        texels of the chunk's blocks (x255, zero outside the image) as [block][16] float4
//   k_bc7_setup<M>     thread per candidate: rough fit, quantise, assign, anchor swap  -> start endpoints + error per region
//   k_bc7_search<M,IM> thread per searcher (grid-stride): the state machine           -> best endpoints + error
//   k_bc7_finish<M>    thread per candidate: first strict minimum over the lsb searches of each region, re-assign, emit,
//                      (error, rank) reduction over the candidates of a block
//
// Trial errors do not depend on the running threshold: map_colors only uses it to leave early with FLT_MAX once the partial
// sum exceeds it, errors are non-negative and fp32 addition of non-negative terms is monotone, so the comparison
// "err < threshold" that follows gives the same answer for the full sum.
#pragma once
#include "bc7.cuh"
#include <type_traits>

#ifdef NVB_EMU_STATS
void emu_bx_stat(int mode, int np, int trials, int iters);
#endif

namespace nvb {

struct Bc7SearchParams {
    Bc7Params P;
    int blk0, nblk;        // this chunk = linear blocks [blk0, blk0 + nblk) of the level
    const float4 *tiles;   // [nblk][16]
    uint4 *setup;          // [nblk][NCAND][NR]          {A, B, lsbs, error bits}
    uint4 *setup_idx;      // [nblk][NCAND]              indices of the start endpoints after the anchor swap, 4 bits per texel
    uint4 *res;            // [nblk][NCAND][NR][NLSB]    {A, B, lsbs, error bits}
    unsigned *perm;        // [nblk][NCAND][NR] (candidate, region) entries ordered by texel count; null = identity
    unsigned *counters;    // [0] next searcher to hand out, [1..17] texel-count histogram, [18..34] scatter cursors
};
#define NVB_BX_COUNTERS 40

// ---- compile-time view shared by the six single-index modes and the two split modes -------------------------------------
template <int M, bool SPLIT = (M == 4 || M == 5)> struct Bc7X;
template <int M> struct Bc7X<M, false> {
    static constexpr bool SPLIT = false, EXHREF = Bc7Cfg<M>::EXHREF;
    static constexpr int NCH = Bc7Cfg<M>::NCH, NR = Bc7Cfg<M>::NR, LSB = Bc7Cfg<M>::LSB;
    static constexpr int NLSB = LSB == 0 ? 1 : (LSB == 1 ? 2 : 4);
    NVB_DEV static int prec(int) { return Bc7Cfg<M>::PREC; }
};
template <int M> struct Bc7X<M, true> {
    static constexpr bool SPLIT = true, EXHREF = false;
    static constexpr int NCH = 4, NR = 1, LSB = 0, NLSB = 1;
    NVB_DEV static int prec(int ch) { return bc7s_prec<M>(ch); }
};

// index arrays are only ever compared for equality inside optimize_one: 4 bits per texel (two arrays for the split modes)
template <bool WIDE> struct Bc7Idx;
template <> struct Bc7Idx<false> {
    unsigned long long lo;
    NVB_DEV void clear() { lo = 0; }
    NVB_DEV bool differs(const Bc7Idx &o) const { return lo != o.lo; }
};
template <> struct Bc7Idx<true> {
    unsigned long long lo, hi;
    NVB_DEV void clear() { lo = hi = 0; }
    NVB_DEV bool differs(const Bc7Idx &o) const { return lo != o.lo || hi != o.hi; }
};

NVB_DEV int bx_get(unsigned v, int ch) { return (int)((v >> (8 * ch)) & 0xFFu); }
NVB_DEV unsigned bx_set(unsigned v, int ch, int x) { return (v & ~(0xFFu << (8 * ch))) | ((unsigned)(x & 0xFF) << (8 * ch)); }

// interpolation weights as compile-time constants (same values as kAvpclW3/7/15)
NVB_DEV constexpr int avpcl_wc(int nidx, int i) {
    constexpr int w3[4] = {0, 21, 43, 64};
    constexpr int w7[8] = {0, 9, 18, 27, 37, 46, 55, 64};
    constexpr int w15[16] = {0, 4, 9, 13, 17, 21, 26, 30, 34, 38, 43, 47, 51, 55, 60, 64};
    return nidx == 16 ? w15[i] : nidx == 8 ? w7[i] : w3[i];
}
template <int NIDX> NVB_DEV void bx_lerp_row(int a, int b, float *out, int stride) {
#pragma unroll
    for (int j = 0; j < NIDX; ++j) out[j * stride] = (float)((a * avpcl_wc(NIDX, NIDX - 1 - j) + b * avpcl_wc(NIDX, j) + 32) >> 6);
}

// "stop at the first increase" scan of the reference, one more palette entry: err_j against the running best
#define NVB_BX_SCAN_STEP(e, j)                     \
    {                                              \
        const bool gt_ = (e) > best, lt_ = (e) < best; \
        live = live && !gt_;                       \
        const bool upd_ = live && lt_;             \
        best = upd_ ? (e) : best;                  \
        bj = upd_ ? (j) : bj;                      \
    }

// map_colors of modes 0,1,2,3,6,7 for one trial over the np texels of the region (px: this thread's shared-memory slot;
// for the RGB modes px[i].w already holds (alpha - 255)^2)
template <int M> NVB_DEV float bx_eval(const float4 *px, int np, unsigned A, unsigned B, int la, int lb, Bc7Idx<false> &idx) {
    using C = Bc7Cfg<M>;
    constexpr int N = C::NIDX;
    float pal[C::NCH][N];
#pragma unroll
    for (int ch = 0; ch < C::NCH; ch++) {
        int a, b;
        if (C::LSB == 0) {
            a = avpcl_unquantize(bx_get(A, ch), C::PREC);
            b = avpcl_unquantize(bx_get(B, ch), C::PREC);
        } else {
            const int lbb = (C::LSB == 1) ? la : lb;
            a = avpcl_unquantize((bx_get(A, ch) << 1) | la, C::PREC + 1);
            b = avpcl_unquantize((bx_get(B, ch) << 1) | lbb, C::PREC + 1);
        }
        bx_lerp_row<N>(a, b, pal[ch], 1);
    }
    // Two palette entries per instruction: the differences and squares issue as FADD2 / FMUL2 (c - p == c + (-p) exactly);
    // the sums that consume the squares are the unfused FFMA2-by-one adds of nvb_common.cuh (a plain packed add of a packed
    // product would be contracted by ptxas), in the reference's order ((x^2 + y^2) + z^2) + w^2.
    float2 npal[C::NCH][N / 2];
#pragma unroll
    for (int ch = 0; ch < C::NCH; ch++)
#pragma unroll
        for (int j = 0; j < N / 2; ++j) npal[ch][j] = make_float2(-pal[ch][2 * j], -pal[ch][2 * j + 1]);
    float tot = 0;
    unsigned long long id = 0;
    for (int i = 0; i < np; ++i) {
        const float4 c = px[i];
        const float ww = c.w;
        const float2 cx = f2splat(c.x), cy = f2splat(c.y), cz = f2splat(c.z), cw = f2splat(c.w);
        float best = 0;
        int bj = 0;
        bool live = true;
#pragma unroll
        for (int jp = 0; jp < N / 2; ++jp) {
            const float2 x = f2add(cx, npal[0][jp]), y = f2add(cy, npal[1][jp]), z = f2add(cz, npal[2][jp]);
            const float2 xx = f2mul(x, x), yy = f2mul(y, y), zz = f2mul(z, z);
            float e0, e1;
            if (C::NCH == 3) {
                const float2 e = f2add_s(f2add_s(f2add_s(xx, yy), zz), f2splat(ww));
                e0 = e.x;
                e1 = e.y;
            } else {
                const float2 w = f2add(cw, npal[C::NCH - 1][jp]);
                const float2 w2 = f2mul(w, w);
                const float2 e = f2add_s(f2add_s(f2add_s(xx, yy), zz), w2);
                e0 = e.x;
                e1 = e.y;
            }
            if (jp == 0) best = e0;
            else NVB_BX_SCAN_STEP(e0, 2 * jp)
            NVB_BX_SCAN_STEP(e1, 2 * jp + 1)
        }
        tot += best;
        id = (id << 4) | (unsigned long long)bj;  // only ever compared for equality
    }
    idx.lo = id;
    return tot;
}

// negated palette of one trial as (entry 2j, entry 2j+1) pairs
template <int M> NVB_DEV void bx_neg_palette(unsigned A, unsigned B, int la, int lb, float2 npal[Bc7Cfg<M>::NCH][Bc7Cfg<M>::NIDX / 2]) {
    using C = Bc7Cfg<M>;
    constexpr int N = C::NIDX;
#pragma unroll
    for (int ch = 0; ch < C::NCH; ch++) {
        int a, b;
        if (C::LSB == 0) {
            a = avpcl_unquantize(bx_get(A, ch), C::PREC);
            b = avpcl_unquantize(bx_get(B, ch), C::PREC);
        } else {
            const int lbb = (C::LSB == 1) ? la : lb;
            a = avpcl_unquantize((bx_get(A, ch) << 1) | la, C::PREC + 1);
            b = avpcl_unquantize((bx_get(B, ch) << 1) | lbb, C::PREC + 1);
        }
        float row[N];
        bx_lerp_row<N>(a, b, row, 1);
#pragma unroll
        for (int j = 0; j < N / 2; ++j) npal[ch][j] = make_float2(-row[2 * j], -row[2 * j + 1]);
    }
}

// map_colors of TWO trials over the same texels (the texel is read and splatted once); same arithmetic per trial as bx_eval
template <int M> NVB_DEV void bx_eval2(const float4 *px, int np, unsigned A0, unsigned B0, unsigned A1, unsigned B1, int la, int lb,
                                       float &err0, float &err1, Bc7Idx<false> &idx0, Bc7Idx<false> &idx1) {
    using C = Bc7Cfg<M>;
    constexpr int N = C::NIDX;
    float2 np0[C::NCH][N / 2], np1[C::NCH][N / 2];
    bx_neg_palette<M>(A0, B0, la, lb, np0);
    bx_neg_palette<M>(A1, B1, la, lb, np1);
    float tot0 = 0, tot1 = 0;
    // the index arrays are only compared for equality: with four palette entries 2 bits per texel fit one 32-bit word, and
    // "id * 4 + j" is one instruction where the 64-bit shift-and-or takes three
    typename std::conditional<(N <= 4), unsigned, unsigned long long>::type id0 = 0, id1 = 0;
    constexpr int IDB = N <= 4 ? 2 : 4;
    for (int i = 0; i < np; ++i) {
        const float4 c = px[i];
        const float ww = c.w;
        const float2 cx = f2splat(c.x), cy = f2splat(c.y), cz = f2splat(c.z), cw = f2splat(c.w);
#pragma unroll
        for (int t = 0; t < 2; t++) {
            float best = 0;
            int bj = 0;
            bool live = true;
#pragma unroll
            for (int jp = 0; jp < N / 2; ++jp) {
                const float2 p0 = t ? np1[0][jp] : np0[0][jp], p1 = t ? np1[1][jp] : np0[1][jp], p2 = t ? np1[2][jp] : np0[2][jp];
                const float2 x = f2add(cx, p0), y = f2add(cy, p1), z = f2add(cz, p2);
                const float2 xx = f2mul(x, x), yy = f2mul(y, y), zz = f2mul(z, z);
                float e0, e1;
                if (C::NCH == 3) {
                    const float2 e = f2add_s(f2add_s(f2add_s(xx, yy), zz), f2splat(ww));
                    e0 = e.x;
                    e1 = e.y;
                } else {
                    const float2 p3 = t ? np1[C::NCH - 1][jp] : np0[C::NCH - 1][jp];
                    const float2 w = f2add(cw, p3);
                    const float2 w2 = f2mul(w, w);
                    const float2 e = f2add_s(f2add_s(f2add_s(xx, yy), zz), w2);
                    e0 = e.x;
                    e1 = e.y;
                }
                if (jp == 0) best = e0;
                else NVB_BX_SCAN_STEP(e0, 2 * jp)
                NVB_BX_SCAN_STEP(e1, 2 * jp + 1)
            }
            if (t == 0) {
                tot0 += best;
                id0 = (id0 << IDB) + (unsigned)bj;
            } else {
                tot1 += best;
                id1 = (id1 << IDB) + (unsigned)bj;
            }
        }
    }
    err0 = tot0;
    err1 = tot1;
    idx0.lo = id0;
    idx1.lo = id1;
}

// map_colors of modes 4,5 for one trial (all 16 texels; rotation applied to the texel as it is read)
template <int M, int IM> NVB_DEV float bx_eval_split(const float4 *px, int rot, unsigned A, unsigned B, Bc7Idx<true> &idx) {
    constexpr int NRGB = (M == 5) ? 4 : (IM == 1 ? 8 : 4), NA = (M == 5) ? 4 : (IM == 1 ? 4 : 8);
    float prgb[3][NRGB], pa[NA];
#pragma unroll
    for (int ch = 0; ch < 3; ch++)
        bx_lerp_row<NRGB>(avpcl_unquantize(bx_get(A, ch), Bc7SplitCfg<M>::PREC_RGB), avpcl_unquantize(bx_get(B, ch), Bc7SplitCfg<M>::PREC_RGB), prgb[ch], 1);
    bx_lerp_row<NA>(avpcl_unquantize(bx_get(A, 3), Bc7SplitCfg<M>::PREC_A), avpcl_unquantize(bx_get(B, 3), Bc7SplitCfg<M>::PREC_A), pa, 1);
    float tot = 0;
    unsigned long long irgb = 0, ia = 0;
    for (int i = 0; i < 16; ++i) {
        const float4 c = px[i];  // already rotated
        float ea, er;
        int ja, jr;
        {
            float best = 0;
            int bj = 0;
            bool live = true;
#pragma unroll
            for (int j = 0; j < NA; ++j) {
                const float d = c.w - pa[j];
                const float e = d * d;
                if (j == 0) best = e;
                else NVB_BX_SCAN_STEP(e, j)
            }
            ea = best;
            ja = bj;
        }
        {
            float best = 0;
            int bj = 0;
            bool live = true;
#pragma unroll
            for (int j = 0; j < NRGB; ++j) {
                const float x = c.x - prgb[0][j], y = c.y - prgb[1][j], z = c.z - prgb[2][j];
                const float e = x * x + y * y + z * z;
                if (j == 0) best = e;
                else NVB_BX_SCAN_STEP(e, j)
            }
            er = best;
            jr = bj;
        }
        const float e1 = rot == 0 ? ea : er, e2 = rot == 0 ? er : ea;
        tot += e1;
        tot += e2;
        irgb = (irgb << 4) | (unsigned long long)jr;
        ia = (ia << 4) | (unsigned long long)ja;
    }
    idx.lo = irgb;
    idx.hi = ia;
    return tot;
}

// ---- the searcher ---------------------------------------------------------------------------------------------------------
enum {
    BXP_LOAD = 0, BXP_INIT_WAIT, BXP_CH_START, BXP_PERT_EMIT, BXP_PERT_WAIT, BXP_PERT_FIN,
    BXP_EXH_CH_START, BXP_EXH_EMIT, BXP_EXH_WAIT, BXP_EXH_CH_END, BXP_STORE, BXP_EXIT
};

// IM: index mode of mode 4 (0 or 1); 0 for every other mode
template <int M, int IM> __global__ void __launch_bounds__(128) k_bc7_search(Bc7SearchParams S) {
    using X = Bc7X<M>;
    using Idx = Bc7Idx<X::SPLIT>;
    constexpr int NCAND_SPLIT = 4 * (M == 4 ? 2 : 1);
    // this thread's texels: 16 float4 + 1 of padding, so that the 128-bit reads of 8 adjacent lanes hit 32 different banks
    __shared__ float4 s_px[128 * 17];
    float4 *px = s_px + threadIdx.x * 17;
    long long s = 0;
    long long total;
    int ncand = 1;
    if constexpr (X::SPLIT) total = (long long)S.nblk * 4;
    else {
        ncand = Bc7Cfg<M>::NITEMS;
        total = (long long)S.nblk * ncand * X::NR * X::NLSB;
    }
    const int nblocks = S.P.lv.bw * S.P.lv.bh;

    // searcher state
    int phase = BXP_LOAD;
    int np = 0, rot = 0;
    long long slot = 0;          // where the result goes
    unsigned A = 0, B = 0;       // current best endpoints ("opt"), 8 bits per channel
    int la = 0, lb = 0;
    float opt_err = 0;
    int ch = 0;
    // perturb_one
    int pk = 0, do_b = 0, pv = 0, step = 0, sgn = -1, beststep = 0;
    bool improved = false;
    float pmin = 0;              // min_err of perturb_one / best_err of exhaustive
    Idx pidx;                    // indices of the best trial of the running perturb_one / exhaustive
    float err0 = 0;
    int va = 0;
    Idx t0, orig_idx, new_idx;
    // exhaustive
    bool first = true, exh_done = false;
    float exh_orig = 0;
    int eo = 0, ei = 0, ta = 0, tb = 0, amin = 0, bmin = 0;
    pidx.clear(); t0.clear(); orig_idx.clear(); new_idx.clear();

#ifdef NVB_EMU_STATS
    int st_trials = 0, st_iters = 0;
#endif
    // next step of perturb_one's (step, sign) walk
#define NVB_BX_ADV_SIGN()                  \
    {                                      \
        if (sgn < 0) {                     \
            sgn = 1;                       \
        } else {                           \
            if (improved) pv += beststep;  \
            improved = false;              \
            step >>= 1;                    \
            sgn = -1;                      \
        }                                  \
    }
    // the trial of the current (step, sign), or — out of range — on to the next one
#define NVB_BX_TRY_PERT()                                  \
    {                                                      \
        const int v_ = pv + sgn * step;                    \
        if (v_ >= 0 && v_ < (1 << X::prec(ch))) {          \
            tA = do_b ? A : bx_set(A, ch, v_);             \
            tB = do_b ? bx_set(B, ch, v_) : B;             \
            phase = BXP_PERT_WAIT;                         \
            have = true;                                   \
        } else                                             \
            NVB_BX_ADV_SIGN()                              \
    }
    // the trial under exhaustive()'s cursor, and the cursor moved on
#define NVB_BX_EXH_EMIT()                                                                                             \
    {                                                                                                                 \
        const int prec_ = X::prec(ch), A0_ = bx_get(A, ch), B0_ = bx_get(B, ch);                                      \
        const int alow_ = max(0, A0_ - 3), ahigh_ = min((1 << prec_) - 1, A0_ + 3);                                   \
        const int blow_ = max(0, B0_ - 3), bhigh_ = min((1 << prec_) - 1, B0_ + 3);                                   \
        const bool a_le_b_ = A0_ <= B0_;                                                                              \
        const int o_end_ = a_le_b_ ? ahigh_ : bhigh_ - 1, i_end_ = a_le_b_ ? bhigh_ - 1 : ahigh_, lowi_ = a_le_b_ ? blow_ : alow_; \
        ta = a_le_b_ ? eo : ei;                                                                                       \
        tb = a_le_b_ ? ei : eo;                                                                                       \
        tA = bx_set(A, ch, ta);                                                                                       \
        tB = bx_set(B, ch, tb);                                                                                       \
        ++ei;                                                                                                         \
        while (!exh_done && ei > i_end_) {                                                                            \
            ++eo;                                                                                                     \
            if (eo > o_end_) exh_done = true;                                                                         \
            else ei = max(eo, lowi_);                                                                                 \
        }                                                                                                             \
        phase = BXP_EXH_WAIT;                                                                                         \
        have = true;                                                                                                  \
    }

    unsigned tA = 0, tB = 0;
    bool have = false;
    for (;;) {
        // slow path: lanes whose next trial is not simply the next step of the running perturb_one / exhaustive walk
        while (!have && phase != BXP_EXIT) {
#ifdef NVB_EMU_STATS
            ++st_iters;
#endif
            switch (phase) {
            case BXP_LOAD: {
                // searchers are handed out one at a time: the searches differ a lot in length (every improvement restarts
                // the channel loop), a static assignment would leave most lanes of a warp waiting for the longest one
                s = (long long)atomicAdd(S.counters, 1u);
                if (s >= total) {
                    phase = BXP_EXIT;
                    break;
                }
                uint4 su;
                if constexpr (X::SPLIT) {
                    const int blk = (int)(s >> 2);
                    rot = (int)(s & 3);
                    slot = (long long)blk * NCAND_SPLIT + rot * (M == 4 ? 2 : 1) + IM;
                    su = S.setup[slot];
                    const float4 *tile = S.tiles + (size_t)blk * 16;
                    np = 16;
                    la = lb = 0;
                    for (int i = 0; i < 16; i++) {
                        float4 c = __ldg(tile + i);
                        // AGBR / RABG / RGAB: swap channel rot-1 with alpha
                        const float w0 = c.w;
                        c.w = rot == 1 ? c.x : rot == 2 ? c.y : rot == 3 ? c.z : c.w;
                        c.x = rot == 1 ? w0 : c.x;
                        c.y = rot == 2 ? w0 : c.y;
                        c.z = rot == 3 ? w0 : c.z;
                        px[i] = c;
                    }
                } else {
                    using C = Bc7Cfg<M>;
                    const int lsbmode = (int)(s % X::NLSB);
                    const long long q = s / X::NLSB;
                    const unsigned e = S.perm ? S.perm[q] : (unsigned)q;
                    const int region = (int)(e % X::NR);
                    const int cand = (int)((e / X::NR) % ncand);
                    const int blk = (int)(e / (X::NR * ncand));
                    slot = (long long)e * X::NLSB + lsbmode;
                    su = S.setup[e];
                    const float4 *tile = S.tiles + (size_t)blk * 16;
                    int shape = 0;
                    if constexpr (C::NSH > 1) shape = S.P.shapes[((size_t)Bc7Slot<M>::v * nblocks + S.blk0 + blk) * 16 + cand];
                    np = 0;
                    for (int i = 0; i < 16; i++)
                        if (bc7_region<C::NR>(shape, i) == region) {
                            float4 c = __ldg(tile + i);
                            if (C::NCH == 3) {
                                const float w = c.w - 255.0f;  // the palette's alpha is 255 in the RGB modes
                                c.w = w * w;
                            }
                            px[np++] = c;
                        }
                    la = lsbmode & 1;
                    lb = (X::LSB == 2) ? (lsbmode >> 1) & 1 : 0;
                }
                A = su.x;
                B = su.y;
                if (X::LSB != 0) {
                    // in_err = map_colors(temp_in with this lsb combination)
                    tA = A;
                    tB = B;
                    phase = BXP_INIT_WAIT;
                    have = true;
                } else {
                    opt_err = __uint_as_float(su.w);
                    ch = 0;
                    phase = BXP_CH_START;
                }
                break;
            }
            case BXP_CH_START:
                if (ch >= X::NCH) {
                    ch = 0;
                    first = true;
                    phase = BXP_EXH_CH_START;
                } else {
                    pk = 0;
                    do_b = 0;
                    pv = bx_get(A, ch);
                    pmin = opt_err;
                    step = 1 << (X::prec(ch) - 1);
                    sgn = -1;
                    improved = false;
                    phase = BXP_PERT_EMIT;
                    NVB_BX_TRY_PERT()  // the first step is half the range: exactly one of -step / +step is in range
                    if (!have) NVB_BX_TRY_PERT()
                }
                break;
            case BXP_PERT_EMIT:
                if (step == 0) phase = BXP_PERT_FIN;
                else NVB_BX_TRY_PERT()
                break;
            case BXP_PERT_FIN: {
                bool again = false;  // start another perturb_one on endpoint do_b
                if (pk == 0) {
                    err0 = pmin;
                    va = pv;
                    t0 = pidx;
                    pk = 1;
                    do_b = 1;
                    again = true;
                } else if (pk == 1) {
                    const float err1 = pmin;
                    if (err0 < err1) {
                        if (!(err0 >= opt_err)) {
                            new_idx = orig_idx = t0;
                            A = bx_set(A, ch, va);
                            opt_err = err0;
                            do_b = 1;
                            again = true;
                        }
                    } else {
                        if (!(err1 >= opt_err)) {
                            new_idx = orig_idx = pidx;
                            B = bx_set(B, ch, pv);
                            opt_err = err1;
                            do_b = 0;
                            again = true;
                        }
                    }
                    pk = 2;
                    if (!again) {  // `continue`: next channel without the restart test
                        ++ch;
                        phase = BXP_CH_START;
                    }
                } else {
                    if (pmin >= opt_err) {
                        if (orig_idx.differs(new_idx)) ch = -1;  // indices changed: start over
                        ++ch;
                        phase = BXP_CH_START;
                    } else {
                        new_idx = pidx;
                        if (do_b == 0) A = bx_set(A, ch, pv);
                        else B = bx_set(B, ch, pv);
                        opt_err = pmin;
                        do_b = 1 - do_b;
                        again = true;
                    }
                }
                if (again) {
                    pv = bx_get(do_b ? B : A, ch);
                    pmin = opt_err;
                    step = 1 << (X::prec(ch) - 1);
                    sgn = -1;
                    improved = false;
                    phase = BXP_PERT_EMIT;
                    NVB_BX_TRY_PERT()
                    if (!have) NVB_BX_TRY_PERT()
                }
                break;
            }
            case BXP_EXH_CH_START: {
                if (ch >= X::NCH) {
                    phase = BXP_STORE;
                    break;
                }
                exh_orig = opt_err;
                pmin = exh_orig;
                if (exh_orig == 0) {  // exhaustive() returns at once; nothing can improve
                    ++ch;
                    break;
                }
                const int prec = X::prec(ch), A0 = bx_get(A, ch), B0 = bx_get(B, ch);
                const int alow = max(0, A0 - 3), ahigh = min((1 << prec) - 1, A0 + 3);
                const int blow = max(0, B0 - 3), bhigh = min((1 << prec) - 1, B0 + 3);
                const bool a_le_b = A0 <= B0;
                const int o_end = a_le_b ? ahigh : bhigh - 1, i_end = a_le_b ? bhigh - 1 : ahigh, lowi = a_le_b ? blow : alow;
                eo = a_le_b ? alow : blow;
                exh_done = eo > o_end;
                if (!exh_done) ei = max(eo, lowi);
                while (!exh_done && ei > i_end) {
                    ++eo;
                    if (eo > o_end) exh_done = true;
                    else ei = max(eo, lowi);
                }
                phase = BXP_EXH_CH_END;
                if (!exh_done) NVB_BX_EXH_EMIT()
                break;
            }
            case BXP_EXH_EMIT:
                if (exh_done) phase = BXP_EXH_CH_END;
                else NVB_BX_EXH_EMIT()
                break;
            case BXP_EXH_CH_END: {
                const float new_err = pmin;
                if (pmin < exh_orig) {
                    A = bx_set(A, ch, amin);
                    B = bx_set(B, ch, bmin);
                    t0 = pidx;
                    if (X::EXHREF) opt_err = pmin;  // modes 0 and 3 pass the error by reference
                }
                if (new_err < opt_err) {
                    opt_err = new_err;
                    if (first) {
                        orig_idx = t0;
                        first = false;
                    } else if (orig_idx.differs(t0)) {
                        ch = -1;
                        first = true;
                    }
                }
                ++ch;
                phase = BXP_EXH_CH_START;
                break;
            }
            case BXP_STORE:
                S.res[slot] = make_uint4(A, B, (unsigned)(la | (lb << 1)), __float_as_uint(opt_err));
#ifdef NVB_EMU_STATS
                emu_bx_stat(M, np, st_trials, st_iters);
                st_trials = st_iters = 0;
#endif
                phase = BXP_LOAD;
                break;
            default:
                break;
            }
        }
        // Every lane of the warp stays in the loop until the last one has run out of work (idle lanes carry np = 0), so the
        // vote below is executed by all 32 lanes: it is the explicit reconvergence point in front of the trial evaluation.
        // (Without it ptxas treats the hand-out loop as a possible spin loop and lets the warp run on in pieces.)
        if (__all_sync(0xffffffffu, phase == BXP_EXIT)) break;
        if (phase == BXP_EXIT) np = 0;

#ifdef NVB_EMU_STATS
        ++st_trials;
#endif
        // ---- the one expensive step: every live lane of the warp is here together ----
        Idx ti;
        float err;
        if constexpr (X::SPLIT) err = bx_eval_split<M, IM>(px, rot, tA, tB, ti);
        else err = bx_eval<M>(px, np, tA, tB, la, lb, ti);

        // consume the result and, on the common path, produce the next trial right here (all lanes together)
        have = false;
        if (phase == BXP_EXIT) {
            // idle lane
        } else if (phase == BXP_INIT_WAIT) {
            opt_err = err;
            ch = 0;
            phase = BXP_CH_START;
        } else if (phase == BXP_PERT_WAIT) {
            if (err < pmin) {
                improved = true;
                pmin = err;
                beststep = sgn * step;
                pidx = ti;
            }
            NVB_BX_ADV_SIGN()
            phase = BXP_PERT_EMIT;
            if (step != 0) NVB_BX_TRY_PERT()
            if (!have && step != 0) NVB_BX_TRY_PERT()
            if (!have && step == 0) phase = BXP_PERT_FIN;
        } else {  // BXP_EXH_WAIT
            if (err < pmin) {
                amin = ta;
                bmin = tb;
                pmin = err;
                pidx = ti;
            }
            phase = BXP_EXH_CH_END;
            if (!exh_done) NVB_BX_EXH_EMIT()
        }
    }
#undef NVB_BX_ADV_SIGN
#undef NVB_BX_TRY_PERT
#undef NVB_BX_EXH_EMIT
}

// Two trials per loop trip (modes 0, 1, 2, 3, 7): the -step / +step pair of a perturb_one level, two consecutive (a, b) pairs
// of exhaustive().  The results are consumed in the reference's order (the second one against the threshold the first one may
// have lowered), so the accepted steps are the same; what changes is that the state transitions - which run with one or
// two active lanes - are paid once per two trials, and the texel is read and splatted once for both.
template <int M> __global__ void __launch_bounds__(128) k_bc7_search2(Bc7SearchParams S) {
    constexpr int IM = 0;
    using X = Bc7X<M>;
    using Idx = Bc7Idx<X::SPLIT>;
    constexpr int NCAND_SPLIT = 4 * (M == 4 ? 2 : 1);
    // this thread's texels: 16 float4 + 1 of padding, so that the 128-bit reads of 8 adjacent lanes hit 32 different banks
    __shared__ float4 s_px[128 * 17];
    float4 *px = s_px + threadIdx.x * 17;
    long long s = 0;
    long long total;
    int ncand = 1;
    if constexpr (X::SPLIT) total = (long long)S.nblk * 4;
    else {
        ncand = Bc7Cfg<M>::NITEMS;
        total = (long long)S.nblk * ncand * X::NR * X::NLSB;
    }
    const int nblocks = S.P.lv.bw * S.P.lv.bh;

    // searcher state
    int phase = BXP_LOAD;
    int np = 0, rot = 0;
    long long slot = 0;          // where the result goes
    unsigned A = 0, B = 0;       // current best endpoints ("opt"), 8 bits per channel
    int la = 0, lb = 0;
    float opt_err = 0;
    int ch = 0;
    // perturb_one
    int pk = 0, do_b = 0, pv = 0, step = 0, sgn = -1, beststep = 0;
    bool improved = false;
    float pmin = 0;              // min_err of perturb_one / best_err of exhaustive
    Idx pidx;                    // indices of the best trial of the running perturb_one / exhaustive
    float err0 = 0;
    int va = 0;
    Idx t0, orig_idx, new_idx;
    // exhaustive
    bool first = true, exh_done = false;
    float exh_orig = 0;
    int eo = 0, ei = 0, ta = 0, tb = 0, amin = 0, bmin = 0;
    pidx.clear(); t0.clear(); orig_idx.clear(); new_idx.clear();

#ifdef NVB_EMU_STATS
    int st_trials = 0, st_iters = 0;
#endif
    // next step of perturb_one's (step, sign) walk
#define NVB_BX_ADV_SIGN()                  \
    {                                      \
        if (improved) pv += beststep;      \
        improved = false;                  \
        step >>= 1;                        \
    }
    // the trial of the current (step, sign), or — out of range — on to the next one
#define NVB_BX_TRY_PERT()                                                                     \
    {                                                                                         \
        /* step <= half the range: at least one of pv - step / pv + step is in range */       \
        const int vm_ = pv - step, vp_ = pv + step;                                           \
        const bool okm_ = vm_ >= 0, okp_ = vp_ < (1 << X::prec(ch));                          \
        const int v0_ = okm_ ? vm_ : vp_;                                                     \
        s0sign = okm_ ? -1 : 1;                                                               \
        have1 = okm_ && okp_;                                                                 \
        const int v1_ = have1 ? vp_ : v0_;                                                    \
        tA = do_b ? A : bx_set(A, ch, v0_);                                                   \
        tB = do_b ? bx_set(B, ch, v0_) : B;                                                   \
        tA1 = do_b ? A : bx_set(A, ch, v1_);                                                  \
        tB1 = do_b ? bx_set(B, ch, v1_) : B;                                                  \
        phase = BXP_PERT_WAIT;                                                                \
        have = true;                                                                          \
    }
    // the trial under exhaustive()'s cursor, and the cursor moved on
#define NVB_BX_EXH_EMIT()                                                                                             \
    {                                                                                                                 \
        const int prec_ = X::prec(ch), A0_ = bx_get(A, ch), B0_ = bx_get(B, ch);                                      \
        const int alow_ = max(0, A0_ - 3), ahigh_ = min((1 << prec_) - 1, A0_ + 3);                                   \
        const int blow_ = max(0, B0_ - 3), bhigh_ = min((1 << prec_) - 1, B0_ + 3);                                   \
        const bool a_le_b_ = A0_ <= B0_;                                                                              \
        const int o_end_ = a_le_b_ ? ahigh_ : bhigh_ - 1, i_end_ = a_le_b_ ? bhigh_ - 1 : ahigh_, lowi_ = a_le_b_ ? blow_ : alow_; \
        ta = a_le_b_ ? eo : ei;                                                                                       \
        tb = a_le_b_ ? ei : eo;                                                                                       \
        tA = bx_set(A, ch, ta);                                                                                       \
        tB = bx_set(B, ch, tb);                                                                                       \
        ++ei;                                                                                                         \
        while (!exh_done && ei > i_end_) {                                                                            \
            ++eo;                                                                                                     \
            if (eo > o_end_) exh_done = true;                                                                         \
            else ei = max(eo, lowi_);                                                                                 \
        }                                                                                                             \
        have1 = !exh_done;                                                                                            \
        ta1 = ta; tb1 = tb; tA1 = tA; tB1 = tB;                                                                       \
        if (have1) {                                                                                                  \
            ta1 = a_le_b_ ? eo : ei;                                                                                  \
            tb1 = a_le_b_ ? ei : eo;                                                                                  \
            tA1 = bx_set(A, ch, ta1);                                                                                 \
            tB1 = bx_set(B, ch, tb1);                                                                                 \
            ++ei;                                                                                                     \
            while (!exh_done && ei > i_end_) {                                                                        \
                ++eo;                                                                                                 \
                if (eo > o_end_) exh_done = true;                                                                     \
                else ei = max(eo, lowi_);                                                                             \
            }                                                                                                         \
        }                                                                                                             \
        phase = BXP_EXH_WAIT;                                                                                         \
        have = true;                                                                                                  \
    }

    unsigned tA = 0, tB = 0, tA1 = 0, tB1 = 0;  // the two trials of a trip
    int s0sign = -1, ta1 = 0, tb1 = 0;
    bool have = false, have1 = false;
    for (;;) {
        // slow path: lanes whose next trial is not simply the next step of the running perturb_one / exhaustive walk
        while (!have && phase != BXP_EXIT) {
#ifdef NVB_EMU_STATS
            ++st_iters;
#endif
            switch (phase) {
            case BXP_LOAD: {
                // searchers are handed out one at a time: the searches differ a lot in length (every improvement restarts
                // the channel loop), a static assignment would leave most lanes of a warp waiting for the longest one
                s = (long long)atomicAdd(S.counters, 1u);
                if (s >= total) {
                    phase = BXP_EXIT;
                    break;
                }
                uint4 su;
                if constexpr (X::SPLIT) {
                    const int blk = (int)(s >> 2);
                    rot = (int)(s & 3);
                    slot = (long long)blk * NCAND_SPLIT + rot * (M == 4 ? 2 : 1) + IM;
                    su = S.setup[slot];
                    const float4 *tile = S.tiles + (size_t)blk * 16;
                    np = 16;
                    la = lb = 0;
                    for (int i = 0; i < 16; i++) {
                        float4 c = __ldg(tile + i);
                        // AGBR / RABG / RGAB: swap channel rot-1 with alpha
                        const float w0 = c.w;
                        c.w = rot == 1 ? c.x : rot == 2 ? c.y : rot == 3 ? c.z : c.w;
                        c.x = rot == 1 ? w0 : c.x;
                        c.y = rot == 2 ? w0 : c.y;
                        c.z = rot == 3 ? w0 : c.z;
                        px[i] = c;
                    }
                } else {
                    using C = Bc7Cfg<M>;
                    const int lsbmode = (int)(s % X::NLSB);
                    const long long q = s / X::NLSB;
                    const unsigned e = S.perm ? S.perm[q] : (unsigned)q;
                    const int region = (int)(e % X::NR);
                    const int cand = (int)((e / X::NR) % ncand);
                    const int blk = (int)(e / (X::NR * ncand));
                    slot = (long long)e * X::NLSB + lsbmode;
                    su = S.setup[e];
                    const float4 *tile = S.tiles + (size_t)blk * 16;
                    int shape = 0;
                    if constexpr (C::NSH > 1) shape = S.P.shapes[((size_t)Bc7Slot<M>::v * nblocks + S.blk0 + blk) * 16 + cand];
                    np = 0;
                    for (int i = 0; i < 16; i++)
                        if (bc7_region<C::NR>(shape, i) == region) {
                            float4 c = __ldg(tile + i);
                            if (C::NCH == 3) {
                                const float w = c.w - 255.0f;  // the palette's alpha is 255 in the RGB modes
                                c.w = w * w;
                            }
                            px[np++] = c;
                        }
                    la = lsbmode & 1;
                    lb = (X::LSB == 2) ? (lsbmode >> 1) & 1 : 0;
                }
                A = su.x;
                B = su.y;
                if (X::LSB != 0) {
                    // in_err = map_colors(temp_in with this lsb combination)
                    tA = tA1 = A;
                    tB = tB1 = B;
                    have1 = false;
                    phase = BXP_INIT_WAIT;
                    have = true;
                } else {
                    opt_err = __uint_as_float(su.w);
                    ch = 0;
                    phase = BXP_CH_START;
                }
                break;
            }
            case BXP_CH_START:
                if (ch >= X::NCH) {
                    ch = 0;
                    first = true;
                    phase = BXP_EXH_CH_START;
                } else {
                    pk = 0;
                    do_b = 0;
                    pv = bx_get(A, ch);
                    pmin = opt_err;
                    step = 1 << (X::prec(ch) - 1);
                    sgn = -1;
                    improved = false;
                    phase = BXP_PERT_EMIT;
                    NVB_BX_TRY_PERT()
                }
                break;
            case BXP_PERT_EMIT:
                if (step == 0) phase = BXP_PERT_FIN;
                else NVB_BX_TRY_PERT()
                break;
            case BXP_PERT_FIN: {
                bool again = false;  // start another perturb_one on endpoint do_b
                if (pk == 0) {
                    err0 = pmin;
                    va = pv;
                    t0 = pidx;
                    pk = 1;
                    do_b = 1;
                    again = true;
                } else if (pk == 1) {
                    const float err1 = pmin;
                    if (err0 < err1) {
                        if (!(err0 >= opt_err)) {
                            new_idx = orig_idx = t0;
                            A = bx_set(A, ch, va);
                            opt_err = err0;
                            do_b = 1;
                            again = true;
                        }
                    } else {
                        if (!(err1 >= opt_err)) {
                            new_idx = orig_idx = pidx;
                            B = bx_set(B, ch, pv);
                            opt_err = err1;
                            do_b = 0;
                            again = true;
                        }
                    }
                    pk = 2;
                    if (!again) {  // `continue`: next channel without the restart test
                        ++ch;
                        phase = BXP_CH_START;
                    }
                } else {
                    if (pmin >= opt_err) {
                        if (orig_idx.differs(new_idx)) ch = -1;  // indices changed: start over
                        ++ch;
                        phase = BXP_CH_START;
                    } else {
                        new_idx = pidx;
                        if (do_b == 0) A = bx_set(A, ch, pv);
                        else B = bx_set(B, ch, pv);
                        opt_err = pmin;
                        do_b = 1 - do_b;
                        again = true;
                    }
                }
                if (again) {
                    pv = bx_get(do_b ? B : A, ch);
                    pmin = opt_err;
                    step = 1 << (X::prec(ch) - 1);
                    sgn = -1;
                    improved = false;
                    phase = BXP_PERT_EMIT;
                    NVB_BX_TRY_PERT()
                }
                break;
            }
            case BXP_EXH_CH_START: {
                if (ch >= X::NCH) {
                    phase = BXP_STORE;
                    break;
                }
                exh_orig = opt_err;
                pmin = exh_orig;
                if (exh_orig == 0) {  // exhaustive() returns at once; nothing can improve
                    ++ch;
                    break;
                }
                const int prec = X::prec(ch), A0 = bx_get(A, ch), B0 = bx_get(B, ch);
                const int alow = max(0, A0 - 3), ahigh = min((1 << prec) - 1, A0 + 3);
                const int blow = max(0, B0 - 3), bhigh = min((1 << prec) - 1, B0 + 3);
                const bool a_le_b = A0 <= B0;
                const int o_end = a_le_b ? ahigh : bhigh - 1, i_end = a_le_b ? bhigh - 1 : ahigh, lowi = a_le_b ? blow : alow;
                eo = a_le_b ? alow : blow;
                exh_done = eo > o_end;
                if (!exh_done) ei = max(eo, lowi);
                while (!exh_done && ei > i_end) {
                    ++eo;
                    if (eo > o_end) exh_done = true;
                    else ei = max(eo, lowi);
                }
                phase = BXP_EXH_CH_END;
                if (!exh_done) NVB_BX_EXH_EMIT()
                break;
            }
            case BXP_EXH_EMIT:
                if (exh_done) phase = BXP_EXH_CH_END;
                else NVB_BX_EXH_EMIT()
                break;
            case BXP_EXH_CH_END: {
                const float new_err = pmin;
                if (pmin < exh_orig) {
                    A = bx_set(A, ch, amin);
                    B = bx_set(B, ch, bmin);
                    t0 = pidx;
                    if (X::EXHREF) opt_err = pmin;  // modes 0 and 3 pass the error by reference
                }
                if (new_err < opt_err) {
                    opt_err = new_err;
                    if (first) {
                        orig_idx = t0;
                        first = false;
                    } else if (orig_idx.differs(t0)) {
                        ch = -1;
                        first = true;
                    }
                }
                ++ch;
                phase = BXP_EXH_CH_START;
                break;
            }
            case BXP_STORE:
                S.res[slot] = make_uint4(A, B, (unsigned)(la | (lb << 1)), __float_as_uint(opt_err));
#ifdef NVB_EMU_STATS
                emu_bx_stat(M, np, st_trials, st_iters);
                st_trials = st_iters = 0;
#endif
                phase = BXP_LOAD;
                break;
            default:
                break;
            }
        }
        // Every lane of the warp stays in the loop until the last one has run out of work (idle lanes carry np = 0), so the
        // vote below is executed by all 32 lanes: it is the explicit reconvergence point in front of the trial evaluation.
        // (Without it ptxas treats the hand-out loop as a possible spin loop and lets the warp run on in pieces.)
        if (__all_sync(0xffffffffu, phase == BXP_EXIT)) break;
        if (phase == BXP_EXIT) np = 0;

#ifdef NVB_EMU_STATS
        ++st_trials;
#endif
        // ---- the one expensive step: every live lane of the warp is here together ----
        Idx ti, ti1;
        float err, err1;
        bx_eval2<M>(px, np, tA, tB, tA1, tB1, la, lb, err, err1, ti, ti1);

        // consume the result and, on the common path, produce the next trial right here (all lanes together)
        have = false;
        if (phase == BXP_EXIT) {
            // idle lane
        } else if (phase == BXP_INIT_WAIT) {
            opt_err = err;
            ch = 0;
            phase = BXP_CH_START;
        } else if (phase == BXP_PERT_WAIT) {
            if (err < pmin) {
                improved = true;
                pmin = err;
                beststep = s0sign * step;
                pidx = ti;
            }
            if (have1 && err1 < pmin) {  // the +step trial, against the threshold the -step trial may have lowered
                improved = true;
                pmin = err1;
                beststep = step;
                pidx = ti1;
            }
            NVB_BX_ADV_SIGN()
            phase = BXP_PERT_FIN;
            if (step != 0) NVB_BX_TRY_PERT()
        } else {  // BXP_EXH_WAIT
            if (err < pmin) {
                amin = ta;
                bmin = tb;
                pmin = err;
                pidx = ti;
            }
            if (have1 && err1 < pmin) {
                amin = ta1;
                bmin = tb1;
                pmin = err1;
                pidx = ti1;
            }
            phase = BXP_EXH_CH_END;
            if (!exh_done) NVB_BX_EXH_EMIT()
        }
    }
#undef NVB_BX_ADV_SIGN
#undef NVB_BX_TRY_PERT
#undef NVB_BX_EXH_EMIT
}


// ---- order the (candidate, region) entries of a chunk by texel count -----------------------------------------------------------
// Lanes of a warp walk the texels of their regions in lock step; a warp whose regions have 3 and 13 texels runs 13 steps
// with most lanes idle.  Searchers are therefore handed out in order of texel count (counting sort, two passes).
template <int M, int NCAND, int PASS> __global__ void __launch_bounds__(256) k_bc7_order(Bc7SearchParams S) {
    using C = Bc7Cfg<M>;
    __shared__ unsigned s_hist[17], s_base[17];
    const int nblocks = S.P.lv.bw * S.P.lv.bh;
    const unsigned total = (unsigned)S.nblk * NCAND * C::NR;
    const unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
    if (threadIdx.x < 17) s_hist[threadIdx.x] = 0;
    __syncthreads();
    int np = 0;
    unsigned rank = 0;
    if (e < total) {
        const int region = (int)(e % C::NR), cand = (int)((e / C::NR) % NCAND), blk = (int)(e / (C::NR * NCAND));
        const int shape = S.P.shapes[((size_t)Bc7Slot<M>::v * nblocks + S.blk0 + blk) * 16 + cand];
        np = __popc(bc7_member_mask<C::NR>(shape, region));
        rank = atomicAdd(&s_hist[np], 1u);
    }
    __syncthreads();
    unsigned *hist = S.counters + 1, *cursor = S.counters + 18;
    if (PASS == 0) {
        if (threadIdx.x < 17 && s_hist[threadIdx.x]) atomicAdd(&hist[threadIdx.x], s_hist[threadIdx.x]);
    } else {
        if (threadIdx.x < 17) {
            unsigned off = 0;
            for (int k = 0; k < (int)threadIdx.x; k++) off += hist[k];
            s_base[threadIdx.x] = off + (s_hist[threadIdx.x] ? atomicAdd(&cursor[threadIdx.x], s_hist[threadIdx.x]) : 0u);
        }
        __syncthreads();
        if (e < total) S.perm[s_base[np] + rank] = e;
    }
}

// ---- tiles -----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bc7_tiles(Bc7SearchParams S, float4 *tiles) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= S.nblk * 16) return;
    const int blk = S.blk0 + (t >> 4), i = t & 15;
    const LevelView &lv = S.P.lv;
    const int x = (blk % lv.bw) * 4 + (i & 3), y = (blk / lv.bw) * 4 + (i >> 2);
    float4 c = make_float4(0.0f, 0.0f, 0.0f, 0.0f);  // texels outside the image are Vector4(0) (BlockCompressor.cpp:152-163)
    if (x < lv.w && y < lv.h) {
        c.x = load_texel(lv, 0, x, y) * 255.0f;
        c.y = load_texel(lv, 1, x, y) * 255.0f;
        c.z = load_texel(lv, 2, x, y) * 255.0f;
        c.w = load_texel(lv, 3, x, y) * 255.0f;
    }
    tiles[t] = c;
}

NVB_DEV void bx_read_tile(const float4 *tiles, int blk, Bc7Tile &t) {
    for (int i = 0; i < 16; i++) {
        const float4 c = tiles[(size_t)blk * 16 + i];
        t.c[i][0] = c.x;
        t.c[i][1] = c.y;
        t.c[i][2] = c.z;
        t.c[i][3] = c.w;
    }
}
NVB_DEV uint4 bx_pack(const Bc7Ep &e, float err) {
    unsigned A = 0, B = 0;
    for (int k = 0; k < 4; k++) {
        A = bx_set(A, k, e.A[k]);
        B = bx_set(B, k, e.B[k]);
    }
    return make_uint4(A, B, (unsigned)(e.a_lsb | (e.b_lsb << 1)), __float_as_uint(err));
}
NVB_DEV uint4 bx_pack_idx(const int *i0, const int *i1) {
    unsigned w[4] = {0, 0, 0, 0};
    for (int i = 0; i < 16; i++) {
        w[i >> 3] |= (unsigned)(i0[i] & 15) << (4 * (i & 7));
        if (i1) w[2 + (i >> 3)] |= (unsigned)(i1[i] & 15) << (4 * (i & 7));
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}
NVB_DEV void bx_unpack_idx(const uint4 &u, int *i0, int *i1) {
    const unsigned w[4] = {u.x, u.y, u.z, u.w};
    for (int i = 0; i < 16; i++) {
        i0[i] = (int)((w[i >> 3] >> (4 * (i & 7))) & 15);
        if (i1) i1[i] = (int)((w[2 + (i >> 3)] >> (4 * (i & 7))) & 15);
    }
}
NVB_DEV void bx_unpack(const uint4 &u, Bc7Ep &e) {
    for (int k = 0; k < 4; k++) {
        e.A[k] = bx_get(u.x, k);
        e.B[k] = bx_get(u.y, k);
    }
    e.a_lsb = (int)(u.z & 1);
    e.b_lsb = (int)((u.z >> 1) & 1);
}

// ---- setup: the part of refine() before optimize_endpts -------------------------------------------------------------------
template <int M, int NCAND> __global__ void __launch_bounds__(128) k_bc7_setup(Bc7SearchParams S) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int blk = t / NCAND, rank = t % NCAND;
    if (blk >= S.nblk) return;
    Bc7Tile tile;
    bx_read_tile(S.tiles, blk, tile);
    if constexpr (M == 4 || M == 5) {
        const int nim = Bc7SplitCfg<M>::NIDXMODES, rotatemode = rank / nim, indexmode = rank % nim;
        Bc7Tile t1;
        bc7_rotate_tile(tile, rotatemode, t1);
        float ep[8];
        bc7s_rough(t1, ep);
        Bc7Ep orig;
        orig.a_lsb = orig.b_lsb = 0;
        for (int k = 0; k < 4; k++) {
            orig.A[k] = avpcl_quantize(ep[k], bc7s_prec<M>(k));
            orig.B[k] = avpcl_quantize(ep[4 + k], bc7s_prec<M>(k));
        }
        int orig_rgb[16], orig_a[16];
        const float orig_err = bc7s_assign_indices<M>(t1, rotatemode, indexmode, orig, orig_rgb, orig_a);
        bc7s_swap_indices<M>(indexmode, orig, orig_rgb, orig_a);
        S.setup[(size_t)blk * NCAND + rank] = bx_pack(orig, orig_err);
        S.setup_idx[(size_t)blk * NCAND + rank] = bx_pack_idx(orig_rgb, orig_a);
    } else {
        using C = Bc7Cfg<M>;
        const int nblocks = S.P.lv.bw * S.P.lv.bh;
        int shape = 0;
        if constexpr (C::NSH > 1) shape = S.P.shapes[((size_t)Bc7Slot<M>::v * nblocks + S.blk0 + blk) * 16 + rank];
        float ep[C::NR][8];
        bc7_rough_endpoints<M>(tile, shape, ep);
        float orig_err[C::NR];
        Bc7Ep orig[C::NR];
        int orig_idx[16];
        bc7_quantize_endpts<M>(ep, orig);
        bc7_assign_indices<M>(tile, shape, orig, orig_idx, orig_err);
        bc7_swap_indices<M>(orig, orig_idx, shape);
        for (int r = 0; r < C::NR; r++) S.setup[((size_t)blk * NCAND + rank) * C::NR + r] = bx_pack(orig[r], orig_err[r]);
        S.setup_idx[(size_t)blk * NCAND + rank] = bx_pack_idx(orig_idx, nullptr);
    }
}

// ---- finish: the part of refine() after optimize_endpts, then the reduction over a block's candidates ---------------------
template <int M, int NCAND> __global__ void __launch_bounds__(128) k_bc7_finish(Bc7SearchParams S) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int blk = t / NCAND, rank = t % NCAND;
    const int nblocks = S.P.lv.bw * S.P.lv.bh;
    float err = FLT_MAX;
    __align__(16) unsigned char out[16];
    *reinterpret_cast<uint4 *>(out) = make_uint4(0, 0, 0, 0);
    if (blk < S.nblk) {
        Bc7Tile tile;
        bx_read_tile(S.tiles, blk, tile);
        if constexpr (M == 4 || M == 5) {
            const int nim = Bc7SplitCfg<M>::NIDXMODES, rotatemode = rank / nim, indexmode = rank % nim;
            Bc7Tile t1;
            bc7_rotate_tile(tile, rotatemode, t1);
            const uint4 su = S.setup[(size_t)blk * NCAND + rank], ru = S.res[(size_t)blk * NCAND + rank];
            Bc7Ep orig, opt;
            bx_unpack(su, orig);
            int orig_rgb[16], orig_a[16], opt_rgb[16], opt_a[16];
            const float orig_err = __uint_as_float(su.w);
            opt = orig;
            const float out_err = __uint_as_float(ru.w);
            if (out_err < orig_err) bx_unpack(ru, opt);
            const float opt_err = bc7s_assign_indices<M>(t1, rotatemode, indexmode, opt, opt_rgb, opt_a);
            bc7s_swap_indices<M>(indexmode, opt, opt_rgb, opt_a);
            const float orig_tot = 0.0f + orig_err, opt_tot = 0.0f + opt_err;
            if (opt_tot < orig_tot) {
                bc7s_emit<M>(opt, opt_rgb, opt_a, rotatemode, indexmode, out);
                err = opt_tot;
            } else {
                bx_unpack_idx(S.setup_idx[(size_t)blk * NCAND + rank], orig_rgb, orig_a);
                bc7s_emit<M>(orig, orig_rgb, orig_a, rotatemode, indexmode, out);
                err = orig_tot;
            }
        } else {
            using C = Bc7Cfg<M>;
            using X = Bc7X<M>;
            int shape = 0;
            if constexpr (C::NSH > 1) shape = S.P.shapes[((size_t)Bc7Slot<M>::v * nblocks + S.blk0 + blk) * 16 + rank];
            float orig_err[C::NR], opt_err[C::NR];
            Bc7Ep orig[C::NR], opt[C::NR];
            int idx[16];
            for (int r = 0; r < C::NR; ++r) {
                const uint4 su = S.setup[((size_t)blk * NCAND + rank) * C::NR + r];
                bx_unpack(su, orig[r]);
                orig_err[r] = __uint_as_float(su.w);
                opt[r] = orig[r];
                float best_err = orig_err[r];
                for (int l = 0; l < X::NLSB; ++l) {
                    const uint4 ru = S.res[(((size_t)blk * NCAND + rank) * C::NR + r) * X::NLSB + l];
                    const float out_err = __uint_as_float(ru.w);
                    if (out_err < best_err) {
                        best_err = out_err;
                        bx_unpack(ru, opt[r]);
                    }
                }
            }
            bc7_assign_indices<M>(tile, shape, opt, idx, opt_err);
            bc7_swap_indices<M>(opt, idx, shape);
            float orig_tot = 0, opt_tot = 0;
            for (int i = 0; i < C::NR; ++i) {
                orig_tot += orig_err[i];
                opt_tot += opt_err[i];
            }
            if (opt_tot < orig_tot) {
                bc7_emit<M>(opt, shape, idx, out);
                err = opt_tot;
            } else {
                bx_unpack_idx(S.setup_idx[(size_t)blk * NCAND + rank], idx, nullptr);
                bc7_emit<M>(orig, shape, idx, out);
                err = orig_tot;
            }
        }
        if (!(err < FLT_MAX)) err = FLT_MAX;
    }
    uint4 b = *reinterpret_cast<uint4 *>(out);
    int r = rank;
    bc7_group_min<NCAND>(err, r, b);
    if (blk < S.nblk && rank == 0) {
        *reinterpret_cast<uint4 *>(S.P.cand + ((size_t)M * nblocks + S.blk0 + blk) * 16) = b;
        S.P.cand_err[(size_t)M * nblocks + S.blk0 + blk] = err;
    }
}

}  // namespace nvb
