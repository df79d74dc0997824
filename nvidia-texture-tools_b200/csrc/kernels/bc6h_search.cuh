// BC6H endpoint search as a per-thread state machine: one thread per *searcher* (same idea as bc7_search.cuh).
//
// refineone / refinetwo (src/bc6h/zohone.cpp:625, zohtwo.cpp:712) pick the first mode whose delta-coded endpoints fit and
// then run optimize_one (zohtwo.cpp:595-668) on every region: a sequential perturbation search of a few hundred
// map_colors() trials.  A block has three such searches (the one-region encoding and the two regions of the best
// two-region shape).  Each is one thread here; the trial evaluation is the single convergence point of the loop, the
// searchers of the two-region kind are ordered by texel count and all of them are handed out dynamically.
//
//   k_bc6_tiles        texels of the level as [block][16] float4 (r, g, b half patterns as floats, importance)
//   k_bc6_setup        thread per (block, kind): mode loop up to the first fitting mode -> start endpoints, errors, indices
//   k_bc6_order        two-region searchers ordered by texel count (counting sort)
//   k_bc6_search<NR>   the state machine
//   k_bc6_finish       thread per (block, kind): re-assign, anchor swap, fit test, emit; then k_bc6_select as before
// k_bc6_refine (bc6h.cuh) is the same computation with one thread per (block, kind); both give identical blocks.
#pragma once
#include "bc6h.cuh"
#include "bc7_search.cuh"  // avpcl_wc: the interpolation weights are the same tables

namespace nvb {

struct Bc6SearchParams {
    Bc6Params P;
    const float4 *tiles;   // [nblocks][16]
    int4 *meta;            // [nblocks][2]   {pattern row or -1, shape, 0, 0}
    int4 *setup;           // [nblocks][3][2] searcher 0 = one-region, 1,2 = regions of the two-region shape: {A0,A1,A2,err}, {B0,B1,B2,0}
    uint2 *setup_idx;      // [nblocks][2]   start indices after the anchor swap, 4 bits per texel
    int4 *res;             // [nblocks][3][2] best endpoints
    unsigned *perm;        // [nblocks][2]   two-region searchers ordered by texel count
    unsigned *counters;    // [0] one-region hand-out, [1] two-region hand-out, [2..18] histogram, [19..35] cursors
};
#define NVB_BC6_COUNTERS 40

__global__ void __launch_bounds__(256) k_bc6_tiles(Bc6SearchParams S, float4 *tiles) {
    const int nblocks = S.P.lv.bw * S.P.lv.bh;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nblocks * 16) return;
    const int blk = t >> 4, i = t & 15;
    float c[3], imp;
    zoh_load_texel(S.P, blk % S.P.lv.bw, blk / S.P.lv.bw, i, c, &imp);
    tiles[t] = make_float4(c[0], c[1], c[2], imp);
}

// roughone for every block of the level: one thread per block (the one-region chain does not wait for the 32-shape ranking)
__global__ void __launch_bounds__(128) k_bc6_rough_one(Bc6SearchParams S) {
    const int nblocks = S.P.lv.bw * S.P.lv.bh;
    const int blk = blockIdx.x * blockDim.x + threadIdx.x;
    if (blk >= nblocks) return;
    float c[16][3], imp[16];
    for (int i = 0; i < 16; i++) {
        const float4 v = S.tiles[(size_t)blk * 16 + i];
        c[i][0] = v.x; c[i][1] = v.y; c[i][2] = v.z;
        imp[i] = v.w;
    }
    float ep1[1][6];
    zoh_rough<1>(c, imp, 0, S.P.is_signed != 0, ep1);
    float *dst = S.P.rough + (size_t)blk * 20;
    for (int k = 0; k < 6; k++) dst[k] = ep1[0][k];
}

NVB_DEV void zs_read_tile(const float4 *tiles, int blk, ZohTile &t) {
    for (int i = 0; i < 16; i++) {
        const float4 c = tiles[(size_t)blk * 16 + i];
        t.c[i][0] = c.x;
        t.c[i][1] = c.y;
        t.c[i][2] = c.z;
        t.imp[i] = c.w;
    }
}

// ---- setup: the mode loop of refineone / refinetwo up to optimize_endpts -------------------------------------------------
template <int NR> NVB_DEV void zs_setup(const Bc6SearchParams &S, int blk, const ZohTile &t, int shape, const float ep[NR][6], bool sgn) {
    constexpr int NPAT = NR == 1 ? 4 : 10;
    constexpr int ROW0 = NR == 1 ? 0 : 4;
    constexpr int KIND = NR - 1;
    float orig_err[NR];
    ZohEndpts orig[NR];
    unsigned c_orig[NR * 2][3];
    int orig_idx[16];
    int found = -1;
    for (int sp = 0; sp < NPAT && found < 0; ++sp) {
        const ZohPattern p = kZohPattern[ROW0 + sp];
        const int prec = p.prec;
        for (int r = 0; r < NR; ++r)
            for (int k = 0; k < 3; k++) {
                orig[r].A[k] = zoh_quantize(ep[r][k], prec, sgn);
                orig[r].B[k] = zoh_quantize(ep[r][3 + k], prec, sgn);
            }
        zoh_assign_indices<NR>(t, shape, orig, prec, sgn, orig_idx, orig_err);
        zoh_swap_indices<NR>(orig, orig_idx, shape);
        zoh_compress_endpts<NR>(orig, p, c_orig);
        if (zoh_endpts_fit<NR>(orig, c_orig, p, sgn)) found = sp;
    }
    S.meta[(size_t)blk * 2 + KIND] = make_int4(found, shape, 0, 0);
    if (found < 0) return;
    unsigned w0 = 0, w1 = 0;
    for (int i = 0; i < 8; i++) {
        w0 |= (unsigned)(orig_idx[i] & 15) << (4 * i);
        w1 |= (unsigned)(orig_idx[8 + i] & 15) << (4 * i);
    }
    S.setup_idx[(size_t)blk * 2 + KIND] = make_uint2(w0, w1);
    for (int r = 0; r < NR; r++) {
        int4 *dst = S.setup + ((size_t)blk * 3 + KIND + r) * 2;
        dst[0] = make_int4(orig[r].A[0], orig[r].A[1], orig[r].A[2], __float_as_int(orig_err[r]));
        dst[1] = make_int4(orig[r].B[0], orig[r].B[1], orig[r].B[2], 0);
    }
}

// thread t < padded: one-region kind of block t; t >= padded: two-region kind of block t - padded
// only_kind < 0: both kinds in one launch (padded layout above); 0 / 1: thread t = block t of that kind
__global__ void __launch_bounds__(128) k_bc6_setup(Bc6SearchParams S, int padded, int only_kind) {
    const int nblocks = S.P.lv.bw * S.P.lv.bh;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int kind = only_kind < 0 ? (t >= padded) : only_kind;
    const int blk = (only_kind < 0 && kind) ? t - padded : t;
    if (blk >= nblocks) return;
    const bool sgn = S.P.is_signed != 0;
    ZohTile tile;
    zs_read_tile(S.tiles, blk, tile);
    const float *src = S.P.rough + (size_t)blk * 20;
    if (kind == 0) {
        float ep[1][6];
        for (int k = 0; k < 6; k++) ep[0][k] = src[k];
        zs_setup<1>(S, blk, tile, 0, ep, sgn);
    } else {
        float ep[2][6];
        for (int k = 0; k < 12; k++) ep[k / 6][k % 6] = src[6 + k];
        zs_setup<2>(S, blk, tile, (int)src[18], ep, sgn);
    }
}

// ---- two-region searchers ordered by texel count ---------------------------------------------------------------------------
template <int PASS> __global__ void __launch_bounds__(256) k_bc6_order(Bc6SearchParams S) {
    __shared__ unsigned s_hist[17], s_base[17];
    const int nblocks = S.P.lv.bw * S.P.lv.bh;
    const unsigned total = (unsigned)nblocks * 2;
    const unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
    if (threadIdx.x < 17) s_hist[threadIdx.x] = 0;
    __syncthreads();
    int np = 0;
    unsigned rank = 0;
    if (e < total) {
        const int blk = (int)(e >> 1), region = (int)(e & 1);
        const int shape = S.meta[(size_t)blk * 2 + 1].y;
        const unsigned m1 = kShape2[shape] & 0xFFFFu;
        np = region ? __popc(m1) : 16 - __popc(m1);
        rank = atomicAdd(&s_hist[np], 1u);
    }
    __syncthreads();
    unsigned *hist = S.counters + 2, *cursor = S.counters + 19;
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

// ---- map_colors of one trial -------------------------------------------------------------------------------------------------
template <int NIDX> NVB_DEV float zs_eval(const float4 *px, int np, int a0, int a1, int a2, int b0, int b1, int b2, int prec, bool sgn) {
    float pal[3][NIDX];
    {
        const int ua[3] = {zoh_unquantize(a0, prec, sgn), zoh_unquantize(a1, prec, sgn), zoh_unquantize(a2, prec, sgn)};
        const int ub[3] = {zoh_unquantize(b0, prec, sgn), zoh_unquantize(b1, prec, sgn), zoh_unquantize(b2, prec, sgn)};
#pragma unroll
        for (int ch = 0; ch < 3; ch++)
#pragma unroll
            for (int j = 0; j < NIDX; ++j)
                pal[ch][j] = (float)zoh_finish_unquantize((ua[ch] * avpcl_wc(NIDX, NIDX - 1 - j) + ub[ch] * avpcl_wc(NIDX, j) + 32) >> 6, sgn);
    }
    // two palette entries per FADD2 / FMUL2; the sum of the squares is the unfused FFMA2-by-one add (see bx_eval in bc7_search.cuh)
    float2 npal[3][NIDX / 2];
#pragma unroll
    for (int ch = 0; ch < 3; ch++)
#pragma unroll
        for (int j = 0; j < NIDX / 2; ++j) npal[ch][j] = make_float2(-pal[ch][2 * j], -pal[ch][2 * j + 1]);
    float tot = 0;
    for (int i = 0; i < np; ++i) {
        const float4 c = px[i];
        const float2 cx = f2splat(c.x), cy = f2splat(c.y), cz = f2splat(c.z), imp = f2splat(c.w);
        float best = 0;
        bool live = true;
#pragma unroll
        for (int jp = 0; jp < NIDX / 2; ++jp) {
            const float2 x = f2add(cx, npal[0][jp]), y = f2add(cy, npal[1][jp]), z = f2add(cz, npal[2][jp]);
            const float2 xx = f2mul(x, x), yy = f2mul(y, y), zz = f2mul(z, z);
            const float2 n = f2add_s(f2add_s(xx, yy), zz);
            const float2 e2 = f2mul(n, imp);
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const float e = h ? e2.y : e2.x;
                if (jp == 0 && h == 0) {
                    best = e;
                } else {
                    // "stop at the first increase (or at error 0)" scan of the reference, without branches
                    const bool gt = e > best, lt = e < best;
                    live = live && !gt;
                    best = (live && lt) ? e : best;
                }
            }
        }
        tot += best;
    }
    return tot;
}

enum { ZSP_LOAD = 0, ZSP_CH_START, ZSP_PERT_EMIT, ZSP_PERT_WAIT, ZSP_PERT_FIN, ZSP_STORE, ZSP_EXIT };

template <int NR> __global__ void __launch_bounds__(128) k_bc6_search(Bc6SearchParams S) {
    constexpr int NIDX = NR == 1 ? 16 : 8;
    constexpr int ROW0 = NR == 1 ? 0 : 4;
    __shared__ float4 s_px[128 * 17];  // 16 texels + 1 of padding per thread: conflict-free 128-bit reads
    float4 *px = s_px + threadIdx.x * 17;
    const int nblocks = S.P.lv.bw * S.P.lv.bh;
    const long long total = (long long)nblocks * NR;
    const bool sgn = S.P.is_signed != 0;

    int phase = ZSP_LOAD;
    int np = 0, prec = 1;
    long long slot = 0;
    int A0 = 0, A1 = 0, A2 = 0, B0 = 0, B1 = 0, B2 = 0;  // current best endpoints ("opt")
    float opt_err = 0;
    int ch = 0;
    int pk = 0, do_b = 0, pv = 0, step = 0, sgn_ = -1, beststep = 0;
    bool improved = false;
    float pmin = 0, err0 = 0;
    int va = 0;
    int tA0 = 0, tA1 = 0, tA2 = 0, tB0 = 0, tB1 = 0, tB2 = 0;
    bool have = false;

#define NVB_ZS_GET(x0, x1, x2) (ch == 0 ? (x0) : ch == 1 ? (x1) : (x2))
#define NVB_ZS_SET(x0, x1, x2, v) \
    {                             \
        if (ch == 0) x0 = (v);    \
        else if (ch == 1) x1 = (v); \
        else x2 = (v);            \
    }
#define NVB_ZS_ADV_SIGN()                 \
    {                                     \
        if (sgn_ < 0) {                   \
            sgn_ = 1;                     \
        } else {                          \
            if (improved) pv += beststep; \
            improved = false;             \
            step >>= 1;                   \
            sgn_ = -1;                    \
        }                                 \
    }
#define NVB_ZS_TRY_PERT()                                        \
    {                                                            \
        const int v_ = pv + sgn_ * step;                         \
        if (v_ >= 0 && v_ < (1 << prec)) {                       \
            tA0 = A0; tA1 = A1; tA2 = A2; tB0 = B0; tB1 = B1; tB2 = B2; \
            if (do_b) NVB_ZS_SET(tB0, tB1, tB2, v_)              \
            else NVB_ZS_SET(tA0, tA1, tA2, v_)                   \
            phase = ZSP_PERT_WAIT;                               \
            have = true;                                         \
        } else                                                   \
            NVB_ZS_ADV_SIGN()                                    \
    }
#define NVB_ZS_BEGIN_PERT()                                                  \
    {                                                                        \
        pv = do_b ? NVB_ZS_GET(B0, B1, B2) : NVB_ZS_GET(A0, A1, A2);         \
        pmin = opt_err;                                                      \
        step = 1 << (prec - 1);                                              \
        sgn_ = -1;                                                           \
        improved = false;                                                    \
        phase = ZSP_PERT_EMIT;                                               \
        NVB_ZS_TRY_PERT()                                                    \
        if (!have) NVB_ZS_TRY_PERT()                                         \
    }

    for (;;) {
        while (!have && phase != ZSP_EXIT) {
            switch (phase) {
            case ZSP_LOAD: {
                const long long s = (long long)atomicAdd(S.counters + (NR - 1), 1u);
                if (s >= total) {
                    phase = ZSP_EXIT;
                    break;
                }
                int blk, region;
                if (NR == 1) {
                    blk = (int)s;
                    region = 0;
                } else {
                    const unsigned e = S.perm[s];
                    blk = (int)(e >> 1);
                    region = (int)(e & 1);
                }
                const int4 mt = S.meta[(size_t)blk * 2 + (NR - 1)];
                if (mt.x < 0) break;  // no mode fits ("should never happen"): nothing to optimise, next searcher
                prec = kZohPattern[ROW0 + mt.x].prec;
                slot = (long long)blk * 3 + (NR - 1) + region;
                const int4 ea = S.setup[slot * 2], eb = S.setup[slot * 2 + 1];
                A0 = ea.x; A1 = ea.y; A2 = ea.z;
                B0 = eb.x; B1 = eb.y; B2 = eb.z;
                opt_err = __int_as_float(ea.w);
                const float4 *tile = S.tiles + (size_t)blk * 16;
                np = 0;
                for (int i = 0; i < 16; i++)
                    if (zoh_region<NR>(mt.y, i) == region) px[np++] = __ldg(tile + i);
                ch = 0;
                phase = ZSP_CH_START;
                break;
            }
            case ZSP_CH_START:
                if (ch >= 3) {
                    phase = ZSP_STORE;
                } else {
                    pk = 0;
                    do_b = 0;
                    NVB_ZS_BEGIN_PERT()
                }
                break;
            case ZSP_PERT_EMIT:
                if (step == 0) phase = ZSP_PERT_FIN;
                else NVB_ZS_TRY_PERT()
                break;
            case ZSP_PERT_FIN: {
                bool again = false;
                if (pk == 0) {
                    err0 = pmin;
                    va = pv;
                    pk = 1;
                    do_b = 1;
                    again = true;
                } else if (pk == 1) {
                    const float err1 = pmin;
                    if (err0 < err1) {
                        if (!(err0 >= opt_err)) {
                            NVB_ZS_SET(A0, A1, A2, va)
                            opt_err = err0;
                            do_b = 1;
                            again = true;
                        }
                    } else {
                        if (!(err1 >= opt_err)) {
                            NVB_ZS_SET(B0, B1, B2, pv)
                            opt_err = err1;
                            do_b = 0;
                            again = true;
                        }
                    }
                    pk = 2;
                    if (!again) {
                        ++ch;
                        phase = ZSP_CH_START;
                    }
                } else {
                    if (pmin >= opt_err) {
                        ++ch;
                        phase = ZSP_CH_START;
                    } else {
                        if (do_b == 0) NVB_ZS_SET(A0, A1, A2, pv)
                        else NVB_ZS_SET(B0, B1, B2, pv)
                        opt_err = pmin;
                        do_b = 1 - do_b;
                        again = true;
                    }
                }
                if (again) NVB_ZS_BEGIN_PERT()
                break;
            }
            case ZSP_STORE:
                S.res[slot * 2] = make_int4(A0, A1, A2, 0);
                S.res[slot * 2 + 1] = make_int4(B0, B1, B2, 0);
                phase = ZSP_LOAD;
                break;
            default:
                break;
            }
        }
        // all 32 lanes vote: explicit reconvergence in front of the trial evaluation (idle lanes carry np = 0)
        if (__all_sync(0xffffffffu, phase == ZSP_EXIT)) break;
        if (phase == ZSP_EXIT) np = 0;

        const float err = zs_eval<NIDX>(px, np, tA0, tA1, tA2, tB0, tB1, tB2, prec, sgn);

        have = false;
        if (phase == ZSP_PERT_WAIT) {
            if (err < pmin) {
                improved = true;
                pmin = err;
                beststep = sgn_ * step;
            }
            NVB_ZS_ADV_SIGN()
            phase = ZSP_PERT_EMIT;
            if (step != 0) NVB_ZS_TRY_PERT()
            if (!have && step != 0) NVB_ZS_TRY_PERT()
            if (!have && step == 0) phase = ZSP_PERT_FIN;
        }
    }
#undef NVB_ZS_GET
#undef NVB_ZS_SET
#undef NVB_ZS_ADV_SIGN
#undef NVB_ZS_TRY_PERT
#undef NVB_ZS_BEGIN_PERT
}

// ---- finish: the rest of refineone / refinetwo --------------------------------------------------------------------------------
template <int NR> NVB_DEV float zs_finish(const Bc6SearchParams &S, int blk, const ZohTile &t, bool sgn, unsigned char *block) {
    constexpr int ROW0 = NR == 1 ? 0 : 4;
    constexpr int KIND = NR - 1;
    const int4 mt = S.meta[(size_t)blk * 2 + KIND];
    if (mt.x < 0) {
        *reinterpret_cast<uint4 *>(block) = make_uint4(0, 0, 0, 0);
        return FLT_MAX;
    }
    const int sp = mt.x, shape = mt.y;
    const ZohPattern p = kZohPattern[ROW0 + sp];
    const int prec = p.prec;
    float orig_err[NR], opt_err[NR];
    ZohEndpts orig[NR], opt[NR];
    unsigned c_orig[NR * 2][3], c_opt[NR * 2][3];
    int idx[16];
    for (int r = 0; r < NR; r++) {
        const size_t slot = (size_t)blk * 3 + KIND + r;
        const int4 ea = S.setup[slot * 2], eb = S.setup[slot * 2 + 1], ra = S.res[slot * 2], rb = S.res[slot * 2 + 1];
        orig[r].A[0] = ea.x; orig[r].A[1] = ea.y; orig[r].A[2] = ea.z;
        orig[r].B[0] = eb.x; orig[r].B[1] = eb.y; orig[r].B[2] = eb.z;
        orig_err[r] = __int_as_float(ea.w);
        opt[r].A[0] = ra.x; opt[r].A[1] = ra.y; opt[r].A[2] = ra.z;
        opt[r].B[0] = rb.x; opt[r].B[1] = rb.y; opt[r].B[2] = rb.z;
    }
    zoh_assign_indices<NR>(t, shape, opt, prec, sgn, idx, opt_err);
    zoh_swap_indices<NR>(opt, idx, shape);
    zoh_compress_endpts<NR>(opt, p, c_opt);
    float orig_tot = 0, opt_tot = 0;
    for (int i = 0; i < NR; ++i) {
        orig_tot += orig_err[i];
        opt_tot += opt_err[i];
    }
    if (zoh_endpts_fit<NR>(opt, c_opt, p, sgn) && opt_tot < orig_tot) {
        zoh_emit<NR>(c_opt, shape, ROW0 + sp, p, idx, block);
        return opt_tot;
    }
    const uint2 oi = S.setup_idx[(size_t)blk * 2 + KIND];
    for (int i = 0; i < 8; i++) {
        idx[i] = (int)((oi.x >> (4 * i)) & 15);
        idx[8 + i] = (int)((oi.y >> (4 * i)) & 15);
    }
    zoh_compress_endpts<NR>(orig, p, c_orig);
    zoh_emit<NR>(c_orig, shape, ROW0 + sp, p, idx, block);
    return orig_tot;
}

__global__ void __launch_bounds__(128) k_bc6_finish(Bc6SearchParams S, int padded) {
    const int nblocks = S.P.lv.bw * S.P.lv.bh;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int kind = t >= padded;
    const int blk = kind ? t - padded : t;
    if (blk >= nblocks) return;
    const bool sgn = S.P.is_signed != 0;
    ZohTile tile;
    zs_read_tile(S.tiles, blk, tile);
    unsigned char *dst = S.P.cand + ((size_t)blk * 2 + kind) * 16;
    const float err = kind == 0 ? zs_finish<1>(S, blk, tile, sgn, dst) : zs_finish<2>(S, blk, tile, sgn, dst);
    S.P.cand_err[(size_t)blk * 2 + kind] = err;
}

}  // namespace nvb
