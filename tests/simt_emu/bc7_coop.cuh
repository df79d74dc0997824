// TEST INFRASTRUCTURE ONLY (CPU emulator cross-check of the BC7 search): never compiled into the product library.
// Warp-cooperative BC7 refinement: one warp per candidate (block, mode, shape rank | rotation x index mode).
//
// Same search as bc7.cuh (and therefore as src/bc7/avpcl_mode*.cpp), re-mapped so that a warp — not a thread — walks the
// sequential endpoint search of one candidate:
//   * the texels of the region being optimised sit one per lane in groups of G = 8 or 16 lanes (G >= texel count);
//   * K = 32 / G trial endpoint pairs are evaluated at once, one per lane group (perturb_one: the -step / +step trials,
//     and endpoint A and B of the first pair of searches when K = 4; exhaustive: K consecutive (a, b) pairs);
//   * a trial's palette is built once per group in shared memory, every lane scans it for its texel without branches, the
//     texel errors are summed in texel order (fp32 addition is not associative) by a shuffle chain, and the search state
//     (endpoints, errors, accepted steps) is replicated in every lane, so control flow never diverges;
//   * the "index" arrays of the reference (compared to decide restarts) are one register per lane.
// Trial errors do not depend on the running threshold (map_colors only uses it to exit early with FLT_MAX, and a partial
// sum above the threshold implies a total above it), so evaluating K trials together and resolving them in the
// reference's order gives the same accepted steps.  Results are bit-identical to the thread-per-candidate kernels.
#pragma once
#include "../../nvidia-texture-tools_b200/csrc/kernels/bc7.cuh"

namespace nvb {

#define NVB_FULL 0xffffffffu

// endpoints of one region, 8 bits per channel (precisions are <= 8 bits)
struct CEp {
    unsigned A, B;
    int a_lsb, b_lsb;
};
NVB_DEV int cep_get(unsigned v, int ch) { return (int)((v >> (8 * ch)) & 0xFFu); }
NVB_DEV unsigned cep_set(unsigned v, int ch, int x) { return (v & ~(0xFFu << (8 * ch))) | ((unsigned)(x & 0xFF) << (8 * ch)); }

struct CoopWarp {
    float (*pal)[16][4];  // shared: [trial group][palette entry][channel]
    int G, K;             // lanes per trial group, groups per warp
    int grp, slot;        // my group and my texel slot inside it
    int np;               // texels in the region being optimised
    float c0, c1, c2, c3; // the texel of my slot (undefined when slot >= np)
};

template <int M> NVB_DEV void coop_palette_entry(const CEp &e, int j, float out[4]) {
    using C = Bc7Cfg<M>;
#pragma unroll
    for (int ch = 0; ch < C::NCH; ch++) {
        int a, b;
        if (C::LSB == 0) {
            a = avpcl_unquantize(cep_get(e.A, ch), C::PREC);
            b = avpcl_unquantize(cep_get(e.B, ch), C::PREC);
        } else {
            const int la = e.a_lsb, lb = (C::LSB == 1) ? e.a_lsb : e.b_lsb;
            a = avpcl_unquantize((cep_get(e.A, ch) << 1) | la, C::PREC + 1);
            b = avpcl_unquantize((cep_get(e.B, ch) << 1) | lb, C::PREC + 1);
        }
        out[ch] = (float)avpcl_lerp(a, b, j, C::NIDX);
    }
    if (C::NCH == 3) out[3] = 255.0f;
}

// the reference's palette scan for one texel ("stop at the first increase or at error 0"), branch-free
template <int NIDX> NVB_DEV void coop_scan(const float (*pal)[4], float c0, float c1, float c2, float c3, float *besterr_out, int *bestj_out) {
    float besterr = FLT_MAX;
    int bestj = 0;
    bool live = true;
#pragma unroll
    for (int j = 0; j < NIDX; ++j) {
        const float x = c0 - pal[j][0], y = c1 - pal[j][1], z = c2 - pal[j][2], w = c3 - pal[j][3];
        const float err = x * x + y * y + z * z + w * w;
        live = live && (besterr > 0);
        const bool stop = live && (err > besterr);
        const bool upd = live && !stop && (err < besterr);
        besterr = upd ? err : besterr;
        bestj = upd ? j : bestj;
        live = live && !stop;
    }
    *besterr_out = besterr;
    *bestj_out = bestj;
}

// map_colors for K trials at once: `e` is the trial of my group.  Returns the group's summed error (identical in all lanes
// of the group) and my texel's palette index.  Every lane of the warp must call it.
template <int M> NVB_DEV float coop_eval(const CoopWarp &W, const CEp &e, int *bestj) {
    using C = Bc7Cfg<M>;
    for (int j = W.slot; j < C::NIDX; j += W.G) coop_palette_entry<M>(e, j, W.pal[W.grp][j]);
    __syncwarp();
    float besterr;
    coop_scan<C::NIDX>(W.pal[W.grp], W.c0, W.c1, W.c2, W.c3, &besterr, bestj);
    float tot = 0;
    for (int i = 0; i < W.G; i++) {
        const float v = __shfl_sync(NVB_FULL, besterr, W.grp * W.G + i);
        if (i < W.np) tot += v;
    }
    __syncwarp();
    return tot;
}

// perturb_one for endpoint A (sel 0) and / or B (sel 1) of channel ch.  nsel = 2 runs both searches in lock step (needs
// K = 4: groups 0,1 = A -/+, groups 2,3 = B -/+); nsel = 1 runs the search of endpoint `only`.
template <int M> NVB_DEV void coop_perturb(const CoopWarp &W, int ch, const CEp &old_e, float old_err, int nsel, int only, CEp new_e[2], float min_err[2],
                                           int idx[2]) {
    using C = Bc7Cfg<M>;
    const int prec = C::PREC;
    new_e[0] = new_e[1] = old_e;
    min_err[0] = min_err[1] = old_err;
    for (int step = 1 << (prec - 1); step; step >>= 1) {
        // my group's trial
        const int s_g = (nsel == 2) ? (W.grp >> 1) : 0;       // which search my group serves
        const int which = (nsel == 2) ? s_g : only;            // 0 = endpoint A, 1 = endpoint B
        const int sign = (W.grp & 1) ? 1 : -1;
        CEp temp = new_e[s_g];
        const int v = cep_get(which ? temp.B : temp.A, ch) + sign * step;
        const bool valid = (W.grp < 2 * nsel) && !(v < 0 || v >= (1 << prec));
        if (which) temp.B = cep_set(temp.B, ch, valid ? v : 0);
        else temp.A = cep_set(temp.A, ch, valid ? v : 0);
        int bj;
        const float tot = coop_eval<M>(W, temp, &bj);
        for (int s = 0; s < nsel; s++) {
            const int g0 = 2 * s, g1 = 2 * s + 1;
            const float em = __shfl_sync(NVB_FULL, tot, g0 * W.G), ep = __shfl_sync(NVB_FULL, tot, g1 * W.G);
            const int vm = __shfl_sync(NVB_FULL, (int)valid, g0 * W.G), vp = __shfl_sync(NVB_FULL, (int)valid, g1 * W.G);
            const int jm = __shfl_sync(NVB_FULL, bj, g0 * W.G + W.slot), jp = __shfl_sync(NVB_FULL, bj, g1 * W.G + W.slot);
            bool improved = false;
            int beststep = 0;
            if (vm && em < min_err[s]) {
                improved = true;
                min_err[s] = em;
                beststep = -step;
                idx[s] = jm;
            }
            if (vp && ep < min_err[s]) {
                improved = true;
                min_err[s] = ep;
                beststep = step;
                idx[s] = jp;
            }
            if (improved) {
                const int w2 = (nsel == 2) ? s : only;
                if (w2) new_e[s].B = cep_set(new_e[s].B, ch, cep_get(new_e[s].B, ch) + beststep);
                else new_e[s].A = cep_set(new_e[s].A, ch, cep_get(new_e[s].A, ch) + beststep);
            }
        }
    }
}

// exhaustive() (+-3 window, DISABLE_EXHAUSTIVE), K consecutive (a, b) pairs per pass, resolved in the reference's order
template <int M> NVB_DEV float coop_exhaustive(const CoopWarp &W, int ch, float &orig_err, CEp &opt, int *idx_out) {
    using C = Bc7Cfg<M>;
    float best_err = orig_err;
    if (orig_err == 0) return orig_err;
    const int prec = C::PREC, delta = 3;
    const int A0 = cep_get(opt.A, ch), B0 = cep_get(opt.B, ch);
    const int alow = max(0, A0 - delta), ahigh = min((1 << prec) - 1, A0 + delta);
    const int blow = max(0, B0 - delta), bhigh = min((1 << prec) - 1, B0 + delta);
    const bool a_le_b = A0 <= B0;
    // cursor over the loop nest: (outer, inner) = (a, b) when a <= b, (b, a) otherwise
    int a, b;
    bool done = false;
    if (a_le_b) {
        a = alow;
        b = max(a, blow);
        while (!done && !(b < bhigh)) {
            ++a;
            if (a > ahigh) done = true;
            else b = max(a, blow);
        }
        if (a > ahigh) done = true;
    } else {
        b = blow;
        a = max(b, alow);
        if (!(b < bhigh)) done = true;
        while (!done && !(a <= ahigh)) {
            ++b;
            if (!(b < bhigh)) done = true;
            else a = max(b, alow);
        }
    }
    int amin = 0, bmin = 0, good = 0;
    while (!done) {
        // the next K trials in loop order
        int ta[4], tb[4];
        bool tv[4];
        for (int t = 0; t < 4; t++) {
            tv[t] = (t < W.K) && !done;
            ta[t] = a;
            tb[t] = b;
            if (tv[t]) {
                if (a_le_b) {
                    ++b;
                    while (!done && !(b < bhigh)) {
                        ++a;
                        if (a > ahigh) done = true;
                        else b = max(a, blow);
                    }
                } else {
                    ++a;
                    while (!done && !(a <= ahigh)) {
                        ++b;
                        if (!(b < bhigh)) done = true;
                        else a = max(b, alow);
                    }
                }
            }
        }
        const int g = W.grp < 4 ? W.grp : 3;
        CEp temp = opt;
        temp.A = cep_set(temp.A, ch, ta[g]);
        temp.B = cep_set(temp.B, ch, tb[g]);
        int bj;
        const float tot = coop_eval<M>(W, temp, &bj);
        for (int t = 0; t < W.K; t++) {
            const float et = __shfl_sync(NVB_FULL, tot, t * W.G);
            const int jt = __shfl_sync(NVB_FULL, bj, t * W.G + W.slot);
            if (tv[t] && et < best_err) {
                amin = ta[t];
                bmin = tb[t];
                best_err = et;
                good = jt;
            }
        }
    }
    if (best_err < orig_err) {
        opt.A = cep_set(opt.A, ch, amin);
        opt.B = cep_set(opt.B, ch, bmin);
        if (C::EXHREF) orig_err = best_err;
        *idx_out = good;
    }
    return best_err;
}

// any texel of the region whose two index registers differ?
NVB_DEV bool coop_indices_differ(const CoopWarp &W, int a, int b) {
    return __any_sync(NVB_FULL, W.grp == 0 && W.slot < W.np && a != b) != 0;
}

template <int M> NVB_DEV float coop_optimize_one(const CoopWarp &W, float orig_err, const CEp &orig, CEp &opt) {
    using C = Bc7Cfg<M>;
    float opt_err = orig_err;
    opt = orig;
    int do_b = 0;
    int orig_idx = 0, new_idx = 0;
    for (int ch = 0; ch < C::NCH; ++ch) {
        CEp ne[2], na[2], nb[2];
        float me[2], ma[2], mb[2];
        int ti[2] = {0, 0}, tia[2] = {0, 0}, tib[2] = {0, 0};
        float err0, err1;
        CEp new_a, new_b;
        int t0, t1;
        if (W.K >= 4) {
            coop_perturb<M>(W, ch, opt, opt_err, 2, 0, ne, me, ti);
            err0 = me[0]; err1 = me[1];
            new_a = ne[0]; new_b = ne[1];
            t0 = ti[0]; t1 = ti[1];
        } else {
            coop_perturb<M>(W, ch, opt, opt_err, 1, 0, na, ma, tia);
            coop_perturb<M>(W, ch, opt, opt_err, 1, 1, nb, mb, tib);
            err0 = ma[0]; err1 = mb[0];
            new_a = na[0]; new_b = nb[0];
            t0 = tia[0]; t1 = tib[0];
        }
        if (err0 < err1) {
            if (err0 >= opt_err) continue;
            new_idx = orig_idx = t0;
            opt.A = cep_set(opt.A, ch, cep_get(new_a.A, ch));
            opt_err = err0;
            do_b = 1;
        } else {
            if (err1 >= opt_err) continue;
            new_idx = orig_idx = t1;
            opt.B = cep_set(opt.B, ch, cep_get(new_b.B, ch));
            opt_err = err1;
            do_b = 0;
        }
        for (;;) {
            coop_perturb<M>(W, ch, opt, opt_err, 1, do_b, ne, me, ti);
            const float err = me[0];
            if (err >= opt_err) break;
            new_idx = ti[0];
            if (do_b == 0) opt.A = cep_set(opt.A, ch, cep_get(ne[0].A, ch));
            else opt.B = cep_set(opt.B, ch, cep_get(ne[0].B, ch));
            opt_err = err;
            do_b = 1 - do_b;
        }
        if (coop_indices_differ(W, orig_idx, new_idx)) ch = -1;
    }
    bool first = true;
    int t0 = 0;
    for (int ch = 0; ch < C::NCH; ++ch) {
        float err_arg = opt_err;
        const float new_err = coop_exhaustive<M>(W, ch, err_arg, opt, &t0);
        if (C::EXHREF) opt_err = err_arg;
        if (new_err < opt_err) {
            opt_err = new_err;
            if (first) {
                orig_idx = t0;
                first = false;
            } else if (coop_indices_differ(W, orig_idx, t0)) {
                ch = -1;
                first = true;
            }
        }
    }
    return opt_err;
}

// ---- whole-tile steps (lanes 0..15 = texels) --------------------------------------------------------------------------
template <int M> NVB_DEV void coop_quantize(const float ep[][8], CEp q[]) {
    using C = Bc7Cfg<M>;
    Bc7Ep tmp[C::NR];
    bc7_quantize_endpts<M>(ep, tmp);
    for (int r = 0; r < C::NR; r++) {
        q[r].A = q[r].B = 0;
        for (int k = 0; k < C::NCH; k++) {
            q[r].A = cep_set(q[r].A, k, tmp[r].A[k]);
            q[r].B = cep_set(q[r].B, k, tmp[r].B[k]);
        }
        q[r].a_lsb = tmp[r].a_lsb;
        q[r].b_lsb = tmp[r].b_lsb;
    }
}

// assign_indices: my texel's index (lanes 0..15) and the per-region error sums (texel order)
template <int M> NVB_DEV void coop_assign(float (*pal)[16][4], int lane, int shape, const CEp e[], float c0, float c1, float c2, float c3, int *idx,
                                          float toterr[]) {
    using C = Bc7Cfg<M>;
    for (int k = lane; k < C::NR * C::NIDX; k += 32) coop_palette_entry<M>(e[k / C::NIDX], k % C::NIDX, pal[k / C::NIDX][k % C::NIDX]);
    __syncwarp();
    const int myreg = bc7_region<C::NR>(shape, lane & 15);
    float besterr;
    coop_scan<C::NIDX>(pal[myreg], c0, c1, c2, c3, &besterr, idx);
    for (int r = 0; r < C::NR; r++) toterr[r] = 0;
    for (int i = 0; i < 16; i++) {
        const float v = __shfl_sync(NVB_FULL, besterr, i);
        const int r = bc7_region<C::NR>(shape, i);
        for (int k = 0; k < C::NR; k++)
            if (k == r) toterr[k] += v;
    }
    __syncwarp();
}

template <int M> NVB_DEV void coop_swap(CEp e[], int *idx, int lane, int shape) {
    using C = Bc7Cfg<M>;
    for (int region = 0; region < C::NR; ++region) {
        const int pos = bc7_anchor<C::NR>(shape, region);
        const int ai = __shfl_sync(NVB_FULL, *idx, pos);
        if (ai & (C::NIDX >> 1)) {
            const unsigned t = e[region].A;
            e[region].A = e[region].B;
            e[region].B = t;
            if (C::LSB == 2) {
                const int tl = e[region].a_lsb;
                e[region].a_lsb = e[region].b_lsb;
                e[region].b_lsb = tl;
            }
            if (bc7_region<C::NR>(shape, lane & 15) == region) *idx = C::NIDX - 1 - *idx;
        }
    }
}

template <int M> NVB_DEV void coop_emit(const CEp e[], int shape, int idx, int lane, unsigned char *block) {
    using C = Bc7Cfg<M>;
    Bc7Ep full[C::NR];
    for (int r = 0; r < C::NR; r++) {
        for (int k = 0; k < 4; k++) {
            full[r].A[k] = cep_get(e[r].A, k);
            full[r].B[k] = cep_get(e[r].B, k);
        }
        full[r].a_lsb = e[r].a_lsb;
        full[r].b_lsb = e[r].b_lsb;
    }
    int indices[16];
#pragma unroll
    for (int i = 0; i < 16; i++) indices[i] = __shfl_sync(NVB_FULL, idx, i);
    if (lane == 0) bc7_emit<M>(full, shape, indices, block);
}

// refine() of one candidate by one warp.  `tile` is this warp's texel tile in shared memory.
template <int M> NVB_DEV float coop_refine(const Bc7Tile &tile, float (*pal)[16][4], int lane, int shape, unsigned char *block) {
    using C = Bc7Cfg<M>;
    // rough endpoints: lane r fits region r, then everybody gets them
    float ep[C::NR][8];
    {
        float mine[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (lane < C::NR) {
            const unsigned mask = bc7_member_mask<C::NR>(shape, lane);
            if (C::NCH == 3) bc7_fit_region_rgb(tile, mask, mine);
            else bc7_fit_region_rgba(tile, mask, mine);
        }
        __syncwarp();
        for (int r = 0; r < C::NR; r++)
            for (int k = 0; k < 8; k++) ep[r][k] = __shfl_sync(NVB_FULL, mine[k], r);
    }
    const float c0 = tile.c[lane & 15][0], c1 = tile.c[lane & 15][1], c2 = tile.c[lane & 15][2], c3 = tile.c[lane & 15][3];
    CEp orig[C::NR], opt[C::NR];
    float orig_err[C::NR], opt_err[C::NR];
    int orig_idx, opt_idx;
    coop_quantize<M>(ep, orig);
    coop_assign<M>(pal, lane, shape, orig, c0, c1, c2, c3, &orig_idx, orig_err);
    coop_swap<M>(orig, &orig_idx, lane, shape);
    for (int region = 0; region < C::NR; ++region) {
        const unsigned mask = bc7_member_mask<C::NR>(shape, region);
        CoopWarp W;
        W.pal = pal;
        W.np = __popc(mask);
        W.G = (W.np <= 8) ? 8 : 16;
        W.K = 32 / W.G;
        W.grp = lane / W.G;
        W.slot = lane % W.G;
        // texel of my slot: the slot-th member of the region
        int ti = 0;
        {
            unsigned m = mask;
            for (int k = 0; k < W.slot; k++) m &= m - 1;
            ti = m ? (__ffs((int)m) - 1) : 0;
        }
        W.c0 = __shfl_sync(NVB_FULL, c0, ti);
        W.c1 = __shfl_sync(NVB_FULL, c1, ti);
        W.c2 = __shfl_sync(NVB_FULL, c2, ti);
        W.c3 = __shfl_sync(NVB_FULL, c3, ti);
        CEp temp_in = orig[region], temp_out;
        opt[region] = temp_in;
        float best_err = orig_err[region];
        if (C::LSB == 0) {
            const float out_err = coop_optimize_one<M>(W, orig_err[region], temp_in, temp_out);
            if (out_err < best_err) {
                best_err = out_err;
                opt[region] = temp_out;
            }
        } else {
            const int nlsb = (C::LSB == 1) ? 2 : 4;
            for (int lsbmode = 0; lsbmode < nlsb; ++lsbmode) {
                temp_in.a_lsb = lsbmode & 1;
                temp_in.b_lsb = (C::LSB == 1) ? 0 : (lsbmode >> 1) & 1;
                int bj;
                const float tot = coop_eval<M>(W, temp_in, &bj);
                const float in_err = __shfl_sync(NVB_FULL, tot, 0);
                const float out_err = coop_optimize_one<M>(W, in_err, temp_in, temp_out);
                if (out_err < best_err) {
                    best_err = out_err;
                    opt[region] = temp_out;
                }
            }
        }
    }
    coop_assign<M>(pal, lane, shape, opt, c0, c1, c2, c3, &opt_idx, opt_err);
    coop_swap<M>(opt, &opt_idx, lane, shape);
    float orig_tot = 0, opt_tot = 0;
    for (int i = 0; i < C::NR; ++i) {
        orig_tot += orig_err[i];
        opt_tot += opt_err[i];
    }
    if (opt_tot < orig_tot) {
        coop_emit<M>(opt, shape, opt_idx, lane, block);
        return opt_tot;
    }
    coop_emit<M>(orig, shape, orig_idx, lane, block);
    return orig_tot;
}

// ---- kernel: WPC warps per CTA, each warp = one candidate; the candidates of a block sit in one CTA ---------------------
template <int M, int NCAND> __global__ void __launch_bounds__((NCAND >= 4 ? NCAND : 4) * 32) k_bc7_refine_coop(Bc7Params P) {
    constexpr int WPC = NCAND >= 4 ? NCAND : 4;   // warps per CTA
    constexpr int BPC = WPC / NCAND;              // blocks per CTA
    __shared__ Bc7Tile s_tile[WPC];
    __shared__ float s_pal[WPC][4][16][4];
    __shared__ float s_err[WPC];
    __shared__ __align__(16) unsigned char s_blk[WPC][16];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nblocks = P.lv.bw * P.lv.bh;
    const int blk = blockIdx.x * BPC + wid / NCAND, rank = wid % NCAND;
    float err = FLT_MAX;
    if (blk < nblocks) {
        if (lane < 16) {
            const int x = (blk % P.lv.bw) * 4 + (lane & 3), y = (blk / P.lv.bw) * 4 + (lane >> 2);
            const bool in = x < P.lv.w && y < P.lv.h;
#pragma unroll
            for (int ch = 0; ch < 4; ch++) s_tile[wid].c[lane][ch] = in ? load_texel(P.lv, ch, x, y) * 255.0f : 0.0f;
        }
        __syncwarp();
        using C = Bc7Cfg<M>;
        int shape = 0;
        if constexpr (C::NSH > 1) shape = P.shapes[((size_t)Bc7Slot<M>::v * nblocks + blk) * 16 + rank];
        err = coop_refine<M>(s_tile[wid], s_pal[wid], lane, shape, s_blk[wid]);
        if (!(err < FLT_MAX)) err = FLT_MAX;
    }
    if (lane == 0) s_err[wid] = err;
    __syncthreads();
    // first strict minimum over the candidates of a block, in rank order
    if (rank == 0 && lane == 0 && blk < nblocks) {
        float best = s_err[wid];
        int bw = wid;
        for (int k = 1; k < NCAND; k++)
            if (s_err[wid + k] < best) {
                best = s_err[wid + k];
                bw = wid + k;
            }
        *reinterpret_cast<uint4 *>(P.cand + ((size_t)M * nblocks + blk) * 16) = *reinterpret_cast<const uint4 *>(s_blk[bw]);
        P.cand_err[(size_t)M * nblocks + blk] = best;
    }
}

}  // namespace nvb
