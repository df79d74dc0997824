// BC7 block encoder — replaces, result-for-result, the reference's AVPCL compressor:
//   CompressorBC7::compressBlock             src/nvtt/CompressorDX11.cpp:80-102  (fp32 RGBA * 255, importance 1)
//   AVPCL::compress                          src/bc7/avpcl.cpp:33-46             (modes 0..7 in order, strict <)
//   compress_mode0..7 / rough / refine / optimize_endpts / optimize_one / perturb_one / exhaustive / map_colors /
//   assign_indices / swap_indices / quantize_endpts / compress_one / emit_block
//                                            src/bc7/avpcl_mode0.cpp .. avpcl_mode7.cpp
//   Utils::{lerp, quantize, unquantize, metric4, metric3, metric1}   src/bc7/avpcl_utils.cpp:27-215
//   flags as set by the caller: flag_premult = flag_nonuniform = flag_nonuniform_ati = false, DISABLE_EXHAUSTIVE,
//   USE_ZOH_INTERP(_ROUNDED) (src/bc7/avpcl.h:19-21)
//
// The eight mode files of the reference are one algorithm instantiated with different constants plus a few
// per-mode quirks; here they are one template (Bc7Cfg<M>) for modes 0,1,2,3,6,7 and one for the two "split" modes
// 4,5 (separate colour / alpha index arrays + channel rotation).  Quirks kept on purpose:
//   * modes 0 and 3 pass the running error to exhaustive() by reference (avpcl_mode0.cpp:546,631, mode3:541,626), so
//     their final restart logic never triggers;
//   * the 4-D eigen solver never tridiagonalises (see eigen.cuh);
//   * Vector4 /= is a true division while Vector3 /= multiplies by the reciprocal.
//
// Work decomposition per level:
//   k_bc7_rough<M>   (M = 0,1,2,3,7) one warp per block: lanes fit the partition shapes (PCA per region + unquantised
//                    palette error); lane 0 runs the reference's partial bubble sort; the NITEMS best shapes are stored.
//   k_bc7_refine<M>  one thread per (block, candidate): candidate = shape rank (or rotation x index mode for modes 4,5);
//                    the thread re-fits its shape and runs the sequential refine; candidates of one block sit in adjacent
//                    lanes and are reduced with (error, rank) so that the first strict minimum wins.
//   k_bc7_select     one thread per block: first strict minimum over the 8 mode results.
#pragma once
#include "../nvb_common.cuh"
#include "bc67_tables.cuh"
#include "eigen.cuh"

namespace nvb {

struct Bc7Params {
    LevelView lv;
    unsigned char *out;        // 16 bytes per block
    unsigned char *shapes;     // scratch: [5][nblocks][16] best shapes of modes 0,1,2,3,7 (slot = kBc7ShapeSlot[mode])
    unsigned char *cand;       // scratch: [8][nblocks][16] best block of each mode
    float *cand_err;           // scratch: [8][nblocks]
};

template <int M> struct Bc7Cfg;
//                                   regions shapes items shapebits channels prec lsb(0 none,1 shared,2 unique) indices exh-by-ref
#define NVB_BC7_CFG(M, NR_, NSH_, NIT_, SB_, NCH_, PREC_, LSB_, NIDX_, EXR_)                                  \
    template <> struct Bc7Cfg<M> {                                                                            \
        static constexpr int NR = NR_, NSH = NSH_, NITEMS = NIT_, SHAPEBITS = SB_, NCH = NCH_, PREC = PREC_, \
                             LSB = LSB_, NIDX = NIDX_, IBITS = (NIDX_ == 16 ? 4 : NIDX_ == 8 ? 3 : 2);        \
        static constexpr bool EXHREF = EXR_;                                                                  \
    };
NVB_BC7_CFG(0, 3, 16, 4, 4, 3, 4, 2, 8, true)
NVB_BC7_CFG(1, 2, 64, 16, 6, 3, 6, 1, 8, false)
NVB_BC7_CFG(2, 3, 64, 16, 6, 3, 5, 0, 4, false)
NVB_BC7_CFG(3, 2, 64, 16, 6, 3, 7, 2, 4, true)
NVB_BC7_CFG(6, 1, 1, 1, 0, 4, 7, 2, 16, false)
NVB_BC7_CFG(7, 2, 64, 16, 6, 4, 5, 2, 4, false)
#undef NVB_BC7_CFG

// interpolation weights out of 64 (avpcl_utils.cpp:27-28); a 4-entry palette uses every fifth 16-entry weight
NVB_TABLE int kAvpclW3[4] = {0, 21, 43, 64};
NVB_TABLE int kAvpclW7[8] = {0, 9, 18, 27, 37, 46, 55, 64};
NVB_TABLE int kAvpclW15[16] = {0, 4, 9, 13, 17, 21, 26, 30, 34, 38, 43, 47, 51, 55, 60, 64};
NVB_DEV int avpcl_weight(int nidx, int i) { return nidx == 16 ? kAvpclW15[i] : nidx == 8 ? kAvpclW7[i] : kAvpclW3[i]; }

// Utils::lerp(int...) with USE_ZOH_INTERP_ROUNDED
NVB_DEV int avpcl_lerp(int a, int b, int i, int nidx) { return (a * avpcl_weight(nidx, nidx - 1 - i) + b * avpcl_weight(nidx, i) + 32) >> 6; }
// Utils::unquantize: bit replication to 8 bits
NVB_DEV int avpcl_unquantize(int q, int prec) { return (q << (8 - prec)) | (q >> (2 * prec - 8)); }
// Utils::quantize
NVB_DEV int avpcl_quantize(float value, int prec) {
    const int unq = x86_ftoi(floorf(value + 0.5f));
    return (unq * ((1 << prec) - 1) + 127) / 255;
}

struct Bc7Tile {
    float c[16][4];
};

NVB_DEV void bc7_load_tile(const LevelView &lv, int bx, int by, Bc7Tile &t) {
    for (int i = 0; i < 16; i++) {
        const int x = bx * 4 + (i & 3), y = by * 4 + (i >> 2);
        if (x < lv.w && y < lv.h) {
#pragma unroll
            for (int ch = 0; ch < 4; ch++) t.c[i][ch] = load_texel(lv, ch, x, y) * 255.0f;
        } else {
            // texels outside the image are Vector4(0) (BlockCompressor.cpp:152-163); importance is 1 for every texel
            t.c[i][0] = t.c[i][1] = t.c[i][2] = t.c[i][3] = 0.0f;
        }
    }
}

template <int NR> NVB_DEV int bc7_region(int shape, int i) {
    if (NR == 1) return 0;
    if (NR == 2) return (kShape2[shape] >> i) & 1;
    return (kShape3[shape] >> (2 * i)) & 3;
}
template <int NR> NVB_DEV int bc7_anchor(int shape, int region) {
    if (region == 0) return 0;
    if (NR == 2) return kAnchor2[shape];
    return kAnchor3[2 * shape + region - 1];
}

NVB_DEV float bc7_metric4(const float a[4], const float b[4]) {
    const float x = a[0] - b[0], y = a[1] - b[1], z = a[2] - b[2], w = a[3] - b[3];
    return x * x + y * y + z * z + w * w;
}

// ---- rough: float endpoints per region (A = ep[r][0..3], B = ep[r][4..7]) ------------------------------------------------
NVB_DEV void bc7_clamp_rgb(float v[4]) {
    for (int k = 0; k < 3; k++) {
        if (v[k] < 0.0f) v[k] = 0.0f;
        if (v[k] > 255.0f) v[k] = 255.0f;
    }
    v[3] = 255.0f;
}
NVB_DEV void bc7_clamp_rgba(float v[4]) {
    for (int k = 0; k < 4; k++) {
        if (v[k] < 0.0f) v[k] = 0.0f;
        if (v[k] > 255.0f) v[k] = 255.0f;
    }
}

// endpoints of one region for the modes whose endpoints are RGB (alpha fixed at 255): modes 0,1,2,3
NVB_DEV void bc7_fit_region_rgb(const Bc7Tile &t, unsigned member_mask, float ep[8]) {
    int np = 0;
    float colors[16][3];
    float alphas[2] = {0.0f, 0.0f};
    float mean[4] = {0, 0, 0, 0};
    for (int i = 0; i < 16; i++)
        if ((member_mask >> i) & 1) {
            colors[np][0] = t.c[i][0];
            colors[np][1] = t.c[i][1];
            colors[np][2] = t.c[i][2];
            if (np < 2) alphas[np] = t.c[i][3];
            for (int k = 0; k < 4; k++) mean[k] += t.c[i][k];
            ++np;
        }
    float *A = ep, *B = ep + 4;
    if (np == 0) {
        A[0] = A[1] = A[2] = B[0] = B[1] = B[2] = 0.0f;
        A[3] = B[3] = 255.0f;
        return;
    }
    if (np == 1 || np == 2) {
        const int j = np - 1;
        for (int k = 0; k < 3; k++) { A[k] = colors[0][k]; B[k] = colors[j][k]; }
        A[3] = alphas[0];
        B[3] = alphas[j];
        return;
    }
    for (int k = 0; k < 4; k++) mean[k] /= (float)np;  // Vector4 /= : true division
    float dir[3];
    principal_axis3(np, colors, dir);
    float minp = FLT_MAX, maxp = -FLT_MAX;
    for (int i = 0; i < np; i++) {
        const float dp = (colors[i][0] - mean[0]) * dir[0] + (colors[i][1] - mean[1]) * dir[1] + (colors[i][2] - mean[2]) * dir[2];
        if (dp < minp) minp = dp;
        if (dp > maxp) maxp = dp;
    }
    for (int k = 0; k < 3; k++) {
        A[k] = mean[k] + dir[k] * minp;
        B[k] = mean[k] + dir[k] * maxp;
    }
    A[3] = B[3] = 255.0f;  // mean.w + 0*minp, then clamp() forces 255
    bc7_clamp_rgb(A);
    bc7_clamp_rgb(B);
}

// endpoints of one region for the RGBA modes 6,7 (4-D PCA)
NVB_DEV void bc7_fit_region_rgba(const Bc7Tile &t, unsigned member_mask, float ep[8]) {
    int np = 0;
    float colors[16][4];
    float mean[4] = {0, 0, 0, 0};
    for (int i = 0; i < 16; i++)
        if ((member_mask >> i) & 1) {
            for (int k = 0; k < 4; k++) {
                colors[np][k] = t.c[i][k];
                mean[k] += t.c[i][k];
            }
            ++np;
        }
    float *A = ep, *B = ep + 4;
    if (np == 0) {
        A[0] = A[1] = A[2] = B[0] = B[1] = B[2] = 0.0f;
        A[3] = B[3] = 255.0f;
        return;
    }
    if (np == 1 || np == 2) {
        const int j = np - 1;
        for (int k = 0; k < 4; k++) { A[k] = colors[0][k]; B[k] = colors[j][k]; }
        return;
    }
    for (int k = 0; k < 4; k++) mean[k] /= (float)np;
    float dir[4];
    principal_axis4(np, colors, dir);
    float minp = FLT_MAX, maxp = -FLT_MAX;
    for (int i = 0; i < np; i++) {
        const float dp = (colors[i][0] - mean[0]) * dir[0] + (colors[i][1] - mean[1]) * dir[1] + (colors[i][2] - mean[2]) * dir[2] +
                         (colors[i][3] - mean[3]) * dir[3];
        if (dp < minp) minp = dp;
        if (dp > maxp) maxp = dp;
    }
    for (int k = 0; k < 4; k++) {
        A[k] = mean[k] + dir[k] * minp;
        B[k] = mean[k] + dir[k] * maxp;
    }
    bc7_clamp_rgba(A);
    bc7_clamp_rgba(B);
}

template <int NR> NVB_DEV unsigned bc7_member_mask(int shape, int region) {
    unsigned m = 0;
    for (int i = 0; i < 16; i++)
        if (bc7_region<NR>(shape, i) == region) m |= 1u << i;
    return m;
}

template <int M> NVB_DEV void bc7_rough_endpoints(const Bc7Tile &t, int shape, float ep[][8]) {
    using C = Bc7Cfg<M>;
    for (int r = 0; r < C::NR; ++r) {
        const unsigned mask = bc7_member_mask<C::NR>(shape, r);
        if (C::NCH == 3) bc7_fit_region_rgb(t, mask, ep[r]);
        else bc7_fit_region_rgba(t, mask, ep[r]);
    }
}

// map_colors(tile, shape, FltEndpts): summed error against the unquantised palette (A*w[d-i] + B*w[i]) / 64
template <int M> NVB_DEV float bc7_rough_error(const Bc7Tile &t, int shape, const float ep[][8]) {
    using C = Bc7Cfg<M>;
    float toterr = 0;
    for (int i = 0; i < 16; i++) {
        const float *A = ep[bc7_region<C::NR>(shape, i)], *B = A + 4;
        float besterr = FLT_MAX;
        for (int j = 0; j < C::NIDX && besterr > 0; ++j) {
            const float wa = (float)avpcl_weight(C::NIDX, C::NIDX - 1 - j), wb = (float)avpcl_weight(C::NIDX, j);
            float pal[4];
            for (int k = 0; k < 4; k++) pal[k] = (A[k] * wa + B[k] * wb) * (1.0f / 64.0f);
            const float err = bc7_metric4(t.c[i], pal);
            if (err > besterr) break;
            if (err < besterr) besterr = err;
        }
        toterr += besterr;
    }
    return toterr;
}

// ---- quantised endpoints ------------------------------------------------------------------------------------------------
struct Bc7Ep {
    int A[4], B[4];
    int a_lsb, b_lsb;  // mode 1 uses a_lsb as the shared lsb
};

template <int M> NVB_DEV void bc7_palette(const Bc7Ep &e, float pal[][4]) {
    using C = Bc7Cfg<M>;
    for (int ch = 0; ch < C::NCH; ch++) {
        int a, b;
        if (C::LSB == 0) {
            a = avpcl_unquantize(e.A[ch], C::PREC);
            b = avpcl_unquantize(e.B[ch], C::PREC);
        } else {
            const int la = e.a_lsb, lb = (C::LSB == 1) ? e.a_lsb : e.b_lsb;
            a = avpcl_unquantize((e.A[ch] << 1) | la, C::PREC + 1);
            b = avpcl_unquantize((e.B[ch] << 1) | lb, C::PREC + 1);
        }
        for (int i = 0; i < C::NIDX; ++i) pal[i][ch] = (float)avpcl_lerp(a, b, i, C::NIDX);
    }
    if (C::NCH == 3)
        for (int i = 0; i < C::NIDX; ++i) pal[i][3] = 255.0f;
}

// map_colors(colors, np, endpts, current_err, indices): indices packed 4 bits per texel
template <int M> NVB_DEV float bc7_map_colors(const float (*colors)[4], int np, const Bc7Ep &e, float current_err, unsigned long long *indices) {
    using C = Bc7Cfg<M>;
    float pal[C::NIDX][4];
    bc7_palette<M>(e, pal);
    float toterr = 0;
    unsigned long long idx = 0;
    for (int i = 0; i < np; ++i) {
        float besterr = FLT_MAX;
        int bestj = 0;
        if (C::NIDX <= 8) {
            // same scan as the reference ("stop at the first increase, or once the error is 0"), written without branches so
            // that the lanes of a warp — which work on different candidates — stay converged
            const float c0 = colors[i][0], c1 = colors[i][1], c2 = colors[i][2], c3 = colors[i][3];
            bool live = true;
#pragma unroll
            for (int j = 0; j < C::NIDX; ++j) {
                const float x = c0 - pal[j][0], y = c1 - pal[j][1], z = c2 - pal[j][2], w = c3 - pal[j][3];
                const float err = x * x + y * y + z * z + w * w;
                live = live && (besterr > 0);
                const bool stop = live && (err > besterr);
                const bool upd = live && !stop && (err < besterr);
                besterr = upd ? err : besterr;
                bestj = upd ? j : bestj;
                live = live && !stop;
            }
        } else {
            for (int j = 0; j < C::NIDX && besterr > 0; ++j) {
                const float err = bc7_metric4(colors[i], pal[j]);
                if (err > besterr) break;
                if (err < besterr) {
                    besterr = err;
                    bestj = j;
                }
            }
        }
        idx |= (unsigned long long)bestj << (4 * i);
        toterr += besterr;
        if (toterr > current_err) return FLT_MAX;
    }
    *indices = idx;
    return toterr;
}

template <int M> NVB_DEV void bc7_assign_indices(const Bc7Tile &t, int shape, const Bc7Ep e[], int indices[16], float toterr[]) {
    using C = Bc7Cfg<M>;
    float pal[C::NR][C::NIDX][4];
    for (int r = 0; r < C::NR; ++r) {
        bc7_palette<M>(e[r], pal[r]);
        toterr[r] = 0;
    }
    for (int i = 0; i < 16; i++) {
        const int region = bc7_region<C::NR>(shape, i);
        float besterr = FLT_MAX;
        int best = 0;
        for (int j = 0; j < C::NIDX && besterr > 0; ++j) {
            const float err = bc7_metric4(t.c[i], pal[region][j]);
            if (err > besterr) break;
            if (err < besterr) {
                besterr = err;
                best = j;
            }
        }
        indices[i] = best;
        toterr[region] += besterr;
    }
}

template <int M> NVB_DEV float bc7_perturb_one(const float (*colors)[4], int np, int ch, const Bc7Ep &old_e, Bc7Ep &new_e, float old_err, int do_b,
                                               unsigned long long *indices) {
    using C = Bc7Cfg<M>;
    Bc7Ep temp = old_e;
    new_e = old_e;
    float min_err = old_err;
    int beststep = 0;
    const int prec = C::PREC;
    for (int step = 1 << (prec - 1); step; step >>= 1) {
        bool improved = false;
        for (int sign = -1; sign <= 1; sign += 2) {
            if (do_b == 0) {
                temp.A[ch] = new_e.A[ch] + sign * step;
                if (temp.A[ch] < 0 || temp.A[ch] >= (1 << prec)) continue;
            } else {
                temp.B[ch] = new_e.B[ch] + sign * step;
                if (temp.B[ch] < 0 || temp.B[ch] >= (1 << prec)) continue;
            }
            unsigned long long ti;
            const float err = bc7_map_colors<M>(colors, np, temp, min_err, &ti);
            if (err < min_err) {
                improved = true;
                min_err = err;
                beststep = sign * step;
                *indices = ti;
            }
        }
        if (improved) {
            if (do_b == 0) new_e.A[ch] += beststep;
            else new_e.B[ch] += beststep;
        }
    }
    return min_err;
}

// exhaustive() with DISABLE_EXHAUSTIVE: +-3 around the current pair, endpoint order preserved
template <int M> NVB_DEV float bc7_exhaustive(const float (*colors)[4], int np, int ch, float &orig_err, Bc7Ep &opt, unsigned long long *indices) {
    using C = Bc7Cfg<M>;
    float best_err = orig_err;
    if (orig_err == 0) return orig_err;
    const int prec = C::PREC, delta = 3;
    Bc7Ep temp = opt;
    const int alow = max(0, opt.A[ch] - delta), ahigh = min((1 << prec) - 1, opt.A[ch] + delta);
    const int blow = max(0, opt.B[ch] - delta), bhigh = min((1 << prec) - 1, opt.B[ch] + delta);
    int amin = 0, bmin = 0;
    unsigned long long good = 0, ti;
    if (opt.A[ch] <= opt.B[ch]) {
        for (int a = alow; a <= ahigh; ++a)
            for (int b = max(a, blow); b < bhigh; ++b) {
                temp.A[ch] = a;
                temp.B[ch] = b;
                const float err = bc7_map_colors<M>(colors, np, temp, best_err, &ti);
                if (err < best_err) {
                    amin = a;
                    bmin = b;
                    best_err = err;
                    good = ti;
                }
            }
    } else {
        for (int b = blow; b < bhigh; ++b)
            for (int a = max(b, alow); a <= ahigh; ++a) {
                temp.A[ch] = a;
                temp.B[ch] = b;
                const float err = bc7_map_colors<M>(colors, np, temp, best_err, &ti);
                if (err < best_err) {
                    amin = a;
                    bmin = b;
                    best_err = err;
                    good = ti;
                }
            }
    }
    if (best_err < orig_err) {
        opt.A[ch] = amin;
        opt.B[ch] = bmin;
        if (C::EXHREF) orig_err = best_err;  // modes 0 and 3 take the error by reference
        *indices = good;
    }
    return best_err;
}

template <int M> NVB_DEV float bc7_optimize_one(const float (*colors)[4], int np, float orig_err, const Bc7Ep &orig, Bc7Ep &opt) {
    using C = Bc7Cfg<M>;
    float opt_err = orig_err;
    opt = orig;
    Bc7Ep new_a, new_b, new_e;
    int do_b;
    unsigned long long orig_idx = 0, new_idx = 0, t0 = 0, t1 = 0;
    for (int ch = 0; ch < C::NCH; ++ch) {
        const float err0 = bc7_perturb_one<M>(colors, np, ch, opt, new_a, opt_err, 0, &t0);
        const float err1 = bc7_perturb_one<M>(colors, np, ch, opt, new_b, opt_err, 1, &t1);
        if (err0 < err1) {
            if (err0 >= opt_err) continue;
            new_idx = orig_idx = t0;
            opt.A[ch] = new_a.A[ch];
            opt_err = err0;
            do_b = 1;
        } else {
            if (err1 >= opt_err) continue;
            new_idx = orig_idx = t1;
            opt.B[ch] = new_b.B[ch];
            opt_err = err1;
            do_b = 0;
        }
        for (;;) {
            const float err = bc7_perturb_one<M>(colors, np, ch, opt, new_e, opt_err, do_b, &t0);
            if (err >= opt_err) break;
            new_idx = t0;
            if (do_b == 0) opt.A[ch] = new_e.A[ch];
            else opt.B[ch] = new_e.B[ch];
            opt_err = err;
            do_b = 1 - do_b;
        }
        if (orig_idx != new_idx) ch = -1;  // indices changed: start over
    }
    bool first = true;
    for (int ch = 0; ch < C::NCH; ++ch) {
        float err_arg = opt_err;
        const float new_err = bc7_exhaustive<M>(colors, np, ch, err_arg, opt, &t0);
        if (C::EXHREF) opt_err = err_arg;
        if (new_err < opt_err) {
            opt_err = new_err;
            if (first) {
                orig_idx = t0;
                first = false;
            } else if (orig_idx != t0) {
                ch = -1;
                first = true;
            }
        }
    }
    return opt_err;
}

template <int M> NVB_DEV void bc7_quantize_endpts(const float ep[][8], Bc7Ep q[]) {
    using C = Bc7Cfg<M>;
    for (int r = 0; r < C::NR; ++r) {
        q[r].a_lsb = q[r].b_lsb = 0;
        for (int k = C::NCH; k < 4; k++) q[r].A[k] = q[r].B[k] = 0;
        if (C::LSB == 0) {
            for (int k = 0; k < C::NCH; k++) {
                q[r].A[k] = avpcl_quantize(ep[r][k], C::PREC);
                q[r].B[k] = avpcl_quantize(ep[r][4 + k], C::PREC);
            }
        } else {
            // compress_one: drop the lsb of every channel, keep the majority lsb (alpha is not counted)
            int ones_a = 0, ones_b = 0;
            for (int k = 0; k < C::NCH; k++) {
                const int fa = avpcl_quantize(ep[r][k], C::PREC + 1), fb = avpcl_quantize(ep[r][4 + k], C::PREC + 1);
                if (k < 3) {
                    ones_a += fa & 1;
                    ones_b += fb & 1;
                }
                q[r].A[k] = fa >> 1;
                q[r].B[k] = fb >> 1;
            }
            if (C::LSB == 1) {
                q[r].a_lsb = (ones_a + ones_b) >= 3;
            } else {
                q[r].a_lsb = ones_a >= 2;
                q[r].b_lsb = ones_b >= 2;
            }
        }
    }
}

template <int M> NVB_DEV void bc7_swap_indices(Bc7Ep e[], int indices[16], int shape) {
    using C = Bc7Cfg<M>;
    for (int region = 0; region < C::NR; ++region) {
        const int pos = bc7_anchor<C::NR>(shape, region);
        if (indices[pos] & (C::NIDX >> 1)) {
            for (int i = 0; i < C::NCH; ++i) {
                const int t = e[region].A[i];
                e[region].A[i] = e[region].B[i];
                e[region].B[i] = t;
            }
            if (C::LSB == 2) {
                const int t = e[region].a_lsb;
                e[region].a_lsb = e[region].b_lsb;
                e[region].b_lsb = t;
            }
            for (int i = 0; i < 16; i++)
                if (bc7_region<C::NR>(shape, i) == region) indices[i] = C::NIDX - 1 - indices[i];
        }
    }
}

struct Bc7Bits {
    unsigned w[4];
    int ptr;
    NVB_DEV void init() { w[0] = w[1] = w[2] = w[3] = 0; ptr = 0; }
    NVB_DEV void write(int value, int nbits) {
        for (int i = 0; i < nbits; ++i) {
            if ((value >> i) & 1) w[ptr >> 5] |= 1u << (ptr & 31);
            ++ptr;
        }
    }
    NVB_DEV void store(unsigned char *block) const { *reinterpret_cast<uint4 *>(block) = make_uint4(w[0], w[1], w[2], w[3]); }
};

template <int M> NVB_DEV void bc7_emit(const Bc7Ep e[], int shape, const int indices[16], unsigned char *block) {
    using C = Bc7Cfg<M>;
    Bc7Bits out;
    out.init();
    out.write(1 << M, M + 1);
    out.write(shape, C::SHAPEBITS);
    for (int j = 0; j < C::NCH; ++j)
        for (int i = 0; i < C::NR; ++i) {
            out.write(e[i].A[j], C::PREC);
            out.write(e[i].B[j], C::PREC);
        }
    if (C::LSB == 1)
        for (int i = 0; i < C::NR; ++i) out.write(e[i].a_lsb, 1);
    if (C::LSB == 2)
        for (int i = 0; i < C::NR; ++i) {
            out.write(e[i].a_lsb, 1);
            out.write(e[i].b_lsb, 1);
        }
    int anchors[3] = {0, -1, -1};
    for (int r = 1; r < C::NR; r++) anchors[r] = bc7_anchor<C::NR>(shape, r);
    for (int pos = 0; pos < 16; ++pos) {
        const bool match = (pos == anchors[0]) || (pos == anchors[1]) || (pos == anchors[2]);
        out.write(indices[pos], C::IBITS - (match ? 1 : 0));
    }
    out.store(block);
}

// refine(): quantise, assign, anchor-swap, optimise every region, re-assign, keep the better of the two
template <int M> NVB_DEV float bc7_refine(const Bc7Tile &t, int shape, const float ep[][8], unsigned char *block) {
    using C = Bc7Cfg<M>;
    float orig_err[C::NR], opt_err[C::NR];
    Bc7Ep orig[C::NR], opt[C::NR];
    int orig_idx[16], opt_idx[16];
    bc7_quantize_endpts<M>(ep, orig);
    bc7_assign_indices<M>(t, shape, orig, orig_idx, orig_err);
    bc7_swap_indices<M>(orig, orig_idx, shape);
    // optimize_endpts
    for (int region = 0; region < C::NR; ++region) {
        float pixels[16][4];
        int np = 0;
        for (int i = 0; i < 16; i++)
            if (bc7_region<C::NR>(shape, i) == region) {
                for (int k = 0; k < 4; k++) pixels[np][k] = t.c[i][k];
                ++np;
            }
        Bc7Ep temp_in = orig[region], temp_out;
        opt[region] = temp_in;
        float best_err = orig_err[region];
        if (C::LSB == 0) {
            const float out_err = bc7_optimize_one<M>(pixels, np, orig_err[region], temp_in, temp_out);
            if (out_err < best_err) {
                best_err = out_err;
                opt[region] = temp_out;
            }
        } else {
            const int nlsb = (C::LSB == 1) ? 2 : 4;
            for (int lsbmode = 0; lsbmode < nlsb; ++lsbmode) {
                temp_in.a_lsb = lsbmode & 1;
                temp_in.b_lsb = (C::LSB == 1) ? 0 : (lsbmode >> 1) & 1;
                unsigned long long ti;
                const float in_err = bc7_map_colors<M>(pixels, np, temp_in, FLT_MAX, &ti);
                const float out_err = bc7_optimize_one<M>(pixels, np, in_err, temp_in, temp_out);
                if (out_err < best_err) {
                    best_err = out_err;
                    opt[region] = temp_out;
                }
            }
        }
    }
    bc7_assign_indices<M>(t, shape, opt, opt_idx, opt_err);
    bc7_swap_indices<M>(opt, opt_idx, shape);
    float orig_tot = 0, opt_tot = 0;
    for (int i = 0; i < C::NR; ++i) {
        orig_tot += orig_err[i];
        opt_tot += opt_err[i];
    }
    if (opt_tot < orig_tot) {
        bc7_emit<M>(opt, shape, opt_idx, block);
        return opt_tot;
    }
    bc7_emit<M>(orig, shape, orig_idx, block);
    return orig_tot;
}

// =====================================================================================================================
// Modes 4 and 5: one region, RGB endpoints + separate alpha endpoints, two index arrays, channel rotation.
// Mode 4: RGB 5 bits / A 6 bits, index mode 0 = 2-bit colour + 3-bit alpha indices, index mode 1 = 3-bit colour + 2-bit alpha.
// Mode 5: RGB 7 bits / A 8 bits, both index arrays 2 bits, no index-mode bit.
// =====================================================================================================================
template <int M> struct Bc7SplitCfg;
template <> struct Bc7SplitCfg<4> { static constexpr int PREC_RGB = 5, PREC_A = 6, NIDXMODES = 2; };
template <> struct Bc7SplitCfg<5> { static constexpr int PREC_RGB = 7, PREC_A = 8, NIDXMODES = 1; };
template <int M> NVB_DEV int bc7s_nidx_rgb(int indexmode) { return M == 5 ? 4 : (indexmode == 1 ? 8 : 4); }
template <int M> NVB_DEV int bc7s_nidx_a(int indexmode) { return M == 5 ? 4 : (indexmode == 1 ? 4 : 8); }
template <int M> NVB_DEV int bc7s_prec(int ch) { return ch == 3 ? Bc7SplitCfg<M>::PREC_A : Bc7SplitCfg<M>::PREC_RGB; }

NVB_DEV void bc7_rotate_tile(const Bc7Tile &in, int rotatemode, Bc7Tile &out) {
    for (int i = 0; i < 16; i++) {
        for (int k = 0; k < 4; k++) out.c[i][k] = in.c[i][k];
        if (rotatemode != 0) {
            const int ch = rotatemode - 1;  // AGBR swaps x, RABG swaps y, RGAB swaps z with w
            const float tmp = out.c[i][ch];
            out.c[i][ch] = out.c[i][3];
            out.c[i][3] = tmp;
        }
    }
}

// rough() of modes 4,5: 3-D PCA on rgb, min/max span on alpha
NVB_DEV void bc7s_rough(const Bc7Tile &t, float ep[8]) {
    float colors[16][3], alphas[16];
    float mean[4] = {0, 0, 0, 0};
    const int np = 16;
    for (int i = 0; i < 16; i++) {
        colors[i][0] = t.c[i][0];
        colors[i][1] = t.c[i][1];
        colors[i][2] = t.c[i][2];
        alphas[i] = t.c[i][3];
        for (int k = 0; k < 4; k++) mean[k] += t.c[i][k];
    }
    for (int k = 0; k < 4; k++) mean[k] /= (float)np;
    float dir[3];
    principal_axis3(np, colors, dir);
    float minp = FLT_MAX, maxp = -FLT_MAX, mina = FLT_MAX, maxa = -FLT_MAX;
    for (int i = 0; i < np; i++) {
        float dp = (colors[i][0] - mean[0]) * dir[0] + (colors[i][1] - mean[1]) * dir[1] + (colors[i][2] - mean[2]) * dir[2];
        if (dp < minp) minp = dp;
        if (dp > maxp) maxp = dp;
        dp = alphas[i] - mean[3];
        if (dp < mina) mina = dp;
        if (dp > maxa) maxa = dp;
    }
    float *A = ep, *B = ep + 4;
    for (int k = 0; k < 3; k++) {
        A[k] = mean[k] + dir[k] * minp;
        B[k] = mean[k] + dir[k] * maxp;
    }
    A[3] = mean[3] + mina;
    B[3] = mean[3] + maxa;
    bc7_clamp_rgba(A);
    bc7_clamp_rgba(B);
}

struct Bc7SplitPal {
    float rgb[8][3];
    float a[8];
};

template <int M> NVB_DEV void bc7s_palette(const Bc7Ep &e, int indexmode, Bc7SplitPal &p) {
    const int nrgb = bc7s_nidx_rgb<M>(indexmode), na = bc7s_nidx_a<M>(indexmode);
    for (int ch = 0; ch < 3; ch++) {
        const int a = avpcl_unquantize(e.A[ch], Bc7SplitCfg<M>::PREC_RGB), b = avpcl_unquantize(e.B[ch], Bc7SplitCfg<M>::PREC_RGB);
        for (int i = 0; i < nrgb; ++i) p.rgb[i][ch] = (float)avpcl_lerp(a, b, i, nrgb);
    }
    const int a = avpcl_unquantize(e.A[3], Bc7SplitCfg<M>::PREC_A), b = avpcl_unquantize(e.B[3], Bc7SplitCfg<M>::PREC_A);
    for (int i = 0; i < na; ++i) p.a[i] = (float)avpcl_lerp(a, b, i, na);
}

// best alpha index then best colour index (rotation 0) or the other way round; both orders give the same two
// independent searches when flag_premult is off, only the order of the two additions into toterr differs.
NVB_DEV void bc7s_best_pair(const float c[4], const Bc7SplitPal &p, int nrgb, int na, float *err_rgb, int *i_rgb, float *err_a, int *i_a) {
    float besterr = FLT_MAX;
    int best = 0;
    for (int j = 0; j < na && besterr > 0; ++j) {
        const float d = c[3] - p.a[j];
        const float err = d * d;
        if (err > besterr) break;
        if (err < besterr) {
            besterr = err;
            best = j;
        }
    }
    *err_a = besterr;
    *i_a = best;
    besterr = FLT_MAX;
    best = 0;
    for (int j = 0; j < nrgb && besterr > 0; ++j) {
        const float x = c[0] - p.rgb[j][0], y = c[1] - p.rgb[j][1], z = c[2] - p.rgb[j][2];
        const float err = x * x + y * y + z * z;
        if (err > besterr) break;
        if (err < besterr) {
            besterr = err;
            best = j;
        }
    }
    *err_rgb = besterr;
    *i_rgb = best;
}

// map_colors of modes 4,5.  indices[0] = colour array, indices[1] = alpha array, 4 bits per texel.
template <int M> NVB_DEV float bc7s_map_colors(const float (*colors)[4], int np, int rotatemode, int indexmode, const Bc7Ep &e, float current_besterr,
                                               unsigned long long indices[2]) {
    Bc7SplitPal p;
    bc7s_palette<M>(e, indexmode, p);
    const int nrgb = bc7s_nidx_rgb<M>(indexmode), na = bc7s_nidx_a<M>(indexmode);
    float toterr = 0;
    unsigned long long irgb = 0, ia = 0;
    for (int i = 0; i < np; ++i) {
        float er, ea;
        int jr, ja;
        bc7s_best_pair(colors[i], p, nrgb, na, &er, &jr, &ea, &ja);
        if (rotatemode == 0) {
            toterr += ea;
            toterr += er;
        } else {
            toterr += er;
            toterr += ea;
        }
        irgb |= (unsigned long long)jr << (4 * i);
        ia |= (unsigned long long)ja << (4 * i);
        if (toterr > current_besterr) return FLT_MAX;
    }
    indices[0] = irgb;
    indices[1] = ia;
    return toterr;
}

template <int M> NVB_DEV float bc7s_assign_indices(const Bc7Tile &t, int rotatemode, int indexmode, const Bc7Ep &e, int idx_rgb[16], int idx_a[16]) {
    Bc7SplitPal p;
    bc7s_palette<M>(e, indexmode, p);
    const int nrgb = bc7s_nidx_rgb<M>(indexmode), na = bc7s_nidx_a<M>(indexmode);
    float toterr = 0;
    for (int i = 0; i < 16; i++) {
        float er, ea;
        bc7s_best_pair(t.c[i], p, nrgb, na, &er, &idx_rgb[i], &ea, &idx_a[i]);
        if (rotatemode == 0) {
            toterr += ea;
            toterr += er;
        } else {
            toterr += er;
            toterr += ea;
        }
    }
    return toterr;
}

template <int M> NVB_DEV float bc7s_perturb_one(const float (*colors)[4], int np, int rotatemode, int indexmode, int ch, const Bc7Ep &old_e, Bc7Ep &new_e,
                                                float old_err, int do_b, unsigned long long indices[2]) {
    Bc7Ep temp = old_e;
    new_e = old_e;
    float min_err = old_err;
    int beststep = 0;
    const int prec = bc7s_prec<M>(ch);
    for (int step = 1 << (prec - 1); step; step >>= 1) {
        bool improved = false;
        for (int sign = -1; sign <= 1; sign += 2) {
            if (do_b == 0) {
                temp.A[ch] = new_e.A[ch] + sign * step;
                if (temp.A[ch] < 0 || temp.A[ch] >= (1 << prec)) continue;
            } else {
                temp.B[ch] = new_e.B[ch] + sign * step;
                if (temp.B[ch] < 0 || temp.B[ch] >= (1 << prec)) continue;
            }
            unsigned long long ti[2];
            const float err = bc7s_map_colors<M>(colors, np, rotatemode, indexmode, temp, min_err, ti);
            if (err < min_err) {
                improved = true;
                min_err = err;
                beststep = sign * step;
                indices[0] = ti[0];
                indices[1] = ti[1];
            }
        }
        if (improved) {
            if (do_b == 0) new_e.A[ch] += beststep;
            else new_e.B[ch] += beststep;
        }
    }
    return min_err;
}

template <int M> NVB_DEV float bc7s_exhaustive(const float (*colors)[4], int np, int rotatemode, int indexmode, int ch, float orig_err, Bc7Ep &opt,
                                               unsigned long long indices[2]) {
    float best_err = orig_err;
    if (orig_err == 0) return orig_err;
    const int prec = bc7s_prec<M>(ch), delta = 3;
    Bc7Ep temp = opt;
    const int alow = max(0, opt.A[ch] - delta), ahigh = min((1 << prec) - 1, opt.A[ch] + delta);
    const int blow = max(0, opt.B[ch] - delta), bhigh = min((1 << prec) - 1, opt.B[ch] + delta);
    int amin = 0, bmin = 0;
    unsigned long long good[2] = {0, 0}, ti[2];
    if (opt.A[ch] <= opt.B[ch]) {
        for (int a = alow; a <= ahigh; ++a)
            for (int b = max(a, blow); b < bhigh; ++b) {
                temp.A[ch] = a;
                temp.B[ch] = b;
                const float err = bc7s_map_colors<M>(colors, np, rotatemode, indexmode, temp, best_err, ti);
                if (err < best_err) {
                    amin = a;
                    bmin = b;
                    best_err = err;
                    good[0] = ti[0];
                    good[1] = ti[1];
                }
            }
    } else {
        for (int b = blow; b < bhigh; ++b)
            for (int a = max(b, alow); a <= ahigh; ++a) {
                temp.A[ch] = a;
                temp.B[ch] = b;
                const float err = bc7s_map_colors<M>(colors, np, rotatemode, indexmode, temp, best_err, ti);
                if (err < best_err) {
                    amin = a;
                    bmin = b;
                    best_err = err;
                    good[0] = ti[0];
                    good[1] = ti[1];
                }
            }
    }
    if (best_err < orig_err) {
        opt.A[ch] = amin;
        opt.B[ch] = bmin;
        indices[0] = good[0];
        indices[1] = good[1];
    }
    return best_err;
}

template <int M> NVB_DEV float bc7s_optimize_one(const float (*colors)[4], int np, int rotatemode, int indexmode, float orig_err, const Bc7Ep &orig, Bc7Ep &opt) {
    float opt_err = orig_err;
    opt = orig;
    Bc7Ep new_a, new_b, new_e;
    int do_b;
    unsigned long long orig_idx[2] = {0, 0}, new_idx[2] = {0, 0}, t0[2] = {0, 0}, t1[2] = {0, 0};
    for (int ch = 0; ch < 4; ++ch) {
        const float err0 = bc7s_perturb_one<M>(colors, np, rotatemode, indexmode, ch, opt, new_a, opt_err, 0, t0);
        const float err1 = bc7s_perturb_one<M>(colors, np, rotatemode, indexmode, ch, opt, new_b, opt_err, 1, t1);
        if (err0 < err1) {
            if (err0 >= opt_err) continue;
            new_idx[0] = orig_idx[0] = t0[0];
            new_idx[1] = orig_idx[1] = t0[1];
            opt.A[ch] = new_a.A[ch];
            opt_err = err0;
            do_b = 1;
        } else {
            if (err1 >= opt_err) continue;
            new_idx[0] = orig_idx[0] = t1[0];
            new_idx[1] = orig_idx[1] = t1[1];
            opt.B[ch] = new_b.B[ch];
            opt_err = err1;
            do_b = 0;
        }
        for (;;) {
            const float err = bc7s_perturb_one<M>(colors, np, rotatemode, indexmode, ch, opt, new_e, opt_err, do_b, t0);
            if (err >= opt_err) break;
            new_idx[0] = t0[0];
            new_idx[1] = t0[1];
            if (do_b == 0) opt.A[ch] = new_e.A[ch];
            else opt.B[ch] = new_e.B[ch];
            opt_err = err;
            do_b = 1 - do_b;
        }
        if (orig_idx[0] != new_idx[0] || orig_idx[1] != new_idx[1]) ch = -1;
    }
    bool first = true;
    for (int ch = 0; ch < 4; ++ch) {
        const float new_err = bc7s_exhaustive<M>(colors, np, rotatemode, indexmode, ch, opt_err, opt, t0);
        if (new_err < opt_err) {
            opt_err = new_err;
            if (first) {
                orig_idx[0] = t0[0];
                orig_idx[1] = t0[1];
                first = false;
            } else if (orig_idx[0] != t0[0] || orig_idx[1] != t0[1]) {
                ch = -1;
                first = true;
            }
        }
    }
    return opt_err;
}

template <int M> NVB_DEV void bc7s_swap_indices(int indexmode, Bc7Ep &e, int idx_rgb[16], int idx_a[16]) {
    const int nrgb = bc7s_nidx_rgb<M>(indexmode), na = bc7s_nidx_a<M>(indexmode);
    if (idx_rgb[0] & (nrgb >> 1)) {
        for (int i = 0; i < 3; ++i) {
            const int t = e.A[i];
            e.A[i] = e.B[i];
            e.B[i] = t;
        }
        for (int i = 0; i < 16; i++) idx_rgb[i] = nrgb - 1 - idx_rgb[i];
    }
    if (idx_a[0] & (na >> 1)) {
        const int t = e.A[3];
        e.A[3] = e.B[3];
        e.B[3] = t;
        for (int i = 0; i < 16; i++) idx_a[i] = na - 1 - idx_a[i];
    }
}

template <int M> NVB_DEV void bc7s_emit(const Bc7Ep &e, const int idx_rgb[16], const int idx_a[16], int rotatemode, int indexmode, unsigned char *block) {
    Bc7Bits out;
    out.init();
    out.write(1 << M, M + 1);
    out.write(rotatemode, 2);
    if (M == 4) out.write(indexmode, 1);
    for (int j = 0; j < 4; ++j) {
        out.write(e.A[j], bc7s_prec<M>(j));
        out.write(e.B[j], bc7s_prec<M>(j));
    }
    // the "2-bit" array first, then the "3-bit" one (both are 2 bits wide in mode 5); texel 0 drops its high bit
    const bool alpha_is_2bits = (M == 4 && indexmode == 1);
    const int *first = alpha_is_2bits ? idx_a : idx_rgb, *second = alpha_is_2bits ? idx_rgb : idx_a;
    const int bits1 = 2, bits2 = (M == 4) ? 3 : 2;
    for (int i = 0; i < 16; ++i) out.write(first[i], bits1 - (i == 0 ? 1 : 0));
    for (int i = 0; i < 16; ++i) out.write(second[i], bits2 - (i == 0 ? 1 : 0));
    out.store(block);
}

// one (rotation, index mode) candidate of compress_mode4 / compress_mode5
template <int M> NVB_DEV float bc7s_candidate(const Bc7Tile &t, int rotatemode, int indexmode, unsigned char *block) {
    Bc7Tile t1;
    bc7_rotate_tile(t, rotatemode, t1);
    float ep[8];
    bc7s_rough(t1, ep);
    Bc7Ep orig, opt;
    orig.a_lsb = orig.b_lsb = 0;
    for (int k = 0; k < 4; k++) {
        orig.A[k] = avpcl_quantize(ep[k], bc7s_prec<M>(k));
        orig.B[k] = avpcl_quantize(ep[4 + k], bc7s_prec<M>(k));
    }
    int orig_rgb[16], orig_a[16], opt_rgb[16], opt_a[16];
    const float orig_err = bc7s_assign_indices<M>(t1, rotatemode, indexmode, orig, orig_rgb, orig_a);
    bc7s_swap_indices<M>(indexmode, orig, orig_rgb, orig_a);
    Bc7Ep temp_out;
    opt = orig;
    const float out_err = bc7s_optimize_one<M>(t1.c, 16, rotatemode, indexmode, orig_err, orig, temp_out);
    if (out_err < orig_err) opt = temp_out;
    const float opt_err = bc7s_assign_indices<M>(t1, rotatemode, indexmode, opt, opt_rgb, opt_a);
    bc7s_swap_indices<M>(indexmode, opt, opt_rgb, opt_a);
    // orig_toterr / opt_toterr are 0 + err
    const float orig_tot = 0.0f + orig_err, opt_tot = 0.0f + opt_err;
    if (opt_tot < orig_tot) {
        bc7s_emit<M>(opt, opt_rgb, opt_a, rotatemode, indexmode, block);
        return opt_tot;
    }
    bc7s_emit<M>(orig, orig_rgb, orig_a, rotatemode, indexmode, block);
    return orig_tot;
}

// ---- kernels ------------------------------------------------------------------------------------------------------------
template <int M> struct Bc7Slot;  // position of the mode's shape list in Bc7Params::shapes
template <> struct Bc7Slot<0> { static constexpr int v = 0; };
template <> struct Bc7Slot<1> { static constexpr int v = 1; };
template <> struct Bc7Slot<2> { static constexpr int v = 2; };
template <> struct Bc7Slot<3> { static constexpr int v = 3; };
template <> struct Bc7Slot<7> { static constexpr int v = 4; };

#define NVB_BC7_ROUGH_WARPS 4

template <int M> __global__ void __launch_bounds__(NVB_BC7_ROUGH_WARPS * 32) k_bc7_rough(Bc7Params P, int blk_begin, int blk_end) {
    using C = Bc7Cfg<M>;
    __shared__ Bc7Tile s_tile[NVB_BC7_ROUGH_WARPS];
    __shared__ float s_mse[NVB_BC7_ROUGH_WARPS][64];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int nblocks = P.lv.bw * P.lv.bh;
    for (int blk = blk_begin + blockIdx.x * NVB_BC7_ROUGH_WARPS + wib; blk < blk_end; blk += gridDim.x * NVB_BC7_ROUGH_WARPS) {
        __syncwarp();
        if (lane < 16) {
            const int x = (blk % P.lv.bw) * 4 + (lane & 3), y = (blk / P.lv.bw) * 4 + (lane >> 2);
            const bool in = x < P.lv.w && y < P.lv.h;
#pragma unroll
            for (int ch = 0; ch < 4; ch++) s_tile[wib].c[lane][ch] = in ? load_texel(P.lv, ch, x, y) * 255.0f : 0.0f;
        }
        __syncwarp();
        for (int s = lane; s < C::NSH; s += 32) {
            float ep[C::NR][8];
            bc7_rough_endpoints<M>(s_tile[wib], s, ep);
            s_mse[wib][s] = bc7_rough_error<M>(s_tile[wib], s, ep);
        }
        __syncwarp();
        if (lane == 0) {
            // "bubble sort -- only need to bubble up the first NITEMS items" (avpcl_mode1.cpp:1029-1033), verbatim order
            unsigned char index[64];
            float *mse = s_mse[wib];
            for (int i = 0; i < C::NSH; ++i) index[i] = (unsigned char)i;
            for (int i = 0; i < C::NITEMS; ++i)
                for (int j = i + 1; j < C::NSH; ++j)
                    if (mse[i] > mse[j]) {
                        const float tf = mse[i]; mse[i] = mse[j]; mse[j] = tf;
                        const unsigned char ti = index[i]; index[i] = index[j]; index[j] = ti;
                    }
            unsigned char *dst = P.shapes + ((size_t)Bc7Slot<M>::v * nblocks + blk) * 16;
            for (int i = 0; i < C::NITEMS; ++i) dst[i] = index[i];
        }
    }
}

// (error, rank) reduction inside groups of GROUP adjacent lanes; every lane of the warp must call it
template <int GROUP> NVB_DEV void bc7_group_min(float &err, int &rank, uint4 &blk) {
#pragma unroll
    for (int d = GROUP >> 1; d >= 1; d >>= 1) {
        const float oe = __shfl_xor_sync(0xffffffffu, err, d);
        const int orank = __shfl_xor_sync(0xffffffffu, rank, d);
        uint4 ob;
        ob.x = __shfl_xor_sync(0xffffffffu, blk.x, d);
        ob.y = __shfl_xor_sync(0xffffffffu, blk.y, d);
        ob.z = __shfl_xor_sync(0xffffffffu, blk.z, d);
        ob.w = __shfl_xor_sync(0xffffffffu, blk.w, d);
        if (oe < err || (oe == err && orank < rank)) {
            err = oe;
            rank = orank;
            blk = ob;
        }
    }
}

#ifdef NVB_EMU  // cross-check design for the CPU emulator tests only (tests/simt_emu): not part of the product library
// One thread per (block, candidate).  NCAND candidates of a block are adjacent lanes.
template <int M, int NCAND> __global__ void __launch_bounds__(128) k_bc7_refine(Bc7Params P) {
    const int nblocks = P.lv.bw * P.lv.bh;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int blk = t / NCAND, rank = t % NCAND;
    float err = FLT_MAX;  // a candidate that is not strictly below FLT_MAX never wins in the reference either
    __align__(16) unsigned char out[16];
    *reinterpret_cast<uint4 *>(out) = make_uint4(0, 0, 0, 0);
    if (blk < nblocks) {
        Bc7Tile tile;
        bc7_load_tile(P.lv, blk % P.lv.bw, blk / P.lv.bw, tile);
        if constexpr (M == 4 || M == 5) {
            const int nim = Bc7SplitCfg<M>::NIDXMODES;
            err = bc7s_candidate<M>(tile, rank / nim, rank % nim, out);
        } else {
            using C = Bc7Cfg<M>;
            int shape = 0;
            if constexpr (C::NSH > 1) shape = P.shapes[((size_t)Bc7Slot<M>::v * nblocks + blk) * 16 + rank];
            float ep[C::NR][8];
            bc7_rough_endpoints<M>(tile, shape, ep);
            err = bc7_refine<M>(tile, shape, ep, out);
        }
        if (!(err < FLT_MAX)) err = FLT_MAX;
    }
    uint4 b = *reinterpret_cast<uint4 *>(out);
    int r = rank;
    bc7_group_min<NCAND>(err, r, b);
    if (blk < nblocks && rank == 0) {
        *reinterpret_cast<uint4 *>(P.cand + ((size_t)M * nblocks + blk) * 16) = b;
        P.cand_err[(size_t)M * nblocks + blk] = err;
    }
}
#endif  // NVB_EMU

__global__ void __launch_bounds__(256) k_bc7_select(Bc7Params P) {
    const int nblocks = P.lv.bw * P.lv.bh;
    for (int blk = blockIdx.x * blockDim.x + threadIdx.x; blk < nblocks; blk += gridDim.x * blockDim.x) {
        float best = FLT_MAX;
        int bm = -1;
        for (int m = 0; m < 8; m++) {
            const float e = P.cand_err[(size_t)m * nblocks + blk];
            if (e < best) {
                best = e;
                bm = m;
            }
        }
        uint4 v = make_uint4(0, 0, 0, 0);
        if (bm >= 0) v = *reinterpret_cast<const uint4 *>(P.cand + ((size_t)bm * nblocks + blk) * 16);
        *reinterpret_cast<uint4 *>(P.out + nvb_out_block(P.lv, blk) * 16) = v;
    }
}

}  // namespace nvb
