// BC1 block encoder: ICBC v1.05 semantics (scalar code path), 16 lanes per 4x4 block.
//
// Replaces, bit-exactly at Quality_Fastest (Level 1) / Normal + Highest (Level 8) / Production (Level 9):
//   FloatColorCompressor task (gather, zero padding, weights)   src/nvtt/BlockCompressor.cpp:116-166
//   CompressorDXT1::compressBlock -> icbc::compress_dxt1         src/nvtt/BlockCompressor.cpp:211-222, icbc.h:3485-3559
//   reduce_colors / skip_blacks / is_black                      src/nvtt/icbc.h:1622-1702
//   computeCentroid / computeCovariance / power method          src/nvtt/icbc.h:1709-1799
//   compute_sat                                                 src/nvtt/icbc.h:1812-1882
//   cluster tables (968 four-cluster / 152 three-cluster)       src/nvtt/icbc.h:1894-1975 (regenerated in host_tables.h)
//   cluster_fit_three / cluster_fit_four (scalar lanes)         src/nvtt/icbc.h:2016-2247 / 2250-2549
//   vector3_to_color16, midpoint rounding                       src/nvtt/icbc.h:1516-1526,1580-1595
//   evaluate_palette (D3D10), evaluate_mse                      src/nvtt/icbc.h:2558-2585,2687-2760
//   compute_indices4 / compute_indices3 / compute_indices       src/nvtt/icbc.h:2803-2941
//   output_block3 / output_block4                               src/nvtt/icbc.h:2944-2980
//   optimize_end_points4, fit_colors_bbox, select_diagonal, inset_bbox   src/nvtt/icbc.h:2984-3016,3096-3151
//   compress_dxt1_single_color_optimal                          src/nvtt/icbc.h:3274-3289
//   compress_dxt1_cluster_fit, refine_endpoints                 src/nvtt/icbc.h:3290-3406
//
// Mapping: a half-warp owns a block, lane l owns texel l.  Per-texel work (distances, indices) is lane-parallel;
// every fp32 reduction whose order matters is evaluated in the reference's order (either redundantly in all
// lanes from shared memory, or as an ordered 16-step shuffle sum); cluster splits are striped over the lanes
// and reduced with (error, split number) so the first strict minimum of the sequential search wins.
#pragma once
#include "../nvb_common.cuh"

namespace nvb {

struct Bc1Params {
    LevelView lv;
    unsigned char *out;   // 8 bytes per block
    int out_stride, out_offset;
    int level;            // ICBC quality level: 1 (Fastest), 8 (Normal/Highest), 9 (Production)
    int transparency;     // AlphaMode_Transparency: weight = saturate(alpha)
    float cw[3];          // colour weights
    const unsigned short *four;   // 968 packed cumulative splits c0 | c1<<5 | c2<<10
    const unsigned short *three;  // 152 packed cumulative splits c0 | c1<<5
    const int *four_total;        // [16]
    const int *three_total;       // [16]
    const float *midpoints5;      // [32]
    const float *midpoints6;      // [64]
    const unsigned char *match5;  // [256][2] ICBC single-colour tables
    const unsigned char *match6;  // [256][2]
    int rgbm = 0;                 // BC3-RGBM colour block: texels become (R,G,B)/M, weights w*M, no 3-colour mode
    float rgbm_min = 0.15f;       //   M = max(R, G, B, rgbm_min)   (CompressorDXT5_RGBM.cpp:23-49)
};

#define NVB_BC1_GROUPS 8
#ifndef NVB_BC1_MINB
#define NVB_BC1_MINB 9  // resident CTAs per SM the register budget is set for (56 registers)
#endif

struct Bc1GroupSmem {
    float4 pts[16];   // reduced colour set (x,y,z,weight)
    float4 sat[17];   // summed-area table along the principal axis: sat[0] = 0, sat[k + 1] = sum of the first k + 1 ordered points
                      // (the split tables store index + 1 with 0 = "nothing", so a split reads sat[field] without a test)
};

struct Bc1Block {
    unsigned c0, c1;  // Color16.u
    unsigned indices;
};

NVB_DEV float icbc_saturate(float x) { return nv_min(nv_max(x, 0.0f), 1.0f); }

// ordered sum over the 16 lanes of the group: ((0 + t0) + t1) + ... + t15
NVB_DEV float group_ordered_sum(unsigned gm, float t) {
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < 16; j++) s += __shfl_sync(gm, t, j, 16);
    return s;
}

NVB_DEV unsigned icbc_vector3_to_color16(const Bc1Params &P, float x, float y, float z) {
    // nv_clamp maps NaN to the lower bound: the operands are always inside [0, 63], plain truncation is the x86 conversion
    unsigned r = (unsigned)__float2int_rz(nv_clamp(x * 31.0f, 0.0f, 31.0f));
    unsigned g = (unsigned)__float2int_rz(nv_clamp(y * 63.0f, 0.0f, 63.0f));
    unsigned b = (unsigned)__float2int_rz(nv_clamp(z * 31.0f, 0.0f, 31.0f));
    r += (x > P.midpoints5[r]) ? 1u : 0u;
    g += (y > P.midpoints6[g]) ? 1u : 0u;
    b += (z > P.midpoints5[b]) ? 1u : 0u;
    return ((r << 11) | (g << 5) | b) & 0xFFFFu;
}

// D3D10 palette as 8-bit channels: pal[i] = r | g<<8 | b<<16 (| 0xFF<<24 unless transparent black)
NVB_DEV void icbc_palette32(unsigned c0, unsigned c1, unsigned pr[4], unsigned pg[4], unsigned pb[4]) {
    const unsigned r0 = (c0 >> 11) & 31, g0 = (c0 >> 5) & 63, b0 = c0 & 31;
    const unsigned r1 = (c1 >> 11) & 31, g1 = (c1 >> 5) & 63, b1 = c1 & 31;
    pr[0] = (r0 << 3) | (r0 >> 2); pg[0] = (g0 << 2) | (g0 >> 4); pb[0] = (b0 << 3) | (b0 >> 2);
    pr[1] = (r1 << 3) | (r1 >> 2); pg[1] = (g1 << 2) | (g1 >> 4); pb[1] = (b1 << 3) | (b1 >> 2);
    if (c0 > c1) {
        pr[2] = (2 * pr[0] + pr[1]) / 3; pg[2] = (2 * pg[0] + pg[1]) / 3; pb[2] = (2 * pb[0] + pb[1]) / 3;
        pr[3] = (2 * pr[1] + pr[0]) / 3; pg[3] = (2 * pg[1] + pg[0]) / 3; pb[3] = (2 * pb[1] + pb[0]) / 3;
    } else {
        pr[2] = (pr[0] + pr[1]) / 2; pg[2] = (pg[0] + pg[1]) / 2; pb[2] = (pb[0] + pb[1]) / 2;
        pr[3] = 0; pg[3] = 0; pb[3] = 0;
    }
}

struct Pal3 {
    float x[4], y[4], z[4];
};
// float(v) / 255.0f for an integer v in [0, 255], correctly rounded: q = v * r with r = fl(1 / 255), one Newton step on the
// remainder - the IEEE division's own fast path without its range check and slow-path call (4 instructions instead of 12;
// the palette is rebuilt for every refinement candidate).  Equal to the division for all 256 inputs: tests/test_capi_boundary.py.
NVB_DEV float icbc_u8_to_float(unsigned v) {
    const float x = (float)v, r = __uint_as_float(0x3b808081u);
    const float q = __fmul_rn(x, r);
    return __fmaf_rn(r, __fmaf_rn(q, -255.0f, x), q);
}
NVB_DEV Pal3 icbc_palette_f(unsigned c0, unsigned c1) {
    unsigned pr[4], pg[4], pb[4];
    icbc_palette32(c0, c1, pr, pg, pb);
    Pal3 p;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        p.x[i] = icbc_u8_to_float(pr[i]);
        p.y[i] = icbc_u8_to_float(pg[i]);
        p.z[i] = icbc_u8_to_float(pb[i]);
    }
    return p;
}

NVB_DEV float dist2w(float cx, float cy, float cz, float px, float py, float pz) {
    // vlen2(vc - vp): both already multiplied by the colour weights
    const float dx = cx - px, dy = cy - py, dz = cz - pz;
    return dx * dx + dy * dy + dz * dz;
}

// U = the colour weights are (1, 1, 1), the reference's default: x * 1.0f is x bit for bit, so the kernels instantiated with
// U leave those multiplications out (3 per split, 3 per texel of every refinement candidate).
template <bool U> NVB_DEV float cw_mul(float x, float w) { return U ? x : x * w; }

// per-texel squared error term of evaluate_mse: |(p - c) * w * 255|^2
template <bool U> NVB_DEV float mse_term(float px, float py, float pz, float cx, float cy, float cz, const float cw[3]) {
    const float dx = cw_mul<U>(px - cx, cw[0]) * 255.0f;
    const float dy = cw_mul<U>(py - cy, cw[1]) * 255.0f;
    const float dz = cw_mul<U>(pz - cz, cw[2]) * 255.0f;
    return dx * dx + dy * dy + dz * dz;
}

NVB_DEV unsigned group_gather_bits2(unsigned gm, unsigned idx, int l) {
    unsigned bits = (idx & 3u) << (2 * l);
#pragma unroll
    for (int d = 8; d >= 1; d >>= 1) bits |= __shfl_xor_sync(gm, bits, d, 16);
    return bits;
}

NVB_DEV float select4(const float v[4], unsigned i) { return (i == 0) ? v[0] : (i == 1) ? v[1] : (i == 2) ? v[2] : v[3]; }

// output_block4 / output_block3: endpoints -> 565, palette, per-texel index, weighted MSE (ordered sum).
// NVB_BC1_CALLS: the two big helpers as real calls instead of 2-3 inlined copies each.  Instruction-cache experiment (ncu shows
// 2.2 warps per issue stalled on "no instruction"): the Level 9 kernel shrinks from 5448 to 4408 instructions but gets slower
// (8192² Production 42.0 -> 43.7 ms), and so does every register budget tried with it (NVB_BC1_MINB 8 / 10: 43.2 / 44.2 ms;
// profiles/r2g_bc1_variants.txt) - so the inlined copies and 9 CTAs per SM stay.
#ifdef NVB_BC1_CALLS
#define NVB_BC1_HELPER NVB_DEV_CALL
#else
#define NVB_BC1_HELPER NVB_DEV
#endif
template <bool U>
NVB_BC1_HELPER float icbc_output_block(const Bc1Params &P, unsigned gm, int l, bool four, bool allow_black, float sx, float sy, float sz,
                                float ex, float ey, float ez, float cx, float cy, float cz, float wt, Bc1Block *blk) {
    unsigned color0 = icbc_vector3_to_color16(P, sx, sy, sz);
    unsigned color1 = icbc_vector3_to_color16(P, ex, ey, ez);
    if (four ? (color0 < color1) : (color0 > color1)) {
        const unsigned t = color0; color0 = color1; color1 = t;
    }
    const Pal3 pal = icbc_palette_f(color0, color1);
    const float vcx = cw_mul<U>(cx, P.cw[0]), vcy = cw_mul<U>(cy, P.cw[1]), vcz = cw_mul<U>(cz, P.cw[2]);
    float d[4];
    unsigned idx;
    if (four) {
#pragma unroll
        for (int i = 0; i < 4; i++)
            d[i] = dist2w(vcx, vcy, vcz, cw_mul<U>(pal.x[i], P.cw[0]), cw_mul<U>(pal.y[i], P.cw[1]), cw_mul<U>(pal.z[i], P.cw[2]));
        const bool b1 = d[1] > d[2], b2 = d[0] > d[2];
        bool x0 = b1 && b2;
        const bool b0 = d[0] > d[3], b3 = d[1] > d[3];
        x0 = x0 || (b0 && b3);
        const bool b4 = d[2] > d[3];
        const bool x1 = b0 && b4;
        idx = (x1 ? 1u : 0u) | (x0 ? 2u : 0u);  // interleave(indices1, indices0): x1 -> bit 0, x0 -> bit 1
    } else {
        // note the operand order vp - vc here (compute_indices3)
#pragma unroll
        for (int i = 0; i < 3; i++)
            d[i] = dist2w(cw_mul<U>(pal.x[i], P.cw[0]), cw_mul<U>(pal.y[i], P.cw[1]), cw_mul<U>(pal.z[i], P.cw[2]), vcx, vcy, vcz);
        const bool i1 = d[1] < d[2];
        const bool i2 = (d[2] <= d[0]) && (d[2] <= d[1]);
        bool i3 = false;
        if (allow_black) {
            const float d3 = vcx * vcx + vcy * vcy + vcz * vcz;
            i3 = (d3 <= d[0]) && (d3 <= d[1]) && (d3 <= d[2]);
        }
        // interleave(indices1, indices0): indices1 = i1|i3 goes to bit 0, indices0 = i2|i3 to bit 1
        idx = ((i1 || i3) ? 1u : 0u) | ((i2 || i3) ? 2u : 0u);
    }
    blk->c0 = color0;
    blk->c1 = color1;
    blk->indices = group_gather_bits2(gm, idx, l);
    const float t = wt * mse_term<U>(select4(pal.x, idx), select4(pal.y, idx), select4(pal.z, idx), cx, cy, cz, P.cw);
    return group_ordered_sum(gm, t);
}

// compute_sat on S.pts[0..n): PCA ordering + summed area table in S.sat.  All lanes of the group call it.
NVB_BC1_HELPER void icbc_compute_sat(Bc1GroupSmem &S, unsigned gm, int l, int n) {
    // centroid
    float total = 0.0f, cx = 0.0f, cy = 0.0f, cz = 0.0f;
    for (int i = 0; i < n; i++) {
        const float4 p = S.pts[i];
        total += p.w;
        cx += p.x * p.w;
        cy += p.y * p.w;
        cz += p.z * p.w;
    }
    {
        const float t = 1.0f / total;
        cx *= t; cy *= t; cz *= t;
    }
    float m0 = 0, m1 = 0, m2 = 0, m3 = 0, m4 = 0, m5 = 0;
    for (int i = 0; i < n; i++) {
        const float4 p = S.pts[i];
        const float ax = p.x - cx, ay = p.y - cy, az = p.z - cz;
        const float bx = ax * p.w, by = ay * p.w, bz = az * p.w;
        m0 += ax * bx;
        m1 += ax * by;
        m2 += ax * bz;
        m3 += ay * by;
        m4 += ay * bz;
        m5 += az * bz;
    }
    float vx = 0.0f, vy = 0.0f, vz = 0.0f;
    if (!(m0 == 0 && m3 == 0 && m5 == 0)) {
        const float r0 = m0 * m0 + m1 * m1 + m2 * m2;
        const float r1 = m1 * m1 + m3 * m3 + m4 * m4;
        const float r2 = m2 * m2 + m4 * m4 + m5 * m5;
        if (r0 > r1 && r0 > r2) { vx = m0; vy = m1; vz = m2; }
        else if (r1 > r2) { vx = m1; vy = m3; vz = m4; }
        else { vx = m2; vy = m4; vz = m5; }
        for (int it = 0; it < 8; it++) {
            const float x = vx * m0 + vy * m1 + vz * m2;
            const float y = vx * m1 + vy * m3 + vz * m4;
            const float z = vx * m2 + vy * m4 + vz * m5;
            const float norm = nv_max(nv_max(x, y), z);
            const float inv = 1.0f / norm;
            vx = x * inv; vy = y * inv; vz = z * inv;
        }
    }
    // order by projection: stable insertion sort == rank by (dps, index)
    float dps = 0.0f;
    float4 mine = make_float4(0.f, 0.f, 0.f, 0.f);
    if (l < n) {
        mine = S.pts[l];
        dps = mine.x * vx + mine.y * vy + mine.z * vz;
    }
    int rank = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const float dj = __shfl_sync(gm, dps, j, 16);
        if (j < n && (dj < dps || (dj == dps && j < l))) rank++;
    }
    __syncwarp(gm);  // everyone is done reading S.sat from a previous call
    if (l < n) S.sat[rank + 1] = make_float4(mine.x * mine.w, mine.y * mine.w, mine.z * mine.w, mine.w);
    __syncwarp(gm);
    // inclusive prefix sums, sequentially (lane 0), in the reference's order
    if (l == 0) {
        S.sat[0] = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 acc = S.sat[1];
        for (int i = 2; i <= n; i++) {
            const float4 s = S.sat[i];
            acc.x = acc.x + s.x; acc.y = acc.y + s.y; acc.z = acc.z + s.z; acc.w = acc.w + s.w;
            S.sat[i] = acc;
        }
    }
    __syncwarp(gm);
}

// refine_endpoints' table of endpoint moves (icbc.h:3329-3348), one word per channel: field k = deltas[k][channel] + 1
constexpr unsigned icbc_delta_word(int ch) {
    constexpr int deltas[16][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {-1, 0, 0}, {0, -1, 0}, {0, 0, -1}, {1, 1, 0}, {1, 0, 1},
                                   {0, 1, 1}, {-1, -1, 0}, {-1, 0, -1}, {0, -1, -1}, {-1, 1, 0}, {1, -1, 0}, {0, -1, 1}, {0, 1, -1}};
    unsigned w = 0;
    for (int k = 0; k < 16; k++) w |= (unsigned)(deltas[k][ch] + 1) << (2 * k);
    return w;
}
constexpr unsigned DWR = icbc_delta_word(0), DWG = icbc_delta_word(1), DWB = icbc_delta_word(2);

struct FitResult {
    float sx, sy, sz, ex, ey, ez;
};

// float(int(v)) with v = saturate(x) * grid + 0.5: v lies in [0.5, grid + 0.5] (the saturate maps NaN to 0), so the x86
// conversion's out-of-range cases cannot occur and the round trip through int is exactly truncf (one FRND.TRUNC)
NVB_DEV float icbc_round5(float x) { return truncf(icbc_saturate(x) * 31.0f + 0.5f) * (1.0f / 31.0f); }
NVB_DEV float icbc_round6(float x) { return truncf(icbc_saturate(x) * 63.0f + 0.5f) * (1.0f / 63.0f); }

// One split of cluster_fit_four (FOUR) or cluster_fit_three.  Returns the error; a/b = snapped endpoints.
template <bool FOUR, bool U>
NVB_DEV float icbc_eval_split(const float4 *sat, unsigned pk, float4 sum, const float msq[3], float a[3], float b[3]) {
    const int c0 = (int)(pk & 31), c1 = (int)((pk >> 5) & 31), c2 = (int)((pk >> 10) & 31);  // index + 1; sat[0] = 0
    const float4 s0 = sat[c0];
    const float4 s1 = sat[c1];
    float alpha2_sum, beta2_sum, alphabeta_sum;
    float ax[3];
    if (FOUR) {
        const float4 s2 = sat[c2];
        const float w3 = sum.w - s2.w;
        const float x2[3] = {s2.x - s1.x, s2.y - s1.y, s2.z - s1.z};
        const float x1[3] = {s1.x - s0.x, s1.y - s0.y, s1.z - s0.z};
        const float x0[3] = {s0.x, s0.y, s0.z};
        const float w2 = s2.w - s1.w;
        const float w1 = s1.w - s0.w;
        const float w0 = s0.w;
        alpha2_sum = w2 * (1.0f / 9.0f) + (w1 * (4.0f / 9.0f) + w0);
        beta2_sum = w1 * (1.0f / 9.0f) + (w2 * (4.0f / 9.0f) + w3);
        alphabeta_sum = (w1 + w2) * (2.0f / 9.0f);
#pragma unroll
        for (int k = 0; k < 3; k++) ax[k] = x2[k] * (1.0f / 3.0f) + (x1[k] * (2.0f / 3.0f) + x0[k]);
    } else {
        const float w2 = sum.w - s1.w;
        const float x1[3] = {s1.x - s0.x, s1.y - s0.y, s1.z - s0.z};
        const float x0[3] = {s0.x, s0.y, s0.z};
        const float w1 = s1.w - s0.w;
        const float w0 = s0.w;
        alphabeta_sum = w1 * 0.25f;
        alpha2_sum = w0 + alphabeta_sum;
        beta2_sum = w2 + alphabeta_sum;
#pragma unroll
        for (int k = 0; k < 3; k++) ax[k] = x0[k] + x1[k] * 0.5f;
    }
    const float factor = 1.0f / (alpha2_sum * beta2_sum - alphabeta_sum * alphabeta_sum);
    const float S[3] = {sum.x, sum.y, sum.z};
    float e1[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float alphax = ax[k];
        const float betax = S[k] - alphax;
        float av = (alphax * beta2_sum - betax * alphabeta_sum) * factor;
        float bv = (betax * alpha2_sum - alphax * alphabeta_sum) * factor;
        av = (k == 1) ? icbc_round6(av) : icbc_round5(av);
        bv = (k == 1) ? icbc_round6(bv) : icbc_round5(bv);
        const float e2 = (av * (bv * alphabeta_sum - alphax) - bv * betax) * 2.0f;
        e1[k] = (av * av) * alpha2_sum + ((bv * bv) * beta2_sum + e2);
        a[k] = av;
        b[k] = bv;
    }
    return cw_mul<U>(e1[0], msq[0]) + cw_mul<U>(e1[1], msq[1]) + cw_mul<U>(e1[2], msq[2]);
}

// Two splits at once (A, B) with packed fp32 (FMUL2 / FADD2 / unfused FFMA2-by-one sums), operation for operation the
// arithmetic of icbc_eval_split; only the errors are returned (the winner is re-evaluated by icbc_eval_split for its
// endpoints).  The pairing follows the registers an LDS.128 of a SAT row fills - (x, y) and (z, w) are aligned register
// pairs already - so nothing has to be moved to form a pair:
//   * the differences s2 - s1 and s1 - s0 are two FADD2 each;
//   * alphax_sum = x2 / 3 + (x1 * 2 / 3 + x0) per channel and alpha2_sum = w2 / 9 + (w1 * 4 / 9 + w0) have the same shape, so
//     the w lane of the (z, w) pair computes alpha2_sum with its own constants while the z lane computes the blue alphax_sum;
//   * the R and G channels of ONE split go through icbc_chan_pair together (grid 31 | 63), with the split's alpha2 / beta2 /
//     alphabeta sums and factor as broadcast scalar operands;
//   * only the B channel and the few weight sums are paired across the two splits, which costs the register moves.
// x = t * factor; saturate(x) = min(max(x, 0), 1) with NaN -> 0 is the .SAT of the scalar multiply (one instruction per
// value instead of half an FMUL2 plus a clamp).
NVB_DEV pf2 icbc_chan_pair(pf2 alphax, pf2 S, pf2 alpha2, pf2 beta2, pf2 ab, float fx, float fy, pf2 g, pf2 gr) {
    const pf2 betax = pf2_sub(S, alphax);
    const pf2 at = pf2_sub_s(pf2_mul(alphax, beta2), pf2_mul(betax, ab));
    const pf2 bt = pf2_sub_s(pf2_mul(betax, alpha2), pf2_mul(alphax, ab));
    pf2 a = pf2_pack(__saturatef(__fmul_rn(pf2_lo(at), fx)), __saturatef(__fmul_rn(pf2_hi(at), fy)));
    pf2 b = pf2_pack(__saturatef(__fmul_rn(pf2_lo(bt), fx)), __saturatef(__fmul_rn(pf2_hi(bt), fy)));
    const pf2 ha = pf2_add_s(pf2_mul(a, g), pf2_splat(0.5f)), hb = pf2_add_s(pf2_mul(b, g), pf2_splat(0.5f));
    a = pf2_mul(pf2_pack(truncf(pf2_lo(ha)), truncf(pf2_hi(ha))), gr);
    b = pf2_mul(pf2_pack(truncf(pf2_lo(hb)), truncf(pf2_hi(hb))), gr);
    const pf2 e2 = pf2_mul(pf2_sub_s(pf2_mul(a, pf2_sub_s(pf2_mul(b, ab), alphax)), pf2_mul(b, betax)), pf2_splat(2.0f));
    return pf2_add_s(pf2_mul(pf2_mul(a, a), alpha2), pf2_add_s(pf2_mul(pf2_mul(b, b), beta2), e2));
}
template <bool FOUR, bool U>
NVB_DEV float2 icbc_eval_pair(const float4 *sat, unsigned pka, unsigned pkb, float4 sum, const float msq[3]) {
    const int a0 = (int)(pka & 31), a1 = (int)((pka >> 5) & 31), a2 = (int)((pka >> 10) & 31);  // index + 1; sat[0] = 0
    const int b0 = (int)(pkb & 31), b1 = (int)((pkb >> 5) & 31), b2 = (int)((pkb >> 10) & 31);
    const float4 sa0 = sat[a0], sb0 = sat[b0];
    const float4 sa1 = sat[a1], sb1 = sat[b1];
    // per split: lo = (x, y), hi = (z, w) of a SAT row
    const pf2 a0l = pf2_pack(sa0.x, sa0.y), a0h = pf2_pack(sa0.z, sa0.w), b0l = pf2_pack(sb0.x, sb0.y), b0h = pf2_pack(sb0.z, sb0.w);
    const pf2 a1l = pf2_pack(sa1.x, sa1.y), a1h = pf2_pack(sa1.z, sa1.w), b1l = pf2_pack(sb1.x, sb1.y), b1h = pf2_pack(sb1.z, sb1.w);
    const pf2 dA1l = pf2_sub(a1l, a0l), dA1h = pf2_sub(a1h, a0h), dB1l = pf2_sub(b1l, b0l), dB1h = pf2_sub(b1h, b0h);  // x1 | w1
    pf2 rAl, rAh, rBl, rBh;  // (alphax_sum.x, .y) and (alphax_sum.z, alpha2_sum) of each split
    pf2 beta2, ab;           // (A, B)
    if (FOUR) {
        const float4 sa2 = sat[a2], sb2 = sat[b2];
        const pf2 a2l = pf2_pack(sa2.x, sa2.y), a2h = pf2_pack(sa2.z, sa2.w), b2l = pf2_pack(sb2.x, sb2.y), b2h = pf2_pack(sb2.z, sb2.w);
        const pf2 dA2l = pf2_sub(a2l, a1l), dA2h = pf2_sub(a2h, a1h), dB2l = pf2_sub(b2l, b1l), dB2h = pf2_sub(b2h, b1h);  // x2 | w2
        const pf2 c23 = pf2_splat(2.0f / 3.0f), c13 = pf2_splat(1.0f / 3.0f);
        const pf2 c23_49 = pf2_pack(2.0f / 3.0f, 4.0f / 9.0f), c13_19 = pf2_pack(1.0f / 3.0f, 1.0f / 9.0f);
        rAl = pf2_add_s(pf2_mul(dA2l, c13), pf2_add_s(pf2_mul(dA1l, c23), a0l));
        rAh = pf2_add_s(pf2_mul(dA2h, c13_19), pf2_add_s(pf2_mul(dA1h, c23_49), a0h));
        rBl = pf2_add_s(pf2_mul(dB2l, c13), pf2_add_s(pf2_mul(dB1l, c23), b0l));
        rBh = pf2_add_s(pf2_mul(dB2h, c13_19), pf2_add_s(pf2_mul(dB1h, c23_49), b0h));
        const pf2 w1 = pf2_pack(pf2_hi(dA1h), pf2_hi(dB1h)), w2 = pf2_pack(pf2_hi(dA2h), pf2_hi(dB2h));
        const pf2 w3 = pf2_sub(pf2_splat(sum.w), pf2_pack(sa2.w, sb2.w));
        beta2 = pf2_add_s(pf2_mul(w1, pf2_splat(1.0f / 9.0f)), pf2_add_s(pf2_mul(w2, pf2_splat(4.0f / 9.0f)), w3));
        ab = pf2_mul(pf2_add(w1, w2), pf2_splat(2.0f / 9.0f));
    } else {
        const pf2 ch = pf2_splat(0.5f), ch_q = pf2_pack(0.5f, 0.25f);
        const pf2 tAh = pf2_mul(dA1h, ch_q), tBh = pf2_mul(dB1h, ch_q);  // (x1.z / 2, alphabeta_sum = w1 / 4)
        rAl = pf2_add_s(a0l, pf2_mul(dA1l, ch));
        rAh = pf2_add_s(a0h, tAh);
        rBl = pf2_add_s(b0l, pf2_mul(dB1l, ch));
        rBh = pf2_add_s(b0h, tBh);
        ab = pf2_pack(pf2_hi(tAh), pf2_hi(tBh));
        const pf2 w2 = pf2_sub(pf2_splat(sum.w), pf2_pack(sa1.w, sb1.w));
        beta2 = pf2_add_s(w2, ab);
    }
    const pf2 alpha2 = pf2_pack(pf2_hi(rAh), pf2_hi(rBh));
    const pf2 det = pf2_sub_s(pf2_mul(alpha2, beta2), pf2_mul(ab, ab));
    const float fA = 1.0f / pf2_lo(det), fB = 1.0f / pf2_hi(det);
    const pf2 gxy = pf2_pack(31.0f, 63.0f), grxy = pf2_pack(1.0f / 31.0f, 1.0f / 63.0f);
    const pf2 sxy = pf2_pack(sum.x, sum.y);
    const pf2 eA = icbc_chan_pair(rAl, sxy, pf2_splat(pf2_lo(alpha2)), pf2_splat(pf2_lo(beta2)), pf2_splat(pf2_lo(ab)), fA, fA, gxy, grxy);
    const pf2 eB = icbc_chan_pair(rBl, sxy, pf2_splat(pf2_hi(alpha2)), pf2_splat(pf2_hi(beta2)), pf2_splat(pf2_hi(ab)), fB, fB, gxy, grxy);
    const pf2 eZ = icbc_chan_pair(pf2_pack(pf2_lo(rAh), pf2_lo(rBh)), pf2_splat(sum.z), alpha2, beta2, ab, fA, fB, pf2_splat(31.0f),
                                  pf2_splat(1.0f / 31.0f));
    if (U)
        return make_float2(__fadd_rn(__fadd_rn(pf2_lo(eA), pf2_hi(eA)), pf2_lo(eZ)), __fadd_rn(__fadd_rn(pf2_lo(eB), pf2_hi(eB)), pf2_hi(eZ)));
    const pf2 mxy = pf2_pack(msq[0], msq[1]);
    const pf2 pA = pf2_mul(eA, mxy), pB = pf2_mul(eB, mxy), pZ = pf2_mul(eZ, pf2_splat(msq[2]));
    return make_float2(__fadd_rn(__fadd_rn(pf2_lo(pA), pf2_hi(pA)), pf2_lo(pZ)), __fadd_rn(__fadd_rn(pf2_lo(pB), pf2_hi(pB)), pf2_hi(pZ)));
}

// PAIR: two splits per trip with packed fp32 (fewer issue slots, more code).  Used at every level; the level-9 kernel used to
// keep the scalar loop because its instruction-cache footprint was the limiter - since the refinement shrank (exact /255 by
// multiply + 2 FMAs, palette look-up through shared memory) the packed loop wins there too (8192² Production 45.8 -> 43.1 ms).
template <bool FOUR, bool PAIR, bool U>
NVB_DEV FitResult icbc_cluster_fit(const Bc1Params &P, const Bc1GroupSmem &S, unsigned gm, int l, int count) {
    const float4 sum = S.sat[count];
    const float msq[3] = {P.cw[0] * P.cw[0], P.cw[1] * P.cw[1], P.cw[2] * P.cw[2]};
    const unsigned short *tab = FOUR ? P.four : P.three;
    const int total = FOUR ? P.four_total[count - 1] : P.three_total[count - 1];
    float besterror = FLT_MAX;
    int besti = 0x7fffffff;
    float a[3], b[3];
    if (!PAIR) {
        for (int i = l; i < total; i += 16) {
            const float e = icbc_eval_split<FOUR, U>(S.sat, __ldg(tab + i), sum, msq, a, b);
            if (e < besterror) {
                besterror = e;
                besti = i;
            }
        }
    }
    // two splits per trip (i and i + 16); the last, unpaired one re-uses split i for the idle half
    for (int i = l; PAIR && i < total; i += 32) {
        const bool two = i + 16 < total;
        const unsigned pka = __ldg(tab + i), pkb = two ? __ldg(tab + i + 16) : pka;
        const float2 e = icbc_eval_pair<FOUR, U>(S.sat, pka, pkb, sum, msq);
        if (e.x < besterror) {
            besterror = e.x;
            besti = i;
        }
        if (two && e.y < besterror) {
            besterror = e.y;
            besti = i + 16;
        }
    }
#pragma unroll
    for (int d = 8; d >= 1; d >>= 1) {
        const float oe = __shfl_xor_sync(gm, besterror, d, 16);
        const int oi = __shfl_xor_sync(gm, besti, d, 16);
        if (oe < besterror || (oe == besterror && oi < besti)) {
            besterror = oe;
            besti = oi;
        }
    }
    FitResult r;
    r.sx = r.sy = r.sz = r.ex = r.ey = r.ez = 0.0f;  // vbeststart / vbestend start at zero
    if (besti != 0x7fffffff) {
        icbc_eval_split<FOUR, U>(S.sat, __ldg(tab + besti), sum, msq, a, b);
        r.sx = a[0]; r.sy = a[1]; r.sz = a[2];
        r.ex = b[0]; r.ey = b[1]; r.ez = b[2];
    }
    return r;
}

// One instantiation per ICBC level (1 = box fit, 8 = cluster fit, 9 = cluster fit + refinement) and for the BC3-RGBM colour
// block: the generic kernel was 177 KB of SASS - more than the instruction cache - and warps in different phases of it
// evicted each other's code (Production got slower when the cluster fit got bigger).
// U: colour weights (1, 1, 1) - see cw_mul.
#define NVB_BC1_PAL_PITCH 5  // float4 per lane: 80 bytes, so that the 128-bit reads of 8 consecutive lanes touch all 32 banks once
template <int LEVEL, bool RGBM, bool U> __global__ void __launch_bounds__(NVB_BC1_GROUPS * 16, NVB_BC1_MINB) k_bc1_icbc_t(Bc1Params P) {
    __shared__ Bc1GroupSmem smem[NVB_BC1_GROUPS];
    // Level 9: the palette of the refinement candidate each lane is measuring, read back by palette index
    __shared__ float4 s_pal[LEVEL == 9 ? NVB_BC1_GROUPS * 16 * NVB_BC1_PAL_PITCH : 1];
    __shared__ __align__(16) int s_off[LEVEL == 9 ? NVB_BC1_GROUPS * 16 : 4];  // per block: byte offset of texel t's palette entry
    const int grp = threadIdx.x >> 4;
    const int l = threadIdx.x & 15;
    const int nblocks = P.lv.bw * P.lv.bh;
    const int blkid = blockIdx.x * NVB_BC1_GROUPS + grp;
    if (blkid >= nblocks) return;
    const unsigned gsh = threadIdx.x & 16;
    const unsigned gm = 0xFFFFu << gsh;
    Bc1GroupSmem &S = smem[grp];

    // ---- gather: fp32 texels, zero padding + zero weight outside the image ----
    const int bx = blkid % P.lv.bw, by = blkid / P.lv.bw;
    const int px = bx * 4 + (l & 3), py = by * 4 + (l >> 2);
    float cx = 0.0f, cy = 0.0f, cz = 0.0f, wt = 0.0f;
    if (px < P.lv.w && py < P.lv.h) {
        cx = load_texel(P.lv, 0, px, py);
        cy = load_texel(P.lv, 1, px, py);
        cz = load_texel(P.lv, 2, px, py);
        wt = 1.0f;
        if (P.transparency) wt = icbc_saturate(load_texel(P.lv, 3, px, py));
    }
    if (RGBM) {
        // convert_to_rgbm (CompressorDXT5_RGBM.cpp:23-49)
        const float R = icbc_saturate(cx), G = icbc_saturate(cy), B = icbc_saturate(cz);
        const float M = nv_max(nv_max(R, G), nv_max(B, P.rgbm_min));
        cx = R / M;
        cy = G / M;
        cz = B / M;
        const float weight_sum = group_ordered_sum(gm, wt);
        wt = (weight_sum == 0) ? 1.0f : wt * M;
    }
    unsigned char *dst = P.out + nvb_out_block(P.lv, blkid) * P.out_stride + P.out_offset;
    Bc1Block out;
    out.c0 = out.c1 = out.indices = 0;
    float error = FLT_MAX;
    int count = 16;
    bool any_black = false;

    if (LEVEL >= 2) {
        // ---- reduce_colors: merge texel i into the first earlier cluster within 1/256 per channel ----
        const float threshold = 1.0f / 256;
        float qx = 0.0f, qy = 0.0f, qz = 0.0f, qw = 0.0f;  // cluster l (valid for l < n)
        int n = 0;
#pragma unroll 1
        for (int i = 0; i < 16; i++) {
            const float tx = __shfl_sync(gm, cx, i, 16), ty = __shfl_sync(gm, cy, i, 16), tz = __shfl_sync(gm, cz, i, 16);
            const float tw = __shfl_sync(gm, wt, i, 16);
            if (tw > 0) {  // group-uniform
                const bool match = (l < n) && (fabsf(qx - tx) < threshold) && (fabsf(qy - ty) < threshold) && (fabsf(qz - tz) < threshold);
                const unsigned mm = (__ballot_sync(gm, match) >> gsh) & 0xFFFFu;
                if (mm) {
                    if (l == __ffs((int)mm) - 1) {
                        const float den = qw + tw;
                        qx = (qx * qw + tx * tw) / den;
                        qy = (qy * qw + ty * tw) / den;
                        qz = (qz * qw + tz * tw) / den;
                        qw += tw;
                    }
                } else {
                    if (l == n) { qx = tx; qy = ty; qz = tz; qw = tw; }
                    n++;
                }
            }
        }
        count = n;
        const bool black = (wt > 0) && (cx < 1.0f / 8) && (cy < 1.0f / 8) && (cz < 1.0f / 8);
        any_black = __ballot_sync(gm, black) & gm;
        if (count == 0) {
            if (l == 0) *reinterpret_cast<uint2 *>(dst) = make_uint2(0u, 0u);
            return;
        }
        if (count == 1) {
            // compress_dxt1_single_color_optimal(vector3_to_color32(colors[0]))
            if (l == 0) {
                const unsigned r = (unsigned)__float2int_rz(icbc_saturate(qx) * 255 + 0.5f) & 0xFF;
                const unsigned g = (unsigned)__float2int_rz(icbc_saturate(qy) * 255 + 0.5f) & 0xFF;
                const unsigned b = (unsigned)__float2int_rz(icbc_saturate(qz) * 255 + 0.5f) & 0xFF;
                unsigned c0 = ((unsigned)P.match5[r * 2 + 0] << 11) | ((unsigned)P.match6[g * 2 + 0] << 5) | P.match5[b * 2 + 0];
                unsigned c1 = ((unsigned)P.match5[r * 2 + 1] << 11) | ((unsigned)P.match6[g * 2 + 1] << 5) | P.match5[b * 2 + 1];
                unsigned indices = 0xaaaaaaaau;
                if (c0 < c1) {
                    const unsigned t = c0; c0 = c1; c1 = t;
                    indices ^= 0x55555555u;
                }
                *reinterpret_cast<uint2 *>(dst) = make_uint2(c0 | (c1 << 16), indices);
            }
            return;
        }
        if (l < count) S.pts[l] = make_float4(qx, qy, qz, qw);
        __syncwarp(gm);
    }

    if (LEVEL == 1) {
        // ---- box fit on all 16 (un-reduced, padding included) colours + least squares refit ----
        // fit_colors_bbox: fold over the texels in order with icbc::max/min semantics
        float c0x = 0.0f, c0y = 0.0f, c0z = 0.0f, c1x = 1.0f, c1y = 1.0f, c1z = 1.0f;
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const float tx = __shfl_sync(gm, cx, j, 16), ty = __shfl_sync(gm, cy, j, 16), tz = __shfl_sync(gm, cz, j, 16);
            c0x = nv_max(c0x, tx); c0y = nv_max(c0y, ty); c0z = nv_max(c0z, tz);
            c1x = nv_min(c1x, tx); c1y = nv_min(c1y, ty); c1z = nv_min(c1z, tz);
        }
        // inset_bbox
        {
            const float bias = (8.0f / 255.0f) / 16.0f;
            const float ix = (c0x - c1x) / 16.0f - bias, iy = (c0y - c1y) / 16.0f - bias, iz = (c0z - c1z) / 16.0f - bias;
            c0x = icbc_saturate(c0x - ix); c0y = icbc_saturate(c0y - iy); c0z = icbc_saturate(c0z - iz);
            c1x = icbc_saturate(c1x + ix); c1y = icbc_saturate(c1y + iy); c1z = icbc_saturate(c1z + iz);
        }
        // select_diagonal
        {
            const float mx = (c0x + c1x) * 0.5f, my = (c0y + c1y) * 0.5f, mz = (c0z + c1z) * 0.5f;
            const float tx = cx - mx, ty = cy - my, tz = cz - mz;
            const float cov_xz = group_ordered_sum(gm, tx * tz);
            const float cov_yz = group_ordered_sum(gm, ty * tz);
            if (cov_xz < 0) { const float t = c0x; c0x = c1x; c1x = t; }
            if (cov_yz < 0) { const float t = c0y; c0y = c1y; c1y = t; }
        }
        error = icbc_output_block<U>(P, gm, l, true, false, c0x, c0y, c0z, c1x, c1y, c1z, cx, cy, cz, wt, &out);
        // optimize_end_points4 on the chosen indices (unweighted, all 16 texels)
        {
            const unsigned bits = out.indices >> (2 * l);
            float beta = (float)(bits & 1);
            if (bits & 2) beta = (1 + beta) / 3.0f;
            const float alpha = 1.0f - beta;
            const float alpha2_sum = group_ordered_sum(gm, alpha * alpha);
            const float beta2_sum = group_ordered_sum(gm, beta * beta);
            const float alphabeta_sum = group_ordered_sum(gm, alpha * beta);
            const float axx = group_ordered_sum(gm, alpha * cx), axy = group_ordered_sum(gm, alpha * cy), axz = group_ordered_sum(gm, alpha * cz);
            const float bxx = group_ordered_sum(gm, beta * cx), bxy = group_ordered_sum(gm, beta * cy), bxz = group_ordered_sum(gm, beta * cz);
            const float denom = alpha2_sum * beta2_sum - alphabeta_sum * alphabeta_sum;
            if (!(fabsf(denom - 0.0f) < 0.0001f)) {
                const float factor = 1.0f / denom;
                const float ax = icbc_saturate((axx * beta2_sum - bxx * alphabeta_sum) * factor);
                const float ay = icbc_saturate((axy * beta2_sum - bxy * alphabeta_sum) * factor);
                const float az = icbc_saturate((axz * beta2_sum - bxz * alphabeta_sum) * factor);
                const float bx2 = icbc_saturate((bxx * alpha2_sum - axx * alphabeta_sum) * factor);
                const float by2 = icbc_saturate((bxy * alpha2_sum - axy * alphabeta_sum) * factor);
                const float bz2 = icbc_saturate((bxz * alpha2_sum - axz * alphabeta_sum) * factor);
                Bc1Block opt;
                const float oe = icbc_output_block<U>(P, gm, l, true, false, ax, ay, az, bx2, by2, bz2, cx, cy, cz, wt, &opt);
                if (oe < error) {
                    error = oe;
                    out = opt;
                }
            }
        }
    } else {
        // ---- compress_dxt1_cluster_fit ----
        icbc_compute_sat(S, gm, l, count);
        FitResult f4 = icbc_cluster_fit<true, true, U>(P, S, gm, l, count);
        Bc1Block cf;
        float best = icbc_output_block<U>(P, gm, l, true, false, f4.sx, f4.sy, f4.sz, f4.ex, f4.ey, f4.ez, cx, cy, cz, wt, &cf);
        // three colour mode (Levels 8/9: always tried; transparent black allowed)
        int sat_count = count;
        bool do_three = !RGBM;  // compress_dxt5_rgbm passes three_color_mode = false
        if (do_three && any_black) {
            // skip_blacks on the reduced set, then a new SAT
            const float4 q = (l < count) ? S.pts[l] : make_float4(1.f, 1.f, 1.f, 0.f);
            const bool keep = (l < count) && !((q.x < 1.0f / 8) && (q.y < 1.0f / 8) && (q.z < 1.0f / 8));
            const unsigned km = (__ballot_sync(gm, keep) >> gsh) & 0xFFFFu;
            const int tmp_count = __popc(km);
            if (tmp_count == 0) {
                do_three = false;
            } else {
                __syncwarp(gm);
                if (keep) S.pts[__popc(km & ((1u << l) - 1u))] = q;
                __syncwarp(gm);
                icbc_compute_sat(S, gm, l, tmp_count);
                sat_count = tmp_count;
            }
        }
        if (do_three) {
            FitResult f3 = icbc_cluster_fit<false, true, U>(P, S, gm, l, sat_count);
            Bc1Block tb;
            const float te = icbc_output_block<U>(P, gm, l, false, true, f3.sx, f3.sy, f3.sz, f3.ex, f3.ey, f3.ez, cx, cy, cz, wt, &tb);
            if (te < best) {
                best = te;
                cf = tb;
            }
        }
        if (best < error) {
            out = cf;
            error = best;
        }
        if (LEVEL == 9) {
            // ---- refine_endpoints (three_color_mode == true: no endpoint re-ordering) ----
            // The reference walks up to 256 single-step endpoint moves one after the other and stops 33 candidates after the
            // last accepted one.  Between two acceptances every candidate is measured against the same block and the same
            // threshold, so 16 consecutive candidates are evaluated at once, one per lane (each lane sums its candidate's
            // 16 texel errors in texel order), and the first improving one in candidate order is accepted.
            float best_error = error;
            int lastImprovement = 0;
            const float vcx = cw_mul<U>(cx, P.cw[0]), vcy = cw_mul<U>(cy, P.cw[1]), vcz = cw_mul<U>(cz, P.cw[2]);
            const char *const my_pal = reinterpret_cast<const char *>(s_pal + threadIdx.x * NVB_BC1_PAL_PITCH);
            int *const g_off = s_off + grp * 16;
            __syncwarp(gm);
            S.pts[l] = make_float4(cx, cy, cz, wt);
            unsigned acc_idx = 0;   // this texel's index at the last acceptance (the reference stores the indices it measured with)
            bool accepted = false;
            int i = 0;
#pragma unroll 1
            for (;;) {
                // indices from the palette of *output* (sic), general 4-way rule: they only change when `out` changes
                const Pal3 pal = icbc_palette_f(out.c0, out.c1);
                float d[4];
#pragma unroll
                for (int q = 0; q < 4; q++)
                    d[q] = dist2w(vcx, vcy, vcz, cw_mul<U>(pal.x[q], P.cw[0]), cw_mul<U>(pal.y[q], P.cw[1]), cw_mul<U>(pal.z[q], P.cw[2]));
                const bool i1 = (d[1] <= d[0]) && (d[1] < d[2]) && (d[1] < d[3]);
                const bool i2 = (d[2] <= d[0]) && (d[2] <= d[1]) && (d[2] < d[3]);
                const bool i3 = (d[3] <= d[0]) && (d[3] <= d[1]) && (d[3] <= d[2]);
                const unsigned idx = ((i1 || i3) ? 1u : 0u) | ((i2 || i3) ? 2u : 0u);
                // texel t's palette entry as a byte offset into a lane's palette slot, shared by the 16 candidates of a round
                __syncwarp(gm);
                g_off[l] = (int)(idx * 16u);
                __syncwarp(gm);
                for (;;) {
                    const int limit = min(255, lastImprovement + 33);  // last candidate the sequential loop reaches
                    const int ci = i + l;
                    // deltas[ci % 16], rows:
                    // (1,0,0)(0,1,0)(0,0,1)(-1,0,0)(0,-1,0)(0,0,-1)(1,1,0)(1,0,1)(0,1,1)(-1,-1,0)(-1,0,-1)(0,-1,-1)(-1,1,0)(1,-1,0)(0,-1,1)(0,1,-1)
                    // as 2-bit fields (delta + 1) of one word per channel
                    const int k2 = 2 * (ci & 15);
                    const unsigned dr = ((DWR >> k2) & 3u) - 1u;
                    const unsigned dg = ((DWG >> k2) & 3u) - 1u;
                    const unsigned db = ((DWB >> k2) & 3u) - 1u;
                    unsigned c0 = out.c0, c1 = out.c1;
                    {
                        unsigned c = ((ci / 16) & 1) ? c0 : c1;
                        const unsigned r = (((c >> 11) & 31) + dr) & 31;
                        const unsigned g = (((c >> 5) & 63) + dg) & 63;
                        const unsigned b = ((c & 31) + db) & 31;
                        c = (r << 11) | (g << 5) | b;
                        if ((ci / 16) & 1) c0 = c; else c1 = c;
                    }
                    // evaluate_mse of the candidate block, texel by texel in order; the candidate's palette goes through this
                    // lane's own shared-memory slot so that "palette[index of texel t]" is one LDS.128 instead of nine selects
                    {
                        const Pal3 rp = icbc_palette_f(c0, c1);
                        float4 *const w = s_pal + threadIdx.x * NVB_BC1_PAL_PITCH;
#pragma unroll
                        for (int q = 0; q < 4; q++) w[q] = make_float4(rp.x[q], rp.y[q], rp.z[q], 0.0f);
                    }
                    float e = 0.0f;
#pragma unroll 2
                    for (int t4 = 0; t4 < 4; t4++) {
                        const int4 o4 = *reinterpret_cast<const int4 *>(g_off + 4 * t4);
                        const int o[4] = {o4.x, o4.y, o4.z, o4.w};
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const float4 q = S.pts[4 * t4 + u];
                            const float4 c = *reinterpret_cast<const float4 *>(my_pal + o[u]);
                            e += q.w * mse_term<U>(c.x, c.y, c.z, q.x, q.y, q.z, P.cw);
                        }
                    }
                    const bool better = (ci <= limit) && (e < best_error);
                    const unsigned bm = (__ballot_sync(gm, better) >> gsh) & 0xFFFFu;
                    if (bm) {
                        const int wl = __ffs((int)bm) - 1;  // first improving candidate in order
                        best_error = __shfl_sync(gm, e, wl, 16);
                        out.c0 = __shfl_sync(gm, c0, wl, 16);
                        out.c1 = __shfl_sync(gm, c1, wl, 16);
                        acc_idx = idx;
                        accepted = true;
                        lastImprovement = i + wl;
                        i += wl + 1;
                        break;  // `out` changed: new indices
                    }
                    i += 16;
                    if (i > limit) break;
                }
                if (i > min(255, lastImprovement + 33)) break;
            }
            if (accepted) out.indices = group_gather_bits2(gm, acc_idx, l);
            error = best_error;
        }
    }
    if (l == 0) *reinterpret_cast<uint2 *>(dst) = make_uint2(out.c0 | (out.c1 << 16), out.indices);
}

// host-side dispatch on (P.level, P.rgbm): LAUNCH(kernel) is expanded with the matching instantiation
#define NVB_BC1_DISPATCH(P, LAUNCH)                                                    \
    do {                                                                               \
        const bool unit_ = (P).cw[0] == 1.0f && (P).cw[1] == 1.0f && (P).cw[2] == 1.0f; \
        if ((P).rgbm) { LAUNCH((k_bc1_icbc_t<8, true, false>)); }                      \
        else if ((P).level == 1) { LAUNCH((k_bc1_icbc_t<1, false, false>)); }          \
        else if ((P).level == 9 && unit_) { LAUNCH((k_bc1_icbc_t<9, false, true>)); }  \
        else if ((P).level == 9) { LAUNCH((k_bc1_icbc_t<9, false, false>)); }          \
        else if (unit_) { LAUNCH((k_bc1_icbc_t<8, false, true>)); }                    \
        else { LAUNCH((k_bc1_icbc_t<8, false, false>)); }                              \
    } while (0)

}  // namespace nvb
