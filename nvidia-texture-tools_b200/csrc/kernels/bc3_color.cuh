// BC2/BC3 colour block encoder: nvsquish weighted cluster fit (4-colour mode), 16 lanes per 4x4 block.
//
// Replaces, bit-exactly (intended semantics = the -O0/-O2 behaviour of the reference):
//   CompressorDXT5::compressBlock (colour part)       src/nvtt/CompressorDX9.cpp:160-175
//   ColorBlock::init(w,h,float*,x,y) / isSingleColor  src/nvimage/ColorBlock.cpp:80-110,136-149
//   nvsquish::ColourSet::ColourSet (minimal set)      src/nvtt/squish/colourset.cpp:35-139
//   ComputeWeightedCovariance / ComputePrincipleComponent (scalar)   src/nvtt/squish/maths.cpp:32-133
//   WeightedClusterFit::SetColourSet / Compress4 (scalar)            src/nvtt/squish/weightedclusterfit.cpp:39-105,476-589
//   WriteColourBlock4 / FloatTo565                    src/nvtt/squish/colourblock.cpp:30-53,108-138
//   OptimalCompress::compressDXT1(Color32)            src/nvtt/OptimalCompressDXT.cpp:254-269
//
// Mapping: a half-warp owns one block, one lane per texel.  Texel de-duplication uses shuffles; the sequential
// fp32 reductions whose order matters (centroid, covariance, power iteration, xsum) are evaluated redundantly
// by every lane from shared memory (broadcast reads) in the reference's order; the <= 969 cluster splits
// (c0,c1,c2) are striped over the 16 lanes, each split reading three running sums from a per-block table
// T[start][len] that is accumulated in exactly the order of the reference's nested loops; the winner is the
// minimum error with the lowest split number (== first strict minimum of the sequential search).
#pragma once
#include "../nvb_common.cuh"

namespace nvb {

struct Bc3ColorParams {
    LevelView lv;
    unsigned char *out;     // BCn level
    int out_stride;         // bytes between blocks (16 for BC2/BC3)
    int out_offset;         // byte offset of the colour block inside a block (8)
    float metric[3];        // CompressionOptions colour weights
    int weight_by_alpha;    // AlphaMode_Transparency => kWeightColourByAlpha
    const unsigned short *cand;  // packed (c0 | c1<<5 | c2<<10) splits, all counts concatenated
    const unsigned *cand_idx;    // same order: positions of the split's three running sums in T, i0 | i1<<8 | i2<<16
    const int *cand_off;    // cand_off[n] .. cand_off[n+1] = splits for a set of n points (n = 1..16), 18 ints
    const unsigned char *omatch5;  // [256][2]
    const unsigned char *omatch6;  // [256][2]
    int dxt5n;              // 1: CompressorDXT5n colour block — tile swizzled to (0xFF, G, 0) and metric (0,1,0) (CompressorDX9.cpp:179-210)
    // BC1a (CompressorDXT1a, CompressorDX9.cpp:83-111) only: two-cluster splits c0|c1<<5 and the alpha-mode match tables
    const unsigned short *cand3;
    const int *cand3_off;
    const unsigned char *omatch5a;  // OMatchAlpha5 [256][2]
    const unsigned char *omatch6a;  // OMatchAlpha6 [256][2]
};

#define NVB_BC3_GROUPS 8  // 4x4 blocks per CTA (128 threads)

struct Bc3GroupSmem {
    float4 pts[16];      // de-duplicated points (x=R,y=G,z=B in [0,1], w=weight)
    float4 sorted[16];   // weighted points in principal-axis order (w*x, w*y, w*z, w)
    float4 T[154];       // running sums: T[off(start)+len] = sum_{k<len} sorted[start+k]
    int rank[16];        // point -> position in the sorted order
};

NVB_DEV int squish_float_to_int(float a, int limit) {
    int i = x86_ftoi(a + 0.5f);
    if (i < 0) i = 0;
    else if (i > limit) i = limit;
    return i;
}

struct SquishSplit {
    float ax, ay, az, bx, by, bz, error;
};

// i0, i1, i2 = positions in T of the three running sums of the split (c0, c1, c2):
//   i0 = c0, i1 = off(c0) + c1, i2 = off(c0 + c1) + c2 with off(s) = s*(n+1) - s*(s-1)/2 (start of row s of T)
NVB_DEV SquishSplit squish_eval_split(const float4 *T, int i0, int i1, int i2, float4 xsum, float mx, float my, float mz) {
    const float4 x0 = T[i0];
    const float4 x1 = T[i1];
    const float4 x2 = T[i2];
    const float w0 = x0.w, w1 = x1.w, w2 = x2.w;
    const float w3 = xsum.w - w0 - w1 - w2;
    const float alpha2_sum = w0 + w1 * (4.0f / 9.0f) + w2 * (1.0f / 9.0f);
    const float beta2_sum = w3 + w2 * (4.0f / 9.0f) + w1 * (1.0f / 9.0f);
    const float alphabeta_sum = (w1 + w2) * (2.0f / 9.0f);
    const float factor = 1.0f / (alpha2_sum * beta2_sum - alphabeta_sum * alphabeta_sum);
    SquishSplit r;
    float e[3];
    const float X0[3] = {x0.x, x0.y, x0.z}, X1[3] = {x1.x, x1.y, x1.z}, X2[3] = {x2.x, x2.y, x2.z};
    const float XS[3] = {xsum.x, xsum.y, xsum.z};
    const float grid[3] = {31.0f, 63.0f, 31.0f};
    const float gridrcp[3] = {1.0f / 31.0f, 1.0f / 63.0f, 1.0f / 31.0f};
    float A[3], B[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float alphax_sum = X0[k] + X1[k] * (2.0f / 3.0f) + X2[k] * (1.0f / 3.0f);
        const float betax_sum = XS[k] - alphax_sum;
        float a = (alphax_sum * beta2_sum - betax_sum * alphabeta_sum) * factor;
        float b = (betax_sum * alpha2_sum - alphax_sum * alphabeta_sum) * factor;
        a = std_min(1.0f, std_max(0.0f, a));
        b = std_min(1.0f, std_max(0.0f, b));
        a = floorf(grid[k] * a + 0.5f) * gridrcp[k];
        b = floorf(grid[k] * b + 0.5f) * gridrcp[k];
        e[k] = a * a * alpha2_sum + b * b * beta2_sum + 2.0f * (a * b * alphabeta_sum - a * alphax_sum - b * betax_sum);
        A[k] = a;
        B[k] = b;
    }
    r.error = e[0] * mx + e[1] * my + e[2] * mz;
    r.ax = A[0]; r.ay = A[1]; r.az = A[2];
    r.bx = B[0]; r.by = B[1]; r.bz = B[2];
    return r;
}

// Two splits at once, as packed fp32 pairs (FMUL2 / FADD2 / unfused FFMA2-by-one sums).  Operation for operation the same
// arithmetic as squish_eval_split - each half of a pair is an independent IEEE round-to-nearest op - only the errors are
// returned.  The pairing follows the registers an LDS.128 of a T row fills: the R and G channels of ONE split are already an
// aligned register pair (x, y), so they go through squish_chan_pair together with the split's weights sums broadcast; only
// the B channel and the weight sums are paired across the two splits (A, B), which costs the register moves.
NVB_DEV pf2 squish_chan_pair(pf2 X0, pf2 X1, pf2 X2, pf2 xs, pf2 alpha2, pf2 beta2, pf2 ab, float fx, float fy, pf2 g, pf2 gr) {
    const pf2 c23 = pf2_splat(2.0f / 3.0f), c13 = pf2_splat(1.0f / 3.0f);
    const pf2 alphax = pf2_add_s(pf2_add_s(X0, pf2_mul(X1, c23)), pf2_mul(X2, c13));
    const pf2 betax = pf2_sub(xs, alphax);
    const pf2 at = pf2_sub_s(pf2_mul(alphax, beta2), pf2_mul(betax, ab));
    const pf2 bt = pf2_sub_s(pf2_mul(betax, alpha2), pf2_mul(alphax, ab));
    // min(1, max(0, t * factor)) with NaN -> 0 is exactly a saturating multiply
    pf2 a = pf2_pack(__saturatef(__fmul_rn(pf2_lo(at), fx)), __saturatef(__fmul_rn(pf2_hi(at), fy)));
    pf2 b = pf2_pack(__saturatef(__fmul_rn(pf2_lo(bt), fx)), __saturatef(__fmul_rn(pf2_hi(bt), fy)));
    const pf2 ha = pf2_add_s(pf2_mul(g, a), pf2_splat(0.5f)), hb = pf2_add_s(pf2_mul(g, b), pf2_splat(0.5f));
    a = pf2_mul(pf2_pack(floorf(pf2_lo(ha)), floorf(pf2_hi(ha))), gr);
    b = pf2_mul(pf2_pack(floorf(pf2_lo(hb)), floorf(pf2_hi(hb))), gr);
    const pf2 s1 = pf2_add_s(pf2_mul(pf2_mul(a, a), alpha2), pf2_mul(pf2_mul(b, b), beta2));
    const pf2 d = pf2_sub_s(pf2_sub_s(pf2_mul(pf2_mul(a, b), ab), pf2_mul(a, alphax)), pf2_mul(b, betax));
    return pf2_add(s1, pf2_add(d, d));
}

NVB_DEV float2 squish_eval_pair(const float4 *T, unsigned pa, unsigned pb, float4 xsum, float mx, float my, float mz) {
    const float4 a0 = T[pa & 0xFF], a1 = T[(pa >> 8) & 0xFF], a2 = T[pa >> 16];
    const float4 b0 = T[pb & 0xFF], b1 = T[(pb >> 8) & 0xFF], b2 = T[pb >> 16];
    const pf2 w0 = pf2_pack(a0.w, b0.w), w1 = pf2_pack(a1.w, b1.w), w2 = pf2_pack(a2.w, b2.w);
    const pf2 w3 = pf2_sub(pf2_sub(pf2_sub(pf2_splat(xsum.w), w0), w1), w2);
    const pf2 c49 = pf2_splat(4.0f / 9.0f), c19 = pf2_splat(1.0f / 9.0f), c29 = pf2_splat(2.0f / 9.0f);
    const pf2 alpha2 = pf2_add_s(pf2_add_s(w0, pf2_mul(w1, c49)), pf2_mul(w2, c19));
    const pf2 beta2 = pf2_add_s(pf2_add_s(w3, pf2_mul(w2, c49)), pf2_mul(w1, c19));
    const pf2 ab = pf2_mul(pf2_add(w1, w2), c29);
    const pf2 det = pf2_sub_s(pf2_mul(alpha2, beta2), pf2_mul(ab, ab));
    const float fA = 1.0f / pf2_lo(det), fB = 1.0f / pf2_hi(det);
    const pf2 gxy = pf2_pack(31.0f, 63.0f), grxy = pf2_pack(1.0f / 31.0f, 1.0f / 63.0f);
    const pf2 xsxy = pf2_pack(xsum.x, xsum.y);
    const pf2 eA = squish_chan_pair(pf2_pack(a0.x, a0.y), pf2_pack(a1.x, a1.y), pf2_pack(a2.x, a2.y), xsxy, pf2_splat(pf2_lo(alpha2)),
                                    pf2_splat(pf2_lo(beta2)), pf2_splat(pf2_lo(ab)), fA, fA, gxy, grxy);
    const pf2 eB = squish_chan_pair(pf2_pack(b0.x, b0.y), pf2_pack(b1.x, b1.y), pf2_pack(b2.x, b2.y), xsxy, pf2_splat(pf2_hi(alpha2)),
                                    pf2_splat(pf2_hi(beta2)), pf2_splat(pf2_hi(ab)), fB, fB, gxy, grxy);
    const pf2 eZ = squish_chan_pair(pf2_pack(a0.z, b0.z), pf2_pack(a1.z, b1.z), pf2_pack(a2.z, b2.z), pf2_splat(xsum.z), alpha2, beta2,
                                    ab, fA, fB, pf2_splat(31.0f), pf2_splat(1.0f / 31.0f));
    const pf2 mxy = pf2_pack(mx, my);
    const pf2 pA = pf2_mul(eA, mxy), pB = pf2_mul(eB, mxy), pZ = pf2_mul(eZ, pf2_splat(mz));
    return make_float2(__fadd_rn(__fadd_rn(pf2_lo(pA), pf2_hi(pA)), pf2_lo(pZ)), __fadd_rn(__fadd_rn(pf2_lo(pB), pf2_hi(pB)), pf2_hi(pZ)));
}

// WeightedClusterFit::Compress3 (weightedclusterfit.cpp:370-472): clusters at 0, 1/2, 1
NVB_DEV SquishSplit squish_eval_split3(const float4 *T, int n, int c0, int c1, float4 xsum, float mx, float my, float mz) {
    const float4 x0 = T[c0];
    const float4 x1 = T[c0 * (n + 1) - ((c0 * (c0 - 1)) >> 1) + c1];
    const float w0 = x0.w, w1 = x1.w;
    const float w2 = xsum.w - w0 - w1;
    const float alpha2_sum = w0 + w1 * 0.25f;
    const float beta2_sum = w2 + w1 * 0.25f;
    const float alphabeta_sum = w1 * 0.25f;
    const float factor = 1.0f / (alpha2_sum * beta2_sum - alphabeta_sum * alphabeta_sum);
    SquishSplit r;
    float e[3];
    const float X0[3] = {x0.x, x0.y, x0.z}, X1[3] = {x1.x, x1.y, x1.z};
    const float XS[3] = {xsum.x, xsum.y, xsum.z};
    const float grid[3] = {31.0f, 63.0f, 31.0f};
    const float gridrcp[3] = {1.0f / 31.0f, 1.0f / 63.0f, 1.0f / 31.0f};
    float A[3], B[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float alphax_sum = X0[k] + X1[k] * 0.5f;
        const float betax_sum = XS[k] - alphax_sum;
        float a = (alphax_sum * beta2_sum - betax_sum * alphabeta_sum) * factor;
        float b = (betax_sum * alpha2_sum - alphax_sum * alphabeta_sum) * factor;
        a = std_min(1.0f, std_max(0.0f, a));
        b = std_min(1.0f, std_max(0.0f, b));
        a = floorf(grid[k] * a + 0.5f) * gridrcp[k];
        b = floorf(grid[k] * b + 0.5f) * gridrcp[k];
        e[k] = a * a * alpha2_sum + b * b * beta2_sum + 2.0f * (a * b * alphabeta_sum - a * alphax_sum - b * betax_sum);
        A[k] = a;
        B[k] = b;
    }
    r.error = e[0] * mx + e[1] * my + e[2] * mz;
    r.ax = A[0]; r.ay = A[1]; r.az = A[2];
    r.bx = B[0]; r.by = B[1]; r.bz = B[2];
    return r;
}

NVB_DEV unsigned squish_to_565(float x, float y, float z) {
    return ((unsigned)squish_float_to_int(31.0f * x, 31) << 11) | ((unsigned)squish_float_to_int(63.0f * y, 63) << 5) |
           (unsigned)squish_float_to_int(31.0f * z, 31);
}

template <bool DXT1A> NVB_DEV void bc3_color_body(const Bc3ColorParams &P) {
    __shared__ Bc3GroupSmem smem[NVB_BC3_GROUPS];
    const int grp = threadIdx.x >> 4;
    const int l = threadIdx.x & 15;
    const int nblocks = P.lv.bw * P.lv.bh;
    const int blk = blockIdx.x * NVB_BC3_GROUPS + grp;
    if (blk >= nblocks) return;  // whole half-warp leaves together; only half-warp-scoped syncs follow
    const unsigned gm = 0xFFFFu << (threadIdx.x & 16);
    Bc3GroupSmem &S = smem[grp];

    // ---- gather + quantise my texel (ColorBlock::init: truncation, partial blocks repeat by modulo) ----
    const int bx = blk % P.lv.bw, by = blk / P.lv.bw;
    const int tw = min(P.lv.w - bx * 4, 4), th = min(P.lv.h - by * 4, 4);
    const int px = bx * 4 + (l & 3) % tw, py = by * 4 + (l >> 2) % th;
    unsigned r8 = 0xFF, b8 = 0;
    const unsigned g8 = quantize_u8_trunc(load_texel(P.lv, 1, px, py));
    if (!P.dxt5n) {
        r8 = quantize_u8_trunc(load_texel(P.lv, 0, px, py));
        b8 = quantize_u8_trunc(load_texel(P.lv, 2, px, py));
    }
    float walpha = 1.0f;
    unsigned a8 = 255;
    if (P.weight_by_alpha || DXT1A) {
        a8 = quantize_u8_trunc(load_texel(P.lv, 3, px, py));
        if (P.weight_by_alpha) walpha = (float)(a8 + 1) / 256.0f;
    }
    const unsigned key = (r8 << 16) | (g8 << 8) | b8;
    // kDxt1: texels with alpha 0 are left out of the colour set (m_remap = -1) and get index 3 (colourset.cpp:49-55,141-152)
    const bool opaque = !(DXT1A && a8 == 0);
    const unsigned transparent_mask = DXT1A ? ((__ballot_sync(gm, !opaque) >> (threadIdx.x & 16)) & 0xFFFFu) : 0u;

    // ---- minimal colour set: first occurrence keeps the point, duplicates add their weight in texel order ----
    unsigned match = 0;
    float weight = 0.0f;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const unsigned kj = __shfl_sync(gm, key, j, 16);
        const float wj = __shfl_sync(gm, walpha, j, 16);
        if (kj == key && opaque && !((transparent_mask >> j) & 1)) {
            weight = (match == 0) ? wj : weight + wj;
            match |= 1u << j;
        }
    }
    const int first = __ffs((int)match) - 1;  // -1 for a transparent texel
    const unsigned uniq = (__ballot_sync(gm, first == l) >> (threadIdx.x & 16)) & 0xFFFFu;
    const int n = __popc(uniq);
    const int myPoint = (first >= 0) ? __popc(uniq & ((1u << first) - 1u)) : 0;  // m_remap[l]

    unsigned char *dst = P.out + nvb_out_block(P.lv, blk) * P.out_stride + P.out_offset;

    if (DXT1A) {
        // CompressorDXT1a: rgba.isSingleColor() looks at the RGB of all 16 texels, transparent ones included
        const unsigned key0 = __shfl_sync(gm, key, 0, 16);
        const bool single = ((__ballot_sync(gm, key != key0) >> (threadIdx.x & 16)) & 0xFFFFu) == 0u;
        if (single) {
            if (l == 0) {
                unsigned c0, c1, indices = 0xaaaaaaaau;
                if (transparent_mask == 0) {  // OptimalCompress::compressDXT1
                    c0 = ((unsigned)P.omatch5[r8 * 2 + 0] << 11) | ((unsigned)P.omatch6[g8 * 2 + 0] << 5) | P.omatch5[b8 * 2 + 0];
                    c1 = ((unsigned)P.omatch5[r8 * 2 + 1] << 11) | ((unsigned)P.omatch6[g8 * 2 + 1] << 5) | P.omatch5[b8 * 2 + 1];
                    if (c0 < c1) {
                        unsigned t = c0; c0 = c1; c1 = t;
                        indices ^= 0x55555555u;
                    }
                } else {  // OptimalCompress::compressDXT1a: 3-colour mode, midpoint tables, index 3 where alpha is 0
                    c0 = ((unsigned)P.omatch5a[r8 * 2 + 0] << 11) | ((unsigned)P.omatch6a[g8 * 2 + 0] << 5) | P.omatch5a[b8 * 2 + 0];
                    c1 = ((unsigned)P.omatch5a[r8 * 2 + 1] << 11) | ((unsigned)P.omatch6a[g8 * 2 + 1] << 5) | P.omatch5a[b8 * 2 + 1];
                    if (c0 > c1) {
                        unsigned t = c0; c0 = c1; c1 = t;
                    }
                    unsigned am = 0;
                    for (int i = 0; i < 16; i++)
                        if ((transparent_mask >> i) & 1) am |= 3u << (2 * i);
                    indices |= am;
                }
                *reinterpret_cast<uint2 *>(dst) = make_uint2(c0 | (c1 << 16), indices);
            }
            return;
        }
        if (n == 0) {
            // every texel transparent (but not all the same RGB): the one split of an empty set has error 0 with both endpoints 0
            if (l == 0) *reinterpret_cast<uint2 *>(dst) = make_uint2(0u, 0xFFFFFFFFu);
            return;
        }
    }
    if (!DXT1A && n == 1) {
        // single colour: optimal endpoints from the match tables, all indices 2.
        if (l == 0 && P.dxt5n) {
            // OptimalCompress::compressDXT1G(uint8 g): red 31, blue 0, green from the 6-bit match table
            unsigned c0 = (31u << 11) | ((unsigned)P.omatch6[g8 * 2 + 0] << 5);
            unsigned c1 = (31u << 11) | ((unsigned)P.omatch6[g8 * 2 + 1] << 5);
            unsigned indices = 0xaaaaaaaau;
            if (c0 < c1) {
                unsigned t = c0; c0 = c1; c1 = t;
                indices ^= 0x55555555u;
            }
            *reinterpret_cast<uint2 *>(dst) = make_uint2(c0 | (c1 << 16), indices);
        } else if (l == 0) {
            unsigned c0 = ((unsigned)P.omatch5[r8 * 2 + 0] << 11) | ((unsigned)P.omatch6[g8 * 2 + 0] << 5) | P.omatch5[b8 * 2 + 0];
            unsigned c1 = ((unsigned)P.omatch5[r8 * 2 + 1] << 11) | ((unsigned)P.omatch6[g8 * 2 + 1] << 5) | P.omatch5[b8 * 2 + 1];
            unsigned indices = 0xaaaaaaaau;
            if (c0 < c1) {
                unsigned t = c0; c0 = c1; c1 = t;
                indices ^= 0x55555555u;
            }
            *reinterpret_cast<uint2 *>(dst) = make_uint2(c0 | (c1 << 16), indices);
        }
        return;
    }

    if (first == l && first >= 0) S.pts[myPoint] = make_float4((float)r8 / 255.0f, (float)g8 / 255.0f, (float)b8 / 255.0f, weight);
    __syncwarp(gm);

    // ---- weighted covariance about the weighted centroid (every lane, reference order) ----
    const float mx = P.metric[0], my = P.metric[1], mz = P.metric[2];
    float total = 0.0f, cx = 0.0f, cy = 0.0f, cz = 0.0f;
    for (int i = 0; i < n; i++) {
        const float4 p = S.pts[i];
        total += p.w;
        cx += p.w * p.x;
        cy += p.w * p.y;
        cz += p.w * p.z;
    }
    {
        const float t = 1.0f / total;
        cx *= t; cy *= t; cz *= t;
    }
    float cov0 = 0, cov1 = 0, cov2 = 0, cov3 = 0, cov4 = 0, cov5 = 0;
    for (int i = 0; i < n; i++) {
        const float4 p = S.pts[i];
        const float ax = (p.x - cx) * mx, ay = (p.y - cy) * my, az = (p.z - cz) * mz;
        const float bxv = p.w * ax, byv = p.w * ay, bzv = p.w * az;
        cov0 += ax * bxv;
        cov1 += ax * byv;
        cov2 += ax * bzv;
        cov3 += ay * byv;
        cov4 += ay * bzv;
        cov5 += az * bzv;
    }
    // ---- principal axis: best row estimate + 8 power iterations normalised by the max component ----
    float vx, vy, vz;
    {
        const float r0 = cov0 * cov0 + cov1 * cov1 + cov2 * cov2;
        const float r1 = cov1 * cov1 + cov3 * cov3 + cov4 * cov4;
        const float r2 = cov2 * cov2 + cov4 * cov4 + cov5 * cov5;
        if (r0 > r1 && r0 > r2) { vx = cov0; vy = cov1; vz = cov2; }
        else if (r1 > r2) { vx = cov1; vy = cov3; vz = cov4; }
        else { vx = cov2; vy = cov4; vz = cov5; }
        for (int it = 0; it < 8; it++) {
            const float x = vx * cov0 + vy * cov1 + vz * cov2;
            const float y = vx * cov1 + vy * cov3 + vz * cov4;
            const float z = vx * cov2 + vy * cov4 + vz * cov5;
            const float norm = std_max(std_max(x, y), z);
            const float iv = 1.0f / norm;
            if (norm == 0.0f) { vx = 0.0f; vy = 0.0f; vz = 0.0f; break; }
            vx = x * iv; vy = y * iv; vz = z * iv;
        }
    }
    // ---- order the points along the axis: stable ascending sort == rank by (dps, index) ----
    float dps = 0.0f;
    if (l < n) {
        const float4 p = S.pts[l];
        dps = p.x * vx + p.y * vy + p.z * vz;
    }
    int rank = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const float dj = __shfl_sync(gm, dps, j, 16);
        if (j < n && (dj < dps || (dj == dps && j < l))) rank++;
    }
    if (l < n) {
        const float4 p = S.pts[l];
        S.sorted[rank] = make_float4(p.w * p.x, p.w * p.y, p.w * p.z, p.w);
        S.rank[l] = rank;
    }
    __syncwarp(gm);
    float4 xsum = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    for (int i = 0; i < n; i++) {
        const float4 s = S.sorted[i];
        xsum.x += s.x; xsum.y += s.y; xsum.z += s.z; xsum.w += s.w;
    }
    // ---- running-sum table, row `start` accumulated front to back like the loop-carried x0/x1/x2 ----
    for (int start = l; start <= n; start += 16) {
        const int off = start * (n + 1) - ((start * (start - 1)) >> 1);
        float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        S.T[off] = acc;
        for (int k = 0; start + k < n; k++) {
            const float4 s = S.sorted[start + k];
            acc.x += s.x; acc.y += s.y; acc.z += s.z; acc.w += s.w;
            S.T[off + k + 1] = acc;
        }
    }
    __syncwarp(gm);

    // ---- all cluster splits, striped over the 16 lanes ----
    const float mqx = mx * mx, mqy = my * my, mqz = mz * mz;
    const int cbeg = P.cand_off[n], ncand = P.cand_off[n + 1] - cbeg;
    float besterror3 = FLT_MAX;
    int bestci3 = 0x7fffffff;
    if (DXT1A) {
        const int c3beg = P.cand3_off[n], n3 = P.cand3_off[n + 1] - c3beg;
        for (int ci = l; ci < n3; ci += 16) {
            const unsigned pk = __ldg(P.cand3 + c3beg + ci);
            const SquishSplit s = squish_eval_split3(S.T, n, pk & 31, (pk >> 5) & 31, xsum, mqx, mqy, mqz);
            if (s.error < besterror3) {
                besterror3 = s.error;
                bestci3 = ci;
            }
        }
#pragma unroll
        for (int d = 8; d >= 1; d >>= 1) {
            const float oe = __shfl_xor_sync(gm, besterror3, d, 16);
            const int oc = __shfl_xor_sync(gm, bestci3, d, 16);
            if (oe < besterror3 || (oe == besterror3 && oc < bestci3)) {
                besterror3 = oe;
                bestci3 = oc;
            }
        }
    }
    float besterror = FLT_MAX;
    int bestci = 0x7fffffff;
    const bool try4 = !DXT1A || transparent_mask == 0;  // ColourFit::Compress: Compress4 only when nothing is transparent
    // two splits per trip (ci and ci + 16); the last, unpaired one re-uses split ci for the idle half
    for (int ci = l; try4 && ci < ncand; ci += 32) {
        const bool two = ci + 16 < ncand;
        const unsigned pa = __ldg(P.cand_idx + cbeg + ci);
        const unsigned pb = two ? __ldg(P.cand_idx + cbeg + ci + 16) : pa;
        const float2 err = squish_eval_pair(S.T, pa, pb, xsum, mqx, mqy, mqz);
        if (err.x < besterror) {
            besterror = err.x;
            bestci = ci;
        }
        if (two && err.y < besterror) {
            besterror = err.y;
            bestci = ci + 16;
        }
    }
#pragma unroll
    for (int d = 8; d >= 1; d >>= 1) {
        const float oe = __shfl_xor_sync(gm, besterror, d, 16);
        const int oc = __shfl_xor_sync(gm, bestci, d, 16);
        if (oe < besterror || (oe == besterror && oc < bestci)) {
            besterror = oe;
            bestci = oc;
        }
    }
    unsigned c565a = 0, c565b = 0, idx = 0;
    // Compress4 replaces the 3-colour result only when strictly better (m_besterror carries over, weightedclusterfit.cpp:555)
    if (DXT1A && bestci3 != 0x7fffffff && !(bestci != 0x7fffffff && besterror < besterror3)) {
        const unsigned pk = __ldg(P.cand3 + P.cand3_off[n] + bestci3);
        const int b0 = pk & 31, b1 = (pk >> 5) & 31;
        const SquishSplit s = squish_eval_split3(S.T, n, b0, b1, xsum, mqx, mqy, mqz);
        if (first >= 0) {
            const int pos = S.rank[myPoint];
            idx = (pos < b0) ? 0u : (pos < b0 + b1) ? 2u : 1u;
        }
        c565a = squish_to_565(s.ax, s.ay, s.az);
        c565b = squish_to_565(s.bx, s.by, s.bz);
        if (c565a > c565b) {  // WriteColourBlock3: a <= b keeps the indices, otherwise swap and exchange 0 <-> 1
            const unsigned t = c565a; c565a = c565b; c565b = t;
            idx = (idx == 0) ? 1u : (idx == 1) ? 0u : idx;
        }
        if (first < 0) idx = 3;
    } else if (bestci != 0x7fffffff) {
        const unsigned pk = __ldg(P.cand + cbeg + bestci);
        const int b0 = pk & 31, b1 = (pk >> 5) & 31, b2 = (pk >> 10) & 31;
        const unsigned pi = __ldg(P.cand_idx + cbeg + bestci);
        const SquishSplit s = squish_eval_split(S.T, pi & 0xFF, (pi >> 8) & 0xFF, pi >> 16, xsum, mqx, mqy, mqz);
        const int pos = S.rank[myPoint];
        idx = (pos < b0) ? 0u : (pos < b0 + b1) ? 2u : (pos < b0 + b1 + b2) ? 3u : 1u;
        c565a = ((unsigned)squish_float_to_int(31.0f * s.ax, 31) << 11) | ((unsigned)squish_float_to_int(63.0f * s.ay, 63) << 5) |
                (unsigned)squish_float_to_int(31.0f * s.az, 31);
        c565b = ((unsigned)squish_float_to_int(31.0f * s.bx, 31) << 11) | ((unsigned)squish_float_to_int(63.0f * s.by, 63) << 5) |
                (unsigned)squish_float_to_int(31.0f * s.bz, 31);
        if (c565a < c565b) {
            const unsigned t = c565a; c565a = c565b; c565b = t;
            idx = (idx ^ 1u) & 3u;
        } else if (c565a == c565b) {
            idx = 0;
        }
    }
    unsigned bits = idx << (2 * l);
#pragma unroll
    for (int d = 8; d >= 1; d >>= 1) bits |= __shfl_xor_sync(gm, bits, d, 16);
    if (l == 0) *reinterpret_cast<uint2 *>(dst) = make_uint2(c565a | (c565b << 16), bits);
}

__global__ void __launch_bounds__(NVB_BC3_GROUPS * 16) k_bc3_color(Bc3ColorParams P) { bc3_color_body<false>(P); }
// BC1a (Normal and above): 3-colour + 4-colour weighted cluster fit with punch-through alpha
__global__ void __launch_bounds__(NVB_BC3_GROUPS * 16) k_bc1a_color(Bc3ColorParams P) { bc3_color_body<true>(P); }

}  // namespace nvb
