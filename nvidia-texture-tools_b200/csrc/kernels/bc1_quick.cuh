// Quality_Fastest colour blocks of BC1a / BC2 / BC3 / BC3n — replaces, bit-exactly:
//   QuickCompress::compressDXT1            src/nvtt/QuickCompressDXT.cpp:653-713  (bounding box, diagonal, inset, 565 rounding,
//                                                                                    nearest-palette indices, one LS refit)
//   QuickCompress::compressDXT1a           src/nvtt/QuickCompressDXT.cpp:716-766  (3-colour mode when a texel has alpha 0)
//   findMinMaxColorsBox / selectDiagonal / insetBBox / roundAndExpand / computeIndices4 / computeIndices3 /
//   optimizeEndPoints4                     src/nvtt/QuickCompressDXT.cpp:70-151,205-237,376-449
//   OptimalCompress::compressDXT1(Color32) src/nvtt/OptimalCompressDXT.cpp:254-269 (single-colour blocks)
//   FastCompressorDXT1a/DXT3/DXT5/DXT5n    src/nvtt/CompressorDX9.cpp:55-81
// One thread per block (~1k fp32/int ops): the kernel is bound by the 16 B/px planar fp32 read, not by arithmetic.
#pragma once
#include "../nvb_common.cuh"

namespace nvb {

struct Dxt1QuickParams {
    LevelView lv;
    unsigned char *out;
    int out_stride, out_offset;
    int dxt1a;   // 1: compressDXT1a (looks at alpha), 0: compressDXT1
    int dxt5n;   // 1: tile swizzled to (0xFF, G, 0) first (FastCompressorDXT5n)
    const unsigned char *omatch5;  // [256][2]
    const unsigned char *omatch6;  // [256][2]
};

struct QVec3 {
    float x, y, z;
};

// roundAndExpand: nearest 565 colour of v (0..255 per channel), v is replaced by the expanded 8-bit colour
NVB_DEV unsigned quick_round_and_expand(QVec3 *v) {
    unsigned r = (unsigned)__float2int_rn(floorf(nv_clamp(v->x * (31.0f / 255.0f), 0.0f, 31.0f)));
    unsigned g = (unsigned)__float2int_rn(floorf(nv_clamp(v->y * (63.0f / 255.0f), 0.0f, 63.0f)));
    unsigned b = (unsigned)__float2int_rn(floorf(nv_clamp(v->z * (31.0f / 255.0f), 0.0f, 31.0f)));
    const float r0 = (float)(((r + 0) << 3) | ((r + 0) >> 2)), r1 = (float)(((r + 1) << 3) | ((r + 1) >> 2));
    if (fabsf(v->x - r1) < fabsf(v->x - r0)) r = min(r + 1, 31u);
    const float g0 = (float)(((g + 0) << 2) | ((g + 0) >> 4)), g1 = (float)(((g + 1) << 2) | ((g + 1) >> 4));
    if (fabsf(v->y - g1) < fabsf(v->y - g0)) g = min(g + 1, 63u);
    const float b0 = (float)(((b + 0) << 3) | ((b + 0) >> 2)), b1 = (float)(((b + 1) << 3) | ((b + 1) >> 2));
    if (fabsf(v->z - b1) < fabsf(v->z - b0)) b = min(b + 1, 31u);
    const unsigned w = (r << 11) | (g << 5) | b;
    r = (r << 3) | (r >> 2);
    g = (g << 2) | (g >> 4);
    b = (b << 3) | (b >> 2);
    v->x = (float)r;
    v->y = (float)g;
    v->z = (float)b;
    return w & 0xffffu;
}

NVB_DEV float quick_dist(const QVec3 &a, const QVec3 &b) {
    const float x = a.x - b.x, y = a.y - b.y, z = a.z - b.z;
    return x * x + y * y + z * z;
}

NVB_DEV QVec3 quick_lerp(const QVec3 &a, const QVec3 &b, float t) {
    const float s = 1.0f - t;
    QVec3 r = {a.x * s + t * b.x, a.y * s + t * b.y, a.z * s + t * b.z};
    return r;
}

NVB_DEV unsigned quick_indices4(const QVec3 block[16], const QVec3 &maxc, const QVec3 &minc) {
    const QVec3 p2 = quick_lerp(maxc, minc, 1.0f / 3.0f), p3 = quick_lerp(maxc, minc, 2.0f / 3.0f);
    unsigned indices = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const float d0 = quick_dist(maxc, block[i]), d1 = quick_dist(minc, block[i]), d2 = quick_dist(p2, block[i]), d3 = quick_dist(p3, block[i]);
        const unsigned b0 = d0 > d3, b1 = d1 > d2, b2 = d0 > d2, b3 = d1 > d3, b4 = d2 > d3;
        const unsigned x0 = b1 & b2, x1 = b0 & b3, x2 = b0 & b4;
        indices |= (x2 | ((x0 | x1) << 1)) << (2 * i);
    }
    return indices;
}

// palette[0] = minColor, [1] = maxColor, [2] = midpoint; n valid colours (texels with alpha > 127, packed first)
NVB_DEV unsigned quick_indices3(const QVec3 block[16], const QVec3 &maxc, const QVec3 &minc) {
    const QVec3 mid = {(minc.x + maxc.x) * 0.5f, (minc.y + maxc.y) * 0.5f, (minc.z + maxc.z) * 0.5f};
    unsigned indices = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const float d0 = quick_dist(minc, block[i]), d1 = quick_dist(maxc, block[i]), d2 = quick_dist(mid, block[i]);
        unsigned index;
        if (d0 < d1 && d0 < d2) index = 0;
        else if (d1 < d2) index = 1;
        else index = 2;
        indices |= index << (2 * i);
    }
    return indices;
}

// bounding box -> diagonal -> inset -> 565 (the shared front half of compressDXT1 / compressDXT1a)
NVB_DEV void quick_bbox_endpoints(const QVec3 *block, int num, QVec3 *maxc, QVec3 *minc) {
    QVec3 mx = {0, 0, 0}, mn = {255, 255, 255};
    for (int i = 0; i < num; i++) {
        mx.x = nv_max(mx.x, block[i].x); mx.y = nv_max(mx.y, block[i].y); mx.z = nv_max(mx.z, block[i].z);
        mn.x = nv_min(mn.x, block[i].x); mn.y = nv_min(mn.y, block[i].y); mn.z = nv_min(mn.z, block[i].z);
    }
    // selectDiagonal
    const QVec3 center = {(mx.x + mn.x) * 0.5f, (mx.y + mn.y) * 0.5f, (mx.z + mn.z) * 0.5f};
    float cx = 0.0f, cy = 0.0f;
    for (int i = 0; i < num; i++) {
        const float tx = block[i].x - center.x, ty = block[i].y - center.y, tz = block[i].z - center.z;
        cx += tx * tz;
        cy += ty * tz;
    }
    float x0 = mx.x, y0 = mx.y, x1 = mn.x, y1 = mn.y;
    if (cx < 0) { const float t = x0; x0 = x1; x1 = t; }
    if (cy < 0) { const float t = y0; y0 = y1; y1 = t; }
    mx.x = x0; mx.y = y0;
    mn.x = x1; mn.y = y1;
    // insetBBox
    const float k = (8.0f / 255.0f) / 16.0f;
    const QVec3 inset = {(mx.x - mn.x) / 16.0f - k, (mx.y - mn.y) / 16.0f - k, (mx.z - mn.z) / 16.0f - k};
    maxc->x = nv_clamp(mx.x - inset.x, 0.0f, 255.0f);
    maxc->y = nv_clamp(mx.y - inset.y, 0.0f, 255.0f);
    maxc->z = nv_clamp(mx.z - inset.z, 0.0f, 255.0f);
    minc->x = nv_clamp(mn.x + inset.x, 0.0f, 255.0f);
    minc->y = nv_clamp(mn.y + inset.y, 0.0f, 255.0f);
    minc->z = nv_clamp(mn.z + inset.z, 0.0f, 255.0f);
}

__global__ void __launch_bounds__(128) k_dxt1_quick(Dxt1QuickParams P) {
    const int nblocks = P.lv.bw * P.lv.bh;
    for (int blk = blockIdx.x * blockDim.x + threadIdx.x; blk < nblocks; blk += gridDim.x * blockDim.x) {
        const int bx = blk % P.lv.bw, by = blk / P.lv.bw;
        const int tw = min(P.lv.w - bx * 4, 4), th = min(P.lv.h - by * 4, 4);
        unsigned rgb[16], alpha[16];
        bool single = true, has_alpha0 = false;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const int px = bx * 4 + (i & 3) % tw, py = by * 4 + (i >> 2) % th;  // ColorBlock::init repeats texels by modulo
            unsigned r = 0xFF, b = 0;
            const unsigned g = quantize_u8_trunc(load_texel(P.lv, 1, px, py));
            if (!P.dxt5n) {
                r = quantize_u8_trunc(load_texel(P.lv, 0, px, py));
                b = quantize_u8_trunc(load_texel(P.lv, 2, px, py));
            }
            rgb[i] = (r << 16) | (g << 8) | b;
            alpha[i] = 255;
            if (P.dxt1a) {
                alpha[i] = quantize_u8_trunc(load_texel(P.lv, 3, px, py));
                has_alpha0 = has_alpha0 || alpha[i] == 0;
            }
            single = single && rgb[i] == rgb[0];
        }
        unsigned c0, c1, indices;
        if (!(P.dxt1a && has_alpha0)) {
            if (single) {
                const unsigned r8 = rgb[0] >> 16, g8 = (rgb[0] >> 8) & 0xFF, b8 = rgb[0] & 0xFF;
                c0 = ((unsigned)P.omatch5[r8 * 2 + 0] << 11) | ((unsigned)P.omatch6[g8 * 2 + 0] << 5) | P.omatch5[b8 * 2 + 0];
                c1 = ((unsigned)P.omatch5[r8 * 2 + 1] << 11) | ((unsigned)P.omatch6[g8 * 2 + 1] << 5) | P.omatch5[b8 * 2 + 1];
                indices = 0xaaaaaaaau;
                if (c0 < c1) {
                    const unsigned t = c0; c0 = c1; c1 = t;
                    indices ^= 0x55555555u;
                }
            } else {
                QVec3 block[16];
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    block[i].x = (float)(rgb[i] >> 16);
                    block[i].y = (float)((rgb[i] >> 8) & 0xFF);
                    block[i].z = (float)(rgb[i] & 0xFF);
                }
                QVec3 maxc, minc;
                quick_bbox_endpoints(block, 16, &maxc, &minc);
                c0 = quick_round_and_expand(&maxc);
                c1 = quick_round_and_expand(&minc);
                if (c0 < c1) {
                    const QVec3 tv = maxc; maxc = minc; minc = tv;
                    const unsigned t = c0; c0 = c1; c1 = t;
                }
                indices = quick_indices4(block, maxc, minc);
                // optimizeEndPoints4: one least-squares refit for these indices
                float alpha2_sum = 0.0f, beta2_sum = 0.0f, alphabeta_sum = 0.0f;
                QVec3 ax = {0, 0, 0}, bxs = {0, 0, 0};
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const unsigned bits = indices >> (2 * i);
                    float beta = (float)(bits & 1);
                    if (bits & 2) beta = (1 + beta) / 3.0f;
                    const float al = 1.0f - beta;
                    alpha2_sum += al * al;
                    beta2_sum += beta * beta;
                    alphabeta_sum += al * beta;
                    ax.x += block[i].x * al; ax.y += block[i].y * al; ax.z += block[i].z * al;
                    bxs.x += block[i].x * beta; bxs.y += block[i].y * beta; bxs.z += block[i].z * beta;
                }
                const float denom = alpha2_sum * beta2_sum - alphabeta_sum * alphabeta_sum;
                // nv::equal(denom, 0.0f): |denom| <= 1e-4 * max(1, |denom|, 0)
                const float m3 = nv_max(nv_max(1.0f, fabsf(denom)), 0.0f);
                if (!(fabsf(denom - 0.0f) <= 0.0001f * m3)) {
                    const float factor = 1.0f / denom;
                    QVec3 a = {(ax.x * beta2_sum - bxs.x * alphabeta_sum) * factor, (ax.y * beta2_sum - bxs.y * alphabeta_sum) * factor,
                               (ax.z * beta2_sum - bxs.z * alphabeta_sum) * factor};
                    QVec3 b = {(bxs.x * alpha2_sum - ax.x * alphabeta_sum) * factor, (bxs.y * alpha2_sum - ax.y * alphabeta_sum) * factor,
                               (bxs.z * alpha2_sum - ax.z * alphabeta_sum) * factor};
                    a.x = nv_clamp(a.x, 0.0f, 255.0f); a.y = nv_clamp(a.y, 0.0f, 255.0f); a.z = nv_clamp(a.z, 0.0f, 255.0f);
                    b.x = nv_clamp(b.x, 0.0f, 255.0f); b.y = nv_clamp(b.y, 0.0f, 255.0f); b.z = nv_clamp(b.z, 0.0f, 255.0f);
                    unsigned k0 = quick_round_and_expand(&a), k1 = quick_round_and_expand(&b);
                    if (k0 < k1) {
                        const QVec3 tv = a; a = b; b = tv;
                        const unsigned t = k0; k0 = k1; k1 = t;
                    }
                    c0 = k0;
                    c1 = k1;
                    indices = quick_indices4(block, a, b);
                }
            }
        } else {
            // compressDXT1a with at least one fully transparent texel: 3-colour mode over the texels with alpha > 127
            QVec3 block[16];
            int num = 0;
#pragma unroll
            for (int i = 0; i < 16; i++) {
                if (alpha[i] > 127) {
                    block[num].x = (float)(rgb[i] >> 16);
                    block[num].y = (float)((rgb[i] >> 8) & 0xFF);
                    block[num].z = (float)(rgb[i] & 0xFF);
                    num++;
                }
            }
            // the reference leaves block[num..15] uninitialised and still indexes all 16 entries in computeIndices3;
            // canonical behaviour here: the unused tail is zero (= what a zero-filled stack frame gives)
            for (int i = num; i < 16; i++) block[i].x = block[i].y = block[i].z = 0.0f;
            QVec3 maxc, minc;
            quick_bbox_endpoints(block, num, &maxc, &minc);
            unsigned k0 = quick_round_and_expand(&maxc), k1 = quick_round_and_expand(&minc);
            if (k0 < k1) {
                const QVec3 tv = maxc; maxc = minc; minc = tv;
                const unsigned t = k0; k0 = k1; k1 = t;
            }
            c0 = k1;
            c1 = k0;
            indices = quick_indices3(block, maxc, minc);
        }
        *reinterpret_cast<uint2 *>(P.out + nvb_out_block(P.lv, blk) * P.out_stride + P.out_offset) = make_uint2(c0 | (c1 << 16), indices);
    }
}

}  // namespace nvb
