// BC4 / BC5 / BC3-alpha block encoders (the "DXT5 alpha block": two 8-bit endpoints + 16 x 3-bit indices).
//
// Replaces, bit-exactly:
//   QuickCompress::compressDXT5A            src/nvtt/QuickCompressDXT.cpp:779-821  (Fastest, Normal)
//     computeAlphaIndices                   src/nvtt/QuickCompressDXT.cpp:505-536
//     optimizeAlpha8                        src/nvtt/QuickCompressDXT.cpp:538-592
//     sameIndices                           src/nvtt/QuickCompressDXT.cpp:643-647
//   AlphaBlockDXT5::evaluatePalette8/6      src/nvimage/BlockDXT.cpp:336-378, layout BlockDXT.h:123-162
//   FastCompressorBC4/BC5::compressBlock    src/nvtt/CompressorDX10.cpp:42-62
//
// One thread per channel-block: the work is ~2k integer/fp32 ops per block, i.e. the kernel is HBM-bound
// (16 B/px read for BC5 from the planar fp32 level, 1 B/px written).
#pragma once
#include "../nvb_common.cuh"

namespace nvb {

// 8-entry palette of an alpha block (D3D10 rounding, bias 0).
NVB_DEV void alpha_palette(unsigned a0, unsigned a1, unsigned pal[8]) {
    pal[0] = a0;
    pal[1] = a1;
    if (a0 > a1) {
        pal[2] = (6 * a0 + 1 * a1) / 7;
        pal[3] = (5 * a0 + 2 * a1) / 7;
        pal[4] = (4 * a0 + 3 * a1) / 7;
        pal[5] = (3 * a0 + 4 * a1) / 7;
        pal[6] = (2 * a0 + 5 * a1) / 7;
        pal[7] = (1 * a0 + 6 * a1) / 7;
    } else {
        pal[2] = (4 * a0 + 1 * a1) / 5;
        pal[3] = (3 * a0 + 2 * a1) / 5;
        pal[4] = (2 * a0 + 3 * a1) / 5;
        pal[5] = (1 * a0 + 4 * a1) / 5;
        pal[6] = 0x00;
        pal[7] = 0xFF;
    }
}

// minimum of the eight keys alpha * B[p] + A[p]: three-input minimum (VIMNMX3, one instruction on sm_90+) - four
// instructions instead of the seven of a chain of two-input minima
NVB_DEV int alpha_min8(int alpha, const int A[8], const int B[8]) {
    int k[8];
#pragma unroll
    for (int p = 0; p < 8; p++) k[p] = alpha * B[p] + A[p];
#ifdef NVB_EMU
    int m = k[0];
    for (int p = 1; p < 8; p++) m = min(m, k[p]);
    return m;
#else
    return min(__vimin3_s32(__vimin3_s32(k[0], k[1], k[2]), k[3], k[4]), __vimin3_s32(k[5], k[6], k[7]));
#endif
}

// Exhaustive nearest-palette-entry search; first minimum wins (strict <).  Returns the summed squared error and
// writes the 16 3-bit indices into bits [16,64) of *blk (bits [0,16) = endpoints are left untouched).
NVB_DEV unsigned alpha_compute_indices(const unsigned src[16], unsigned a0, unsigned a1, unsigned long long *blk) {
    unsigned pal[8];
    alpha_palette(a0, a1, pal);
    // The reference scans the eight entries and keeps the first strictly smaller squared distance.  Same result without the
    // compare / select chain: key_p = 8 * (pal_p - a)^2 + p orders by error first and by entry number among equal errors, and
    // key_p = (8 pal_p^2 + p) - 16 pal_p * a + 8 a^2, whose last term does not depend on p - so one multiply-add per entry
    // and a running minimum (exact integer arithmetic, |terms| < 2^21).
    int A[8], B[8];
#pragma unroll
    for (int p = 0; p < 8; p++) {
        A[p] = (int)(pal[p] * pal[p] * 8u) + p;
        B[p] = -16 * (int)pal[p];
    }
    unsigned total = 0;
    unsigned long long bits = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const int alpha = (int)src[i];
        const int m = alpha_min8(alpha, A, B);
        const unsigned key = (unsigned)(m + 8 * alpha * alpha);
        total += key >> 3;
        bits |= (unsigned long long)(key & 7u) << (3 * i);
    }
    *blk = (*blk & 0xFFFFull) | (bits << 16);
    return total;
}

// k / 7.0f for k = 1..6, correctly rounded (i.e. bit-identical to the IEEE division of the reference) without the division
// sequence: q = k * RN(1/7), r = fma(-7, q, k) (exact), q' = fma(r, RN(1/7), q).  Checked against k / 7.0f for the six values
// it is used for (tests/test_capi_boundary.py: test_div7_identity); the plain product alone is off by one ulp for k = 3 and 6.
NVB_DEV float alpha_div7(float k) {
    const float y = 1.0f / 7.0f;  // folded at compile time, correctly rounded
    const float q = __fmul_rn(k, y);
    const float r = __fmaf_rn(-7.0f, q, k);
    return __fmaf_rn(r, y, q);
}

// Least-squares endpoint refit for the current indices (always with the 8-step weights, also when the block is
// in 6-step mode — that is what the reference does).
NVB_DEV void alpha_optimize8(const unsigned src[16], unsigned long long *blk) {
    float alpha2_sum = 0, beta2_sum = 0, alphabeta_sum = 0, alphax_sum = 0, betax_sum = 0;
    unsigned long long bits = *blk >> 16;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        unsigned idx = (unsigned)(bits >> (3 * i)) & 7u;
        float alpha;
        if (idx < 2) alpha = 1.0f - (float)idx;
        else alpha = alpha_div7(8.0f - (float)idx);
        float beta = 1 - alpha;
        float x = (float)src[i];
        alpha2_sum += alpha * alpha;
        beta2_sum += beta * beta;
        alphabeta_sum += alpha * beta;
        alphax_sum += alpha * x;
        betax_sum += beta * x;
    }
    const float factor = 1.0f / (alpha2_sum * beta2_sum - alphabeta_sum * alphabeta_sum);
    float a = (alphax_sum * beta2_sum - betax_sum * alphabeta_sum) * factor;
    float b = (betax_sum * alpha2_sum - alphax_sum * alphabeta_sum) * factor;
    // uint(min(max(a, 0.0f), 255.0f)) with nv::min/max NaN semantics (NaN -> 0).
    unsigned alpha0 = (unsigned)(int)nv_min(nv_max(a, 0.0f), 255.0f);
    unsigned alpha1 = (unsigned)(int)nv_min(nv_max(b, 0.0f), 255.0f);
    if (alpha0 < alpha1) {
        unsigned t = alpha0;
        alpha0 = alpha1;
        alpha1 = t;
        unsigned long long nb = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            unsigned idx = (unsigned)(bits >> (3 * i)) & 7u;
            unsigned n = (idx < 2) ? (1 - idx) : (9 - idx);
            nb |= (unsigned long long)(n & 7) << (3 * i);
        }
        bits = nb;
    } else if (alpha0 == alpha1) {
        bits = 0;
    }
    *blk = (bits << 16) | ((unsigned long long)alpha1 << 8) | alpha0;
}

// QuickCompress::compressDXT5A with iterationCount = 8.
NVB_DEV unsigned long long alpha_quick_compress(const unsigned src[16]) {
    unsigned amax = 0, amin = 255;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        amax = max(amax, src[i]);
        amin = min(amin, src[i]);
    }
    // uint8 arithmetic promoted to int; results stored back into 8-bit fields.
    unsigned a0 = (amax - (amax - amin) / 34) & 0xFF;
    unsigned a1 = (amin + (amax - amin) / 34) & 0xFF;
    unsigned long long block = ((unsigned long long)a1 << 8) | a0;
    unsigned besterror = alpha_compute_indices(src, a0, a1, &block);
    unsigned long long best = block;
    for (int it = 0; it < 8; it++) {
        alpha_optimize8(src, &block);
        unsigned error = alpha_compute_indices(src, (unsigned)(block & 0xFF), (unsigned)((block >> 8) & 0xFF), &block);
        if (error >= besterror) break;
        if ((block >> 16) == (best >> 16)) {
            best = block;
            break;
        }
        besterror = error;
        best = block;
    }
    return best;
}

// One thread per 4x4 block of one channel.  `channel` is the FloatImage plane (0=R,1=G,2=B,3=A): BC4 -> R,
// BC5 -> R then G (second launch, out_offset 8), BC3 alpha -> A.  mode 0 = QuickCompress (Fastest/Normal).
struct AlphaBlocksParams {
    LevelView lv;
    int channel;
    unsigned char *out;
    int out_stride;  // bytes between consecutive blocks in the output (8 for BC4, 16 for BC5/BC3)
    int out_offset;  // byte offset of this alpha block inside the output block
    int mode;        // 0 quick, 1 optimal (Production / Highest)
    const unsigned char *omatch6 = nullptr;  // k_dxt1g_optimal only: OMatch6[g][0..1] (single-colour green)
};

NVB_DEV void alpha_gather_block(const LevelView &lv, int channel, int bx, int by, unsigned src[16]) {
    const int x0 = bx * 4, y0 = by * 4;
    const float *plane = lv.data + (size_t)channel * lv.plane;
    const bool gam = (lv.to_gamma_table != nullptr && channel < 3);
    // 128-bit row loads need 16-byte aligned rows: width a multiple of 4 and an aligned plane (callers may pass any device pointer)
    if (x0 + 4 <= lv.w && y0 + 4 <= lv.h && (lv.w & 3) == 0 && ((size_t)plane & 15) == 0) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float4 v = *reinterpret_cast<const float4 *>(plane + (size_t)(y0 + i) * lv.w + x0);
            float t[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; e++) {
                float f = t[e];
                if (gam) f = nvb_powf_5_11(f, lv.to_gamma_table);
                src[i * 4 + e] = quantize_u8_trunc(f);
            }
        }
    } else {
        const int tw = min(lv.w - x0, 4), th = min(lv.h - y0, 4);
#pragma unroll
        for (int i = 0; i < 4; i++) {
#pragma unroll
            for (int e = 0; e < 4; e++) {
                float f = plane[(size_t)(y0 + i % th) * lv.w + x0 + e % tw];
                if (gam) f = nvb_powf_5_11(f, lv.to_gamma_table);
                src[i * 4 + e] = quantize_u8_trunc(f);
            }
        }
    }
}

// The refit loop of QuickCompress::compressDXT5A runs 1 to 8 times per block (early-outs), so a warp that gives every lane
// one block and waits for the slowest runs half empty (16.8 active lanes measured).  Here a lane that finishes its block
// stores it and gathers its next one while the others keep iterating: one loop trip = [refit] + exhaustive index search for
// every lane that has a block (same functions, same order of operations per block as alpha_quick_compress).
__global__ void __launch_bounds__(128) k_alpha_blocks(AlphaBlocksParams P) {
    const int nblocks = P.lv.bw * P.lv.bh;
    const int stride = gridDim.x * blockDim.x;
    int blk = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned src[16];
    unsigned long long block = 0, best = 0;
    unsigned besterror = 0;
    int it = -1;  // -1: the initial index search of a fresh block
    bool have = false;
    for (;;) {
        if (!have && blk < nblocks) {
            alpha_gather_block(P.lv, P.channel, blk % P.lv.bw, blk / P.lv.bw, src);
            unsigned amax = 0, amin = 255;
#pragma unroll
            for (int i = 0; i < 16; i++) {
                amax = max(amax, src[i]);
                amin = min(amin, src[i]);
            }
            // uint8 arithmetic promoted to int; results stored back into 8-bit fields
            const unsigned a0 = (amax - (amax - amin) / 34) & 0xFF, a1 = (amin + (amax - amin) / 34) & 0xFF;
            block = ((unsigned long long)a1 << 8) | a0;
            it = -1;
            have = true;
        }
        if (__all_sync(0xffffffffu, !have)) break;  // every lane stays until the warp has run out of blocks
        if (have) {
            if (it >= 0) alpha_optimize8(src, &block);
            const unsigned error = alpha_compute_indices(src, (unsigned)(block & 0xFF), (unsigned)((block >> 8) & 0xFF), &block);
            bool done = false;
            if (it < 0) {
                besterror = error;
                best = block;
                it = 0;
            } else if (error >= besterror) {
                done = true;
            } else if ((block >> 16) == (best >> 16)) {
                best = block;
                done = true;
            } else {
                besterror = error;
                best = block;
                done = ++it >= 8;
            }
            if (done) {
                *reinterpret_cast<uint2 *>(P.out + nvb_out_block(P.lv, blk) * P.out_stride + P.out_offset) =
                    make_uint2((unsigned)(best & 0xFFFFFFFFu), (unsigned)(best >> 32));
                blk += stride;
                have = false;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// OptimalCompress::compressDXT5A (src/nvtt/OptimalCompressDXT.cpp:512-607; computeAlphaError :189-217,
// computeAlphaIndices :219-244) — BC4/BC5 at Production/Highest, BC3 alpha at Highest.
// Brute force over every (alpha0, alpha1) pair of the 8-step and then the 6-step encoding; the error of a pair is
// an exact integer (weights are 1), so the early-outs of the reference only save time and the result is the first
// strict minimum in loop order.  One warp per channel-block: the pairs are flattened in loop order and striped
// over the 32 lanes, then reduced with (error, order number).
// Canonical behaviour for the reference's read of uninitialised output bytes (:546, besterror is first computed
// from whatever the output buffer held): the output is treated as zero-filled (alpha0 = alpha1 = 0).
// ---------------------------------------------------------------------------------------------------------
NVB_DEV unsigned alpha_pair_error(const unsigned src[16], unsigned a0, unsigned a1) {
    unsigned pal[8];
    alpha_palette(a0, a1, pal);
    // min_p (a - pal_p)^2 = a^2 + min_p (pal_p^2 - 2 pal_p a): one multiply-add per entry and a running minimum
    int A[8], B[8];
#pragma unroll
    for (int p = 0; p < 8; p++) {
        A[p] = (int)(pal[p] * pal[p]);
        B[p] = -2 * (int)pal[p];
    }
    unsigned total = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const int alpha = (int)src[i];
        const int m = alpha_min8(alpha, A, B);
        total += (unsigned)(m + alpha * alpha);
    }
    return total;
}

__global__ void __launch_bounds__(128) k_alpha_optimal(AlphaBlocksParams P) {
    const int lane = threadIdx.x & 31;
    const int warps_per_cta = blockDim.x >> 5;
    const int nblocks = P.lv.bw * P.lv.bh;
    for (int blk = blockIdx.x * warps_per_cta + (threadIdx.x >> 5); blk < nblocks; blk += gridDim.x * warps_per_cta) {
        unsigned src[16];
        alpha_gather_block(P.lv, P.channel, blk % P.lv.bw, blk / P.lv.bw, src);  // every lane holds the whole block
        int mina = 255, maxa = 0, mina_no01 = 255, maxa_no01 = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const int a = (int)src[i];
            mina = min(mina, a);
            maxa = max(maxa, a);
            if (a != 0 && a != 255) {
                mina_no01 = min(mina_no01, a);
                maxa_no01 = max(maxa_no01, a);
            }
        }
        unsigned a0, a1;
        if (maxa - mina < 8) {
            a0 = (unsigned)maxa;
            a1 = (unsigned)mina;
        } else if (maxa_no01 - mina_no01 < 6) {
            a0 = (unsigned)mina_no01;
            a1 = (unsigned)maxa_no01;
        } else {
            unsigned besterror = alpha_pair_error(src, 0, 0);  // zero-filled output block
            unsigned bestk = 0xffffffffu;                      // "no candidate improved": keep (maxa, mina)
            unsigned bestpair = ((unsigned)maxa << 8) | (unsigned)mina;  // alpha0 << 8 | alpha1
            // 8-step pairs: for a0 in [lo+9, hi) for a1 in [lo, a0-8)
            const int lo8 = (mina <= 8) ? 0 : mina - 8, hi8 = (maxa >= 255 - 8) ? 255 : maxa + 8;
            const int R8 = hi8 - lo8;
            const int n8 = (R8 > 9) ? ((R8 - 9) * (R8 - 8)) / 2 : 0;
            const int lo6 = (mina_no01 <= 6) ? 0 : mina_no01 - 6, hi6 = (maxa_no01 >= 255 - 6) ? 255 : maxa_no01 + 6;
            const int R6 = hi6 - lo6;
            const int n6 = (R6 > 9) ? ((R6 - 9) * (R6 - 8)) / 2 : 0;
#pragma unroll 1
            for (int pass = 0; pass < 2; pass++) {
                const int lo = pass ? lo6 : lo8, hi = pass ? hi6 : hi8, n = pass ? n6 : n8;
                const unsigned kbase = pass ? (unsigned)n8 : 0u;
                // walk the flattened (row = x0, column = x1) triangle; row r (x0 = lo+9+r) has r+1 entries
                int row = 0, col = lane;
                while (row < hi - lo - 9 && col > row) { col -= row + 1; row++; }
                for (int k = lane; k < n; k += 32) {
                    const unsigned x0 = (unsigned)(lo + 9 + row), x1 = (unsigned)(lo + col);
                    const unsigned e = pass ? alpha_pair_error(src, x1, x0) : alpha_pair_error(src, x0, x1);
                    if (e < besterror) {
                        besterror = e;
                        bestk = kbase + (unsigned)k;
                        bestpair = pass ? ((x1 << 8) | x0) : ((x0 << 8) | x1);
                    }
                    col += 32;
                    while (col > row) { col -= row + 1; row++; }
                }
            }
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) {
                const unsigned oe = __shfl_xor_sync(0xffffffffu, besterror, d);
                const unsigned ok = __shfl_xor_sync(0xffffffffu, bestk, d);
                const unsigned op = __shfl_xor_sync(0xffffffffu, bestpair, d);
                if (oe < besterror || (oe == besterror && ok < bestk)) {
                    besterror = oe;
                    bestk = ok;
                    bestpair = op;
                }
            }
            a0 = bestpair >> 8;
            a1 = bestpair & 0xFF;
        }
        if (lane == 0) {
            unsigned long long b = ((unsigned long long)a1 << 8) | a0;
            alpha_compute_indices(src, a0, a1, &b);
            *reinterpret_cast<uint2 *>(P.out + nvb_out_block(P.lv, blk) * P.out_stride + P.out_offset) =
                make_uint2((unsigned)(b & 0xFFFFFFFFu), (unsigned)(b >> 32));
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// BC2 explicit alpha: OptimalCompress::compressDXT3A (src/nvtt/OptimalCompressDXT.cpp:140-157,485-503) — each texel
// takes the nearest of the three 4-bit levels around a>>4 (bit-replicated to 8 bits), first on ties as written.
// One thread per block; texel i lands in bits [4i, 4i+4) (AlphaBlockDXT3, src/nvimage/BlockDXT.h:77-99).
// ---------------------------------------------------------------------------------------------------------
NVB_DEV unsigned alpha_quantize4(unsigned a) {
    int q0 = max((int)(a >> 4) - 1, 0), q1 = (int)(a >> 4), q2 = min((int)(a >> 4) + 1, 0xF);
    q0 = (q0 << 4) | q0;
    q1 = (q1 << 4) | q1;
    q2 = (q2 << 4) | q2;
    const int d0 = (q0 - (int)a) * (q0 - (int)a), d1 = (q1 - (int)a) * (q1 - (int)a), d2 = (q2 - (int)a) * (q2 - (int)a);
    if (d0 < d1 && d0 < d2) return (unsigned)(q0 >> 4);
    if (d1 < d2) return (unsigned)(q1 >> 4);
    return (unsigned)(q2 >> 4);
}

__global__ void __launch_bounds__(128) k_alpha_dxt3(AlphaBlocksParams P) {
    const int nblocks = P.lv.bw * P.lv.bh;
    for (int blk = blockIdx.x * blockDim.x + threadIdx.x; blk < nblocks; blk += gridDim.x * blockDim.x) {
        unsigned src[16];
        alpha_gather_block(P.lv, P.channel, blk % P.lv.bw, blk / P.lv.bw, src);
        unsigned long long b = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) b |= (unsigned long long)alpha_quantize4(src[i]) << (4 * i);
        *reinterpret_cast<uint2 *>(P.out + nvb_out_block(P.lv, blk) * P.out_stride + P.out_offset) =
            make_uint2((unsigned)(b & 0xFFFFFFFFu), (unsigned)(b >> 32));
    }
}

// ---- BC3n Quality_Highest colour block: OptimalCompress::compressDXT1G (OptimalCompressDXT.cpp:294-381) -------------------
// Brute force over the (g0 > g1) green endpoint pairs of [min-4, max+4] in 6-bit space; integer errors.  One warp per block:
// the pairs are flattened in loop order and striped over the lanes, (error, order) reduction = first strict minimum.
NVB_DEV int green_pair_error(const unsigned src[16], int g0, int g1) {
    int pal[4];
    pal[0] = (g0 << 2) | (g0 >> 4);
    pal[1] = (g1 << 2) | (g1 >> 4);
    pal[2] = (2 * pal[0] + pal[1]) / 3;
    pal[3] = (2 * pal[1] + pal[0]) / 3;
    int total = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const int g = (int)src[i];
        int e = (g - pal[0]) * (g - pal[0]);
        e = min(e, (g - pal[1]) * (g - pal[1]));
        e = min(e, (g - pal[2]) * (g - pal[2]));
        e = min(e, (g - pal[3]) * (g - pal[3]));
        total += e;  // the reference leaves early once total > best; such a total never passes "error < bestError" either
    }
    return total;
}

__global__ void __launch_bounds__(128) k_dxt1g_optimal(AlphaBlocksParams P) {
    const int lane = threadIdx.x & 31;
    const int warps_per_cta = blockDim.x >> 5;
    const int nblocks = P.lv.bw * P.lv.bh;
    for (int blk = blockIdx.x * warps_per_cta + (threadIdx.x >> 5); blk < nblocks; blk += gridDim.x * warps_per_cta) {
        unsigned src[16];
        alpha_gather_block(P.lv, 1, blk % P.lv.bw, blk / P.lv.bw, src);  // green channel, ColorBlock quantisation
        int ming = 63, maxg = 0;
        bool single = true;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const int green = ((int)src[i] + 1) >> 2;  // may be 64 for g = 255: the 6-bit field assignments below wrap it to 0
            ming = min(ming, green);
            maxg = max(maxg, green);
            if (src[i] != src[0]) single = false;
        }
        unsigned c0, c1, indices;
        if (single) {
            // compressDXT1G(uint8 g)
            c0 = (31u << 11) | ((unsigned)P.omatch6[src[0] * 2 + 0] << 5);
            c1 = (31u << 11) | ((unsigned)P.omatch6[src[0] * 2 + 1] << 5);
            indices = 0xaaaaaaaau;
            if (c0 < c1) {
                const unsigned t = c0; c0 = c1; c1 = t;
                indices ^= 0x55555555u;
            }
        } else {
            int besterror = green_pair_error(src, maxg & 63, ming & 63);
            unsigned bestk = 0xffffffffu;
            int bestg0 = maxg, bestg1 = ming;
            const int lo = (ming <= 4) ? 0 : ming - 4, hi = (maxg >= 63 - 4) ? 63 : maxg + 4;
            // for g0 in [lo+1, hi] for g1 in [lo, g0): row r (g0 = lo+1+r) has r+1 entries
            const int rows = hi - lo;
            const int n = rows > 0 ? rows * (rows + 1) / 2 : 0;
            int row = 0, col = lane;
            while (row < rows && col > row) { col -= row + 1; row++; }
            for (int k = lane; k < n; k += 32) {
                const int g0 = lo + 1 + row, g1 = lo + col;
                const int e = green_pair_error(src, g0, g1);
                if (e < besterror) {
                    besterror = e;
                    bestk = (unsigned)k;
                    bestg0 = g0;
                    bestg1 = g1;
                }
                col += 32;
                while (col > row) { col -= row + 1; row++; }
            }
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) {
                const int oe = __shfl_xor_sync(0xffffffffu, besterror, d);
                const unsigned ok = __shfl_xor_sync(0xffffffffu, bestk, d);
                const int o0 = __shfl_xor_sync(0xffffffffu, bestg0, d), o1 = __shfl_xor_sync(0xffffffffu, bestg1, d);
                if (oe < besterror || (oe == besterror && ok < bestk)) {
                    besterror = oe;
                    bestk = ok;
                    bestg0 = o0;
                    bestg1 = o1;
                }
            }
            c0 = (31u << 11) | ((unsigned)(bestg0 & 63) << 5);
            c1 = (31u << 11) | ((unsigned)(bestg1 & 63) << 5);
            // block->evaluatePalette(palette, false): four colours when col0.u > col1.u, else three + transparent black (g = 0)
            const int p0 = (int)(((c0 >> 5) & 63) << 2 | ((c0 >> 5) & 63) >> 4), p1 = (int)(((c1 >> 5) & 63) << 2 | ((c1 >> 5) & 63) >> 4);
            int p2, p3;
            if (c0 > c1) {
                p2 = (2 * p0 + p1) / 3;
                p3 = (2 * p1 + p0) / 3;
            } else {
                p2 = (p0 + p1) / 2;
                p3 = 0;
            }
            // computeGreenIndices
            indices = 0;
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const int c = (int)src[i];
                const unsigned d0 = (unsigned)((p0 - c) * (p0 - c)), d1 = (unsigned)((p1 - c) * (p1 - c));
                const unsigned d2 = (unsigned)((p2 - c) * (p2 - c)), d3 = (unsigned)((p3 - c) * (p3 - c));
                const unsigned b0 = d0 > d3, b1 = d1 > d2, b2 = d0 > d2, b3 = d1 > d3, b4 = d2 > d3;
                const unsigned x0 = b1 & b2, x1 = b0 & b3, x2 = b0 & b4;
                indices |= (x2 | ((x0 | x1) << 1)) << (2 * i);
            }
        }
        if (lane == 0)
            *reinterpret_cast<uint2 *>(P.out + nvb_out_block(P.lv, blk) * P.out_stride + P.out_offset) = make_uint2(c0 | (c1 << 16), indices);
    }
}

// ---- BC3-RGBM multiplier block (compress_dxt5_rgbm, CompressorDXT5_RGBM.cpp:54-118) ---------------------------------------
// After the colour block (R,G,B)/M has been encoded: decode it, solve the per-texel multiplier m in the least-squares sense,
// quantise it to 8 bits and run OptimalCompress::compressDXT5A on those values *with the texel weights* (fp32 error sums in
// texel order, computeAlphaError :189-217).  One warp per block; every lane holds the block, the (alpha0, alpha1) pairs are
// striped over the lanes as in k_alpha_optimal.
struct RgbmAlphaParams {
    LevelView lv;
    unsigned char *out;   // 16-byte BC3 blocks: alpha block at +0, colour block (already written) at +8
    float min_m;
    int transparency;     // AlphaMode_Transparency: texel weight = saturate(alpha), else 1 (0 outside the image)
};

NVB_DEV float alpha_pair_error_w(const unsigned src[16], const float w[16], unsigned a0, unsigned a1) {
    unsigned pal[8];
    alpha_palette(a0, a1, pal);
    float total = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        int best = 0x7fffffff;
#pragma unroll
        for (int p = 0; p < 8; p++) {
            const int d = (int)src[i] - (int)pal[p];
            best = min(best, d * d);
        }
        total += (float)best * w[i];
    }
    return total;
}

__global__ void __launch_bounds__(128) k_rgbm_alpha(RgbmAlphaParams P) {
    const int lane = threadIdx.x & 31;
    const int warps_per_cta = blockDim.x >> 5;
    const int nblocks = P.lv.bw * P.lv.bh;
    for (int blk = blockIdx.x * warps_per_cta + (threadIdx.x >> 5); blk < nblocks; blk += gridDim.x * warps_per_cta) {
        unsigned char *dst = P.out + nvb_out_block(P.lv, blk) * 16;
        // colour block -> 8-bit palette (BlockDXT1::evaluatePalette, D3D10)
        const uint2 cb = *reinterpret_cast<const uint2 *>(dst + 8);
        const unsigned c0 = cb.x & 0xFFFFu, c1 = cb.x >> 16;
        int pr[4], pg[4], pb[4];
        {
            const int r0 = (c0 >> 11) & 31, g0 = (c0 >> 5) & 63, b0 = c0 & 31, r1 = (c1 >> 11) & 31, g1 = (c1 >> 5) & 63, b1 = c1 & 31;
            pr[0] = (r0 << 3) | (r0 >> 2); pg[0] = (g0 << 2) | (g0 >> 4); pb[0] = (b0 << 3) | (b0 >> 2);
            pr[1] = (r1 << 3) | (r1 >> 2); pg[1] = (g1 << 2) | (g1 >> 4); pb[1] = (b1 << 3) | (b1 >> 2);
            if (c0 > c1) {
                pr[2] = (2 * pr[0] + pr[1]) / 3; pg[2] = (2 * pg[0] + pg[1]) / 3; pb[2] = (2 * pb[0] + pb[1]) / 3;
                pr[3] = (2 * pr[1] + pr[0]) / 3; pg[3] = (2 * pg[1] + pg[0]) / 3; pb[3] = (2 * pb[1] + pb[0]) / 3;
            } else {
                pr[2] = (pr[0] + pr[1]) / 2; pg[2] = (pg[0] + pg[1]) / 2; pb[2] = (pb[0] + pb[1]) / 2;
                pr[3] = pg[3] = pb[3] = 0;
            }
        }
        unsigned src[16];
        float w[16];
        const int bx = blk % P.lv.bw, by = blk / P.lv.bw;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const int x = bx * 4 + (i & 3), y = by * 4 + (i >> 2);
            float cx = 0.0f, cy = 0.0f, cz = 0.0f, wt = 0.0f;
            if (x < P.lv.w && y < P.lv.h) {
                cx = load_texel(P.lv, 0, x, y);
                cy = load_texel(P.lv, 1, x, y);
                cz = load_texel(P.lv, 2, x, y);
                wt = P.transparency ? nv_clamp(load_texel(P.lv, 3, x, y), 0.0f, 1.0f) : 1.0f;
            }
            const float R = nv_clamp(cx, 0.0f, 1.0f), G = nv_clamp(cy, 0.0f, 1.0f), B = nv_clamp(cz, 0.0f, 1.0f);
            const unsigned idx = (cb.y >> (2 * i)) & 3u;
            const float rm = (float)pr[idx] / 255.0f, gm = (float)pg[idx] / 255.0f, bm = (float)pb[idx] / 255.0f;
            float m = (rm * R + gm * G + bm * B) / (rm * rm + gm * gm + bm * bm);
            m = (m - P.min_m) / (1 - P.min_m);
            // U8(ftoi_round(saturate(m) * 255.0f)): cvtss2si, round to nearest even
            src[i] = (unsigned)__float2int_rn(nv_clamp(m, 0.0f, 1.0f) * 255.0f) & 0xFFu;
            w[i] = wt;
        }
        int mina = 255, maxa = 0, mina_no01 = 255, maxa_no01 = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const int a = (int)src[i];
            mina = min(mina, a);
            maxa = max(maxa, a);
            if (a != 0 && a != 255) {
                mina_no01 = min(mina_no01, a);
                maxa_no01 = max(maxa_no01, a);
            }
        }
        unsigned a0, a1;
        if (maxa - mina < 8) {
            a0 = (unsigned)maxa;
            a1 = (unsigned)mina;
        } else if (maxa_no01 - mina_no01 < 6) {
            a0 = (unsigned)mina_no01;
            a1 = (unsigned)maxa_no01;
        } else {
            float besterror = alpha_pair_error_w(src, w, 0, 0);  // zero-filled output block (see k_alpha_optimal)
            unsigned bestk = 0xffffffffu;
            unsigned bestpair = ((unsigned)maxa << 8) | (unsigned)mina;
            const int lo8 = (mina <= 8) ? 0 : mina - 8, hi8 = (maxa >= 255 - 8) ? 255 : maxa + 8;
            const int R8 = hi8 - lo8;
            const int n8 = (R8 > 9) ? ((R8 - 9) * (R8 - 8)) / 2 : 0;
            const int lo6 = (mina_no01 <= 6) ? 0 : mina_no01 - 6, hi6 = (maxa_no01 >= 255 - 6) ? 255 : maxa_no01 + 6;
            const int R6 = hi6 - lo6;
            const int n6 = (R6 > 9) ? ((R6 - 9) * (R6 - 8)) / 2 : 0;
#pragma unroll 1
            for (int pass = 0; pass < 2; pass++) {
                const int lo = pass ? lo6 : lo8, hi = pass ? hi6 : hi8, n = pass ? n6 : n8;
                const unsigned kbase = pass ? (unsigned)n8 : 0u;
                int row = 0, col = lane;
                while (row < hi - lo - 9 && col > row) { col -= row + 1; row++; }
                for (int k = lane; k < n; k += 32) {
                    const unsigned x0 = (unsigned)(lo + 9 + row), x1 = (unsigned)(lo + col);
                    const float e = pass ? alpha_pair_error_w(src, w, x1, x0) : alpha_pair_error_w(src, w, x0, x1);
                    if (e < besterror) {
                        besterror = e;
                        bestk = kbase + (unsigned)k;
                        bestpair = pass ? ((x1 << 8) | x0) : ((x0 << 8) | x1);
                    }
                    col += 32;
                    while (col > row) { col -= row + 1; row++; }
                }
            }
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) {
                const float oe = __shfl_xor_sync(0xffffffffu, besterror, d);
                const unsigned ok = __shfl_xor_sync(0xffffffffu, bestk, d);
                const unsigned op = __shfl_xor_sync(0xffffffffu, bestpair, d);
                if (oe < besterror || (oe == besterror && ok < bestk)) {
                    besterror = oe;
                    bestk = ok;
                    bestpair = op;
                }
            }
            a0 = bestpair >> 8;
            a1 = bestpair & 0xFF;
        }
        if (lane == 0) {
            unsigned long long b = ((unsigned long long)a1 << 8) | a0;
            alpha_compute_indices(src, a0, a1, &b);
            *reinterpret_cast<uint2 *>(dst) = make_uint2((unsigned)(b & 0xFFFFFFFFu), (unsigned)(b >> 32));
        }
    }
}

}  // namespace nvb
