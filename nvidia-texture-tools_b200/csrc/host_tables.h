// Host-side (plain C++) builders for the small lookup tables the kernels consume.  Everything here is computed
// from the published formulae, not copied from the reference's generated tables.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <stdio.h>
#include <vector>

namespace nvb {

// ---- gamma 2.2 exponent tables: table[k] = float(2^((k-127)*p/q)), k = sign|exponent bits ------------------
// src/nvmath/Gamma.cpp:33-300 lists the same numbers; entries with the sign bit set are 0, k=0 (zero/denormal)
// is 0 and k=255 (inf/NaN) is +inf.
inline void build_gamma_tables(float to_gamma[512], float to_linear[512]) {
    for (int k = 0; k < 512; k++) {
        to_gamma[k] = 0.0f;
        to_linear[k] = 0.0f;
    }
    for (int k = 1; k < 255; k++) {
        const double e = (double)(k - 127);
        to_gamma[k] = (float)pow(2.0, e * 5.0 / 11.0);
        to_linear[k] = (float)pow(2.0, e * 11.0 / 5.0);  // underflows to 0 / overflows to +inf at the ends
    }
    to_gamma[255] = INFINITY;
    to_linear[255] = INFINITY;
}

// ---- squish cluster splits (src/nvtt/squish/weightedclusterfit.cpp:496-553 loop nest, flattened) -----------
// For a colour set of n points: all (c0,c1,c2) with c0+c1+c2 <= n in lexicographic order, packed c0|c1<<5|c2<<10.
inline void build_squish_splits(std::vector<uint16_t> &cand, int off[18]) {
    cand.clear();
    off[0] = 0;
    for (int n = 1; n <= 16; n++) {
        off[n] = (int)cand.size();
        for (int c0 = 0; c0 <= n; c0++)
            for (int c1 = 0; c1 <= n - c0; c1++)
                for (int c2 = 0; c2 <= n - c0 - c1; c2++) cand.push_back((uint16_t)(c0 | (c1 << 5) | (c2 << 10)));
    }
    off[17] = (int)cand.size();
}

// Same splits, as positions in the kernel's running-sum table T (row s of T starts at off(s) = s*(n+1) - s*(s-1)/2 and
// holds the sums of 0..n-s consecutive sorted points starting at point s): i0 | i1<<8 | i2<<16.
inline void build_squish_split_indices(std::vector<uint32_t> &idx) {
    idx.clear();
    for (int n = 1; n <= 16; n++) {
        auto off = [n](int s) { return s * (n + 1) - (s * (s - 1)) / 2; };
        for (int c0 = 0; c0 <= n; c0++)
            for (int c1 = 0; c1 <= n - c0; c1++)
                for (int c2 = 0; c2 <= n - c0 - c1; c2++)
                    idx.push_back((uint32_t)(c0 | ((off(c0) + c1) << 8) | ((off(c0 + c1) + c2) << 16)));
    }
}

// Two-cluster splits of WeightedClusterFit::Compress3 (weightedclusterfit.cpp:391-446): (c0,c1), c0+c1 <= n, packed c0|c1<<5.
inline void build_squish_splits3(std::vector<uint16_t> &cand, int off[18]) {
    cand.clear();
    off[0] = 0;
    for (int n = 1; n <= 16; n++) {
        off[n] = (int)cand.size();
        for (int c0 = 0; c0 <= n; c0++)
            for (int c1 = 0; c1 <= n - c0; c1++) cand.push_back((uint16_t)(c0 | (c1 << 5)));
    }
    off[17] = (int)cand.size();
}

// ---- single-colour endpoint match tables (src/nvtt/SingleColorLookup.cpp:34-89, non-alpha mode) ------------
// For every 8-bit value find (max,min) 5/6-bit endpoints whose 2/3-1/3 interpolant is closest, with a small
// penalty on the endpoint distance; first best in (min,max) scan order wins.
// alpha_mode (OMatchAlpha5/6, used by the 3-colour single-colour block of BC1a): the interpolant is the midpoint.
inline void build_omatch(uint8_t *table /*[256][2]*/, int size, bool alpha_mode = false) {
    std::vector<int> expand(size);
    for (int i = 0; i < size; i++) expand[i] = (size == 32) ? ((i << 3) | (i >> 2)) : ((i << 2) | (i >> 4));
    for (int i = 0; i < 256; i++) {
        int bestErr = 256 * 100;
        for (int mn = 0; mn < size; mn++) {
            for (int mx = 0; mx < size; mx++) {
                const int mine = expand[mn], maxe = expand[mx];
                int err = (alpha_mode ? abs((maxe + mine) / 2 - i) : abs((maxe * 2 + mine) / 3 - i)) * 100;
                err += abs(mx - mn) * 3;
                if (err < bestErr) {
                    table[i * 2 + 0] = (uint8_t)mx;
                    table[i * 2 + 1] = (uint8_t)mn;
                    bestErr = err;
                }
            }
        }
    }
}

// ---- ICBC tables (src/nvtt/icbc.h) ------------------------------------------------------------------------
// Cluster splits (icbc.h:1906-1975): cumulative cluster sizes (c0, c0+c1, c0+c1+c2) packed c0|c1<<5|c2<<10, grouped by
// the total t = last cumulative count (1..16) and, inside a group, in (c0, c1) order; total[t-1] = entries usable by a
// colour set of t points.  968 four-cluster and 152 three-cluster splits.
inline void build_icbc_splits(std::vector<uint16_t> &four, int four_total[16], std::vector<uint16_t> &three, int three_total[16]) {
    four.clear();
    three.clear();
    for (int t = 1; t <= 16; t++) {
        for (int c0 = 0; c0 <= t; c0++)
            for (int c1 = 0; c1 <= t - c0; c1++) four.push_back((uint16_t)(c0 | ((c0 + c1) << 5) | (t << 10)));
        four_total[t - 1] = (int)four.size();
        for (int c0 = 0; c0 <= t; c0++) three.push_back((uint16_t)(c0 | (t << 5)));
        three_total[t - 1] = (int)three.size();
    }
}
// 565 rounding midpoints (icbc.h:1516-1526): the arithmetic mean of two neighbouring bit-expanded codes / 255,
// written with six decimals; the last entry is FLT_MAX.
inline void build_icbc_midpoints(float mid5[32], float mid6[64]) {
    char buf[32];
    for (int i = 0; i < 31; i++) {
        const int e0 = (i << 3) | (i >> 2), e1 = ((i + 1) << 3) | ((i + 1) >> 2);
        snprintf(buf, sizeof buf, "%.6f", (e0 + e1) / 510.0);
        mid5[i] = strtof(buf, nullptr);
    }
    mid5[31] = 3.402823466e+38F;
    for (int i = 0; i < 63; i++) {
        const int e0 = (i << 2) | (i >> 4), e1 = ((i + 1) << 2) | ((i + 1) >> 4);
        snprintf(buf, sizeof buf, "%.6f", (e0 + e1) / 510.0);
        mid6[i] = strtof(buf, nullptr);
    }
    mid6[63] = 3.402823466e+38F;
}
// ICBC single-colour tables for Decoder_D3D10 (icbc.h:3166-3270): the error actually minimised is
// max(|amd - i|, |nv - i|) of the AMD and NVIDIA hardware interpolants; first best in (mn, mx) scan order.
inline void build_icbc_match(uint8_t *table /*[256][2]*/, int size) {
    std::vector<int> expand(size);
    for (int i = 0; i < size; i++) expand[i] = (size == 32) ? ((i << 3) | (i >> 2)) : ((i << 2) | (i >> 4));
    for (int i = 0; i < 256; i++) {
        int bestErr = 256 * 100;
        for (int mn = 0; mn < size; mn++) {
            for (int mx = 0; mx < size; mx++) {
                const int mine = expand[mn], maxe = expand[mx];
                const int amd = (43 * maxe + 21 * mine + 32) >> 6;
                int nv;
                if (size == 32) nv = ((2 * mx + mn) * 22) / 8;
                else nv = (256 * mine + (maxe - mine) / 4 + 128 + (maxe - mine) * 80) / 256;
                const int amd_err = abs(amd - i), nv_err = abs(nv - i);
                const int err = amd_err > nv_err ? amd_err : nv_err;
                if (err < bestErr) {
                    bestErr = err;
                    table[i * 2 + 0] = (uint8_t)mx;
                    table[i * 2 + 1] = (uint8_t)mn;
                }
            }
        }
    }
}

// ---- polyphase kernels (src/nvimage/Filter.cpp:115-131,157-271,563-608) ------------------------------------
enum FilterKind { Filter_Box = 0, Filter_Triangle = 1, Filter_Kaiser = 2, Filter_Mitchell = 3 };

struct FilterDesc {
    int kind;
    float width;
    float p0, p1;  // Kaiser: alpha, stretch.  Mitchell: B, C.
};

inline float filt_sincf(const float x) {
    if (fabs(x) < 0.0001f) return 1.0f + x * x * (-1.0f / 6.0f + x * x * 1.0f / 120.0f);
    return sinf(x) / x;
}
inline float filt_bessel0(float x) {
    const float EPSILON_RATIO = 1e-6f;
    float xh = 0.5f * x, sum = 1.0f, pw = 1.0f, ds = 1.0;
    int k = 0;
    while (ds > sum * EPSILON_RATIO) {
        ++k;
        pw = pw * (xh / k);
        ds = pw * pw;
        sum = sum + ds;
    }
    return sum;
}
inline float filter_eval(const FilterDesc &f, float x) {
    switch (f.kind) {
    case Filter_Box:
        return (fabsf(x) <= f.width) ? 1.0f : 0.0f;
    case Filter_Triangle:
        x = fabsf(x);
        return (x < f.width) ? f.width - x : 0.0f;
    case Filter_Kaiser: {
        const float PI_F = 3.14159265358979323846f;
        const float sinc_value = filt_sincf(PI_F * x * f.p1);
        const float t = x / f.width;
        if ((1 - t * t) >= 0) return sinc_value * filt_bessel0(f.p0 * sqrtf(1 - t * t)) / filt_bessel0(f.p0);
        return 0;
    }
    default: {  // Mitchell
        const float b = f.p0, c = f.p1;
        const float p0 = (6.0f - 2.0f * b) / 6.0f;
        const float p2 = (-18.0f + 12.0f * b + 6.0f * c) / 6.0f;
        const float p3 = (12.0f - 9.0f * b - 6.0f * c) / 6.0f;
        const float q0 = (8.0f * b + 24.0f * c) / 6.0f;
        const float q1 = (-12.0f * b - 48.0f * c) / 6.0f;
        const float q2 = (6.0f * b + 30.0f * c) / 6.0f;
        const float q3 = (-b - 6.0f * c) / 6.0f;
        x = fabsf(x);
        if (x < 1.0f) return p0 + x * x * (p2 + x * p3);
        if (x < 2.0f) return q0 + x * (q1 + x * (q2 + x * q3));
        return 0.0f;
    }
    }
}
inline float filter_sample_box(const FilterDesc &f, float x, float scale, int samples) {
    double sum = 0;
    const float isamples = 1.0f / float(samples);
    for (int s = 0; s < samples; s++) {
        const float p = (x + (float(s) + 0.5f) * isamples) * scale;
        sum += filter_eval(f, p);
    }
    return float(sum * isamples);
}

struct PolyphaseTable {
    int length = 0, window = 0;
    float width = 0;
    std::vector<float> weights;  // [length][window], each row normalised
    std::vector<int> left;       // [length]
};

inline void build_polyphase(const FilterDesc &f, unsigned srcLength, unsigned dstLength, PolyphaseTable &t) {
    int samples = 32;
    float scale = float(dstLength) / float(srcLength);
    const float iscale = 1.0f / scale;
    if (scale > 1) {
        samples = 1;
        scale = 1;
    }
    t.length = (int)dstLength;
    t.width = f.width * iscale;
    t.window = (int)ceilf(t.width * 2) + 1;
    t.weights.assign((size_t)t.window * t.length, 0.0f);
    t.left.assign(t.length, 0);
    for (int i = 0; i < t.length; i++) {
        const float center = (0.5f + i) * iscale;
        const int left = (int)floorf(center - t.width);
        t.left[i] = left;
        float total = 0.0f;
        for (int j = 0; j < t.window; j++) {
            const float sample = filter_sample_box(f, left + j - center, scale, samples);
            t.weights[(size_t)i * t.window + j] = sample;
            total += sample;
        }
        for (int j = 0; j < t.window; j++) t.weights[(size_t)i * t.window + j] /= total;
    }
}

// Kernel2(9)::initBlendedSobel(scale) + Kernel2::normalize (src/nvimage/Filter.cpp:494-560,358-369).
// A (2k+1)^2 Sobel-like derivative kernel has element sign(e-k) * ((k+1-|e-k|) + (k-|i-k|)) at row i, column e; the four
// sizes 9,7,5,3 are blended with weights scale.w, .z, .y, .x (accumulated in that order) and L1-normalised.
static inline void build_blended_sobel(const float scale[4], float out[81]) {
    const float wts[4] = {scale[3], scale[2], scale[1], scale[0]};
    for (int s = 0; s < 4; s++) {
        const int k = 4 - s, n = 2 * k + 1, off = s;
        for (int i = 0; i < n; i++)
            for (int e = 0; e < n; e++) {
                const int de = e - k, ae = de < 0 ? -de : de, ai = (i - k) < 0 ? (k - i) : (i - k);
                const float elem = (de == 0) ? 0.0f : (float)((de < 0 ? -1 : 1) * ((k + 1 - ae) + (k - ai)));
                float &dst = out[(i + off) * 9 + e + off];
                if (s == 0) dst = elem * wts[s];
                else dst += elem * wts[s];
            }
    }
    float total = 0.0f;
    for (int i = 0; i < 81; i++) total += fabsf(out[i]);
    const float inv = 1.0f / total;
    for (int i = 0; i < 81; i++) out[i] *= inv;
}

}  // namespace nvb
