/*
 * oracle.c — TEST INFRASTRUCTURE ONLY.  Plain-C, single-threaded restatement of the reference algorithms on the
 * BCn + mip hot path (NVTT 2.1.2).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may link or call this file; the product (nvidia-texture-tools_b200/) never does.
 *
 * Restated here: image ops (setImage, gamma, box / polyphase mips, renormalise), BC1 (ICBC levels 1 / 8 / 9), BC3 colour (squish
 * weighted cluster fit), BC4 / BC5 / BC3 alpha (quick, iterative and brute-force), and the Format_RGBA pixel-format writer
 * (orc_convert_level).  BC2, BC3n, BC1a, BC6H and BC7 are checked against oracle/_ref directly.
 *
 * Parity status: PINNED.  Every function here is checked bit-for-bit against oracle/_ref (the unmodified reference
 * built with -O2 -ffp-contract=off -DICBC_SIMD=0 -DSQUISH_USE_SSE=0, see oracle/build_ref.sh) by tests/test_oracle.py
 * when /root/reference is available, and against the golden vectors in tests/golden/ (generated from that build by
 * tests/golden/make_golden.py) everywhere else.  The reference ships no known-answer vectors of its own.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared oracle.c -lm   (no FMA contraction: the arithmetic below must round
 * exactly like the reference's scalar code).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* nv::max / nv::min / nv::clamp (src/nvcore/Utils.h:158-204): NaN-asymmetric ternaries */
static float fmax_nv(float a, float b) { return (b < a) ? a : b; }
static float fmin_nv(float a, float b) { return (a < b) ? a : b; }
static float fclamp_nv(float x, float a, float b) { return fmin_nv(fmax_nv(x, a), b); }
static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

/* ---------------------------------------------------------------------------------------------------------------
 * Gamma 2.2 (src/nvmath/Gamma.cpp:311-354): 512-entry exponent table x degree-4 mantissa polynomial.
 * ------------------------------------------------------------------------------------------------------------- */
static float g_tab_5_11[512], g_tab_11_5[512];
static int g_tabs_ready = 0;
static void init_gamma_tables(void) {
    if (g_tabs_ready) return;
    for (int k = 0; k < 512; k++) g_tab_5_11[k] = g_tab_11_5[k] = 0.0f;
    for (int k = 1; k < 255; k++) {
        g_tab_5_11[k] = (float)pow(2.0, (k - 127) * 5.0 / 11.0);
        g_tab_11_5[k] = (float)pow(2.0, (k - 127) * 11.0 / 5.0);
    }
    g_tab_5_11[255] = INFINITY;
    g_tab_11_5[255] = INFINITY;
    g_tabs_ready = 1;
}
static float powf_5_11(float x) { /* Gamma.cpp:311-327 */
    union { float f; uint32_t u; } m;
    m.f = x;
    int k = m.u >> 23;
    m.u = (m.u & ((1u << 23) - 1)) | (127u << 23);
    float pe = g_tab_5_11[k];
    float pm = (((-0.0110083047f * m.f + 0.0905038750f) * m.f - 0.324697506f) * m.f + 0.876040946f) * m.f + 0.369160989f;
    return pe * pm;
}
static float powf_11_5(float x) { /* Gamma.cpp:338-354 */
    union { float f; uint32_t u; } m;
    m.f = x;
    int k = m.u >> 23;
    m.u = (m.u & ((1u << 23) - 1)) | (127u << 23);
    float pe = g_tab_11_5[k];
    float pm = (((-0.00916587552f * m.f + 0.119315466f) * m.f + 1.01847068f) * m.f - 0.158338739f) * m.f + 0.0297184721f;
    return pe * pm;
}
static int nv_equal(float f0, float f1) { /* nvmath.h:139-143 */
    return fabs(f0 - f1) <= 0.0001f * fmax_nv(1.0f, fmax_nv(fabsf(f0), fabsf(f1)));
}
/* Surface::toLinear / toGamma (Surface.cpp:1470-1488 -> FloatImage.cpp:259-298): RGB planes only */
void orc_to_linear(float *img, int w, int h, float gamma) {
    init_gamma_tables();
    if (nv_equal(gamma, 1.0f)) return;
    size_t n = (size_t)3 * w * h;
    if (gamma == 2.2f) for (size_t i = 0; i < n; i++) img[i] = powf_11_5(img[i]);
    else for (size_t i = 0; i < n; i++) img[i] = powf(fmax_nv(0.0f, img[i]), gamma);
}
void orc_to_gamma(float *img, int w, int h, float gamma) {
    init_gamma_tables();
    if (nv_equal(gamma, 1.0f)) return;
    size_t n = (size_t)3 * w * h;
    if (gamma == 2.2f) for (size_t i = 0; i < n; i++) img[i] = powf_5_11(img[i]);
    else for (size_t i = 0; i < n; i++) img[i] = powf(fmax_nv(0.0f, img[i]), 1.0f / gamma);
}

/* ---------------------------------------------------------------------------------------------------------------
 * Surface::setImage (Surface.cpp:728-815); half_to_float (Half.cpp:443-498) is the exact widening conversion.
 * ------------------------------------------------------------------------------------------------------------- */
static float half_to_float(uint16_t h) {
    uint32_t s = (uint32_t)(h & 0x8000) << 16, e = (h >> 10) & 31, m = h & 0x3FF, r;
    if (e == 0) {
        if (m == 0) r = 0;
        else {
            int sh = 0;
            while (!(m & 0x400)) { m <<= 1; sh++; }
            r = ((uint32_t)(113 - sh) << 23) | ((m & 0x3FF) << 13);
        }
    } else if (e == 31) r = 0x7F800000u | (m << 13);
    else r = ((e + 112) << 23) | (m << 13);
    union { uint32_t u; float f; } c;
    c.u = s | r;
    return c.f;
}
void orc_set_image(int format, int w, int h, const void *data, float *dst) {
    size_t n = (size_t)w * h;
    float *r = dst, *g = dst + n, *b = dst + 2 * n, *a = dst + 3 * n;
    if (format == 0) {
        const uint8_t *s = (const uint8_t *)data; /* B,G,R,A */
        for (size_t i = 0; i < n; i++) {
            r[i] = (float)s[4 * i + 2] / 255.0f;
            g[i] = (float)s[4 * i + 1] / 255.0f;
            b[i] = (float)s[4 * i + 0] / 255.0f;
            a[i] = (float)s[4 * i + 3] / 255.0f;
        }
    } else if (format == 1) {
        const uint16_t *s = (const uint16_t *)data;
        for (size_t i = 0; i < n; i++) {
            r[i] = half_to_float(s[4 * i]); g[i] = half_to_float(s[4 * i + 1]);
            b[i] = half_to_float(s[4 * i + 2]); a[i] = half_to_float(s[4 * i + 3]);
        }
    } else if (format == 2) {
        const float *s = (const float *)data;
        for (size_t i = 0; i < n; i++) { r[i] = s[4 * i]; g[i] = s[4 * i + 1]; b[i] = s[4 * i + 2]; a[i] = s[4 * i + 3]; }
    } else {
        const float *s = (const float *)data;
        for (size_t i = 0; i < n; i++) { r[i] = s[i]; g[i] = 0; b[i] = 0; a[i] = 0; }
    }
}

/* ---------------------------------------------------------------------------------------------------------------
 * FloatImage::fastDownSample (FloatImage.cpp:559-737)
 * ------------------------------------------------------------------------------------------------------------- */
void orc_box_down(const float *src, int sw, int sh, float *dst) {
    const int w = imax(1, sw / 2), h = imax(1, sh / 2);
    for (int c = 0; c < 4; c++) {
        const float *s = src + (size_t)c * sw * sh;
        float *d = dst + (size_t)c * w * h;
        if (sw == 1 || sh == 1) {
            const unsigned n = (unsigned)(w * h);
            if ((sw * sh) & 1) {
                const float scale = 1.0f / (2 * n + 1);
                for (unsigned x = 0; x < n; x++) {
                    const float w0 = (float)(n - x), w1 = (float)(n - 0), w2 = (float)(1 + x);
                    *d++ = scale * (w0 * s[0] + w1 * s[1] + w2 * s[2]);
                    s += 2;
                }
            } else {
                for (unsigned x = 0; x < n; x++) { *d++ = 0.5f * (s[0] + s[1]); s += 2; }
            }
        } else if (!(sw & 1) && !(sh & 1)) {
            for (int y = 0; y < h; y++) {
                for (int x = 0; x < w; x++) { *d++ = 0.25f * (s[0] + s[1] + s[sw] + s[sw + 1]); s += 2; }
                s += sw;
            }
        } else if ((sw & 1) && (sh & 1)) {
            const float scale = 1.0f / (sw * sh);
            for (int y = 0; y < h; y++) {
                const float v0 = (float)(h - y), v1 = (float)(h - 0), v2 = (float)(1 + y);
                for (int x = 0; x < w; x++) {
                    const float w0 = (float)(w - x), w1 = (float)(w - 0), w2 = (float)(1 + x);
                    float f = 0.0f;
                    f += v0 * (w0 * s[0 * sw + 2 * x] + w1 * s[0 * sw + 2 * x + 1] + w2 * s[0 * sw + 2 * x + 2]);
                    f += v1 * (w0 * s[1 * sw + 2 * x] + w1 * s[1 * sw + 2 * x + 1] + w2 * s[1 * sw + 2 * x + 2]);
                    f += v2 * (w0 * s[2 * sw + 2 * x] + w1 * s[2 * sw + 2 * x + 1] + w2 * s[2 * sw + 2 * x + 2]);
                    *d++ = f * scale;
                }
                s += 2 * sw;
            }
        } else if (sw & 1) {
            const float scale = 1.0f / (2 * sw);
            for (int y = 0; y < h; y++) {
                for (int x = 0; x < w; x++) {
                    const float w0 = (float)(w - x), w1 = (float)(w - 0), w2 = (float)(1 + x);
                    float f = 0.0f;
                    f += w0 * (s[2 * x + 0] + s[sw + 2 * x + 0]);
                    f += w1 * (s[2 * x + 1] + s[sw + 2 * x + 1]);
                    f += w2 * (s[2 * x + 2] + s[sw + 2 * x + 2]);
                    *d++ = f * scale;
                }
                s += 2 * sw;
            }
        } else {
            const float scale = 1.0f / (2 * sh);
            for (int y = 0; y < h; y++) {
                const float v0 = (float)(h - y), v1 = (float)(h - 0), v2 = (float)(1 + y);
                for (int x = 0; x < w; x++) {
                    float f = 0.0f;
                    f += v0 * (s[0 * sw + 2 * x] + s[0 * sw + 2 * x + 1]);
                    f += v1 * (s[1 * sw + 2 * x] + s[1 * sw + 2 * x + 1]);
                    f += v2 * (s[2 * sw + 2 * x] + s[2 * sw + 2 * x + 1]);
                    *d++ = f * scale;
                }
                s += 2 * sw;
            }
        }
    }
}

/* ---------------------------------------------------------------------------------------------------------------
 * Polyphase resize (Filter.cpp:50-131,157-271,563-608; FloatImage.cpp:761-808,1115-1176; wrap FloatImage.h:298-351)
 * kind: 0 Box, 1 Triangle, 2 Kaiser(p0=alpha,p1=stretch), 3 Mitchell(p0=B,p1=C)
 * ------------------------------------------------------------------------------------------------------------- */
static float sincf_nv(float x) {
    if (fabs(x) < 0.0001f) return 1.0f + x * x * (-1.0f / 6.0f + x * x * 1.0f / 120.0f);
    return sinf(x) / x;
}
static float bessel0_nv(float x) {
    float xh = 0.5f * x, sum = 1.0f, pw = 1.0f, ds = 1.0;
    int k = 0;
    while (ds > sum * 1e-6f) { ++k; pw = pw * (xh / k); ds = pw * pw; sum = sum + ds; }
    return sum;
}
static float filter_eval(int kind, float width, float p0, float p1, float x) {
    if (kind == 0) return (fabsf(x) <= width) ? 1.0f : 0.0f;
    if (kind == 1) { x = fabsf(x); return (x < width) ? width - x : 0.0f; }
    if (kind == 2) {
        const float PI_F = (float)(3.1415926535897932384626433833);
        const float sinc_value = sincf_nv(PI_F * x * p1);
        const float t = x / width;
        if ((1 - t * t) >= 0) return sinc_value * bessel0_nv(p0 * sqrtf(1 - t * t)) / bessel0_nv(p0);
        return 0;
    }
    {
        const float b = p0, c = p1;
        const float m0 = (6.0f - 2.0f * b) / 6.0f, m2 = (-18.0f + 12.0f * b + 6.0f * c) / 6.0f, m3 = (12.0f - 9.0f * b - 6.0f * c) / 6.0f;
        const float q0 = (8.0f * b + 24.0f * c) / 6.0f, q1 = (-12.0f * b - 48.0f * c) / 6.0f, q2 = (6.0f * b + 30.0f * c) / 6.0f, q3 = (-b - 6.0f * c) / 6.0f;
        x = fabsf(x);
        if (x < 1.0f) return m0 + x * x * (m2 + x * m3);
        if (x < 2.0f) return q0 + x * (q1 + x * (q2 + x * q3));
        return 0.0f;
    }
}
static float sample_box(int kind, float width, float p0, float p1, float x, float scale, int samples) {
    double sum = 0;
    float isamples = 1.0f / (float)samples;
    for (int s = 0; s < samples; s++) {
        float p = (x + ((float)s + 0.5f) * isamples) * scale;
        sum += filter_eval(kind, width, p0, p1, p);
    }
    return (float)(sum * isamples);
}
typedef struct { int length, window; float width; float *data; } PolyKernel;
static void poly_build(PolyKernel *k, int kind, float fwidth, float p0, float p1, unsigned srcLength, unsigned dstLength) {
    int samples = 32;
    float scale = (float)dstLength / (float)srcLength;
    const float iscale = 1.0f / scale;
    if (scale > 1) { samples = 1; scale = 1; }
    k->length = (int)dstLength;
    k->width = fwidth * iscale;
    k->window = (int)ceilf(k->width * 2) + 1;
    k->data = (float *)calloc((size_t)k->window * k->length, sizeof(float));
    for (int i = 0; i < k->length; i++) {
        const float center = (0.5f + i) * iscale;
        const int left = (int)floorf(center - k->width);
        float total = 0.0f;
        for (int j = 0; j < k->window; j++) {
            const float sample = sample_box(kind, fwidth, p0, p1, left + j - center, scale, samples);
            k->data[i * k->window + j] = sample;
            total += sample;
        }
        for (int j = 0; j < k->window; j++) k->data[i * k->window + j] /= total;
    }
}
static int wrap_idx(int x, int w, int mode) {
    if (mode == 0) return imin(imax(x, 0), w - 1);
    if (mode == 1) { if (x >= 0) return x % w; return (x + 1) % w + w - 1; }
    if (w == 1) x = 0;
    x = abs(x);
    while (x >= w) x = abs(w + w - x - 2);
    return x;
}
void orc_resize(const float *src, int sw, int sh, float *dst, int dw, int dh, int kind, float fwidth, float p0, float p1, int wrap) {
    PolyKernel kx, ky;
    poly_build(&kx, kind, fwidth, p0, p1, sw, dw);
    poly_build(&ky, kind, fwidth, p0, p1, sh, dh);
    float *tmp = (float *)malloc(sizeof(float) * (size_t)dw * sh);
    for (int c = 0; c < 4; c++) {
        const float *ch = src + (size_t)c * sw * sh;
        {
            const float scale = (float)kx.length / (float)sw, iscale = 1.0f / scale;
            for (int y = 0; y < sh; y++)
                for (int i = 0; i < kx.length; i++) {
                    const float center = (0.5f + i) * iscale;
                    const int left = (int)floorf(center - kx.width);
                    float sum = 0;
                    for (int j = 0; j < kx.window; j++) sum += kx.data[i * kx.window + j] * ch[(size_t)y * sw + wrap_idx(left + j, sw, wrap)];
                    tmp[(size_t)y * dw + i] = sum;
                }
        }
        {
            float *out = dst + (size_t)c * dw * dh;
            const float scale = (float)ky.length / (float)sh, iscale = 1.0f / scale;
            for (int x = 0; x < dw; x++)
                for (int i = 0; i < ky.length; i++) {
                    const float center = (0.5f + i) * iscale;
                    const int left = (int)floorf(center - ky.width);
                    float sum = 0;
                    for (int j = 0; j < ky.window; j++) sum += ky.data[i * ky.window + j] * tmp[(size_t)wrap_idx(left + j, sh, wrap) * dw + x];
                    out[(size_t)i * dw + x] = sum;
                }
        }
    }
    free(tmp);
    free(kx.data);
    free(ky.data);
}
/* Surface::buildNextMipmap (Surface.cpp:1344-1406): filter 0 Box / 1 Triangle / 2 Kaiser */
void orc_next_mipmap(const float *src, int sw, int sh, float *dst, int filter, float fwidth, float p0, float p1, int wrap, int alphaMode) {
    const int dw = imax(1, sw / 2), dh = imax(1, sh / 2);
    if (filter == 0 && fwidth == 0.5f && alphaMode != 1) { orc_box_down(src, sw, sh, dst); return; }
    orc_resize(src, sw, sh, dst, dw, dh, filter, fwidth, p0, p1, wrap);
}
/* expandNormals -> normalizeNormalMap -> packNormals (Context.cpp:329-334; FloatImage.cpp:201-242) */
void orc_renormalize(float *img, int w, int h) {
    size_t n = (size_t)w * h;
    for (size_t i = 0; i < 3 * n; i++) img[i] = 2.0f * img[i] + -1.0f;
    for (size_t i = 0; i < n; i++) {
        float x = img[i], y = img[n + i], z = img[2 * n + i];
        float l = sqrtf(x * x + y * y + z * z);
        if (fabs(l) <= 0.0f) { x = y = z = 0.0f; }
        else { float s = 1.0f / l; x = x * s; y = y * s; z = z * s; }
        img[i] = x; img[n + i] = y; img[2 * n + i] = z;
    }
    for (size_t i = 0; i < 3 * n; i++) img[i] = 0.5f * img[i] + 0.5f;
}

/* ---------------------------------------------------------------------------------------------------------------
 * ColorBlock::init (ColorBlock.cpp:80-110): truncating quantiser, partial blocks repeat texels by modulo.
 * out: 16 x {b,g,r,a}
 * ------------------------------------------------------------------------------------------------------------- */
static void colorblock_init(int w, int h, const float *data, int x, int y, uint8_t bgra[64]) {
    const int bw = imin(w - x, 4), bh = imin(h - y, 4);
    const size_t plane = (size_t)w * h;
    for (int i = 0; i < 4; i++) {
        const int by = i % bh;
        for (int e = 0; e < 4; e++) {
            const int bx = e % bw;
            const size_t idx = (size_t)(y + by) * w + x + bx;
            uint8_t *c = bgra + 4 * (i * 4 + e);
            c[2] = (uint8_t)(255 * fclamp_nv(data[idx + 0 * plane], 0.0f, 1.0f));
            c[1] = (uint8_t)(255 * fclamp_nv(data[idx + 1 * plane], 0.0f, 1.0f));
            c[0] = (uint8_t)(255 * fclamp_nv(data[idx + 2 * plane], 0.0f, 1.0f));
            c[3] = (uint8_t)(255 * fclamp_nv(data[idx + 3 * plane], 0.0f, 1.0f));
        }
    }
}

/* ---------------------------------------------------------------------------------------------------------------
 * QuickCompress::compressDXT5A (QuickCompressDXT.cpp:505-592,643-647,779-821; BlockDXT.cpp:336-378,408-416)
 * ------------------------------------------------------------------------------------------------------------- */
static void alpha_palette(unsigned a0, unsigned a1, uint8_t p[8]) {
    p[0] = (uint8_t)a0; p[1] = (uint8_t)a1;
    if (a0 > a1) for (int i = 2; i < 8; i++) p[i] = (uint8_t)(((8 - i) * a0 + (i - 1) * a1) / 7);
    else { for (int i = 2; i < 6; i++) p[i] = (uint8_t)(((6 - i) * a0 + (i - 1) * a1) / 5); p[6] = 0; p[7] = 255; }
}
static unsigned blk_index(uint64_t u, int i) { return (unsigned)((u >> (3 * i + 16)) & 7); }
static uint64_t blk_set_index(uint64_t u, int i, unsigned v) {
    const int off = 3 * i + 16;
    return (u & ~((uint64_t)7 << off)) | ((uint64_t)v << off);
}
static unsigned alpha_indices(const uint8_t src[16], uint64_t *blk) {
    uint8_t pal[8];
    alpha_palette((unsigned)(*blk & 0xFF), (unsigned)((*blk >> 8) & 0xFF), pal);
    unsigned total = 0;
    for (int i = 0; i < 16; i++) {
        unsigned besterror = 256 * 256, best = 8;
        for (unsigned p = 0; p < 8; p++) {
            int d = pal[p] - src[i];
            unsigned e = (unsigned)(d * d);
            if (e < besterror) { besterror = e; best = p; }
        }
        total += besterror;
        *blk = blk_set_index(*blk, i, best);
    }
    return total;
}
static void alpha_optimize8(const uint8_t src[16], uint64_t *blk) {
    float alpha2_sum = 0, beta2_sum = 0, alphabeta_sum = 0, alphax_sum = 0, betax_sum = 0;
    for (int i = 0; i < 16; i++) {
        unsigned idx = blk_index(*blk, i);
        float alpha;
        if (idx < 2) alpha = 1.0f - idx;
        else alpha = (8.0f - idx) / 7.0f;
        float beta = 1 - alpha;
        alpha2_sum += alpha * alpha;
        beta2_sum += beta * beta;
        alphabeta_sum += alpha * beta;
        alphax_sum += alpha * src[i];
        betax_sum += beta * src[i];
    }
    const float factor = 1.0f / (alpha2_sum * beta2_sum - alphabeta_sum * alphabeta_sum);
    float a = (alphax_sum * beta2_sum - betax_sum * alphabeta_sum) * factor;
    float b = (betax_sum * alpha2_sum - alphax_sum * alphabeta_sum) * factor;
    unsigned alpha0 = (unsigned)fmin_nv(fmax_nv(a, 0.0f), 255.0f);
    unsigned alpha1 = (unsigned)fmin_nv(fmax_nv(b, 0.0f), 255.0f);
    if (alpha0 < alpha1) {
        unsigned t = alpha0; alpha0 = alpha1; alpha1 = t;
        for (int i = 0; i < 16; i++) {
            unsigned idx = blk_index(*blk, i);
            *blk = blk_set_index(*blk, i, idx < 2 ? 1 - idx : 9 - idx);
        }
    } else if (alpha0 == alpha1) {
        for (int i = 0; i < 16; i++) *blk = blk_set_index(*blk, i, 0);
    }
    *blk = (*blk & ~(uint64_t)0xFFFF) | alpha0 | ((uint64_t)alpha1 << 8);
}
static uint64_t alpha_quick(const uint8_t src[16]) {
    uint8_t a0 = 0, a1 = 255;
    for (int i = 0; i < 16; i++) { if (src[i] > a0) a0 = src[i]; if (src[i] < a1) a1 = src[i]; }
    uint64_t block = 0;
    block |= (uint8_t)(a0 - (a0 - a1) / 34);
    block |= (uint64_t)(uint8_t)(a1 + (a0 - a1) / 34) << 8;
    unsigned besterror = alpha_indices(src, &block);
    uint64_t best = block;
    for (int i = 0; i < 8; i++) {
        alpha_optimize8(src, &block);
        unsigned error = alpha_indices(src, &block);
        if (error >= besterror) break;
        if ((block | 0xFFFF) == (best | 0xFFFF)) { best = block; break; }
        besterror = error;
        best = block;
    }
    return best;
}

/* ---------------------------------------------------------------------------------------------------------------
 * OptimalCompress::compressDXT5A (OptimalCompressDXT.cpp:189-244,512-607): brute force over (alpha0, alpha1).
 * The reference computes the first `besterror` from the not-yet-written output block (:546); the canonical behaviour
 * restated here is a zero-filled output buffer (alpha0 = alpha1 = 0), see DESIGN.md.
 * ------------------------------------------------------------------------------------------------------------- */
static float alpha_error_opt(const uint8_t src[16], unsigned a0, unsigned a1, float bestError) {
    uint8_t pal[8];
    alpha_palette(a0, a1, pal);
    float total = 0;
    for (int i = 0; i < 16; i++) {
        int minDist = 0x7fffffff;
        for (int p = 0; p < 8; p++) {
            int d = (int)src[i] - (int)pal[p];
            d *= d;
            if (d < minDist) minDist = d;
        }
        total += minDist * 1.0f;
        if (total > bestError) return total;
    }
    return total;
}
static uint64_t alpha_optimal(const uint8_t src[16]) {
    uint8_t mina = 255, maxa = 0, mina_no01 = 255, maxa_no01 = 0;
    for (int i = 0; i < 16; i++) {
        uint8_t a = src[i];
        if (a < mina) mina = a;
        if (a > maxa) maxa = a;
        if (a != 0 && a != 255) { if (a < mina_no01) mina_no01 = a; if (a > maxa_no01) maxa_no01 = a; }
    }
    unsigned alpha0, alpha1;
    if (maxa - mina < 8) { alpha0 = maxa; alpha1 = mina; }
    else if (maxa_no01 - mina_no01 < 6) { alpha0 = mina_no01; alpha1 = maxa_no01; }
    else {
        float besterror = alpha_error_opt(src, 0, 0, FLT_MAX);
        int besta0 = maxa, besta1 = mina;
        mina = (uint8_t)((mina <= 8) ? 0 : mina - 8);
        maxa = (uint8_t)((maxa >= 255 - 8) ? 255 : maxa + 8);
        for (int a0 = mina + 9; a0 < maxa; a0++)
            for (int a1 = mina; a1 < a0 - 8; a1++) {
                float e = alpha_error_opt(src, (unsigned)a0, (unsigned)a1, besterror);
                if (e < besterror) { besterror = e; besta0 = a0; besta1 = a1; }
            }
        mina_no01 = (uint8_t)((mina_no01 <= 6) ? 0 : mina_no01 - 6);
        maxa_no01 = (uint8_t)((maxa_no01 >= 255 - 6) ? 255 : maxa_no01 + 6);
        for (int a0 = mina_no01 + 9; a0 < maxa_no01; a0++)
            for (int a1 = mina_no01; a1 < a0 - 8; a1++) {
                float e = alpha_error_opt(src, (unsigned)a1, (unsigned)a0, besterror);
                if (e < besterror) { besterror = e; besta0 = a1; besta1 = a0; }
            }
        alpha0 = (unsigned)besta0; alpha1 = (unsigned)besta1;
    }
    uint64_t block = (uint64_t)alpha0 | ((uint64_t)alpha1 << 8);
    alpha_indices(src, &block);
    return block;
}

/* ---------------------------------------------------------------------------------------------------------------
 * BC3 colour: nvsquish WeightedClusterFit, scalar path, intended (-O2) semantics
 * (colourset.cpp:35-139, maths.cpp:32-133, weightedclusterfit.cpp:39-105,476-589, colourblock.cpp:30-138,
 *  OptimalCompressDXT.cpp:254-269 + SingleColorLookup.cpp:34-89)
 * ------------------------------------------------------------------------------------------------------------- */
static uint8_t g_om5[256][2], g_om6[256][2];
static int g_om_ready = 0;
static void prepare_opt_table(uint8_t *table, int size) {
    int expand[64];
    for (int i = 0; i < size; i++) expand[i] = size == 32 ? ((i << 3) | (i >> 2)) : ((i << 2) | (i >> 4));
    for (int i = 0; i < 256; i++) {
        int bestErr = 256 * 100;
        for (int mn = 0; mn < size; mn++)
            for (int mx = 0; mx < size; mx++) {
                int err = abs((expand[mx] * 2 + expand[mn]) / 3 - i) * 100 + abs(mx - mn) * 3;
                if (err < bestErr) { table[i * 2] = (uint8_t)mx; table[i * 2 + 1] = (uint8_t)mn; bestErr = err; }
            }
    }
}
static void write_block(uint8_t *out, unsigned c0, unsigned c1, uint32_t indices) {
    out[0] = c0 & 0xFF; out[1] = c0 >> 8; out[2] = c1 & 0xFF; out[3] = c1 >> 8;
    out[4] = indices & 0xFF; out[5] = (indices >> 8) & 0xFF; out[6] = (indices >> 16) & 0xFF; out[7] = indices >> 24;
}
static void single_color_dxt1(unsigned r, unsigned g, unsigned b, const uint8_t m5[256][2], const uint8_t m6[256][2], uint8_t *out) {
    unsigned c0 = ((unsigned)m5[r][0] << 11) | ((unsigned)m6[g][0] << 5) | m5[b][0];
    unsigned c1 = ((unsigned)m5[r][1] << 11) | ((unsigned)m6[g][1] << 5) | m5[b][1];
    uint32_t indices = 0xaaaaaaaau;
    if (c0 < c1) { unsigned t = c0; c0 = c1; c1 = t; indices ^= 0x55555555u; }
    write_block(out, c0, c1, indices);
}
static int squish_ftoi(float a, int limit) {
    int i = (int)(a + 0.5f);
    if (i < 0) i = 0; else if (i > limit) i = limit;
    return i;
}
static void squish_compress4(const uint8_t bgra[64], const float metric[3], int weightByAlpha, uint8_t *out) {
    float px[16], py[16], pz[16], pw[16];
    int remap[16], count = 0;
    for (int i = 0; i < 16; i++) {
        for (int j = 0;; j++) {
            if (j == i) {
                px[count] = (float)bgra[4 * i + 2] / 255.0f;
                py[count] = (float)bgra[4 * i + 1] / 255.0f;
                pz[count] = (float)bgra[4 * i + 0] / 255.0f;
                float w = (float)(bgra[4 * i + 3] + 1) / 256.0f;
                pw[count] = weightByAlpha ? w : 1.0f;
                remap[i] = count++;
                break;
            }
            if (bgra[4 * i] == bgra[4 * j] && bgra[4 * i + 1] == bgra[4 * j + 1] && bgra[4 * i + 2] == bgra[4 * j + 2]) {
                int index = remap[j];
                float w = (float)(bgra[4 * i + 3] + 1) / 256.0f;
                pw[index] += weightByAlpha ? w : 1.0f;
                remap[i] = index;
                break;
            }
        }
    }
    /* covariance */
    float total = 0.0f, cx = 0, cy = 0, cz = 0;
    for (int i = 0; i < count; i++) { total += pw[i]; cx += pw[i] * px[i]; cy += pw[i] * py[i]; cz += pw[i] * pz[i]; }
    { float t = 1.0f / total; cx *= t; cy *= t; cz *= t; }
    float cov[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < count; i++) {
        float ax = (px[i] - cx) * metric[0], ay = (py[i] - cy) * metric[1], az = (pz[i] - cz) * metric[2];
        float bx = pw[i] * ax, by = pw[i] * ay, bz = pw[i] * az;
        cov[0] += ax * bx; cov[1] += ax * by; cov[2] += ax * bz; cov[3] += ay * by; cov[4] += ay * bz; cov[5] += az * bz;
    }
    float vx, vy, vz;
    {
        float r0 = cov[0] * cov[0] + cov[1] * cov[1] + cov[2] * cov[2];
        float r1 = cov[1] * cov[1] + cov[3] * cov[3] + cov[4] * cov[4];
        float r2 = cov[2] * cov[2] + cov[4] * cov[4] + cov[5] * cov[5];
        if (r0 > r1 && r0 > r2) { vx = cov[0]; vy = cov[1]; vz = cov[2]; }
        else if (r1 > r2) { vx = cov[1]; vy = cov[3]; vz = cov[4]; }
        else { vx = cov[2]; vy = cov[4]; vz = cov[5]; }
        for (int it = 0; it < 8; it++) {
            float x = vx * cov[0] + vy * cov[1] + vz * cov[2];
            float y = vx * cov[1] + vy * cov[3] + vz * cov[4];
            float z = vx * cov[2] + vy * cov[4] + vz * cov[5];
            float m = x < y ? y : x;       /* std::max(std::max(x,y),z) */
            float norm = m < z ? z : m;
            float iv = 1.0f / norm;
            if (norm == 0.0f) { vx = vy = vz = 0.0f; break; }
            vx = x * iv; vy = y * iv; vz = z * iv;
        }
    }
    float dps[16];
    int order[16];
    for (int i = 0; i < count; i++) { dps[i] = px[i] * vx + py[i] * vy + pz[i] * vz; order[i] = i; }
    for (int i = 0; i < count; i++)
        for (int j = i; j > 0 && dps[j] < dps[j - 1]; --j) {
            float t = dps[j]; dps[j] = dps[j - 1]; dps[j - 1] = t;
            int o = order[j]; order[j] = order[j - 1]; order[j - 1] = o;
        }
    float wx[17], wy[17], wz[17], ww[17], xs[3] = {0, 0, 0}, wsum = 0.0f;
    for (int i = 0; i < count; i++) {
        int p = order[i];
        wx[i] = pw[p] * px[p]; wy[i] = pw[p] * py[p]; wz[i] = pw[p] * pz[p]; ww[i] = pw[p];
        xs[0] += wx[i]; xs[1] += wy[i]; xs[2] += wz[i]; wsum += ww[i];
    }
    wx[count] = wy[count] = wz[count] = ww[count] = 0.0f; /* the reference reads one past the end; the value is dead */
    const float grid[3] = {31.0f, 63.0f, 31.0f}, gridrcp[3] = {1.0f / 31.0f, 1.0f / 63.0f, 1.0f / 31.0f};
    const float msq[3] = {metric[0] * metric[0], metric[1] * metric[1], metric[2] * metric[2]};
    float beststart[3] = {0, 0, 0}, bestend[3] = {0, 0, 0}, besterror = FLT_MAX;
    int b0 = 0, b1 = 0, b2 = 0;
    float x0[3] = {0, 0, 0}, w0 = 0.0f;
    for (int c0 = 0; c0 <= count; c0++) {
        float x1[3] = {0, 0, 0}, w1 = 0.0f;
        for (int c1 = 0; c1 <= count - c0; c1++) {
            float x2[3] = {0, 0, 0}, w2 = 0.0f;
            for (int c2 = 0; c2 <= count - c0 - c1; c2++) {
                float w3 = wsum - w0 - w1 - w2;
                const float alpha2_sum = w0 + w1 * (4.0f / 9.0f) + w2 * (1.0f / 9.0f);
                const float beta2_sum = w3 + w2 * (4.0f / 9.0f) + w1 * (1.0f / 9.0f);
                const float alphabeta_sum = (w1 + w2) * (2.0f / 9.0f);
                const float factor = 1.0f / (alpha2_sum * beta2_sum - alphabeta_sum * alphabeta_sum);
                float a[3], b[3], e1[3];
                for (int k = 0; k < 3; k++) {
                    const float alphax_sum = x0[k] + x1[k] * (2.0f / 3.0f) + x2[k] * (1.0f / 3.0f);
                    const float betax_sum = xs[k] - alphax_sum;
                    float av = (alphax_sum * beta2_sum - betax_sum * alphabeta_sum) * factor;
                    float bv = (betax_sum * alpha2_sum - alphax_sum * alphabeta_sum) * factor;
                    av = (0.0f < av) ? av : 0.0f; av = (av < 1.0f) ? av : 1.0f;   /* Min(one, Max(zero, a)) with std:: semantics */
                    bv = (0.0f < bv) ? bv : 0.0f; bv = (bv < 1.0f) ? bv : 1.0f;
                    av = floorf(grid[k] * av + 0.5f) * gridrcp[k];
                    bv = floorf(grid[k] * bv + 0.5f) * gridrcp[k];
                    e1[k] = av * av * alpha2_sum + bv * bv * beta2_sum + 2.0f * (av * bv * alphabeta_sum - av * alphax_sum - bv * betax_sum);
                    a[k] = av; b[k] = bv;
                }
                float error = e1[0] * msq[0] + e1[1] * msq[1] + e1[2] * msq[2];
                if (error < besterror) {
                    besterror = error;
                    memcpy(beststart, a, sizeof a); memcpy(bestend, b, sizeof b);
                    b0 = c0; b1 = c1; b2 = c2;
                }
                x2[0] += wx[c0 + c1 + c2]; x2[1] += wy[c0 + c1 + c2]; x2[2] += wz[c0 + c1 + c2]; w2 += ww[c0 + c1 + c2];
            }
            x1[0] += wx[c0 + c1]; x1[1] += wy[c0 + c1]; x1[2] += wz[c0 + c1]; w1 += ww[c0 + c1];
        }
        x0[0] += wx[c0]; x0[1] += wy[c0]; x0[2] += wz[c0]; w0 += ww[c0];
    }
    if (!(besterror < FLT_MAX)) { memset(out, 0, 8); return; }
    uint8_t bestindices[16], ordered[16], idx[16];
    {
        int i = 0;
        for (; i < b0; i++) bestindices[i] = 0;
        for (; i < b0 + b1; i++) bestindices[i] = 2;
        for (; i < b0 + b1 + b2; i++) bestindices[i] = 3;
        for (; i < count; i++) bestindices[i] = 1;
    }
    for (int i = 0; i < count; i++) ordered[order[i]] = bestindices[i];
    for (int i = 0; i < 16; i++) idx[i] = ordered[remap[i]];
    int a565 = (squish_ftoi(31.0f * beststart[0], 31) << 11) | (squish_ftoi(63.0f * beststart[1], 63) << 5) | squish_ftoi(31.0f * beststart[2], 31);
    int b565 = (squish_ftoi(31.0f * bestend[0], 31) << 11) | (squish_ftoi(63.0f * bestend[1], 63) << 5) | squish_ftoi(31.0f * bestend[2], 31);
    if (a565 < b565) { int t = a565; a565 = b565; b565 = t; for (int i = 0; i < 16; i++) idx[i] = (idx[i] ^ 1) & 3; }
    else if (a565 == b565) memset(idx, 0, 16);
    uint32_t bits = 0;
    for (int i = 0; i < 16; i++) bits |= (uint32_t)idx[i] << (2 * i);
    write_block(out, (unsigned)a565, (unsigned)b565, bits);
}

/* ---------------------------------------------------------------------------------------------------------------
 * BC1: ICBC v1.05 compress_dxt1, scalar path (icbc.h; line numbers in each function)
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct { float x, y, z; } V3;
static float g_mid5[32], g_mid6[64];
static uint8_t g_match5[256][2], g_match6[256][2];
static uint16_t g_four[968][3], g_three[152][2];
static int g_four_total[16], g_three_total[16];
static int g_icbc_ready = 0;
static void icbc_init(void) {
    if (g_icbc_ready) return;
    char buf[32];
    for (int i = 0; i < 31; i++) { /* icbc.h:1516-1526: six-decimal literals */
        snprintf(buf, sizeof buf, "%.6f", ((((i) << 3) | ((i) >> 2)) + (((i + 1) << 3) | ((i + 1) >> 2))) / 510.0);
        g_mid5[i] = strtof(buf, NULL);
    }
    g_mid5[31] = FLT_MAX;
    for (int i = 0; i < 63; i++) {
        snprintf(buf, sizeof buf, "%.6f", ((((i) << 2) | ((i) >> 4)) + (((i + 1) << 2) | ((i + 1) >> 4))) / 510.0);
        g_mid6[i] = strtof(buf, NULL);
    }
    g_mid6[63] = FLT_MAX;
    for (int pass = 0; pass < 2; pass++) { /* icbc.h:3166-3270, Decoder_D3D10: err = max(amd_err, nv_err) */
        const int size = pass ? 64 : 32;
        uint8_t(*table)[2] = pass ? g_match6 : g_match5;
        for (int i = 0; i < 256; i++) {
            int bestErr = 256 * 100;
            for (int mn = 0; mn < size; mn++)
                for (int mx = 0; mx < size; mx++) {
                    int mine = pass ? ((mn << 2) | (mn >> 4)) : ((mn << 3) | (mn >> 2));
                    int maxe = pass ? ((mx << 2) | (mx >> 4)) : ((mx << 3) | (mx >> 2));
                    int amd = (43 * maxe + 21 * mine + 32) >> 6;
                    int nv = pass ? (256 * mine + (maxe - mine) / 4 + 128 + (maxe - mine) * 80) / 256 : ((2 * mx + mn) * 22) / 8;
                    int err = imax(abs(amd - i), abs(nv - i));
                    if (err < bestErr) { bestErr = err; table[i][0] = (uint8_t)mx; table[i][1] = (uint8_t)mn; }
                }
        }
    }
    int n4 = 0, n3 = 0; /* icbc.h:1906-1975 */
    for (int t = 1; t <= 16; t++) {
        for (int c0 = 0; c0 <= t; c0++)
            for (int c1 = 0; c1 <= t - c0; c1++) { g_four[n4][0] = c0; g_four[n4][1] = c0 + c1; g_four[n4][2] = t; n4++; }
        g_four_total[t - 1] = n4;
        for (int c0 = 0; c0 <= t; c0++) { g_three[n3][0] = c0; g_three[n3][1] = t; n3++; }
        g_three_total[t - 1] = n3;
    }
    g_icbc_ready = 1;
}
static float sat01(float x) { return fclamp_nv(x, 0.0f, 1.0f); }
static unsigned v3_to_color16(V3 v) { /* icbc.h:1580-1595 */
    unsigned r = (unsigned)fclamp_nv(v.x * 31.0f, 0.0f, 31.0f);
    unsigned g = (unsigned)fclamp_nv(v.y * 63.0f, 0.0f, 63.0f);
    unsigned b = (unsigned)fclamp_nv(v.z * 31.0f, 0.0f, 31.0f);
    r += (v.x > g_mid5[r]); g += (v.y > g_mid6[g]); b += (v.z > g_mid5[b]);
    return (uint16_t)((r << 11) | (g << 5) | b);
}
static void palette_d3d10(unsigned c0, unsigned c1, V3 pal[4]) { /* icbc.h:2558-2585 + color_to_vector3 */
    unsigned p[4][3];
    p[0][0] = ((c0 >> 11) & 31); p[0][1] = ((c0 >> 5) & 63); p[0][2] = c0 & 31;
    p[1][0] = ((c1 >> 11) & 31); p[1][1] = ((c1 >> 5) & 63); p[1][2] = c1 & 31;
    for (int i = 0; i < 2; i++) {
        p[i][0] = (p[i][0] << 3) | (p[i][0] >> 2); p[i][1] = (p[i][1] << 2) | (p[i][1] >> 4); p[i][2] = (p[i][2] << 3) | (p[i][2] >> 2);
    }
    for (int k = 0; k < 3; k++) {
        if (c0 > c1) { p[2][k] = (2 * p[0][k] + p[1][k]) / 3; p[3][k] = (2 * p[1][k] + p[0][k]) / 3; }
        else { p[2][k] = (p[0][k] + p[1][k]) / 2; p[3][k] = 0; }
    }
    for (int i = 0; i < 4; i++) { pal[i].x = p[i][0] / 255.0f; pal[i].y = p[i][1] / 255.0f; pal[i].z = p[i][2] / 255.0f; }
}
static float mse_term(V3 p, V3 c, const float cw[3]) { /* icbc.h:2687-2690 */
    float dx = (p.x - c.x) * cw[0] * 255, dy = (p.y - c.y) * cw[1] * 255, dz = (p.z - c.z) * cw[2] * 255;
    return dx * dx + dy * dy + dz * dz;
}
static float block_mse(const V3 col[16], const float wts[16], const float cw[3], const V3 pal[4], uint32_t indices) { /* icbc.h:2737-2760 */
    float error = 0.0f;
    for (int i = 0; i < 16; i++) error += wts[i] * mse_term(pal[(indices >> (2 * i)) & 3], col[i], cw);
    return error;
}
static float len2w(V3 a, V3 b) { float x = a.x - b.x, y = a.y - b.y, z = a.z - b.z; return x * x + y * y + z * z; }
static V3 mulw(V3 v, const float cw[3]) { V3 r = {v.x * cw[0], v.y * cw[1], v.z * cw[2]}; return r; }
typedef struct { unsigned c0, c1; uint32_t indices; } Blk;
static float output_block4(const V3 col[16], const float wts[16], const float cw[3], V3 v0, V3 v1, Blk *blk) { /* icbc.h:2963-2980, 2803-2837 */
    unsigned color0 = v3_to_color16(v0), color1 = v3_to_color16(v1);
    if (color0 < color1) { unsigned t = color0; color0 = color1; color1 = t; }
    V3 pal[4];
    palette_d3d10(color0, color1, pal);
    V3 p0 = mulw(pal[0], cw), p1 = mulw(pal[1], cw), p2 = mulw(pal[2], cw), p3 = mulw(pal[3], cw);
    uint32_t indices = 0;
    for (int i = 0; i < 16; i++) {
        V3 vc = mulw(col[i], cw);
        float d0 = len2w(vc, p0), d1 = len2w(vc, p1), d2 = len2w(vc, p2), d3 = len2w(vc, p3);
        int b1 = d1 > d2, b2 = d0 > d2, x0 = b1 & b2, b0 = d0 > d3, b3 = d1 > d3;
        x0 = x0 | (b0 & b3);
        int b4 = d2 > d3, x1 = b0 & b4;
        indices |= (uint32_t)(x1 | (x0 << 1)) << (2 * i);
    }
    blk->c0 = color0; blk->c1 = color1; blk->indices = indices;
    return block_mse(col, wts, cw, pal, indices);
}
static float output_block3(const V3 col[16], const float wts[16], const float cw[3], int allow_black, V3 v0, V3 v1, Blk *blk) { /* icbc.h:2944-2961, 2839-2882 */
    unsigned color0 = v3_to_color16(v0), color1 = v3_to_color16(v1);
    if (color0 > color1) { unsigned t = color0; color0 = color1; color1 = t; }
    V3 pal[4];
    palette_d3d10(color0, color1, pal);
    V3 p0 = mulw(pal[0], cw), p1 = mulw(pal[1], cw), p2 = mulw(pal[2], cw);
    uint32_t indices = 0;
    for (int i = 0; i < 16; i++) {
        V3 vc = mulw(col[i], cw);
        float d0 = len2w(p0, vc), d1 = len2w(p1, vc), d2 = len2w(p2, vc);
        int i1 = d1 < d2, i2 = (d2 <= d0) & (d2 <= d1), i3 = 0;
        if (allow_black) { float d3 = vc.x * vc.x + vc.y * vc.y + vc.z * vc.z; i3 = (d3 <= d0) & (d3 <= d1) & (d3 <= d2); }
        indices |= (uint32_t)((i1 | i3) | ((i2 | i3) << 1)) << (2 * i);
    }
    blk->c0 = color0; blk->c1 = color1; blk->indices = indices;
    return block_mse(col, wts, cw, pal, indices);
}
typedef struct { float r[16], g[16], b[16], w[16]; } Sat;
static int compute_sat(const V3 *colors, const float *weights, int count, Sat *sat) { /* icbc.h:1709-1882 */
    V3 centroid = {0, 0, 0};
    float total = 0.0f;
    for (int i = 0; i < count; i++) {
        total += weights[i];
        centroid.x += colors[i].x * weights[i]; centroid.y += colors[i].y * weights[i]; centroid.z += colors[i].z * weights[i];
    }
    { float t = 1.0f / total; centroid.x *= t; centroid.y *= t; centroid.z *= t; }
    float m[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < count; i++) {
        V3 a = {colors[i].x - centroid.x, colors[i].y - centroid.y, colors[i].z - centroid.z};
        V3 b = {a.x * weights[i], a.y * weights[i], a.z * weights[i]};
        m[0] += a.x * b.x; m[1] += a.x * b.y; m[2] += a.x * b.z; m[3] += a.y * b.y; m[4] += a.y * b.z; m[5] += a.z * b.z;
    }
    V3 v = {0, 0, 0};
    if (!(m[0] == 0 && m[3] == 0 && m[5] == 0)) {
        float r0 = m[0] * m[0] + m[1] * m[1] + m[2] * m[2], r1 = m[1] * m[1] + m[3] * m[3] + m[4] * m[4], r2 = m[2] * m[2] + m[4] * m[4] + m[5] * m[5];
        if (r0 > r1 && r0 > r2) { v.x = m[0]; v.y = m[1]; v.z = m[2]; }
        else if (r1 > r2) { v.x = m[1]; v.y = m[3]; v.z = m[4]; }
        else { v.x = m[2]; v.y = m[4]; v.z = m[5]; }
        for (int i = 0; i < 8; i++) {
            float x = v.x * m[0] + v.y * m[1] + v.z * m[2];
            float y = v.x * m[1] + v.y * m[3] + v.z * m[4];
            float z = v.x * m[2] + v.y * m[4] + v.z * m[5];
            float norm = fmax_nv(fmax_nv(x, y), z);
            float inv = 1.0f / norm;
            v.x = x * inv; v.y = y * inv; v.z = z * inv;
        }
    }
    int order[16];
    float dps[16];
    for (int i = 0; i < count; i++) { order[i] = i; dps[i] = colors[i].x * v.x + colors[i].y * v.y + colors[i].z * v.z; }
    for (int i = 0; i < count; i++)
        for (int j = i; j > 0 && dps[j] < dps[j - 1]; --j) {
            float t = dps[j]; dps[j] = dps[j - 1]; dps[j - 1] = t;
            int o = order[j]; order[j] = order[j - 1]; order[j - 1] = o;
        }
    float w = weights[order[0]];
    sat->r[0] = colors[order[0]].x * w; sat->g[0] = colors[order[0]].y * w; sat->b[0] = colors[order[0]].z * w; sat->w[0] = w;
    for (int i = 1; i < count; i++) {
        w = weights[order[i]];
        sat->r[i] = sat->r[i - 1] + colors[order[i]].x * w;
        sat->g[i] = sat->g[i - 1] + colors[order[i]].y * w;
        sat->b[i] = sat->b[i - 1] + colors[order[i]].z * w;
        sat->w[i] = sat->w[i - 1] + w;
    }
    for (int i = count; i < 16; i++) sat->r[i] = sat->g[i] = sat->b[i] = sat->w[i] = FLT_MAX;
    return count;
}
static float round5(float x) { return (float)(int)(sat01(x) * 31.0f + 0.5f) * (1.0f / 31.0f); }
static float round6(float x) { return (float)(int)(sat01(x) * 63.0f + 0.5f) * (1.0f / 63.0f); }
static void cluster_fit(const Sat *sat, int count, const float msq[3], int four, V3 *start, V3 *end) { /* icbc.h:2016-2549 scalar lanes */
    const float sum[3] = {sat->r[count - 1], sat->g[count - 1], sat->b[count - 1]}, w_sum = sat->w[count - 1];
    float besterror = FLT_MAX;
    V3 bs = {0, 0, 0}, be = {0, 0, 0};
    const int total = four ? g_four_total[count - 1] : g_three_total[count - 1];
    for (int i = 0; i < total; i++) {
        float x0[3] = {0, 0, 0}, x1[3] = {0, 0, 0}, x2[3] = {0, 0, 0}, w0 = 0, w1 = 0, w2 = 0;
        int c0 = (four ? g_four[i][0] : g_three[i][0]) - 1, c1 = (four ? g_four[i][1] : g_three[i][1]) - 1, c2 = four ? g_four[i][2] - 1 : -1;
        if (c0 >= 0) { x0[0] = sat->r[c0]; x0[1] = sat->g[c0]; x0[2] = sat->b[c0]; w0 = sat->w[c0]; }
        if (c1 >= 0) { x1[0] = sat->r[c1]; x1[1] = sat->g[c1]; x1[2] = sat->b[c1]; w1 = sat->w[c1]; }
        if (c2 >= 0) { x2[0] = sat->r[c2]; x2[1] = sat->g[c2]; x2[2] = sat->b[c2]; w2 = sat->w[c2]; }
        float alpha2_sum, beta2_sum, alphabeta_sum, ax[3];
        if (four) {
            float w3 = w_sum - w2;
            for (int k = 0; k < 3; k++) { x2[k] = x2[k] - x1[k]; x1[k] = x1[k] - x0[k]; }
            w2 = w2 - w1; w1 = w1 - w0;
            alpha2_sum = w2 * (1.0f / 9.0f) + (w1 * (4.0f / 9.0f) + w0);
            beta2_sum = w1 * (1.0f / 9.0f) + (w2 * (4.0f / 9.0f) + w3);
            alphabeta_sum = (w1 + w2) * (2.0f / 9.0f);
            for (int k = 0; k < 3; k++) ax[k] = x2[k] * (1.0f / 3.0f) + (x1[k] * (2.0f / 3.0f) + x0[k]);
        } else {
            float w2b = w_sum - w1;
            for (int k = 0; k < 3; k++) x1[k] = x1[k] - x0[k];
            w1 = w1 - w0;
            alphabeta_sum = w1 * 0.25f;
            alpha2_sum = w0 + alphabeta_sum;
            beta2_sum = w2b + alphabeta_sum;
            for (int k = 0; k < 3; k++) ax[k] = x0[k] + x1[k] * 0.5f;
        }
        float factor = 1.0f / (alpha2_sum * beta2_sum - alphabeta_sum * alphabeta_sum);
        float a[3], b[3], e1[3];
        for (int k = 0; k < 3; k++) {
            float alphax = ax[k], betax = sum[k] - alphax;
            float av = (alphax * beta2_sum - betax * alphabeta_sum) * factor;
            float bv = (betax * alpha2_sum - alphax * alphabeta_sum) * factor;
            av = k == 1 ? round6(av) : round5(av);
            bv = k == 1 ? round6(bv) : round5(bv);
            float e2 = (av * (bv * alphabeta_sum - alphax) - bv * betax) * 2.0f;
            e1[k] = (av * av) * alpha2_sum + ((bv * bv) * beta2_sum + e2);
            a[k] = av; b[k] = bv;
        }
        float error = e1[0] * msq[0] + e1[1] * msq[1] + e1[2] * msq[2];
        if (error < besterror) { besterror = error; bs.x = a[0]; bs.y = a[1]; bs.z = a[2]; be.x = b[0]; be.y = b[1]; be.z = b[2]; }
    }
    *start = bs; *end = be;
}
static int is_black(V3 c) { return c.x < 1.0f / 8 && c.y < 1.0f / 8 && c.z < 1.0f / 8; }
/* level: 1, 8 or 9.  col/wts: the 16 gathered texels.  out: 8 bytes. */
static void icbc_compress(int level, const V3 col[16], const float wts[16], const float cw[3], uint8_t *out) { /* icbc.h:3485-3559 */
    icbc_init();
    V3 colors[16];
    float weights[16];
    int any_black = 0, count;
    if (level >= 2) { /* reduce_colors, icbc.h:1638-1683, threshold 1/256 */
        const float threshold = 1.0f / 256;
        int n = 0;
        for (int i = 0; i < 16; i++) {
            V3 ci = col[i];
            float wi = wts[i];
            if (wi > 0) {
                int j;
                for (j = 0; j < n; j++) {
                    if (fabsf(colors[j].x - ci.x) < threshold && fabsf(colors[j].y - ci.y) < threshold && fabsf(colors[j].z - ci.z) < threshold) {
                        float den = weights[j] + wi;
                        colors[j].x = (colors[j].x * weights[j] + ci.x * wi) / den;
                        colors[j].y = (colors[j].y * weights[j] + ci.y * wi) / den;
                        colors[j].z = (colors[j].z * weights[j] + ci.z * wi) / den;
                        weights[j] += wi;
                        break;
                    }
                }
                if (j == n) { colors[n] = ci; weights[n] = wi; n++; }
                if (is_black(ci)) any_black = 1;
            }
        }
        count = n;
    } else {
        for (int i = 0; i < 16; i++) colors[i] = col[i];
        count = 16;
    }
    if (count == 0) { memset(out, 0, 8); return; }
    if (count == 1) {
        unsigned r = (uint8_t)(sat01(colors[0].x) * 255 + 0.5f), g = (uint8_t)(sat01(colors[0].y) * 255 + 0.5f), b = (uint8_t)(sat01(colors[0].z) * 255 + 0.5f);
        single_color_dxt1(r, g, b, g_match5, g_match6, out);
        return;
    }
    float error = FLT_MAX;
    Blk output = {0, 0, 0};
    if (level == 1) { /* box fit + least squares, icbc.h:3096-3151, 2984-3016 */
        V3 c0 = {0, 0, 0}, c1 = {1, 1, 1};
        for (int i = 0; i < count; i++) {
            c0.x = fmax_nv(c0.x, colors[i].x); c0.y = fmax_nv(c0.y, colors[i].y); c0.z = fmax_nv(c0.z, colors[i].z);
            c1.x = fmin_nv(c1.x, colors[i].x); c1.y = fmin_nv(c1.y, colors[i].y); c1.z = fmin_nv(c1.z, colors[i].z);
        }
        {
            float bias = (8.0f / 255.0f) / 16.0f;
            V3 inset = {(c0.x - c1.x) / 16.0f - bias, (c0.y - c1.y) / 16.0f - bias, (c0.z - c1.z) / 16.0f - bias};
            c0.x = sat01(c0.x - inset.x); c0.y = sat01(c0.y - inset.y); c0.z = sat01(c0.z - inset.z);
            c1.x = sat01(c1.x + inset.x); c1.y = sat01(c1.y + inset.y); c1.z = sat01(c1.z + inset.z);
        }
        {
            V3 center = {(c0.x + c1.x) * 0.5f, (c0.y + c1.y) * 0.5f, (c0.z + c1.z) * 0.5f};
            float cov_xz = 0.0f, cov_yz = 0.0f;
            for (int i = 0; i < count; i++) {
                V3 t = {colors[i].x - center.x, colors[i].y - center.y, colors[i].z - center.z};
                cov_xz += t.x * t.z; cov_yz += t.y * t.z;
            }
            if (cov_xz < 0) { float t = c0.x; c0.x = c1.x; c1.x = t; }
            if (cov_yz < 0) { float t = c0.y; c0.y = c1.y; c1.y = t; }
        }
        error = output_block4(col, wts, cw, c0, c1, &output);
        float alpha2_sum = 0, beta2_sum = 0, alphabeta_sum = 0;
        V3 axs = {0, 0, 0}, bxs = {0, 0, 0};
        for (int i = 0; i < 16; i++) {
            const unsigned bits = output.indices >> (2 * i);
            float beta = (float)(bits & 1);
            if (bits & 2) beta = (1 + beta) / 3.0f;
            float alpha = 1.0f - beta;
            alpha2_sum += alpha * alpha; beta2_sum += beta * beta; alphabeta_sum += alpha * beta;
            axs.x += col[i].x * alpha; axs.y += col[i].y * alpha; axs.z += col[i].z * alpha;
            bxs.x += col[i].x * beta; bxs.y += col[i].y * beta; bxs.z += col[i].z * beta;
        }
        float denom = alpha2_sum * beta2_sum - alphabeta_sum * alphabeta_sum;
        if (!(fabsf(denom - 0.0f) < 0.0001f)) {
            float factor = 1.0f / denom;
            V3 a = {sat01((axs.x * beta2_sum - bxs.x * alphabeta_sum) * factor), sat01((axs.y * beta2_sum - bxs.y * alphabeta_sum) * factor), sat01((axs.z * beta2_sum - bxs.z * alphabeta_sum) * factor)};
            V3 b = {sat01((bxs.x * alpha2_sum - axs.x * alphabeta_sum) * factor), sat01((bxs.y * alpha2_sum - axs.y * alphabeta_sum) * factor), sat01((bxs.z * alpha2_sum - axs.z * alphabeta_sum) * factor)};
            Blk opt;
            float oe = output_block4(col, wts, cw, a, b, &opt);
            if (oe < error) { error = oe; output = opt; }
        }
    } else { /* compress_dxt1_cluster_fit, icbc.h:3290-3326 */
        const float msq[3] = {cw[0] * cw[0], cw[1] * cw[1], cw[2] * cw[2]};
        Sat sat;
        int sat_count = compute_sat(colors, weights, count, &sat);
        V3 start, end;
        cluster_fit(&sat, sat_count, msq, 1, &start, &end);
        Blk cf;
        float best = output_block4(col, wts, cw, start, end, &cf);
        int do_three = 1;
        if (any_black) {
            V3 tc[16];
            float tw[16];
            int n = 0;
            for (int i = 0; i < count; i++) if (!is_black(colors[i])) { tc[n] = colors[i]; tw[n] = weights[i]; n++; }
            if (!n) do_three = 0;
            else sat_count = compute_sat(tc, tw, n, &sat);
        }
        if (do_three) {
            cluster_fit(&sat, sat_count, msq, 0, &start, &end);
            Blk tb;
            float te = output_block3(col, wts, cw, 1, start, end, &tb);
            if (te < best) { best = te; cf = tb; }
        }
        if (best < error) { output = cf; error = best; }
        if (level == 9) { /* refine_endpoints, icbc.h:3329-3406 (three_color_mode == true) */
            static const int8_t deltas[16][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {-1, 0, 0}, {0, -1, 0}, {0, 0, -1}, {1, 1, 0}, {1, 0, 1},
                                                 {0, 1, 1}, {-1, -1, 0}, {-1, 0, -1}, {0, -1, -1}, {-1, 1, 0}, {1, -1, 0}, {0, -1, 1}, {0, 1, -1}};
            float best_error = error;
            int lastImprovement = 0;
            for (int i = 0; i < 256; i++) {
                Blk refined = output;
                unsigned *c = ((i / 16) & 1) ? &refined.c0 : &refined.c1;
                unsigned r = (((*c >> 11) & 31) + deltas[i % 16][0]) & 31, g = (((*c >> 5) & 63) + deltas[i % 16][1]) & 63, b = ((*c & 31) + deltas[i % 16][2]) & 31;
                *c = (r << 11) | (g << 5) | b;
                V3 pal[4];
                palette_d3d10(output.c0, output.c1, pal); /* sic: palette of *output */
                V3 p0 = mulw(pal[0], cw), p1 = mulw(pal[1], cw), p2 = mulw(pal[2], cw), p3 = mulw(pal[3], cw);
                uint32_t indices = 0;
                for (int t = 0; t < 16; t++) {
                    V3 vc = mulw(col[t], cw);
                    float d0 = len2w(vc, p0), d1 = len2w(vc, p1), d2 = len2w(vc, p2), d3 = len2w(vc, p3);
                    int i1 = (d1 <= d0) & (d1 < d2) & (d1 < d3), i2 = (d2 <= d0) & (d2 <= d1) & (d2 < d3), i3 = (d3 <= d0) & (d3 <= d1) & (d3 <= d2);
                    indices |= (uint32_t)((i1 | i3) | ((i2 | i3) << 1)) << (2 * t);
                }
                refined.indices = indices;
                V3 rp[4];
                palette_d3d10(refined.c0, refined.c1, rp);
                float refined_error = block_mse(col, wts, cw, rp, refined.indices);
                if (refined_error < best_error) { best_error = refined_error; output = refined; lastImprovement = i; }
                if (i - lastImprovement > 32) break;
            }
        }
    }
    write_block(out, output.c0, output.c1, output.indices);
}

/* ---------------------------------------------------------------------------------------------------------------
 * Format_RGB / Format_RGBA: PixelFormatConverter::compress (src/nvtt/CompressorRGB.cpp:410-575) with its BitStream
 * (:332-404), toFloat11 / toFloat10 (:129-160), PixelFormat::convert and maskShiftAndSize (src/nvimage/PixelFormat.h:37-76),
 * nv::half_from_float (src/nvmath/Half.cpp:378-441) and computeBytePitch (src/nvimage/nvimage.h:11-24).
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct {
    int pixelType;                       /* nvtt::PixelType */
    unsigned bitcount;                   /* != 0: mask form */
    unsigned rmask, gmask, bmask, amask;
    unsigned rsize, gsize, bsize, asize; /* size form and float channel widths */
    int pitchAlignment;
    int width, height;
} OrcPixelFormatDesc;

static int msb_set(uint32_t v) { return (int32_t)v < 0; }
static uint32_t sels(uint32_t test, uint32_t a, uint32_t b) { return msb_set(test) ? a : b; }
/* the select network of half_from_float, same data flow; x86 masks shift counts to 5 bits */
static uint16_t half_from_float_bits(uint32_t f) {
    const uint32_t f_s = f & 0x80000000u, f_e = f & 0x7f800000u, f_m = f & 0x007fffffu;
    const uint32_t h_s = (f_s >> 16) & 0xffffu;
    const uint32_t e_amount = (f_e >> 23) & 0xffffu;
    const uint32_t e_half_bias = e_amount - 0x70u;
    const uint32_t f_snan = f & 0x7fc00000u;
    const uint32_t m_rounded = f_m + ((f_m & 0x00001000u) << 1);
    const uint32_t denorm_sa = 1u - e_half_bias;
    const uint32_t h_m_denorm = ((m_rounded | 0x00800000u) >> (denorm_sa & 31u)) >> 13;
    const uint32_t m_nan = f_m >> 13;
    const uint32_t h_em_norm = (e_half_bias << 10) | (m_rounded >> 13);
    const uint32_t flagged = 0x8fu - e_half_bias;
    uint32_t r = sels(0u - (m_rounded & 0x00800000u), (e_half_bias + 1u) << 10, h_em_norm);
    r = sels(flagged, 0x7c00u | m_nan, r);
    r = sels(flagged & (m_nan - 1u), 0x7c01u, r);
    r = sels((0x1fu - e_half_bias) | (flagged & (f_m - 1u)), 0x7c00u, r);
    r = sels(~(0x70u - e_amount), h_m_denorm, r);
    r = sels(~(f_snan - 0x7fc00000u), 0x7e00u, r);
    return (uint16_t)(h_s | r);
}

typedef struct { uint8_t *ptr, *end; uint8_t buffer, bits; } OrcBitStream;
static void bs_byte(OrcBitStream *s, unsigned v) { if (s->ptr < s->end) *s->ptr = (uint8_t)v; s->ptr++; }
static void bs_put_bits(OrcBitStream *s, uint32_t p, unsigned n) {
    uint64_t buffer = (uint32_t)(((uint32_t)s->buffer << (n & 31u)) | p);  /* int << int | uint: 32-bit, then widened */
    unsigned bits = s->bits + n;
    while (bits >= 8) { bs_byte(s, (unsigned)(buffer & 0xFF)); buffer >>= 8; bits -= 8; }
    s->buffer = (uint8_t)buffer;
    s->bits = (uint8_t)bits;
}
static void bs_put_raw(OrcBitStream *s, uint32_t v, int nbytes) { for (int i = 0; i < nbytes; i++) bs_byte(s, (v >> (8 * i)) & 0xFF); }
static uint32_t float_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static uint32_t to_float11(float f) {
    if (f < 0) f = 0;
    if (f > 65024) f = 65024;
    const uint32_t u = float_bits(f);
    return ((((u >> 23) & 0xFF) - 127 + 15) << 6) | ((u & 0x7FFFFF) >> 17);
}
static uint32_t to_float10(float f) {
    if (f < 0) f = 0;
    if (f > 64512) f = 64512;
    const uint32_t u = float_bits(f);
    return ((((u >> 23) & 0xFF) - 127 + 15) << 5) | ((u & 0x7FFFFF) >> 18);
}
static void bs_put_float_channel(OrcBitStream *s, float v, unsigned size) {
    if (size == 32) bs_put_raw(s, float_bits(v), 4);
    else if (size == 16) bs_put_raw(s, half_from_float_bits(float_bits(v)), 2);
    else if (size == 11) bs_put_bits(s, to_float11(v), 11);
    else if (size == 10) bs_put_bits(s, to_float10(v), 10);
    else bs_put_bits(s, 0, size);
}
static uint32_t pf_convert(uint32_t c, unsigned inbits, unsigned outbits) { /* PixelFormat.h:37-53 */
    if (inbits == 0) return 0;
    if (inbits >= outbits) return c >> (inbits - outbits);
    return (c << (outbits - inbits)) | pf_convert(c, inbits, outbits - inbits);
}
static void mask_shift_size(uint32_t mask, unsigned *shift, unsigned *size) {
    *shift = 0; *size = 0;
    if (!mask) return;
    while ((mask & 1) == 0) { ++*shift; mask >>= 1; }
    while ((mask & 1) == 1) { ++*size; mask >>= 1; }
}
static int iround_nv(float f) { return (int)floorf(f + 0.5f); }

/* toFloat3SE (CompressorRGB.cpp:231-269) as the reference's x86-64 build computes it: ftoi_round = cvtss2si (nearest even,
 * INT_MIN for NaN / out of range), and 1 << n with a negative n takes the count modulo 32 (so every pixel below 256 gets
 * zero mantissas - kept, the reference is the contract). */
static int ftoi_round_x86(float f) { return (f >= -2147483648.0f && f < 2147483648.0f) ? (int)lrintf(f) : (int)0x80000000; }
static float nvmax_f(float a, float b) { return (b < a) ? a : b; }
static float nvmin_f(float a, float b) { return (a < b) ? a : b; }
static uint32_t to_float3se(float r, float g, float b) {
    const int N = 9, B = 15;
    const float sharedexp_max = 65408.0f;
    r = nvmax_f(0.0f, nvmin_f(sharedexp_max, r));
    g = nvmax_f(0.0f, nvmin_f(sharedexp_max, g));
    b = nvmax_f(0.0f, nvmin_f(sharedexp_max, b));
    const float max_c = nvmax_f(r, nvmax_f(g, b));
    int fl = ftoi_round_x86(floorf(log2f(max_c)));
    if (fl < -B - 1) fl = -B - 1;
    const int exp_shared_p = fl + 1 + B;
    const float dp = (float)(int)(1u << ((unsigned)(exp_shared_p - B - N) & 31u));
    const int max_s = ftoi_round_x86(max_c / dp);
    int exp_shared = exp_shared_p;
    if (max_s == (1 << N)) exp_shared++;
    const float ds = (float)(int)(1u << ((unsigned)(exp_shared - B - N) & 31u));
    const uint32_t xm = (uint32_t)ftoi_round_x86(r / ds) & 0x1FFu, ym = (uint32_t)ftoi_round_x86(g / ds) & 0x1FFu;
    const uint32_t zm = (uint32_t)ftoi_round_x86(b / ds) & 0x1FFu;
    return xm | (ym << 9) | (zm << 18) | (((uint32_t)exp_shared & 31u) << 27);
}

/* out == NULL: returns the level size (0 = layout the reference asserts on) */
long orc_convert_level(const OrcPixelFormatDesc *d, const float *planar, uint8_t *out) {
    unsigned size[4] = {d->rsize, d->gsize, d->bsize, d->asize}, shift[4] = {0, 0, 0, 0}, bitCount;
    const int isFloat = d->pixelType == 4;
    if (isFloat) {
        bitCount = size[0] + size[1] + size[2] + size[3];
    } else if (d->bitcount != 0) {
        bitCount = d->bitcount;
        const uint32_t mask[4] = {d->rmask, d->gmask, d->bmask, d->amask};
        for (int i = 0; i < 4; i++) mask_shift_size(mask[i], &shift[i], &size[i]);
    } else {
        bitCount = size[0] + size[1] + size[2] + size[3];
        shift[3] = 0; shift[2] = size[3]; shift[1] = shift[2] + size[2]; shift[0] = shift[1] + size[1];
    }
    if (bitCount == 0 || (!isFloat && bitCount > 32)) return 0;
    const int rgb9e5 = d->pixelType == 6 && size[0] == 9 && size[1] == 9 && size[2] == 9 && size[3] == 5;
    if (rgb9e5 && bitCount != 32) return 0; /* putBits(v.v, 32) whatever bitCount says: only 32-bit layouts are consistent */
    const unsigned alignBits = 8u * (unsigned)d->pitchAlignment;
    const unsigned pitch = ((((unsigned)d->width * bitCount + alignBits - 1) / alignBits) * alignBits + 7) / 8;
    const long total = (long)pitch * d->height;
    if (!out) return total;
    const size_t whd = (size_t)d->width * d->height;
    for (int y = 0; y < d->height; y++) {
        OrcBitStream s = {out + (size_t)y * pitch, out + (size_t)(y + 1) * pitch, 0, 0};
        const float *src = planar + (size_t)y * d->width;
        for (int x = 0; x < d->width; x++) {
            const float c[4] = {src[x], src[x + whd], src[x + 2 * whd], src[x + 3 * whd]};
            if (isFloat) {
                for (int i = 0; i < 4; i++) bs_put_float_channel(&s, c[i], size[i]);
            } else if (d->pixelType == 0 || d->pixelType == 2) {
                uint32_t p = 0;
                for (int i = 0; i < 4; i++) {
                    const float v = (d->pixelType == 0) ? c[i] * 65535.0f : c[i];
                    const int iv = iround_nv(fclamp_nv(v, 0.0f, 65535.0f));
                    p |= pf_convert((uint32_t)iv, 16, size[i]) << (shift[i] & 31u);
                }
                bs_put_bits(&s, p, bitCount);
            } else if (rgb9e5) {
                bs_put_bits(&s, to_float3se(c[0], c[1], c[2]), 32);
            } else {
                bs_put_bits(&s, 0, bitCount); /* signed types: zero components; other SharedExp layouts: zeros */
            }
        }
        if (s.bits) { bs_byte(&s, s.buffer); s.buffer = 0; s.bits = 0; } /* flush + zero padding of align() */
        while (s.ptr < s.end) bs_byte(&s, 0);
    }
    return total;
}

/* ---------------------------------------------------------------------------------------------------------------
 * One level: Compressor::Private::compress + ColorBlockCompressor / FloatColorCompressor (BlockCompressor.cpp:60-205)
 * format: nvtt::Format (1 BC1, 4 BC3, 6 BC4, 7 BC5); returns bytes written, 0 if unsupported.
 * ------------------------------------------------------------------------------------------------------------- */
long orc_compress_level(int format, int quality, int alphaMode, int w, int h, const float *rgba, const float cw4[4], uint8_t *out) {
    const int bw = (w + 3) / 4, bh = (h + 3) / 4;
    const size_t plane = (size_t)w * h;
    const float one[4] = {1, 1, 1, 1};
    const float *cw = cw4 ? cw4 : one;
    if (!g_om_ready) { prepare_opt_table(&g_om5[0][0], 32); prepare_opt_table(&g_om6[0][0], 64); g_om_ready = 1; }
    if (format == 1) {
        const int level = quality == 0 ? 1 : (quality == 2 ? 9 : 8);
        for (int by = 0; by < bh; by++)
            for (int bx = 0; bx < bw; bx++) {
                V3 col[16];
                float wts[16];
                for (int i = 0; i < 16; i++) {
                    int x = bx * 4 + (i & 3), y = by * 4 + (i >> 2);
                    if (x < w && y < h) {
                        size_t idx = (size_t)y * w + x;
                        col[i].x = rgba[idx]; col[i].y = rgba[plane + idx]; col[i].z = rgba[2 * plane + idx];
                        wts[i] = alphaMode == 1 ? sat01(rgba[3 * plane + idx]) : 1.0f;
                    } else { col[i].x = col[i].y = col[i].z = 0; wts[i] = 0.0f; }
                }
                icbc_compress(level, col, wts, cw, out + 8 * ((size_t)by * bw + bx));
            }
        return (long)bw * bh * 8;
    }
    if (format == 4 || format == 6 || format == 7) {
        if (format == 4 && quality == 0) return 0; /* QuickCompress::compressDXT5 (fast DXT1) is not restated */
        /* BC4/BC5: Production/Highest -> OptimalCompress (Context.cpp:1098-1115); BC3: Highest only (CompressorDX9.cpp:149) */
        const int optimal = (format == 4) ? (quality == 3) : (quality >= 2);
        const int bs = format == 6 ? 8 : 16;
        for (int by = 0; by < bh; by++)
            for (int bx = 0; bx < bw; bx++) {
                uint8_t bgra[64], ch[16], *dst = out + bs * ((size_t)by * bw + bx);
                colorblock_init(w, h, rgba, bx * 4, by * 4, bgra);
                uint64_t a;
                if (format == 6 || format == 7) {
                    for (int i = 0; i < 16; i++) ch[i] = bgra[4 * i + 2]; /* red */
                    a = optimal ? alpha_optimal(ch) : alpha_quick(ch);
                    memcpy(dst, &a, 8);
                    if (format == 7) {
                        for (int i = 0; i < 16; i++) ch[i] = bgra[4 * i + 1]; /* green */
                        a = optimal ? alpha_optimal(ch) : alpha_quick(ch);
                        memcpy(dst + 8, &a, 8);
                    }
                } else {
                    for (int i = 0; i < 16; i++) ch[i] = bgra[4 * i + 3];
                    a = optimal ? alpha_optimal(ch) : alpha_quick(ch);
                    memcpy(dst, &a, 8);
                    int single = 1;
                    for (int i = 1; i < 16; i++) if (memcmp(bgra, bgra + 4 * i, 3)) single = 0;
                    if (single) single_color_dxt1(bgra[2], bgra[1], bgra[0], g_om5, g_om6, dst + 8);
                    else squish_compress4(bgra, cw, alphaMode == 1, dst + 8);
                }
            }
        return (long)bw * bh * bs;
    }
    return 0;
}

/* ---------------------------------------------------------------------------------------------------------------
 * Whole pipeline for one face: Compressor::Private::compress(InputOptions...) (Context.cpp:260-345), no header.
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct {
    int inputFormat, width, height, wrapMode, mipmapFilter, generateMipmaps, maxLevel;
    float kaiserWidth, kaiserAlpha, kaiserStretch, inputGamma, outputGamma;
    int isNormalMap, normalizeMipmaps, alphaMode, format, quality;
    float colorWeights[4];
} OrcProcessDesc;

long orc_process_face(const OrcProcessDesc *d, const void *image, uint8_t *out) {
    int w = d->width, h = d->height, mips = 1;
    if (d->generateMipmaps) {
        int tw = w, th = h;
        while (tw != 1 || th != 1) { tw = imax(1, tw / 2); th = imax(1, th / 2); mips++; }
        if (d->maxLevel > 0) mips = imin(mips, d->maxLevel);
    }
    float *img = (float *)malloc(sizeof(float) * 4 * (size_t)w * h), *tmp = (float *)malloc(sizeof(float) * 4 * (size_t)w * h);
    orc_set_image(d->inputFormat, w, h, image, img);
    if (!d->isNormalMap) orc_to_linear(img, w, h, d->inputGamma);
    long total = 0;
    for (int m = 0; m < mips; m++) {
        if (m > 0) {
            const int dw = imax(1, w / 2), dh = imax(1, h / 2);
            float fw = d->mipmapFilter == 0 ? 0.5f : (d->mipmapFilter == 1 ? 1.0f : d->kaiserWidth);
            orc_next_mipmap(img, w, h, tmp, d->mipmapFilter, fw, d->kaiserAlpha, d->kaiserStretch, d->wrapMode, d->alphaMode);
            float *t = img; img = tmp; tmp = t;
            w = dw; h = dh;
            if (d->isNormalMap && d->normalizeMipmaps) orc_renormalize(img, w, h);
        }
        memcpy(tmp, img, sizeof(float) * 4 * (size_t)w * h);
        if (!d->isNormalMap) orc_to_gamma(tmp, w, h, d->outputGamma);
        long n = orc_compress_level(d->format, d->quality, d->alphaMode, w, h, tmp, d->colorWeights, out + total);
        if (n == 0) { total = -1; break; }
        total += n;
    }
    free(img);
    free(tmp);
    return total;
}
