// Format_RGB / Format_RGBA: the uncompressed pixel-format writer (planar fp32 -> packed scanlines).
//
// Replaces, byte for byte:
//   PixelFormatConverter::compress                    src/nvtt/CompressorRGB.cpp:410-575
//   BitStream::putBits / putFloat / putHalf / flush    src/nvtt/CompressorRGB.cpp:332-404
//   toFloat11 / toFloat10                              src/nvtt/CompressorRGB.cpp:129-160
//   PixelFormat::convert (16 -> n bits)                src/nvimage/PixelFormat.h:37-53
//   nv::half_from_float                                src/nvmath/Half.cpp:378-441  (bc6h.cuh)
//   toFloat3SE (R9G9B9E5, PixelType_SharedExp)         src/nvtt/CompressorRGB.cpp:231-269
//
// The reference walks every scanline through a small bit stream.  A pixel whose bit count is a whole number of bytes
// (and, for float types, whose channels are 0/16/32 bits wide) always starts on an empty stream, so those formats - all
// the ones nvcompress can produce - run one thread per pixel (k_pixel_format).  Anything else (10/11-bit floats, odd bit
// counts) carries stream state from pixel to pixel and is run one thread per scanline (k_pixel_format_rows) with the same
// state machine, quirks included: the stream's pending bits sit ABOVE the new value ((buffer << n) | p in 32-bit
// arithmetic), bytes leave from the low end, toFloat11/10 do not mask their exponent.
// HBM-bound: 16 B read + bitCount/8 B written per pixel.
#pragma once
#include "../nvb_common.cuh"
#include "bc6h.cuh"

namespace nvb {

struct PixelFormatParams {
    LevelView lv;
    unsigned char *out;   // h scanlines of `pitch` bytes
    unsigned pitch;       // computeBytePitch(w, bitCount, pitchAlignment)
    unsigned bitCount;
    int kind;             // 0 UnsignedNorm, 1 UnsignedInt, 2 Float, 3 signed types / other SharedExp layouts (zeros, as the reference writes), 4 R9G9B9E5
    unsigned size[4];     // r, g, b, a
    unsigned shift[4];
};

// BitStream of the reference; `ptr` is an offset into the scanline
struct PfStream {
    unsigned char *base;
    unsigned ptr;
    unsigned buffer;  // uint8 in the reference
    unsigned bits;
    unsigned limit;   // pitch: never write past the scanline (the reference's malloc'ed line is exactly pitch bytes)
};

NVB_DEV void pf_put_byte(PfStream &s, unsigned v) {
    if (s.ptr < s.limit) s.base[s.ptr] = (unsigned char)v;
    s.ptr++;
}

NVB_DEV void pf_put_bits(PfStream &s, unsigned p, unsigned bitCount) {
    // (this->buffer << bitCount) | p : int << int, or'ed with a uint => 32-bit unsigned, then widened
    unsigned long long buffer = (unsigned long long)((s.buffer << (bitCount & 31u)) | p);
    unsigned bits = s.bits + bitCount;
    while (bits >= 8) {
        pf_put_byte(s, (unsigned)(buffer & 0xFF));
        buffer >>= 8;
        bits -= 8;
    }
    s.buffer = (unsigned)(buffer & 0xFF);
    s.bits = bits;
}

NVB_DEV void pf_put_raw(PfStream &s, unsigned v, int nbytes) {  // putFloat / putHalf: stored at ptr whatever the bit state
    for (int i = 0; i < nbytes; i++) pf_put_byte(s, (v >> (8 * i)) & 0xFF);
}

NVB_DEV unsigned pf_to_float11(float f) {
    // the clamps select BITS: a float min/max would replace a NaN's payload by the canonical one, the reference's compares keep it
    const unsigned u = (f < 0.0f) ? 0u : (f > 65024.0f) ? 0x477E0000u : __float_as_uint(f);
    const unsigned E = ((u >> 23) & 0xFF) - 127 + 15;
    const unsigned M = (u & 0x7FFFFF) >> (23 - 6);
    return (E << 6) | M;
}
NVB_DEV unsigned pf_to_float10(float f) {
    const unsigned u = (f < 0.0f) ? 0u : (f > 64512.0f) ? 0x477C0000u : __float_as_uint(f);
    const unsigned E = ((u >> 23) & 0xFF) - 127 + 15;
    const unsigned M = (u & 0x7FFFFF) >> (23 - 5);
    return (E << 5) | M;
}

NVB_DEV void pf_put_float_channel(PfStream &s, float v, unsigned size) {
    if (size == 32) pf_put_raw(s, __float_as_uint(v), 4);
    else if (size == 16) pf_put_raw(s, half_from_float_bits(__float_as_uint(v)) & 0xFFFFu, 2);
    else if (size == 11) pf_put_bits(s, pf_to_float11(v), 11);
    else if (size == 10) pf_put_bits(s, pf_to_float10(v), 10);
    else pf_put_bits(s, 0, size);
}

// ftoi_round = cvtss2si under the default rounding mode: nearest even; NaN and out-of-range give INT_MIN
NVB_DEV int pf_ftoi_round(float f) {
    return (f >= -2147483648.0f && f < 2147483648.0f) ? __float2int_rn(f) : (int)0x80000000;
}
// toFloat3SE as the reference's x86-64 build computes it.  Two things are kept on purpose:
//   * the divisor 1 << (exp_shared - B - N) has a NEGATIVE shift count for every value below 256; x86 takes the count modulo
//     32, so those pixels get the right exponent and zero mantissas (1 << 31 is INT_MIN: the quotient is a tiny negative);
//   * the exponent comes from floor(log2f(max)), and a correctly rounded log2f returns k for the last few floats below 2^k.
//     log2 in double, rounded to float, reproduces that (checked against glibc's log2f around every power of two:
//     tests/test_oracle.py); the float pipeline never sees the double otherwise (this layout is rare: rank 4 of SURVEY 8f).
NVB_DEV unsigned pf_to_float3se(float r, float g, float b) {
    const int N = 9, B = 15;
    const float sharedexp_max = 65408.0f;
    r = nv_max(0.0f, nv_min(sharedexp_max, r));
    g = nv_max(0.0f, nv_min(sharedexp_max, g));
    b = nv_max(0.0f, nv_min(sharedexp_max, b));
    const float max_c = nv_max(r, nv_max(g, b));
    const float lg = (float)log2((double)max_c);  // log2f(0) = -inf -> floorf -> cvtss2si = INT_MIN
    const int exp_shared_p = nv_max(-B - 1, pf_ftoi_round(floorf(lg))) + 1 + B;
    const float dp = (float)(int)(1u << ((unsigned)(exp_shared_p - B - N) & 31u));
    const int max_s = pf_ftoi_round(max_c / dp);
    int exp_shared = exp_shared_p;
    if (max_s == (1 << N)) exp_shared++;
    const float ds = (float)(int)(1u << ((unsigned)(exp_shared - B - N) & 31u));
    const unsigned xm = (unsigned)pf_ftoi_round(r / ds) & 0x1FFu, ym = (unsigned)pf_ftoi_round(g / ds) & 0x1FFu;
    const unsigned zm = (unsigned)pf_ftoi_round(b / ds) & 0x1FFu;
    return xm | (ym << 9) | (zm << 18) | (((unsigned)exp_shared & 31u) << 27);
}

// PixelFormat::convert(c, 16, outbits)
NVB_DEV unsigned pf_convert16(unsigned c, unsigned outbits) {
    unsigned inbits = 16, r = 0;
    // bitexpand: (c << (out - in)) | convert(c, in, out - in), unrolled; ends with a truncation
    while (inbits < outbits) {
        r |= c << (outbits - inbits);
        outbits -= inbits;
    }
    return r | (c >> (inbits - outbits));
}

// iround(clamp(v, 0, 65535)) = int(floorf(x + 0.5f)); nv::clamp = min(max(x, a), b) with nv::max(a,b) = a>b?a:b, so NaN -> lower bound
NVB_DEV unsigned pf_to_u16(float v) {
    float c = nv_clamp(v, 0.0f, 65535.0f);
    return (unsigned)__float2int_rz(floorf(c + 0.5f));  // c is inside [0, 65535]: no out-of-range case
}

// the fixed-point pixel: 16-bit components converted to their field widths and or'ed together
NVB_DEV unsigned pf_fixed_pixel(const PixelFormatParams &P, float r, float g, float b, float a) {
    unsigned ir, ig, ib, ia;
    if (P.kind == 0) {
        ir = pf_to_u16(r * 65535.0f); ig = pf_to_u16(g * 65535.0f); ib = pf_to_u16(b * 65535.0f); ia = pf_to_u16(a * 65535.0f);
    } else {
        ir = pf_to_u16(r); ig = pf_to_u16(g); ib = pf_to_u16(b); ia = pf_to_u16(a);
    }
    unsigned p = 0;
    p |= pf_convert16(ir, P.size[0]) << (P.shift[0] & 31u);
    p |= pf_convert16(ig, P.size[1]) << (P.shift[1] & 31u);
    p |= pf_convert16(ib, P.size[2]) << (P.shift[2] & 31u);
    p |= pf_convert16(ia, P.size[3]) << (P.shift[3] & 31u);
    return p;
}

NVB_DEV void pf_put_pixel(const PixelFormatParams &P, PfStream &s, int x, int y) {
    const float r = load_texel(P.lv, 0, x, y), g = load_texel(P.lv, 1, x, y), b = load_texel(P.lv, 2, x, y), a = load_texel(P.lv, 3, x, y);
    if (P.kind == 2) {
        pf_put_float_channel(s, r, P.size[0]);
        pf_put_float_channel(s, g, P.size[1]);
        pf_put_float_channel(s, b, P.size[2]);
        pf_put_float_channel(s, a, P.size[3]);
    } else if (P.kind == 3) {
        pf_put_bits(s, 0, P.bitCount);
    } else if (P.kind == 4) {
        pf_put_bits(s, pf_to_float3se(r, g, b), 32);
    } else {
        pf_put_bits(s, pf_fixed_pixel(P, r, g, b, a), P.bitCount);
    }
}

// Four pixels per thread for the layouts that dominate in practice: 8/16/32-bit fixed point, RGBA16F and RGBA32F, on rows
// whose width is a multiple of 4 and whose pitch keeps 16-byte alignment.  128-bit plane loads, one 4..64-byte store run.
// mode: 1, 2, 4 = bytes of a fixed-point pixel; 8 = four halfs; 16 = four floats.
__global__ void __launch_bounds__(256) k_pixel_format_x4(PixelFormatParams P, int mode) {
    const int x4 = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x4 * 4 >= P.lv.w) return;
    const size_t off = (size_t)y * P.lv.w + (size_t)x4 * 4;
    const float4 R = *reinterpret_cast<const float4 *>(P.lv.data + off);
    const float4 G = *reinterpret_cast<const float4 *>(P.lv.data + P.lv.plane + off);
    const float4 B = *reinterpret_cast<const float4 *>(P.lv.data + 2 * P.lv.plane + off);
    const float4 A = *reinterpret_cast<const float4 *>(P.lv.data + 3 * P.lv.plane + off);
    unsigned char *row = P.out + (size_t)y * P.pitch;
    if (mode <= 4) {
        const unsigned p0 = pf_fixed_pixel(P, R.x, G.x, B.x, A.x), p1 = pf_fixed_pixel(P, R.y, G.y, B.y, A.y);
        const unsigned p2 = pf_fixed_pixel(P, R.z, G.z, B.z, A.z), p3 = pf_fixed_pixel(P, R.w, G.w, B.w, A.w);
        if (mode == 4) reinterpret_cast<uint4 *>(row)[x4] = make_uint4(p0, p1, p2, p3);
        else if (mode == 2) reinterpret_cast<uint2 *>(row)[x4] = make_uint2((p0 & 0xFFFFu) | (p1 << 16), (p2 & 0xFFFFu) | (p3 << 16));
        else reinterpret_cast<unsigned *>(row)[x4] = (p0 & 0xFFu) | ((p1 & 0xFFu) << 8) | ((p2 & 0xFFu) << 16) | (p3 << 24);
    } else if (mode == 8) {
        const float r[4] = {R.x, R.y, R.z, R.w}, g[4] = {G.x, G.y, G.z, G.w}, b[4] = {B.x, B.y, B.z, B.w}, a[4] = {A.x, A.y, A.z, A.w};
        unsigned w[8];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            w[2 * i] = (half_from_float_bits(__float_as_uint(r[i])) & 0xFFFFu) | (half_from_float_bits(__float_as_uint(g[i])) << 16);
            w[2 * i + 1] = (half_from_float_bits(__float_as_uint(b[i])) & 0xFFFFu) | (half_from_float_bits(__float_as_uint(a[i])) << 16);
        }
        uint4 *o = reinterpret_cast<uint4 *>(row) + (size_t)x4 * 2;
        o[0] = make_uint4(w[0], w[1], w[2], w[3]);
        o[1] = make_uint4(w[4], w[5], w[6], w[7]);
    } else {
        float4 *o = reinterpret_cast<float4 *>(row) + (size_t)x4 * 4;
        o[0] = make_float4(R.x, G.x, B.x, A.x);
        o[1] = make_float4(R.y, G.y, B.y, A.y);
        o[2] = make_float4(R.z, G.z, B.z, A.z);
        o[3] = make_float4(R.w, G.w, B.w, A.w);
    }
}

// One thread per pixel: byte-aligned pixels.  The last thread of a row also writes the zero padding up to `pitch`.
__global__ void __launch_bounds__(256) k_pixel_format(PixelFormatParams P) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= P.lv.w) return;
    const unsigned bytes = P.bitCount >> 3;
    unsigned char *row = P.out + (size_t)y * P.pitch;
    if ((P.kind <= 1 || P.kind == 4) && bytes == 4 && (P.pitch & 3u) == 0 && ((size_t)P.out & 3u) == 0) {
        // the common 32-bit case: one aligned store
        unsigned char tmp[4];
        PfStream s{tmp, 0, 0, 0, 4};
        pf_put_pixel(P, s, x, y);
        reinterpret_cast<unsigned *>(row)[x] = tmp[0] | (tmp[1] << 8) | (tmp[2] << 16) | ((unsigned)tmp[3] << 24);
    } else {
        PfStream s{row, x * bytes, 0, 0, P.pitch};
        pf_put_pixel(P, s, x, y);
    }
    if (x == P.lv.w - 1)
        for (unsigned i = (unsigned)P.lv.w * bytes; i < P.pitch; i++) row[i] = 0;
}

// One thread per scanline: the general bit stream.
__global__ void __launch_bounds__(64) k_pixel_format_rows(PixelFormatParams P) {
    const int y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= P.lv.h) return;
    unsigned char *row = P.out + (size_t)y * P.pitch;
    PfStream s{row, 0, 0, 0, P.pitch};
    for (int x = 0; x < P.lv.w; x++) pf_put_pixel(P, s, x, y);
    // align(): flush the pending bits, then zero bytes up to the pitch
    if (s.bits) {
        pf_put_byte(s, s.buffer);
        s.buffer = 0;
        s.bits = 0;
    }
    while (s.ptr < P.pitch) pf_put_byte(s, 0);
}

}  // namespace nvb
