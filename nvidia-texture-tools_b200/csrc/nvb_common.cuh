// Shared device helpers for the sm_100a BCn + mip kernels.
//
// Parity rules: the whole library is compiled with -fmad=false so that nvcc never contracts
// a*b+c into an FMA — every +,-,*,/ and sqrtf below is a single IEEE-754 round-to-nearest operation, exactly
// like the pinned reference build (-O2 -ffp-contract=off, scalar code paths).  min/max/clamp and float->int
// conversions reproduce the *reference's* semantics (nvcore/Utils.h:158-204, x86 cvttss2si), not CUDA's.
#pragma once
#ifndef NVB_EMU
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>
#endif

#define NVB_DEV __device__ __forceinline__
// a real call instead of another inlined copy: for big helpers used at several places of one kernel (instruction-cache footprint)
#ifdef NVB_EMU
#define NVB_DEV_CALL static inline
#else
#define NVB_DEV_CALL __device__ __noinline__
#endif
// read-only lookup tables defined in headers (global memory, L1/L2-cached; plain static data under the emulator)
#ifdef NVB_EMU
#define NVB_TABLE static const
#else
#define NVB_TABLE static __device__ const
#endif

// nv::max(a,b) = (b < a) ? a : b  /  nv::min(a,b) = (a < b) ? a : b   (src/nvcore/Utils.h:161-188)
// NaN behaviour: max(NaN, x) = x, max(x, NaN) = NaN; min(NaN, x) = x; min(x, NaN) = NaN.
template <class T> NVB_DEV T nv_max(T a, T b) { return (b < a) ? a : b; }
template <class T> NVB_DEV T nv_min(T a, T b) { return (a < b) ? a : b; }
template <class T> NVB_DEV T nv_clamp(T x, T a, T b) { return nv_min(nv_max(x, a), b); }
// std::max(a,b) = (a < b) ? b : a ; std::min(a,b) = (b < a) ? b : a   (used by squish maths.h:163-178)
template <class T> NVB_DEV T std_max(T a, T b) { return (a < b) ? b : a; }
template <class T> NVB_DEV T std_min(T a, T b) { return (b < a) ? b : a; }

// (int)f as compiled for x86-64 (cvttss2si): NaN and out-of-range give INT_MIN ("integer indefinite").
// CUDA's cvt.rzi.s32.f32 saturates and maps NaN to 0, so the difference must be made explicit.
NVB_DEV int x86_ftoi(float f) {
    return (f >= -2147483648.0f && f < 2147483648.0f) ? (int)f : (int)0x80000000;
}
// (uint32)f on x86-64 is compiled as a 64-bit cvttss2si followed by truncation to 32 bits.
NVB_DEV unsigned x86_ftou(float f) {
    if (f >= -9223372036854775808.0f && f < 9223372036854775808.0f) return (unsigned)(long long)f;
    return 0u;  // 0x8000000000000000 truncated
}

// ---- packed fp32 pairs (sm_100 FMUL2 / FADD2: two IEEE round-to-nearest results per issue slot) -----------------------
// ptxas 12.9 contracts mul.rn.f32x2 -> add.rn.f32x2 into FFMA2 even with --fmad=false (measured; scalar ops are left
// alone), which would change results.  So a sum that consumes a packed product is always done with two scalar FADDs
// (f2add_s / f2sub_s); f2add / f2sub (one FADD2) are only used where neither operand is a product.
#ifdef NVB_EMU
NVB_DEV float2 f2mul(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
NVB_DEV float2 f2add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
NVB_DEV float2 f2sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
#else
NVB_DEV float2 f2mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
NVB_DEV float2 f2add(float2 a, float2 b) { return __fadd2_rn(a, b); }
NVB_DEV float2 f2sub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
#endif
#ifdef NVB_EMU
NVB_DEV float2 f2add_s(float2 a, float2 b) { return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)); }
NVB_DEV float2 f2sub_s(float2 a, float2 b) { return make_float2(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y)); }
#elif defined(NVB_FASTMATH)
// the fast-math build (libnvtt_b200_fastmath.so, outside the parity contract) lets ptxas contract these into FFMA2
NVB_DEV float2 f2add_s(float2 a, float2 b) { return f2add(a, b); }
NVB_DEV float2 f2sub_s(float2 a, float2 b) { return f2sub(a, b); }
#else
// An unfused packed sum in ONE issue slot: fma(a, 1, b) rounds a*1 + b once, i.e. it IS add.rn(a, b) bit for bit (and
// fma(b, -1, a) is sub.rn(a, b)), and an FFMA2 cannot be contracted with the FMUL2 that produced its operand.  The ones
// live in constant memory so that neither NVVM nor ptxas can see their value and fold the fma back into an add.
static __constant__ float2 nvb_one2 = {1.0f, 1.0f};
static __constant__ float2 nvb_mone2 = {-1.0f, -1.0f};
NVB_DEV float2 f2add_s(float2 a, float2 b) { return __ffma2_rn(a, nvb_one2, b); }
NVB_DEV float2 f2sub_s(float2 a, float2 b) { return __ffma2_rn(b, nvb_mone2, a); }
#endif
NVB_DEV float2 f2splat(float v) { return make_float2(v, v); }

// The same pairs kept PACKED in one 64-bit register across operations (pf2): the float2 intrinsics above re-pack their
// operands at every call and ptxas then materialises a fresh register pair per use; with pf2 a pair is packed once.
//   pf2_add / pf2_sub: plain FADD2, only for operands that are not products.  pf2_add_s / pf2_sub_s: the unfused
//   FFMA2-by-one sum, safe after a product.
#ifdef NVB_EMU
typedef float2 pf2;
NVB_DEV pf2 pf2_pack(float lo, float hi) { return make_float2(lo, hi); }
NVB_DEV float pf2_lo(pf2 v) { return v.x; }
NVB_DEV float pf2_hi(pf2 v) { return v.y; }
NVB_DEV pf2 pf2_mul(pf2 a, pf2 b) { return f2mul(a, b); }
NVB_DEV pf2 pf2_add(pf2 a, pf2 b) { return f2add(a, b); }
NVB_DEV pf2 pf2_sub(pf2 a, pf2 b) { return f2sub(a, b); }
NVB_DEV pf2 pf2_add_s(pf2 a, pf2 b) { return f2add_s(a, b); }
NVB_DEV pf2 pf2_sub_s(pf2 a, pf2 b) { return f2sub_s(a, b); }
#else
typedef unsigned long long pf2;
NVB_DEV pf2 pf2_pack(float lo, float hi) {
    pf2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
NVB_DEV float pf2_lo(pf2 v) { return __uint_as_float((unsigned)v); }
NVB_DEV float pf2_hi(pf2 v) { return __uint_as_float((unsigned)(v >> 32)); }
NVB_DEV pf2 pf2_mul(pf2 a, pf2 b) {
    pf2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
NVB_DEV pf2 pf2_fma(pf2 a, pf2 b, pf2 c) {
    pf2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
NVB_DEV pf2 pf2_add(pf2 a, pf2 b) {
    pf2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
NVB_DEV pf2 pf2_sub(pf2 a, pf2 b) {
    pf2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
#ifdef NVB_FASTMATH
NVB_DEV pf2 pf2_add_s(pf2 a, pf2 b) { return pf2_add(a, b); }
NVB_DEV pf2 pf2_sub_s(pf2 a, pf2 b) { return pf2_sub(a, b); }
#else
NVB_DEV pf2 pf2_add_s(pf2 a, pf2 b) { return pf2_fma(a, *reinterpret_cast<const pf2 *>(&nvb_one2), b); }
NVB_DEV pf2 pf2_sub_s(pf2 a, pf2 b) { return pf2_fma(b, *reinterpret_cast<const pf2 *>(&nvb_mone2), a); }
#endif
#endif
NVB_DEV pf2 pf2_splat(float v) { return pf2_pack(v, v); }

// ---- gamma 2.2 approximations (src/nvmath/Gamma.cpp:311-354) -------------------------------------------
// table[k] = float(2^((k-127)*p/q)) for the 9 "sign|exponent" bits; the tables are generated on the host by
// nvb::build_gamma_tables() (capi.cu) and uploaded once per context.
struct GammaTables {
    const float *to_gamma;   // pow_5_11 table, 512 floats (linear -> gamma 2.2)
    const float *to_linear;  // pow_11_5 table, 512 floats (gamma 2.2 -> linear)
};

NVB_DEV float nvb_powf_5_11(float x, const float *__restrict__ table) {
    unsigned u = __float_as_uint(x);
    int k = (int)(u >> 23);
    float m = __uint_as_float((u & 0x7FFFFFu) | (127u << 23));
    float pe = table[k];
    float pm = (((-0.0110083047f * m + 0.0905038750f) * m - 0.324697506f) * m + 0.876040946f) * m + 0.369160989f;
    return pe * pm;
}
NVB_DEV float nvb_powf_11_5(float x, const float *__restrict__ table) {
    unsigned u = __float_as_uint(x);
    int k = (int)(u >> 23);
    float m = __uint_as_float((u & 0x7FFFFFu) | (127u << 23));
    float pe = table[k];
    float pm = (((-0.00916587552f * m + 0.119315466f) * m + 1.01847068f) * m - 0.158338739f) * m + 0.0297184721f;
    return pe * pm;
}

// uint8(255 * clamp(v, 0, 1))  — truncating quantiser of ColorBlock::init (src/nvimage/ColorBlock.cpp:104-107)
NVB_DEV unsigned quantize_u8_trunc(float v) {
    float c = nv_clamp(v, 0.0f, 1.0f);
    return (unsigned)(int)(255.0f * c);
}

// How an encoder kernel reads one mip level: planar fp32 [c][y][x] (FloatImage layout, FloatImage.h:193-229).
struct LevelView {
    const float *__restrict__ data;  // 4 planes; row y of plane c starts at data + c*plane + y*w
    size_t plane;                 // floats between planes (w*h of the whole level; a row band of a level keeps the level's stride)
    int w, h;
    int bw, bh;                   // blocks per row / column = (w+3)/4, (h+3)/4
    const float *to_gamma_table;  // non-null => apply powf_5_11 to R,G,B while loading (fused Surface::toGamma)
    // Block-row sharding of one image over several GPUs: this view is the concatenation of one GPU's row chunks; chunk j
    // (cyc_rpc block rows of the view) is chunk j * cyc_n + cyc_i of the whole level.  cyc_rpc == 0: the view is the level.
    int cyc_rpc = 0, cyc_n = 1, cyc_i = 0;
};

// index of block `blk` of the view (row-major) inside the whole level's output
NVB_DEV size_t nvb_out_block(const LevelView &lv, int blk) {
    if (lv.cyc_rpc == 0) return (size_t)blk;
    const int by = blk / lv.bw, bx = blk - by * lv.bw;
    const int c = by / lv.cyc_rpc, r = by - c * lv.cyc_rpc;
    return (size_t)((c * lv.cyc_n + lv.cyc_i) * lv.cyc_rpc + r) * lv.bw + bx;
}

NVB_DEV float load_texel(const LevelView &lv, int c, int x, int y) {
    float v = lv.data[(size_t)c * lv.plane + (size_t)y * lv.w + x];
    if (lv.to_gamma_table != nullptr && c < 3) v = nvb_powf_5_11(v, lv.to_gamma_table);
    return v;
}
