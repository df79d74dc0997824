// Image-op kernels of the mip chain: input conversion, gamma, box / polyphase down-sampling, normal-map
// renormalisation.  All are HBM-bound streaming kernels over planar fp32 [c][y][x]; arithmetic follows the
// reference operation-for-operation (no FMA contraction, same summation order) so results are bit-identical.
//
// Replaces:
//   Surface::setImage (BGRA8 / RGBA16F / RGBA32F / R32F)   src/nvtt/Surface.cpp:728-815
//   half_to_float                                          src/nvmath/Half.cpp:443-498
//   FloatImage::toLinear / toGamma / exponentiate          src/nvimage/FloatImage.cpp:259-298 (Gamma.cpp:311-354)
//   FloatImage::fastDownSample                             src/nvimage/FloatImage.cpp:559-737
//   FloatImage::applyKernelX / applyKernelY                src/nvimage/FloatImage.cpp:1115-1176 (wrap: FloatImage.h:298-351)
//   FloatImage::scaleBias / normalize                      src/nvimage/FloatImage.cpp:201-242
#pragma once
#include "../nvb_common.cuh"

namespace nvb {

// ---------------------------------------------------------------------------------------------------------
// Input conversion (+ optional fused toLinear on R,G,B).
// ---------------------------------------------------------------------------------------------------------
NVB_DEV float half_bits_to_float(unsigned h) {
    const unsigned s = (h & 0x8000u) << 16;
    const unsigned e = (h >> 10) & 0x1Fu;
    const unsigned m = h & 0x3FFu;
    unsigned r;
    if (e == 0) {
        if (m == 0) r = 0;
        else {
            // denormal half: normalise
            const int nlz = __clz((int)m) - 21;  // leading zeros within the 11-bit field (1..10)
            const unsigned mm = (m << nlz) & 0x3FFu;
            r = ((unsigned)(113 - nlz) << 23) | (mm << 13);
        }
    } else if (e == 31) {
        r = 0x7F800000u | (m << 13);
    } else {
        r = ((e + 112u) << 23) | (m << 13);
    }
    return __uint_as_float(s | r);
}

struct SetImageParams {
    const void *src;  // interleaved input texels
    float *dst;       // planar fp32 RGBA
    int count;        // pixels converted by this launch
    int format;       // nvtt::InputFormat: 0 BGRA_8UB, 1 RGBA_16F, 2 RGBA_32F, 3 R_32F
    const float *to_linear_table;  // non-null => fused Surface::toLinear(2.2) on R,G,B
    size_t plane;     // floats between output planes (= pixels of the whole image; a row band converts `count` < plane pixels)
};

__global__ void __launch_bounds__(256) k_set_image(SetImageParams P) {
    const size_t n = (size_t)P.count;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float r, g, b, a;
        if (P.format == 0) {
            const uchar4 c = reinterpret_cast<const uchar4 *>(P.src)[i];  // memory order B,G,R,A (Color32)
            r = (float)c.z / 255.0f;
            g = (float)c.y / 255.0f;
            b = (float)c.x / 255.0f;
            a = (float)c.w / 255.0f;
        } else if (P.format == 1) {
            const ushort4 c = reinterpret_cast<const ushort4 *>(P.src)[i];
            r = half_bits_to_float(c.x);
            g = half_bits_to_float(c.y);
            b = half_bits_to_float(c.z);
            a = half_bits_to_float(c.w);
        } else if (P.format == 2) {
            const float4 c = reinterpret_cast<const float4 *>(P.src)[i];
            r = c.x; g = c.y; b = c.z; a = c.w;
        } else {
            r = reinterpret_cast<const float *>(P.src)[i];
            g = 0.0f; b = 0.0f; a = 0.0f;
        }
        if (P.to_linear_table != nullptr) {
            r = nvb_powf_11_5(r, P.to_linear_table);
            g = nvb_powf_11_5(g, P.to_linear_table);
            b = nvb_powf_11_5(b, P.to_linear_table);
        }
        P.dst[i] = r;
        P.dst[P.plane + i] = g;
        P.dst[2 * P.plane + i] = b;
        P.dst[3 * P.plane + i] = a;
    }
}

// BGRA8 input, four pixels per thread: a byte has 256 values, so "float(c) / 255.0f" (an IEEE division, ~12 instructions) and the
// fused toLinear(2.2) (table x degree-4 polynomial) are looked up in two 256-entry shared-memory tables that every CTA fills
// with exactly those expressions first - the same bits, ~20 instructions per pixel instead of ~100, and the kernel becomes
// the copy it should be: one 128-bit load, four 128-bit stores per thread.  Needs 16-byte aligned src / dst / plane and a
// pixel count that is a multiple of 4.
__global__ void __launch_bounds__(256) k_set_image_bgra8_x4(SetImageParams P) {
    __shared__ float s_lin[256], s_unit[256];
    {
        const float u = (float)threadIdx.x / 255.0f;
        s_unit[threadIdx.x] = u;
        s_lin[threadIdx.x] = P.to_linear_table != nullptr ? nvb_powf_11_5(u, P.to_linear_table) : u;
    }
    __syncthreads();
    const size_t n4 = (size_t)P.count >> 2;
    const uint4 *src = reinterpret_cast<const uint4 *>(P.src);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 c = __ldg(src + i);  // four Color32: memory order B, G, R, A
        const unsigned px[4] = {c.x, c.y, c.z, c.w};
        float r[4], g[4], b[4], a[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            b[k] = s_lin[px[k] & 0xFFu];
            g[k] = s_lin[(px[k] >> 8) & 0xFFu];
            r[k] = s_lin[(px[k] >> 16) & 0xFFu];
            a[k] = s_unit[px[k] >> 24];
        }
        float4 *d0 = reinterpret_cast<float4 *>(P.dst) + i;
        const size_t p4 = P.plane >> 2;
        d0[0] = make_float4(r[0], r[1], r[2], r[3]);
        d0[p4] = make_float4(g[0], g[1], g[2], g[3]);
        d0[2 * p4] = make_float4(b[0], b[1], b[2], b[3]);
        d0[3 * p4] = make_float4(a[0], a[1], a[2], a[3]);
    }
}

// In-place gamma on the first 3 planes.  mode 0: powf_11_5 (toLinear 2.2), 1: powf_5_11 (toGamma 2.2),
// 2: powf(max(0,x), power) (any other gamma; libm vs CUDA powf => tolerance, not bit-exact).
struct GammaParams {
    float *data;
    size_t count;  // 3 * pixels
    int mode;
    const float *table;
    float power;
};

__global__ void __launch_bounds__(256) k_gamma(GammaParams P) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < P.count; i += (size_t)gridDim.x * blockDim.x) {
        float v = P.data[i];
        if (P.mode == 0) v = nvb_powf_11_5(v, P.table);
        else if (P.mode == 1) v = nvb_powf_5_11(v, P.table);
        else v = powf(nv_max(0.0f, v), P.power);
        P.data[i] = v;
    }
}

// ---------------------------------------------------------------------------------------------------------
// fastDownSample: 2x2 box for even sizes, 3-tap polyphase box for odd sizes, 1-D when one side is 1.
// ---------------------------------------------------------------------------------------------------------
struct BoxDownParams {
    const float *src;
    float *dst;
    int sw, sh;  // source size
    int dw, dh;  // max(1, sw/2), max(1, sh/2)
    int planes;  // 4
};

// The common case of a power-of-two chain: even x even source whose width is a multiple of 4.  Two outputs per thread from two
// 128-bit loads, no index divisions; each output is 0.25f * (((s00 + s01) + s10) + s11), the expression of the generic kernel.
__global__ void __launch_bounds__(256) k_box_down_even4(BoxDownParams P) {
    const int x2 = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, c = blockIdx.z;
    if (x2 * 2 >= P.dw) return;
    const float *s = P.src + (size_t)c * P.sw * P.sh + (size_t)(2 * y) * P.sw + 4 * (size_t)x2;
    const float4 a = __ldg(reinterpret_cast<const float4 *>(s));
    const float4 b = __ldg(reinterpret_cast<const float4 *>(s + P.sw));
    float2 o;
    o.x = 0.25f * (a.x + a.y + b.x + b.y);
    o.y = 0.25f * (a.z + a.w + b.z + b.w);
    *reinterpret_cast<float2 *>(P.dst + (size_t)c * P.dw * P.dh + (size_t)y * P.dw + 2 * (size_t)x2) = o;
}

__global__ void __launch_bounds__(256) k_box_down(BoxDownParams P) {
    const int sw = P.sw, sh = P.sh, w = P.dw, h = P.dh;
    const size_t per_plane = (size_t)w * h;
    const size_t total = per_plane * P.planes;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i / per_plane);
        const size_t r = i - (size_t)c * per_plane;
        const int y = (int)(r / w), x = (int)(r - (size_t)y * w);
        const float *src = P.src + (size_t)c * sw * sh;
        float out;
        if (sw == 1 || sh == 1) {
            const unsigned n = (unsigned)(w * h);
            const unsigned k = (unsigned)r;  // linear index along the long side
            const float *s = src + 2 * (size_t)k;
            if ((sw * sh) & 1) {
                const float scale = 1.0f / (float)(2 * n + 1);
                const float w0 = (float)(n - k), w1 = (float)(n - 0), w2 = (float)(1 + k);
                out = scale * (w0 * s[0] + w1 * s[1] + w2 * s[2]);
            } else {
                out = 0.5f * (s[0] + s[1]);
            }
        } else if ((sw & 1) == 0 && (sh & 1) == 0) {
            const float *s = src + (size_t)(2 * y) * sw + 2 * x;
            out = 0.25f * (s[0] + s[1] + s[sw] + s[sw + 1]);
        } else if ((sw & 1) && (sh & 1)) {
            const float scale = 1.0f / (float)(sw * sh);
            const float v0 = (float)(h - y), v1 = (float)(h - 0), v2 = (float)(1 + y);
            const float w0 = (float)(w - x), w1 = (float)(w - 0), w2 = (float)(1 + x);
            const float *s = src + (size_t)(2 * y) * sw + 2 * x;
            float f = 0.0f;
            f += v0 * (w0 * s[0] + w1 * s[1] + w2 * s[2]);
            f += v1 * (w0 * s[sw] + w1 * s[sw + 1] + w2 * s[sw + 2]);
            f += v2 * (w0 * s[2 * sw] + w1 * s[2 * sw + 1] + w2 * s[2 * sw + 2]);
            out = f * scale;
        } else if (sw & 1) {
            const float scale = 1.0f / (float)(2 * sw);
            const float w0 = (float)(w - x), w1 = (float)(w - 0), w2 = (float)(1 + x);
            const float *s = src + (size_t)(2 * y) * sw + 2 * x;
            float f = 0.0f;
            f += w0 * (s[0] + s[sw]);
            f += w1 * (s[1] + s[sw + 1]);
            f += w2 * (s[2] + s[sw + 2]);
            out = f * scale;
        } else {
            const float scale = 1.0f / (float)(2 * sh);
            const float v0 = (float)(h - y), v1 = (float)(h - 0), v2 = (float)(1 + y);
            const float *s = src + (size_t)(2 * y) * sw + 2 * x;
            float f = 0.0f;
            f += v0 * (s[0] + s[1]);
            f += v1 * (s[sw] + s[sw + 1]);
            f += v2 * (s[2 * sw] + s[2 * sw + 1]);
            out = f * scale;
        }
        P.dst[i] = out;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Polyphase separable resize.  Per output coordinate i the host precomputes (Filter.cpp:563-608 restated in
// host/polyphase.cpp): left[i] and `win` normalised weights.  sum = ((0 + w0*s0) + w1*s1) + ... in tap order.
// ---------------------------------------------------------------------------------------------------------
NVB_DEV int wrap_coord(int x, int w, int mode) {
    if (mode == 0) {  // clamp
        return nv_min(nv_max(x, 0), w - 1);
    } else if (mode == 1) {  // repeat
        if (x >= 0) return x % w;
        return (x + 1) % w + w - 1;
    } else {  // mirror
        if (w == 1) x = 0;
        x = abs(x);
        while (x >= w) x = abs(w + w - x - 2);
        return x;
    }
}

struct PolyphaseParams {
    const float *src;
    float *dst;
    int sw, sh;  // source plane size
    int dw, dh;  // destination plane size (X pass: dh == sh; Y pass: dw == sw)
    int planes;
    int win;            // taps per output
    const float *weights;  // [len][win]
    const int *left;       // [len]
    int wrap;              // nvtt::WrapMode: 0 clamp, 1 repeat, 2 mirror
};

__global__ void __launch_bounds__(256) k_polyphase_x(PolyphaseParams P) {
    const size_t per_plane = (size_t)P.dw * P.dh;
    const size_t total = per_plane * P.planes;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i / per_plane);
        const size_t r = i - (size_t)c * per_plane;
        const int y = (int)(r / P.dw), x = (int)(r - (size_t)y * P.dw);
        const float *row = P.src + (size_t)c * P.sw * P.sh + (size_t)y * P.sw;
        const float *wt = P.weights + (size_t)x * P.win;
        const int left = P.left[x];
        float sum = 0;
        for (int j = 0; j < P.win; j++) {
            const int idx = wrap_coord(left + j, P.sw, P.wrap);
            sum += wt[j] * row[idx];
        }
        P.dst[i] = sum;
    }
}

__global__ void __launch_bounds__(256) k_polyphase_y(PolyphaseParams P) {
    const size_t per_plane = (size_t)P.dw * P.dh;
    const size_t total = per_plane * P.planes;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i / per_plane);
        const size_t r = i - (size_t)c * per_plane;
        const int y = (int)(r / P.dw), x = (int)(r - (size_t)y * P.dw);
        const float *col = P.src + (size_t)c * P.sw * P.sh + x;
        const float *wt = P.weights + (size_t)y * P.win;
        const int left = P.left[y];
        float sum = 0;
        for (int j = 0; j < P.win; j++) {
            const int idx = wrap_coord(left + j, P.sh, P.wrap);
            sum += wt[j] * col[(size_t)idx * P.sw];
        }
        P.dst[i] = sum;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Fused separable resize: one CTA produces a 32x32 tile of one destination plane.  The source footprint of the tile is
// staged in shared memory once (wrap mode applied while loading), the X pass is evaluated for the footprint's rows
// into a second shared buffer and the Y pass reads that — the `dw x sh` intermediate image of FloatImage::resize
// (FloatImage.cpp:761-808) never goes to HBM.  Every output is the same tap-ordered sum as in k_polyphase_x / _y, so the
// result is bit-identical; HBM traffic drops from 16+8+8+4 to 16+4 bytes per source texel and channel group.
// ---------------------------------------------------------------------------------------------------------
#define NVB_PF_TILE 32
#define NVB_PF_EXT 88   // max source rows / columns a tile may need (2:1 Kaiser width 3 needs 75)
#define NVB_PF_MAXWIN 32

struct Polyphase2DParams {
    const float *src;
    float *dst;
    int sw, sh, dw, dh;
    int winx, winy;
    const float *wx;   // [dw][winx]
    const int *leftx;  // [dw]
    const float *wy;   // [dh][winy]
    const int *lefty;  // [dh]
    int wrap;
};

// source column cc of a staged row lives at NVB_PF_COL(cc): even and odd columns are kept apart so that the stride-2
// reads of a 2:1 X pass (lane ox reads column 2*ox + const) hit 32 different banks
#define NVB_PF_HALF ((NVB_PF_EXT + 1) / 2 + 1)
#define NVB_PF_COL(cc) ((((cc) & 1) ? NVB_PF_HALF : 0) + ((cc) >> 1))

// WX / WY: compile-time window sizes of the X and Y kernels (13 = Kaiser, 9 = Mitchell, 5 = Triangle at 2:1), 0 = run time.
// With known windows the tap loops are fully unrolled, the X weights of a lane sit in registers and the staged columns are
// addressed as (even base | odd base) + constant; the sums add the same products in the same ascending tap order.
template <int WX, int WY> __global__ void __launch_bounds__(256) k_polyphase_2d_t(Polyphase2DParams P) {
    __shared__ float s_in[NVB_PF_EXT][2 * NVB_PF_HALF];
    __shared__ float s_tmp[NVB_PF_EXT][NVB_PF_TILE + 1];
    const int c = blockIdx.z;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int tx0 = blockIdx.x * NVB_PF_TILE, ty0 = blockIdx.y * NVB_PF_TILE;
    const int tw = min(NVB_PF_TILE, P.dw - tx0), th = min(NVB_PF_TILE, P.dh - ty0);
    const int winx = WX ? WX : P.winx, winy = WY ? WY : P.winy;
    const int x_lo = P.leftx[tx0], x_hi = P.leftx[tx0 + tw - 1] + winx;
    const int y_lo = P.lefty[ty0], y_hi = P.lefty[ty0 + th - 1] + winy;
    const int ncols = x_hi - x_lo, nrows = y_hi - y_lo;
    const float *plane = P.src + (size_t)c * P.sw * P.sh;
    // 1. stage the footprint: one warp per row, lanes along x (coalesced); wrap only for tiles that touch the border
    const bool interior = x_lo >= 0 && x_hi <= P.sw && y_lo >= 0 && y_hi <= P.sh;
    // Four rows x three column groups = up to 12 independent loads in flight per thread before the first store (the plain
    // load -> store loop left the kernel waiting on one global load at a time: long-scoreboard 11 per issue, 13 % of HBM peak).
    static_assert(NVB_PF_EXT <= 96, "three column groups of 32 cover a staged row");
    for (int r0 = wid; r0 < nrows; r0 += 32) {
        float v[4][3];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int r = r0 + 8 * k;
            const bool rok = r < nrows;
            const int sy = !rok ? 0 : interior ? y_lo + r : wrap_coord(y_lo + r, P.sh, P.wrap);
            const float *row = plane + (size_t)sy * P.sw;
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const int cc = lane + 32 * j;
                const bool ok = rok && cc < ncols;
                const int sx = !ok ? 0 : interior ? x_lo + cc : wrap_coord(x_lo + cc, P.sw, P.wrap);
                v[k][j] = ok ? __ldg(row + sx) : 0.0f;
            }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int r = r0 + 8 * k;
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const int cc = lane + 32 * j;
                if (r < nrows && cc < ncols) s_in[r][NVB_PF_COL(cc)] = v[k][j];
            }
        }
    }
    __syncthreads();
    // 2. X pass for every staged row: lane = output column, warp w owns rows w, w+8, ...
    if (lane < tw) {
        const float *wt = P.wx + (size_t)(tx0 + lane) * winx;
        const int off = P.leftx[tx0 + lane] - x_lo;
        if (WX) {
            float w[WX ? WX : 1];
#pragma unroll
            for (int j = 0; j < WX; j++) w[j] = wt[j];
            const int cE = NVB_PF_COL(off), cO = NVB_PF_COL(off + 1);  // column off + 2m lives at cE + m, off + 2m + 1 at cO + m
            for (int r = wid; r < nrows; r += 8) {
                const float *row = s_in[r];
                float a = 0.0f;
#pragma unroll
                for (int j = 0; j < WX; j++) a += w[j] * row[((j & 1) ? cO : cE) + (j >> 1)];
                s_tmp[r][lane] = a;
            }
        } else {
            // the taps are walked once and every row keeps its own running sum, so each sum still adds its taps in ascending order
            float acc[(NVB_PF_EXT + 7) / 8];
#pragma unroll
            for (int k = 0; k < (NVB_PF_EXT + 7) / 8; k++) acc[k] = 0.0f;
            for (int j = 0; j < winx; j++) {
                const float w = wt[j];
                const int col = NVB_PF_COL(off + j);
#pragma unroll
                for (int k = 0; k < (NVB_PF_EXT + 7) / 8; k++) {
                    const int r = wid + 8 * k;
                    if (r < nrows) acc[k] += w * s_in[r][col];
                }
            }
#pragma unroll
            for (int k = 0; k < (NVB_PF_EXT + 7) / 8; k++) {
                const int r = wid + 8 * k;
                if (r < nrows) s_tmp[r][lane] = acc[k];
            }
        }
    }
    __syncthreads();
    // 3. Y pass: warp w owns output rows w, w+8, w+16, w+24
    if (lane < tw) {
        for (int oy = wid; oy < th; oy += 8) {
            const float *wt = P.wy + (size_t)(ty0 + oy) * winy;
            const int off = P.lefty[ty0 + oy] - y_lo;
            float sum = 0;
            if (WY) {
#pragma unroll
                for (int j = 0; j < WY; j++) sum += wt[j] * s_tmp[off + j][lane];
            } else {
                for (int j = 0; j < winy; j++) sum += wt[j] * s_tmp[off + j][lane];
            }
            P.dst[(size_t)c * P.dw * P.dh + (size_t)(ty0 + oy) * P.dw + tx0 + lane] = sum;
        }
    }
}
// ---------------------------------------------------------------------------------------------------------
// Normal-map helpers.  k_scale_bias: ptr = scale*ptr + bias on planes 0..2.  k_normalize: normalizeSafe with
// epsilon 0 (zero vector stays zero).  k_renormalize = expandNormals -> normalizeNormalMap -> packNormals
// fused (element-wise, so fusing does not change a single bit), src/nvtt/Context.cpp:329-334.
// ---------------------------------------------------------------------------------------------------------
struct ScaleBiasParams {
    float *data;
    size_t count;  // 3 * pixels
    float scale, bias;
};
__global__ void __launch_bounds__(256) k_scale_bias(ScaleBiasParams P) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < P.count; i += (size_t)gridDim.x * blockDim.x)
        P.data[i] = P.scale * P.data[i] + P.bias;
}

// FloatImage::clamp(channel, 1, low, high)  (src/nvimage/FloatImage.cpp:245-256): nv::clamp = min(max(x, low), high)
struct ClampParams {
    float *data;
    size_t count;
    float low, high;
};
__global__ void __launch_bounds__(256) k_clamp(ClampParams P) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < P.count; i += (size_t)gridDim.x * blockDim.x)
        P.data[i] = nv_clamp(P.data[i], P.low, P.high);
}

// Surface::range(channel, &min, &max, alpha_channel, alpha_ref)  (src/nvtt/Surface.cpp:526-566): running `if (f < lo) lo = f;
// if (f > hi) hi = f` from (FLT_MAX, -FLT_MAX), optionally only where alpha > alpha_ref.  Comparisons with NaN are false
// there and here, so the result does not depend on the order of the scan.
struct RangeParams {
    const float *data;
    const float *alpha;  // null: every texel counts
    size_t count;
    float alpha_ref;
    float2 *partial;     // one (min, max) per CTA
};
__global__ void __launch_bounds__(256) k_range(RangeParams P) {
    float lo = FLT_MAX, hi = -FLT_MAX;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < P.count; i += (size_t)gridDim.x * blockDim.x) {
        if (P.alpha != nullptr && !(P.alpha[i] > P.alpha_ref)) continue;
        const float f = P.data[i];
        if (f < lo) lo = f;
        if (f > hi) hi = f;
    }
    __shared__ float slo[8], shi[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float l2 = __shfl_xor_sync(0xFFFFFFFFu, lo, o), h2 = __shfl_xor_sync(0xFFFFFFFFu, hi, o);
        if (l2 < lo) lo = l2;
        if (h2 > hi) hi = h2;
    }
    if ((threadIdx.x & 31) == 0) {
        slo[threadIdx.x >> 5] = lo;
        shi[threadIdx.x >> 5] = hi;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; i++) {
            if (slo[i] < lo) lo = slo[i];
            if (shi[i] > hi) hi = shi[i];
        }
        P.partial[blockIdx.x] = make_float2(lo, hi);
    }
}

// Surface::toneMap  (src/nvtt/Surface.cpp:2444-2494).  mode 0 = Linear and 3 = Lightmap (same code there: scale r, g, b by
// 1 / max3 when it exceeds 1), 1 = Reindhart (c / (c + 1)), 2 = Halo (1 - exp2f(-c); CUDA's exp2f, 2 ulp from glibc's).
struct ToneMapParams {
    float *data;
    size_t pixels;
    int mode;
};
__global__ void __launch_bounds__(256) k_tone_map(ToneMapParams P) {
    const size_t n = P.pixels;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float r = P.data[i], g = P.data[n + i], b = P.data[2 * n + i];
        if (P.mode == 0 || P.mode == 3) {
            const float m = nv_max(r, nv_max(g, b));
            if (m > 1.0f) {
                const float s = 1.0f / m;
                r *= s;
                g *= s;
                b *= s;
            }
        } else if (P.mode == 1) {
            r /= r + 1;
            g /= g + 1;
            b /= b + 1;
        } else {
            r = 1 - exp2f(-r);
            g = 1 - exp2f(-g);
            b = 1 - exp2f(-b);
        }
        P.data[i] = r;
        P.data[n + i] = g;
        P.data[2 * n + i] = b;
    }
}

// Surface::toRGBM(range, threshold)  (src/nvtt/Surface.cpp:1862-1946): per texel, search the 8-bit multiplier within +-16
// of the analytic one for the smallest reconstruction error of the 8-bit quantised colour; first strict minimum wins.
// (The `range` argument is shadowed by a local 255 in the reference, so it has no effect there either.)
struct ToRgbmParams {
    float *data;
    size_t pixels;
    float threshold;  // already clamped to [1e-6, 1]
};
__global__ void __launch_bounds__(256) k_to_rgbm(ToRgbmParams P) {
    const size_t n = P.pixels;
    const float threshold = P.threshold, range = 255.0f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float R = nv_clamp(P.data[i], 0.0f, 1.0f), G = nv_clamp(P.data[n + i], 0.0f, 1.0f), B = nv_clamp(P.data[2 * n + i], 0.0f, 1.0f);
        float M = nv_max(nv_max(R, G), nv_max(B, threshold));
        const int iM = __float2int_rn(ceilf((M - threshold) / (1 - threshold) * range));  // ftoi_ceil = cvtss2si(ceilf(x))
        float bestM = 0.0f, bestError = FLT_MAX;
        const int m0 = iM - 16 > 0 ? iM - 16 : 0, m1 = iM + 16 < 256 ? iM + 16 : 256;
        for (int m = m0; m < m1; m++) {
            const float fm = float(m) / range;
            const float Mq = fm * (1 - threshold) + threshold;
            const int ir = __float2int_rn(range * nv_clamp(R / Mq, 0.0f, 1.0f));
            const int ig = __float2int_rn(range * nv_clamp(G / Mq, 0.0f, 1.0f));
            const int ib = __float2int_rn(range * nv_clamp(B / Mq, 0.0f, 1.0f));
            const float fr = (float(ir) / range) * Mq, fg = (float(ig) / range) * Mq, fb = (float(ib) / range) * Mq;
            const float error = (R - fr) * (R - fr) + (G - fg) * (G - fg) + (B - fb) * (B - fb);
            if (error < bestError) {
                bestError = error;
                bestM = Mq;
            }
        }
        M = bestM;
        P.data[i] = nv_clamp(R / M, 0.0f, 1.0f);
        P.data[n + i] = nv_clamp(G / M, 0.0f, 1.0f);
        P.data[2 * n + i] = nv_clamp(B / M, 0.0f, 1.0f);
        P.data[3 * n + i] = (M - threshold) / (1 - threshold);
    }
}

// Surface::binarize(channel, threshold, dither = false): c = float(c > threshold)   (src/nvtt/Surface.cpp:2656-2670)
struct BinarizeParams {
    float *data;   // the channel's plane
    size_t count;
    float threshold;
};
__global__ void __launch_bounds__(256) k_binarize(BinarizeParams P) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < P.count; i += (size_t)gridDim.x * blockDim.x)
        P.data[i] = (P.data[i] > P.threshold) ? 1.0f : 0.0f;
}

// Surface::quantize(channel, bits, exactEndPoints, dither) and Surface::binarize(channel, threshold, dither = true)
// (src/nvtt/Surface.cpp:2656-2775).  Without dithering: element-wise.  With dithering: the reference's Floyd-Steinberg scan,
// where the value of texel (x, y) depends on the quantisation error d of (x-1, y) and of (x-1, y-1), (x, y-1), (x+1, y-1):
//     E(x, y) = (((0 + 1/16 d(x-1,y-1)) + 5/16 d(x,y-1)) + 3/16 d(x+1,y-1)) + 7/16 d(x-1,y)        [row1[] / row0[] cells]
//     q = Q(f + E),  d = f - q   (sic: the error that is propagated is f - q, not (f + E) - q)
// That recurrence is a wavefront with skew 2: one CTA per channel, one thread per row of a band of 1024 rows, thread r
// works on x = t - 2r at step t and reads the three errors of the row above from a 4-deep history in shared memory (one
// __syncthreads per step).  The last row of a band leaves its errors in global memory for the first row of the next band.
// Every cell is accumulated in the reference's order, so the result is bit-identical.
struct QuantizeParams {
    float *data;       // the channel's plane
    int w, h;
    float scale, offset0, offset1;  // quantize: q = saturate((floorf(v * scale + offset0) + offset1) / scale)
    float threshold;   // binarize: q = float(v > threshold)
    int binarize;
    int dither;
    float *carry;      // dither: w floats of scratch (errors of the last row of the previous band)
};
NVB_DEV float quantize_value(const QuantizeParams &P, float v) {
    if (P.binarize) return (v > P.threshold) ? 1.0f : 0.0f;
    return nv_clamp((floorf(v * P.scale + P.offset0) + P.offset1) / P.scale, 0.0f, 1.0f);
}
__global__ void __launch_bounds__(256) k_quantize(QuantizeParams P) {
    const size_t n = (size_t)P.w * P.h;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        P.data[i] = quantize_value(P, P.data[i]);
}
#define NVB_FS_ROWS 1024
__global__ void __launch_bounds__(NVB_FS_ROWS) k_quantize_dither(QuantizeParams P) {
    __shared__ float s_hist[NVB_FS_ROWS][4];  // s_hist[r][t & 3] = d of row r computed at step t
    const int r = threadIdx.x;
    int band = 0;
    for (int y0 = 0; y0 < P.h; y0 += NVB_FS_ROWS, ++band) {
        const int rows = min(NVB_FS_ROWS, P.h - y0);
        const int y = y0 + r;
        const bool active = r < rows;
        float *row = P.data + (size_t)(active ? y : 0) * P.w;
        const float *carry_in = P.carry + (size_t)(band & 1) * P.w;         // errors of the last row of the previous band
        float *carry_out = P.carry + (size_t)((band + 1) & 1) * P.w;
        float dl = 0.0f;  // d(x-1, y)
        const int steps = P.w + 2 * (rows - 1);
        for (int t = 0; t < steps; t++) {
            const int x = t - 2 * r;
            float d = 0.0f;
            if (active && x >= 0 && x < P.w) {
                float e = 0.0f;  // the row1[] cell: zero-initialised, then += in the order of the reference's scan
                if (y > 0) {
                    // thread r-1 computed x+1 at step t-1, x at t-2 and x-1 at t-3
                    if (x - 1 >= 0) e += (1.0f / 16.0f) * (r > 0 ? s_hist[r - 1][(t - 3) & 3] : carry_in[x - 1]);
                    e += (5.0f / 16.0f) * (r > 0 ? s_hist[r - 1][(t - 2) & 3] : carry_in[x]);
                    if (x + 1 < P.w) e += (3.0f / 16.0f) * (r > 0 ? s_hist[r - 1][(t - 1) & 3] : carry_in[x + 1]);
                }
                if (x > 0) e += (7.0f / 16.0f) * dl;
                const float f = row[x];
                const float q = quantize_value(P, f + e);
                d = f - q;
                row[x] = q;
                dl = d;
                if (r == rows - 1) carry_out[x] = d;
            }
            // slot t & 3 was last read during step t-1 (as "t-4"... i.e. never again): safe to overwrite before the barrier
            s_hist[r][t & 3] = d;
            __syncthreads();
        }
    }
}

struct NormalizeParams {
    float *data;
    size_t pixels;
    int expand_pack;  // 1 => x = 2x-1 before, x = 0.5x+0.5 after
};
__global__ void __launch_bounds__(256) k_normalize(NormalizeParams P) {
    const size_t n = P.pixels;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float x = P.data[i], y = P.data[n + i], z = P.data[2 * n + i];
        if (P.expand_pack) {
            x = 2.0f * x + -1.0f;
            y = 2.0f * y + -1.0f;
            z = 2.0f * z + -1.0f;
        }
        const float l = sqrtf(x * x + y * y + z * z);
        if (fabsf(l) <= 0.0f) {
            x = 0.0f; y = 0.0f; z = 0.0f;
        } else {
            const float s = 1.0f / l;
            x = x * s; y = y * s; z = z * s;
        }
        if (P.expand_pack) {
            x = 0.5f * x + 0.5f;
            y = 0.5f * y + 0.5f;
            z = 0.5f * z + 0.5f;
        }
        P.data[i] = x;
        P.data[n + i] = y;
        P.data[2 * n + i] = z;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Surface::toGreyScale (src/nvtt/Surface.cpp:1732-1756): grey = r*rs + g*gs + b*bs + a*as written to all four planes.
// The scales are normalised by their sum on the host exactly like the reference does.
// ---------------------------------------------------------------------------------------------------------
struct GreyScaleParams {
    float *data;    // planar fp32 RGBA, in place
    size_t pixels;
    float scale[4];
};

__global__ void __launch_bounds__(256) k_grey_scale(GreyScaleParams P) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < P.pixels; i += (size_t)gridDim.x * blockDim.x) {
        const float grey = P.data[i] * P.scale[0] + P.data[P.pixels + i] * P.scale[1] + P.data[2 * P.pixels + i] * P.scale[2] +
                           P.data[3 * P.pixels + i] * P.scale[3];
        P.data[i] = grey;
        P.data[P.pixels + i] = grey;
        P.data[2 * P.pixels + i] = grey;
        P.data[3 * P.pixels + i] = grey;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Surface::toNormalMap -> nv::createNormalMap(FloatImage*, wm, filterWeights) (src/nvimage/NormalMap.cpp:82-124,183-198):
// du/dv = 9x9 blended-Sobel response of the alpha plane (FloatImage::applyKernelXY, FloatImage.cpp:1019-1044: rows
// outer, columns inner, one running fp32 sum), n = normalize(du, dv, 1/16); alpha is copied.  The 9x9 kernel
// (Kernel2::initBlendedSobel + L1 normalize, Filter.cpp:494-560,358-369) is built on the host; kdv is its transpose.
// 162 taps per texel from a plane that lives in L2 after the first touch: HBM-bound at 4 B in + 16 B out per texel.
// ---------------------------------------------------------------------------------------------------------
struct NormalMapParams {
    const float *src;  // planar fp32 RGBA
    float *dst;        // planar fp32 RGBA (different buffer)
    int w, h;
    int wrap;
    float kdu[81];     // kdu[i*9 + e] = valueAt(e, i)
};

__global__ void __launch_bounds__(256) k_to_normal_map(NormalMapParams P) {
    const size_t pixels = (size_t)P.w * P.h;
    const float *alpha = P.src + 3 * pixels;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < pixels; p += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(p % P.w), y = (int)(p / P.w);
        float du = 0.0f, dv = 0.0f;
        int xs[9];
#pragma unroll
        for (int e = 0; e < 9; e++) xs[e] = wrap_coord(x + e - 4, P.w, P.wrap);
        for (int i = 0; i < 9; i++) {
            const float *row = alpha + (size_t)wrap_coord(y + i - 4, P.h, P.wrap) * P.w;
#pragma unroll
            for (int e = 0; e < 9; e++) {
                const float v = row[xs[e]];
                du += P.kdu[i * 9 + e] * v;
                dv += P.kdu[e * 9 + i] * v;  // kdv = transpose(kdu)
            }
        }
        const float hs = 1.0f / 16.0f;
        const float l = sqrtf(du * du + dv * dv + hs * hs);
        const float il = 1.0f / l;
        P.dst[p] = du * il;
        P.dst[pixels + p] = dv * il;
        P.dst[2 * pixels + p] = hs * il;
        P.dst[3 * pixels + p] = alpha[p];
    }
}

}  // namespace nvb
