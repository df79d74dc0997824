// 2:1 polyphase mip filter (Kaiser 13 taps, Mitchell 9, Triangle 5) as ONE persistent kernel per level, Blackwell style:
//   * the source footprint of a 32x32 output tile (75x75 texels for Kaiser) is fetched by TMA (cp.async.bulk.tensor, 3-D tensor
//     map over [plane][y][x]) into a two-stage shared-memory ring; an mbarrier per stage carries the transaction count, so the
//     fetch of the next (tile, plane) runs under the arithmetic of the current one and no thread spends registers or issue
//     slots on staging.  Where a footprint crosses the image border TMA writes zeros; those few positions are then patched
//     with the texels the wrap mode (clamp / repeat / mirror) selects;
//   * X pass (FloatImage::applyKernelX, FloatImage.cpp:1115-1144) and Y pass (applyKernelY, :1146-1176) are register blocked:
//     a thread produces four neighbouring outputs from one sliding window (6 LDS.128 for 52 multiply-adds in X, 19 LDS.32
//     for 52 in Y) instead of one shared-memory load per multiply-add - the old kernel was LSU bound (58 % LSU pipe);
//   * the four planes of a tile are consecutive ring entries, so the normal-map renormalisation of the mip
//     (expandNormals -> normalizeNormalMap -> packNormals, Context.cpp:329-334) is applied to the tile before it is written:
//     no separate pass over the new level;
//   * FUSE (opt-in, NVTT_B200_FUSED_MIP_ENCODE): the tile of the NEW level is block-encoded where it lies - BC4 / BC5 at
//     Quality_Fastest / Normal (QuickCompress::compressDXT5A, bc_alpha.cuh) from the quantised red (and green) plane kept in
//     shared memory, 64 blocks x channels per tile, one thread each - so the level is never read back for its encode.
//     Byte-identical to k_alpha_blocks; measured slower than the two kernels side by side, hence opt-in.
// Every output is the same ascending-tap sum of single-rounded products as in k_polyphase_x / _y (no FMA), so results are
// bit-identical to the reference.  At an exact 2:1 ratio every output column (row) has the same 13 weights and
// left[i] = 2 i + left0 - the host checks both on the tables it built from the reference's formulae before taking this path.
#pragma once
#include "../nvb_common.cuh"
#include "image_ops.cuh"
#include "bc_alpha.cuh"
#ifndef NVB_EMU
#include <cuda.h>
#endif

namespace nvb {

#define NVB_PT_TW 32
#define NVB_PT_TH 32
#define NVB_PT_THREADS 320
#define NVB_PT_MAXW 13

template <int W> struct PtGeom {
    static constexpr int IN_W = 2 * NVB_PT_TW + W - 2;  // source columns under a tile
    static constexpr int IN_H = 2 * NVB_PT_TH + W - 2;
    // The innermost TMA coordinate must be a multiple of 16 bytes (measured on B200: any other start faults with "illegal
    // instruction", profiles/microbench/tma_probe.cu), so the box starts SHIFT texels left of the footprint: left0 = -(W - 3) / 2
    // for a centred 2:1 filter, and 2 tx0 + left0 - SHIFT is a multiple of 4.
    static constexpr int LEFT0 = -(W - 3) / 2;
    static constexpr int SHIFT = ((LEFT0 % 4) + 4) % 4;
    // box rows: a multiple of 16 bytes, and an ODD number of 16-byte groups so that lanes walking down the rows hit different banks
    static constexpr int BOX_W0 = (SHIFT + IN_W + 3) & ~3;
    static constexpr int BOX_W = ((BOX_W0 / 4) & 1) ? BOX_W0 : BOX_W0 + 4;
    static constexpr int TMP_PITCH = NVB_PT_TW + 4;
    static constexpr int STAGE_BYTES = ((BOX_W * IN_H * 4) + 127) & ~127;
    static constexpr int TMP_BYTES = ((IN_H * TMP_PITCH * 4) + 127) & ~127;
    static constexpr int OUT_PITCH = NVB_PT_TW + 1;
    static constexpr int OUT_BYTES = 3 * NVB_PT_TH * OUT_PITCH * 4;
    static constexpr int SMEM_BYTES = 2 * STAGE_BYTES + TMP_BYTES + OUT_BYTES;
    static constexpr int NV = (SHIFT + 2 * 4 + W - 2 + 3) / 4;   // float4 loads per X-pass window
    static_assert(8 * (NVB_PT_TW / 4 - 1) + 4 * NV <= BOX_W, "the last window stays inside the staged row");
};

struct PolyTmaParams {
    const float *src;
    float *dst;
    int sw, sh, dw, dh;
    int wrap;
    int left0x, left0y;  // left[i] = 2 i + left0
    float wx[NVB_PT_MAXW], wy[NVB_PT_MAXW];
    int tiles_x, tiles_y;
    int normalize;       // planes 0..2 of the new level: x = 2x - 1, normalise (zero stays zero), x = 0.5x + 0.5
    // FUSE: block-encode the new level's first enc_channels planes (1 = BC4, 2 = BC5) as DXT5 alpha blocks
    unsigned char *enc_out = nullptr;
    int enc_channels = 0, enc_stride = 8;
    const float *enc_to_gamma = nullptr;  // fused Surface::toGamma(2.2) of the colour pipeline (not for normal maps)
};

#ifndef NVB_EMU
NVB_DEV unsigned pt_smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
NVB_DEV void pt_mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pt_smem_addr(bar)), "r"(count));
}
NVB_DEV void pt_mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pt_smem_addr(bar)), "r"(bytes) : "memory");
}
NVB_DEV void pt_mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(pt_smem_addr(bar)),
        "r"(parity)
        : "memory");
}
NVB_DEV void pt_tma_load_3d(void *smem_dst, const CUtensorMap *map, unsigned long long *bar, int x, int y, int z) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                     pt_smem_addr(smem_dst)),
                 "l"(map), "r"(pt_smem_addr(bar)), "r"(x), "r"(y), "r"(z)
                 : "memory");
}

template <int W, bool FUSE = false> __global__ void __launch_bounds__(NVB_PT_THREADS, 3) k_polyphase_tma(const __grid_constant__ CUtensorMap tmap, PolyTmaParams P) {
    using G = PtGeom<W>;
    extern __shared__ __align__(128) unsigned char pt_smem[];
    __shared__ __align__(8) unsigned long long mbar[2];
    // FUSE: the tile's encoded channels, quantised like ColorBlock::init does (uint8(255 * clamp(v, 0, 1)), truncation)
    __shared__ unsigned char s_q[FUSE ? 2 : 1][FUSE ? NVB_PT_TH : 1][FUSE ? NVB_PT_TW + 4 : 4];
    float *const s_in0 = reinterpret_cast<float *>(pt_smem);
    float *const s_in1 = reinterpret_cast<float *>(pt_smem + G::STAGE_BYTES);
    float *const s_tmp = reinterpret_cast<float *>(pt_smem + 2 * G::STAGE_BYTES);
    float *const s_out = reinterpret_cast<float *>(pt_smem + 2 * G::STAGE_BYTES + G::TMP_BYTES);
    const int tid = threadIdx.x;
    const int ntiles = P.tiles_x * P.tiles_y;
    const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int nitems = my_tiles * 4;  // (tile, plane) ring entries of this CTA
    float wx[W], wy[W];
#pragma unroll
    for (int j = 0; j < W; j++) {
        wx[j] = P.wx[j];
        wy[j] = P.wy[j];
    }
    // weight pairs for the packed products: X pairs start at the first tap whose sample sits at an even position of the window
    pf2 wxp[W / 2 + 1], wyp[W / 2 + 1];
#pragma unroll
    for (int j = (G::SHIFT & 1); j + 1 < W; j += 2) wxp[(j - (G::SHIFT & 1)) / 2] = pf2_pack(P.wx[j], P.wx[j + 1]);
#pragma unroll
    for (int j = 0; j + 1 < W; j += 2) wyp[j / 2] = pf2_pack(P.wy[j], P.wy[j + 1]);
    if (tid == 0) {
        pt_mbar_init(&mbar[0], 1);
        pt_mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // plain copies of what the loops need (kernel parameters would be re-read from the constant bank at every use)
    const int sw = P.sw, sh = P.sh, dw = P.dw, dh = P.dh, wrap = P.wrap, tiles_x = P.tiles_x, normalize = P.normalize;
    const int left0x = P.left0x, left0y = P.left0y;
    const float *const src = P.src;
    float *const dst = P.dst;
    const size_t dn = (size_t)dw * dh;

    struct Tile {
        int tx0, ty0, tw, th, x_lo, y_lo;
        bool interior;
    };
    auto tile_of = [&](int k) {  // k-th tile of this CTA
        Tile T;
        const int tile = (int)blockIdx.x + k * (int)gridDim.x;
        const int tyi = tile / tiles_x;
        T.tx0 = (tile - tyi * tiles_x) * NVB_PT_TW;
        T.ty0 = tyi * NVB_PT_TH;
        T.tw = min(NVB_PT_TW, dw - T.tx0);
        T.th = min(NVB_PT_TH, dh - T.ty0);
        T.x_lo = 2 * T.tx0 + left0x;
        T.y_lo = 2 * T.ty0 + left0y;
        // TMA fills what lies outside the image with zeros; the wrap modes need real texels there
        T.interior = T.tw == NVB_PT_TW && T.th == NVB_PT_TH && T.x_lo >= 0 && T.y_lo >= 0 && T.x_lo + G::IN_W <= sw && T.y_lo + G::IN_H <= sh;
        return T;
    };
    auto issue = [&](const Tile &T, int plane, int st) {  // thread 0 only
        // the stage was read (generic proxy) by the X pass two entries ago; order those reads before the async-proxy write
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        pt_mbar_expect_tx(&mbar[st], (unsigned)(G::BOX_W * G::IN_H * 4));
        pt_tma_load_3d(st ? s_in1 : s_in0, &tmap, &mbar[st], T.x_lo - G::SHIFT, T.y_lo, plane);
    };

    // X-pass work split: thread -> (row, first column group); rows run along the lanes (conflict-free 128-bit accesses).
    // NVB_PT_THREADS = 4 * XROWS: a thread owns one staged row and the column groups xg0 and xg0 + 4.
    constexpr int XROWS = NVB_PT_THREADS / 4;
    static_assert(XROWS >= G::IN_H && NVB_PT_TW / 4 == 8, "every staged row has a thread");
    const int xr = tid % XROWS, xg0 = tid / XROWS;
    const int yox = tid & (NVB_PT_TW - 1), ygy = tid >> 5;  // Y pass: column, group of four rows (warps 0..7)

    unsigned phase_bits = 0u;  // bit st = parity the next wait on stage st expects
    Tile T = tile_of(0), Tn = T;
    if (my_tiles > 0 && tid == 0) issue(T, 0, 0);
    for (int it = 0; it < nitems; it++) {
        const int plane = it & 3, st = it & 1;
        if (plane == 3 && (it >> 2) + 1 < my_tiles) Tn = tile_of((it >> 2) + 1);
        if (it + 1 < nitems && tid == 0) issue(plane == 3 ? Tn : T, (it + 1) & 3, st ^ 1);
        float *const s_in = st ? s_in1 : s_in0;
        pt_mbar_wait(&mbar[st], (phase_bits >> st) & 1u);
        phase_bits ^= 1u << st;
        if (!T.interior) {
            // border tile: TMA has filled what lies outside the image with zeros; the wrap mode wants real texels there.
            // Only the few out-of-image positions of the footprint are patched (generic stores after the barrier completed).
            const int nrows = 2 * T.th + W - 2, ncols = 2 * T.tw + W - 2;
            const float *pl = src + (size_t)plane * sw * sh;
            for (int i = tid; i < G::IN_H * G::BOX_W; i += NVB_PT_THREADS) {
                const int r = i / G::BOX_W, c = i - r * G::BOX_W - G::SHIFT;
                const int sy = T.y_lo + r, sx = T.x_lo + c;
                if (((unsigned)sy >= (unsigned)sh || (unsigned)sx >= (unsigned)sw) && r < nrows && c >= 0 && c < ncols)
                    s_in[i] = __ldg(pl + (size_t)wrap_coord(sy, sh, wrap) * sw + wrap_coord(sx, sw, wrap));
            }
            __syncthreads();
        }
        // X pass
        if (xr < G::IN_H) {
#pragma unroll
            for (int half = 0; half < 2; half++) {
                const int g = xg0 + 4 * half;
                const float4 *row = reinterpret_cast<const float4 *>(s_in + xr * G::BOX_W + 8 * g);
                float s[4 * G::NV];
#pragma unroll
                for (int k = 0; k < G::NV; k++) {
                    const float4 v = row[k];
                    s[4 * k] = v.x;
                    s[4 * k + 1] = v.y;
                    s[4 * k + 2] = v.z;
                    s[4 * k + 3] = v.w;
                }
                // the products of two neighbouring taps are one FMUL2 (the staged values are aligned register pairs already, the
                // weight pairs are packed once); the sum stays the reference's ascending chain of single-rounded adds
                float o[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    float a = 0.0f;
                    constexpr int J0 = G::SHIFT & 1;  // first tap whose sample sits at an even position
                    if (J0) a += wx[0] * s[G::SHIFT + 2 * q];
#pragma unroll
                    for (int j = J0; j + 1 < W; j += 2) {
                        const pf2 pp = pf2_mul(wxp[(j - J0) / 2], pf2_pack(s[G::SHIFT + 2 * q + j], s[G::SHIFT + 2 * q + j + 1]));
                        a += pf2_lo(pp);
                        a += pf2_hi(pp);
                    }
                    if (((W - J0) & 1) != 0) a += wx[W - 1] * s[G::SHIFT + 2 * q + W - 1];
                    o[q] = a;
                }
                *reinterpret_cast<float4 *>(s_tmp + xr * G::TMP_PITCH + 4 * g) = make_float4(o[0], o[1], o[2], o[3]);
            }
        }
        __syncthreads();
        // Y pass
        const bool keep = normalize && plane < 3;
        if (ygy < NVB_PT_TH / 4 && 4 * ygy < T.th) {
            float s[2 * 4 + W - 2];
#pragma unroll
            for (int k = 0; k < 2 * 4 + W - 2; k++) s[k] = s_tmp[(8 * ygy + k) * G::TMP_PITCH + yox];
            float *const drow = dst + (size_t)plane * dn + (size_t)(T.ty0 + 4 * ygy) * dw + T.tx0 + yox;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                float a = 0.0f;
#pragma unroll
                for (int j = 0; j + 1 < W; j += 2) {
                    const pf2 pp = pf2_mul(wyp[j / 2], pf2_pack(s[2 * q + j], s[2 * q + j + 1]));
                    a += pf2_lo(pp);
                    a += pf2_hi(pp);
                }
                if (W & 1) a += wy[W - 1] * s[2 * q + W - 1];
                const int oy = 4 * ygy + q;
                if (oy < T.th && yox < T.tw) {
                    if (keep) s_out[(plane * NVB_PT_TH + oy) * G::OUT_PITCH + yox] = a;
                    else drow[(size_t)q * dw] = a;
                    if (FUSE && !normalize && plane < P.enc_channels)
                        s_q[plane][oy][yox] = (unsigned char)quantize_u8_trunc(P.enc_to_gamma ? nvb_powf_5_11(a, P.enc_to_gamma) : a);
                }
            }
        }
        __syncthreads();
        if (normalize && plane == 2) {
            // the tile's x, y, z are complete: FloatImage::scaleBias(2, -1), normalize (normalizeSafe, epsilon 0), scaleBias(0.5, 0.5)
            for (int i = tid; i < NVB_PT_TW * NVB_PT_TH; i += NVB_PT_THREADS) {
                const int ox = i & (NVB_PT_TW - 1), oy = i >> 5;
                if (ox >= T.tw || oy >= T.th) continue;
                float x = s_out[(0 * NVB_PT_TH + oy) * G::OUT_PITCH + ox], y = s_out[(1 * NVB_PT_TH + oy) * G::OUT_PITCH + ox],
                      z = s_out[(2 * NVB_PT_TH + oy) * G::OUT_PITCH + ox];
                x = 2.0f * x + -1.0f;
                y = 2.0f * y + -1.0f;
                z = 2.0f * z + -1.0f;
                const float l = sqrtf(x * x + y * y + z * z);
                if (fabsf(l) <= 0.0f) {
                    x = 0.0f; y = 0.0f; z = 0.0f;
                } else {
                    const float sc = 1.0f / l;
                    x = x * sc; y = y * sc; z = z * sc;
                }
                x = 0.5f * x + 0.5f;
                y = 0.5f * y + 0.5f;
                z = 0.5f * z + 0.5f;
                const size_t o = (size_t)(T.ty0 + oy) * dw + T.tx0 + ox;
                dst[o] = x;
                dst[dn + o] = y;
                dst[2 * dn + o] = z;
                if (FUSE) {  // normal maps are not gamma corrected (Context.cpp:337-343)
                    s_q[0][oy][ox] = (unsigned char)quantize_u8_trunc(x);
                    if (P.enc_channels > 1) s_q[1][oy][ox] = (unsigned char)quantize_u8_trunc(y);
                }
            }
            // s_out is next written by the Y pass of the following tile's plane 0, two barriers from here
        }
        if (FUSE && plane == (normalize ? 2 : P.enc_channels - 1)) {
            // the encoded channels of the tile are complete: one thread per (4x4 block, channel), partial blocks at the image
            // edge repeat their texels by modulo like alpha_gather_block (ColorBlock::init)
            if (normalize) __syncthreads();  // the renormalised values (the plain path is behind the Y pass's barrier already)
            const int nbx = (T.tw + 3) >> 2, nby = (T.th + 3) >> 2;
            for (int u = tid; u < 64 * P.enc_channels; u += NVB_PT_THREADS) {
                const int chan = u >> 6, bx = u & 7, by = (u >> 3) & 7;
                if (bx >= nbx || by >= nby) continue;
                const int twb = min(T.tw - 4 * bx, 4), thb = min(T.th - 4 * by, 4);
                unsigned v[16];
#pragma unroll
                for (int i = 0; i < 16; i++) v[i] = s_q[chan][4 * by + (i >> 2) % thb][4 * bx + (i & 3) % twb];
                const unsigned long long blk = alpha_quick_compress(v);
                const size_t bi = (size_t)(T.ty0 / 4 + by) * ((dw + 3) >> 2) + (T.tx0 / 4 + bx);
                *reinterpret_cast<uint2 *>(P.enc_out + bi * P.enc_stride + 8 * chan) = make_uint2((unsigned)(blk & 0xFFFFFFFFu), (unsigned)(blk >> 32));
            }
            // s_q is next written two (plain) or three (normal map) barriers from here
        }
        if (plane == 3) T = Tn;
    }
}
#endif  // NVB_EMU

}  // namespace nvb
