// 2:1 polyphase mip filter (Kaiser 13 taps, Mitchell 9, Triangle 5) as ONE persistent kernel per level, Blackwell style:
//   * the source footprint of a 32x32 output tile (75x75 texels for Kaiser) is fetched by TMA (cp.async.bulk.tensor, 3-D tensor
//     map over [plane][y][x]) into a two-stage shared-memory ring; an mbarrier per stage carries the transaction count, so the
//     fetch of the next (tile, plane) runs under the arithmetic of the current one and no thread spends registers or issue
//     slots on staging.  Tiles whose footprint crosses the image border (wrap modes) are staged by hand instead;
//   * X pass (FloatImage::applyKernelX, FloatImage.cpp:1115-1144) and Y pass (applyKernelY, :1146-1176) are register blocked:
//     a thread produces four neighbouring outputs from one sliding window (5 LDS.128 for 52 multiply-adds in X, 19 LDS.32
//     for 52 in Y) instead of one shared-memory load per multiply-add - the old kernel was LSU bound (58 % LSU pipe);
//   * the four planes of a tile are consecutive ring entries, so the normal-map renormalisation of the mip
//     (expandNormals -> normalizeNormalMap -> packNormals, Context.cpp:329-334) is applied to the tile before it is written:
//     no separate pass over the new level.
// Every output is the same ascending-tap sum of single-rounded products as in k_polyphase_x / _y (no FMA), so results are
// bit-identical to the reference.  At an exact 2:1 ratio every output column (row) has the same 13 weights and
// left[i] = 2 i + left0 - the host checks both on the tables it built from the reference's formulae before taking this path.
#pragma once
#include "../nvb_common.cuh"
#include "image_ops.cuh"
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
    static constexpr int BOX_W = (IN_W + 3) & ~3;        // TMA box rows are multiples of 16 bytes
    static constexpr int TMP_PITCH = NVB_PT_TW + 4;
    static constexpr int STAGE_BYTES = ((BOX_W * IN_H * 4) + 127) & ~127;
    static constexpr int TMP_BYTES = ((IN_H * TMP_PITCH * 4) + 127) & ~127;
    static constexpr int OUT_PITCH = NVB_PT_TW + 1;
    static constexpr int OUT_BYTES = 3 * NVB_PT_TH * OUT_PITCH * 4;
    static constexpr int SMEM_BYTES = 2 * STAGE_BYTES + TMP_BYTES + OUT_BYTES;
    static constexpr int NV = (2 * 4 + W - 2 + 3) / 4;   // float4 loads per X-pass window
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

template <int W> __global__ void __launch_bounds__(NVB_PT_THREADS, 3) k_polyphase_tma(const __grid_constant__ CUtensorMap tmap, PolyTmaParams P) {
    using G = PtGeom<W>;
    extern __shared__ __align__(128) unsigned char pt_smem[];
    __shared__ __align__(8) unsigned long long mbar[2];
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
    if (tid == 0) {
        pt_mbar_init(&mbar[0], 1);
        pt_mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    struct Item {
        int tx0, ty0, tw, th, plane, x_lo, y_lo;
        bool interior;
    };
    auto item_of = [&](int it) {
        Item I;
        const int tile = (int)blockIdx.x + (it >> 2) * (int)gridDim.x;
        I.plane = it & 3;
        I.tx0 = (tile % P.tiles_x) * NVB_PT_TW;
        I.ty0 = (tile / P.tiles_x) * NVB_PT_TH;
        I.tw = min(NVB_PT_TW, P.dw - I.tx0);
        I.th = min(NVB_PT_TH, P.dh - I.ty0);
        I.x_lo = 2 * I.tx0 + P.left0x;
        I.y_lo = 2 * I.ty0 + P.left0y;
        // TMA fills what lies outside the image with zeros; the wrap modes need real texels there
        I.interior = I.tw == NVB_PT_TW && I.th == NVB_PT_TH && I.x_lo >= 0 && I.y_lo >= 0 && I.x_lo + G::IN_W <= P.sw && I.y_lo + G::IN_H <= P.sh;
        return I;
    };
    auto issue = [&](int it) {  // thread 0 only
        const Item I = item_of(it);
        if (!I.interior) return;
        const int st = it & 1;
        // the stage was read (generic proxy) by the X pass two entries ago; order those reads before the async-proxy write
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        pt_mbar_expect_tx(&mbar[st], (unsigned)(G::BOX_W * G::IN_H * 4));
        pt_tma_load_3d(st ? s_in1 : s_in0, &tmap, &mbar[st], I.x_lo, I.y_lo, I.plane);
    };

    unsigned phase_bits = 0u;  // bit st = parity the next wait on stage st expects
    if (nitems > 0 && tid == 0) issue(0);
    for (int it = 0; it < nitems; it++) {
        if (it + 1 < nitems && tid == 0) issue(it + 1);
        const Item I = item_of(it);
        const int st = it & 1;
        float *const s_in = st ? s_in1 : s_in0;
        const int nrows = 2 * I.th + W - 2, ncols = 2 * I.tw + W - 2;
        if (I.interior) {
            pt_mbar_wait(&mbar[st], (phase_bits >> st) & 1u);
            phase_bits ^= 1u << st;
        } else {
            const float *plane = P.src + (size_t)I.plane * P.sw * P.sh;
            for (int i = tid; i < nrows * G::BOX_W; i += NVB_PT_THREADS) {
                const int r = i / G::BOX_W, c = i - r * G::BOX_W;
                float v = 0.0f;
                if (c < ncols) v = __ldg(plane + (size_t)wrap_coord(I.y_lo + r, P.sh, P.wrap) * P.sw + wrap_coord(I.x_lo + c, P.sw, P.wrap));
                s_in[i] = v;
            }
            __syncthreads();
        }
        // X pass: item = (source row, group of four output columns); lanes run along the rows (conflict-free 128-bit accesses)
        for (int i = tid; i < nrows * (NVB_PT_TW / 4); i += NVB_PT_THREADS) {
            const int g = i / nrows, r = i - g * nrows;
            const float4 *row = reinterpret_cast<const float4 *>(s_in + r * G::BOX_W + 8 * g);
            float s[4 * G::NV];
#pragma unroll
            for (int k = 0; k < G::NV; k++) {
                const float4 v = row[k];
                s[4 * k] = v.x;
                s[4 * k + 1] = v.y;
                s[4 * k + 2] = v.z;
                s[4 * k + 3] = v.w;
            }
            float o[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                float a = 0.0f;
#pragma unroll
                for (int j = 0; j < W; j++) a += wx[j] * s[2 * q + j];
                o[q] = a;
            }
            *reinterpret_cast<float4 *>(s_tmp + r * G::TMP_PITCH + 4 * g) = make_float4(o[0], o[1], o[2], o[3]);
        }
        __syncthreads();
        // Y pass: item = (output column, group of four output rows); lanes run along the columns
        const bool keep = P.normalize && I.plane < 3;
        float *const dplane = P.dst + (size_t)I.plane * P.dw * P.dh;
        for (int i = tid; i < NVB_PT_TW * (NVB_PT_TH / 4); i += NVB_PT_THREADS) {
            const int ox = i & (NVB_PT_TW - 1), gy = i >> 5;
            if (4 * gy >= I.th) continue;
            float s[2 * 4 + W - 2];
#pragma unroll
            for (int k = 0; k < 2 * 4 + W - 2; k++) s[k] = s_tmp[(8 * gy + k) * G::TMP_PITCH + ox];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                float a = 0.0f;
#pragma unroll
                for (int j = 0; j < W; j++) a += wy[j] * s[2 * q + j];
                const int oy = 4 * gy + q;
                if (oy < I.th && ox < I.tw) {
                    if (keep) s_out[(I.plane * NVB_PT_TH + oy) * G::OUT_PITCH + ox] = a;
                    else dplane[(size_t)(I.ty0 + oy) * P.dw + I.tx0 + ox] = a;
                }
            }
        }
        __syncthreads();
        if (P.normalize && I.plane == 2) {
            // the tile's x, y, z are complete: FloatImage::scaleBias(2, -1), normalize (normalizeSafe, epsilon 0), scaleBias(0.5, 0.5)
            const size_t dn = (size_t)P.dw * P.dh;
            for (int i = tid; i < NVB_PT_TW * NVB_PT_TH; i += NVB_PT_THREADS) {
                const int ox = i & (NVB_PT_TW - 1), oy = i >> 5;
                if (ox >= I.tw || oy >= I.th) continue;
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
                const size_t o = (size_t)(I.ty0 + oy) * P.dw + I.tx0 + ox;
                P.dst[o] = x;
                P.dst[dn + o] = y;
                P.dst[2 * dn + o] = z;
            }
            // s_out is next written by the Y pass of the following tile's plane 0, two barriers from here
        }
    }
}
#endif  // NVB_EMU

}  // namespace nvb
