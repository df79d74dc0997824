// TEST / DEBUG INFRASTRUCTURE ONLY — runs the product's device code (csrc/kernels/*.cuh) under the lock-step
// CPU emulator so kernel logic can be checked against the oracle in a container without a GPU.
#include "cuda_emu.h"
#include <algorithm>
#include "../../nvidia-texture-tools_b200/csrc/kernels/bc_alpha.cuh"
#include "../../nvidia-texture-tools_b200/csrc/kernels/bc3_color.cuh"
#include "../../nvidia-texture-tools_b200/csrc/kernels/bc1_icbc.cuh"
#include "../../nvidia-texture-tools_b200/csrc/kernels/bc1_quick.cuh"
#include "../../nvidia-texture-tools_b200/csrc/kernels/bc6h.cuh"
#include "../../nvidia-texture-tools_b200/csrc/kernels/bc7.cuh"
#include "bc7_coop.cuh"
#include "../../nvidia-texture-tools_b200/csrc/kernels/bc7_search.cuh"
#include "../../nvidia-texture-tools_b200/csrc/kernels/bc6h_search.cuh"
#include "../../nvidia-texture-tools_b200/csrc/kernels/image_ops.cuh"
#include "../../nvidia-texture-tools_b200/csrc/kernels/pixel_format.cuh"
#include "../../nvidia-texture-tools_b200/csrc/host_tables.h"

using namespace nvb;

static float g_to_gamma[512], g_to_linear[512];
static std::vector<uint16_t> g_cand;
static int g_cand_off[18];
static uint8_t g_om5[512], g_om6[512], g_om5a[512], g_om6a[512];
static std::vector<uint16_t> g_cand3;
static std::vector<uint32_t> g_cand_idx;
static int g_cand3_off[18];
static std::vector<uint16_t> g_four, g_three;
static int g_four_total[16], g_three_total[16];
static float g_mid5[32], g_mid6[64];
static uint8_t g_match5[512], g_match6[512];
static bool g_init = false;
static void init_tables() {
    if (g_init) return;
    build_gamma_tables(g_to_gamma, g_to_linear);
    build_squish_splits(g_cand, g_cand_off);
    build_omatch(g_om5, 32);
    build_omatch(g_om6, 64);
    build_omatch(g_om5a, 32, true);
    build_omatch(g_om6a, 64, true);
    build_squish_splits3(g_cand3, g_cand3_off);
    build_squish_split_indices(g_cand_idx);
    build_icbc_splits(g_four, g_four_total, g_three, g_three_total);
    build_icbc_midpoints(g_mid5, g_mid6);
    build_icbc_match(g_match5, 32);
    build_icbc_match(g_match6, 64);
    g_init = true;
}
static LevelView make_lv(const float *data, int w, int h, int gamma) {
    LevelView lv;
    lv.data = data; lv.plane = (size_t)w * h; lv.w = w; lv.h = h; lv.bw = (w + 3) / 4; lv.bh = (h + 3) / 4;
    lv.to_gamma_table = gamma ? g_to_gamma : nullptr;
    return lv;
}

// NVB_EMU_BC7 = "scalar" (thread per candidate), "coop" (warp per candidate) or unset: the product's searcher state machine
template <int M, int NCAND> static void emu_bc7_mode(Bc7Params &P, int nb) {
    if constexpr (M == 0 || M == 1 || M == 2 || M == 3 || M == 7)
        emu::launch(dim3((nb + NVB_BC7_ROUGH_WARPS - 1) / NVB_BC7_ROUGH_WARPS), dim3(NVB_BC7_ROUGH_WARPS * 32), 0, [&] { k_bc7_rough<M>(P, 0, nb); });
    const char *how = getenv("NVB_EMU_BC7");
    if (how && !strcmp(how, "scalar")) {
        emu::launch(dim3((nb * NCAND + 127) / 128), dim3(128), 0, [&] { k_bc7_refine<M, NCAND>(P); });
        return;
    }
    if constexpr (M != 4 && M != 5) {
        if (how && !strcmp(how, "coop")) {
            constexpr int WPC = NCAND >= 4 ? NCAND : 4, BPC = WPC / NCAND;
            emu::launch(dim3((nb + BPC - 1) / BPC), dim3(WPC * 32), 0, [&] { k_bc7_refine_coop<M, NCAND>(P); });
            return;
        }
    }
    // two chunks, to exercise the chunk offsets
    for (int blk0 = 0; blk0 < nb;) {
        const int n = blk0 == 0 ? (nb + 1) / 2 : nb - blk0;
        using X = Bc7X<M>;
        std::vector<float4> tiles((size_t)n * 16);
        std::vector<uint4> setup((size_t)n * NCAND * X::NR), sidx((size_t)n * NCAND), res((size_t)n * NCAND * X::NR * X::NLSB);
        Bc7SearchParams S;
        std::vector<unsigned> perm((size_t)n * NCAND * X::NR), counters(NVB_BX_COUNTERS, 0);
        S.P = P; S.blk0 = blk0; S.nblk = n; S.tiles = tiles.data(); S.setup = setup.data(); S.setup_idx = sidx.data(); S.res = res.data();
        S.perm = X::NR > 1 ? perm.data() : nullptr; S.counters = counters.data();
        emu::launch(dim3((n * 16 + 255) / 256), dim3(256), 0, [&] { k_bc7_tiles(S, tiles.data()); });
        emu::launch(dim3((n * NCAND + 127) / 128), dim3(128), 0, [&] { k_bc7_setup<M, NCAND>(S); });
        if constexpr (X::NR > 1) {
            const int og = (n * NCAND * X::NR + 255) / 256;
            emu::launch(dim3(og), dim3(256), 0, [&] { k_bc7_order<M, NCAND, 0>(S); });
            emu::launch(dim3(og), dim3(256), 0, [&] { k_bc7_order<M, NCAND, 1>(S); });
        }
        const int grid = 3;  // few threads: every thread walks several searchers
        if constexpr (M == 0 || M == 1 || M == 2 || M == 3 || M == 7) {
            if (how && !strcmp(how, "search1")) emu::launch(dim3(grid), dim3(128), 0, [&] { k_bc7_search<M, 0>(S); });
            else emu::launch(dim3(grid), dim3(128), 0, [&] { k_bc7_search2<M>(S); });
        } else {
            emu::launch(dim3(grid), dim3(128), 0, [&] { k_bc7_search<M, 0>(S); });
        }
        if constexpr (M == 4) {
            counters[0] = 0;
            emu::launch(dim3(grid), dim3(128), 0, [&] { k_bc7_search<M, 1>(S); });
        }
        emu::launch(dim3((n * NCAND + 127) / 128), dim3(128), 0, [&] { k_bc7_finish<M, NCAND>(S); });
        blk0 += n;
    }
}

extern "C" {

void emu_alpha_blocks(const float *planar, int w, int h, int channel, unsigned char *out, int stride, int offset, int gamma, int mode) {
    init_tables();
    AlphaBlocksParams P;
    P.lv = make_lv(planar, w, h, gamma);
    P.channel = channel; P.out = out; P.out_stride = stride; P.out_offset = offset; P.mode = mode;
    int nb = P.lv.bw * P.lv.bh;
    if (mode == 2) emu::launch(dim3((nb + 127) / 128), dim3(128), 0, [&] { k_alpha_dxt3(P); });
    else if (mode == 1) emu::launch(dim3((nb + 3) / 4), dim3(128), 0, [&] { k_alpha_optimal(P); });
    else emu::launch(dim3((nb + 127) / 128), dim3(128), 0, [&] { k_alpha_blocks(P); });
}

void emu_bc3_color_ex(const float *planar, int w, int h, const float *metric, int weight_by_alpha, unsigned char *out, int stride, int offset, int gamma, int dxt5n);
void emu_bc3_color(const float *planar, int w, int h, const float *metric, int weight_by_alpha, unsigned char *out, int stride, int offset, int gamma) {
    emu_bc3_color_ex(planar, w, h, metric, weight_by_alpha, out, stride, offset, gamma, 0);
}
void emu_bc3_color_ex(const float *planar, int w, int h, const float *metric, int weight_by_alpha, unsigned char *out, int stride, int offset, int gamma, int dxt5n) {
    init_tables();
    Bc3ColorParams P;
    P.lv = make_lv(planar, w, h, gamma);
    P.out = out; P.out_stride = stride; P.out_offset = offset;
    P.metric[0] = metric[0]; P.metric[1] = metric[1]; P.metric[2] = metric[2];
    P.weight_by_alpha = weight_by_alpha;
    P.dxt5n = dxt5n & 1;
    P.cand3 = g_cand3.data(); P.cand3_off = g_cand3_off; P.omatch5a = g_om5a; P.omatch6a = g_om6a;
    P.cand_idx = g_cand_idx.data();
    P.cand = g_cand.data(); P.cand_off = g_cand_off; P.omatch5 = g_om5; P.omatch6 = g_om6;
    int nb = P.lv.bw * P.lv.bh;
    if (dxt5n & 2) emu::launch(dim3((nb + NVB_BC3_GROUPS - 1) / NVB_BC3_GROUPS), dim3(NVB_BC3_GROUPS * 16), 0, [&] { k_bc1a_color(P); });
    else emu::launch(dim3((nb + NVB_BC3_GROUPS - 1) / NVB_BC3_GROUPS), dim3(NVB_BC3_GROUPS * 16), 0, [&] { k_bc3_color(P); });
}

void emu_bc1(const float *planar, int w, int h, const float *cw, int level, int transparency, unsigned char *out, int gamma) {
    init_tables();
    Bc1Params P;
    P.lv = make_lv(planar, w, h, gamma);
    P.out = out; P.out_stride = 8; P.out_offset = 0;
    P.level = level; P.transparency = transparency;
    P.cw[0] = cw[0]; P.cw[1] = cw[1]; P.cw[2] = cw[2];
    P.four = g_four.data(); P.three = g_three.data(); P.four_total = g_four_total; P.three_total = g_three_total;
    P.midpoints5 = g_mid5; P.midpoints6 = g_mid6; P.match5 = g_match5; P.match6 = g_match6;
    int nb = P.lv.bw * P.lv.bh;
    emu::launch(dim3((nb + NVB_BC1_GROUPS - 1) / NVB_BC1_GROUPS), dim3(NVB_BC1_GROUPS * 16), 0, [&] {
#define EMU_BC1_GO(KFN) KFN(P)
        NVB_BC1_DISPATCH(P, EMU_BC1_GO);
#undef EMU_BC1_GO
    });
}

void emu_bc6(const float *planar, int w, int h, int is_signed, int transparency, unsigned char *out, int gamma) {
    init_tables();
    Bc6Params P;
    P.lv = make_lv(planar, w, h, gamma);
    P.out = out; P.is_signed = is_signed; P.transparency = transparency;
    int nb = P.lv.bw * P.lv.bh;
    std::vector<float> rough((size_t)nb * 20), err((size_t)nb * 2);
    std::vector<unsigned char> cand((size_t)nb * 32);
    P.rough = rough.data(); P.cand = cand.data(); P.cand_err = err.data();
    emu::launch(dim3((nb + NVB_BC6_ROUGH_WARPS - 1) / NVB_BC6_ROUGH_WARPS), dim3(NVB_BC6_ROUGH_WARPS * 32), 0, [&] { k_bc6_rough(P, 3); });
    int padded = (nb + 127) / 128 * 128;
    const char *how = getenv("NVB_EMU_BC6");
    if (how && !strcmp(how, "scalar")) {  // thread per (block, kind)
        emu::launch(dim3(2 * padded / 128), dim3(128), 0, [&] { k_bc6_refine(P, padded); });
    } else {  // the product's searcher state machine
        std::vector<float4> tiles((size_t)nb * 16);
        std::vector<int4> meta((size_t)nb * 2), setup((size_t)nb * 6), res((size_t)nb * 6);
        std::vector<uint2> sidx((size_t)nb * 2);
        std::vector<unsigned> perm((size_t)nb * 2), counters(NVB_BC6_COUNTERS, 0);
        Bc6SearchParams S;
        S.P = P; S.tiles = tiles.data(); S.meta = meta.data(); S.setup = setup.data(); S.setup_idx = sidx.data(); S.res = res.data();
        S.perm = perm.data(); S.counters = counters.data();
        emu::launch(dim3((nb * 16 + 255) / 256), dim3(256), 0, [&] { k_bc6_tiles(S, tiles.data()); });
        emu::launch(dim3(2 * padded / 128), dim3(128), 0, [&] { k_bc6_setup(S, padded, -1); });
        emu::launch(dim3((nb * 2 + 255) / 256), dim3(256), 0, [&] { k_bc6_order<0>(S); });
        emu::launch(dim3((nb * 2 + 255) / 256), dim3(256), 0, [&] { k_bc6_order<1>(S); });
        emu::launch(dim3(2), dim3(128), 0, [&] { k_bc6_search<1>(S); });
        emu::launch(dim3(2), dim3(128), 0, [&] { k_bc6_search<2>(S); });
        emu::launch(dim3(2 * padded / 128), dim3(128), 0, [&] { k_bc6_finish(S, padded); });
    }
    emu::launch(dim3((nb + 255) / 256), dim3(256), 0, [&] { k_bc6_select(P); });
}

// mode_mask: bit m set = run mode m (the others keep error FLT_MAX).  errs (optional): [8][nb] per-mode errors.
void emu_bc7(const float *planar, int w, int h, unsigned char *out, int gamma, int mode_mask, float *errs, unsigned char *cands) {
    init_tables();
    Bc7Params P;
    P.lv = make_lv(planar, w, h, gamma);
    P.out = out;
    int nb = P.lv.bw * P.lv.bh;
    std::vector<unsigned char> shapes((size_t)nb * 5 * 16), cand((size_t)nb * 8 * 16);
    std::vector<float> err((size_t)nb * 8, FLT_MAX);
    P.shapes = shapes.data(); P.cand = cand.data(); P.cand_err = err.data();
    if (mode_mask & 1) emu_bc7_mode<0, 4>(P, nb);
    if (mode_mask & 2) emu_bc7_mode<1, 16>(P, nb);
    if (mode_mask & 4) emu_bc7_mode<2, 16>(P, nb);
    if (mode_mask & 8) emu_bc7_mode<3, 16>(P, nb);
    if (mode_mask & 16) emu_bc7_mode<4, 8>(P, nb);
    if (mode_mask & 32) emu_bc7_mode<5, 4>(P, nb);
    if (mode_mask & 64) emu_bc7_mode<6, 1>(P, nb);
    if (mode_mask & 128) emu_bc7_mode<7, 16>(P, nb);
    emu::launch(dim3((nb + 255) / 256), dim3(256), 0, [&] { k_bc7_select(P); });
    if (errs) memcpy(errs, err.data(), err.size() * sizeof(float));
    if (cands) memcpy(cands, cand.data(), cand.size());
}

unsigned emu_half_from_float(unsigned f) { return half_from_float_bits(f); }

void emu_dxt1_quick(const float *planar, int w, int h, int dxt1a, int dxt5n, unsigned char *out, int stride, int offset, int gamma) {
    init_tables();
    Dxt1QuickParams Q;
    Q.lv = make_lv(planar, w, h, gamma);
    Q.out = out; Q.out_stride = stride; Q.out_offset = offset; Q.dxt1a = dxt1a; Q.dxt5n = dxt5n;
    Q.omatch5 = g_om5; Q.omatch6 = g_om6;
    int nb = Q.lv.bw * Q.lv.bh;
    emu::launch(dim3((nb + 127) / 128), dim3(128), 0, [&] { k_dxt1_quick(Q); });
}

void emu_set_image(const void *src, float *dst, int count, int format, int to_linear) {
    init_tables();
    SetImageParams P{src, dst, count, format, to_linear ? g_to_linear : nullptr, (size_t)count};
    emu::launch(dim3((count + 255) / 256), dim3(256), 0, [&] { k_set_image(P); });
}

void emu_set_image_x4(const void *src, float *dst, int count, int to_linear) {
    init_tables();
    SetImageParams P{src, dst, count, 0, to_linear ? g_to_linear : nullptr, (size_t)count};
    emu::launch(dim3(3), dim3(256), 0, [&] { k_set_image_bgra8_x4(P); });  // few CTAs: exercises the grid-stride loop
}

void emu_gamma(float *data, size_t pixels, int mode, float power) {
    init_tables();
    GammaParams P{data, 3 * pixels, mode, mode == 0 ? g_to_linear : g_to_gamma, power};
    emu::launch(dim3((unsigned)((P.count + 255) / 256)), dim3(256), 0, [&] { k_gamma(P); });
}

void emu_box_down(const float *src, int sw, int sh, float *dst) {
    BoxDownParams P{src, dst, sw, sh, sw / 2 > 1 ? sw / 2 : 1, sh / 2 > 1 ? sh / 2 : 1, 4};
    size_t total = (size_t)P.dw * P.dh * 4;
    emu::launch(dim3((unsigned)((total + 255) / 256)), dim3(256), 0, [&] { k_box_down(P); });
}

// full 2-D resize: X pass into tmp (dw x sh) then Y pass (FloatImage::resize, FloatImage.cpp:761-808)
void emu_resize(const float *src, int sw, int sh, float *dst, int dw, int dh, int kind, float width, float p0, float p1, int wrap) {
    FilterDesc f{kind, width, p0, p1};
    PolyphaseTable tx, ty;
    build_polyphase(f, sw, dw, tx);
    build_polyphase(f, sh, dh, ty);
    auto ext = [](const PolyphaseTable &t) {
        int m = 0;
        for (int i0 = 0; i0 < t.length; i0 += NVB_PF_TILE) {
            int i1 = std::min(i0 + NVB_PF_TILE - 1, t.length - 1);
            m = std::max(m, t.left[i1] + t.window - t.left[i0]);
        }
        return m;
    };
    if (!getenv("NVB_EMU_UNFUSED") && ext(tx) <= NVB_PF_EXT && ext(ty) <= NVB_PF_EXT && tx.window <= NVB_PF_MAXWIN && ty.window <= NVB_PF_MAXWIN) {
        Polyphase2DParams Q{src, dst, sw, sh, dw, dh, tx.window, ty.window, tx.weights.data(), tx.left.data(), ty.weights.data(), ty.left.data(), wrap};
        emu::launch(dim3((dw + NVB_PF_TILE - 1) / NVB_PF_TILE, (dh + NVB_PF_TILE - 1) / NVB_PF_TILE, 4), dim3(256), 0, [&] {
            if (Q.winx == 13 && Q.winy == 13) k_polyphase_2d_t<13, 13>(Q);
            else if (Q.winx == 9 && Q.winy == 9) k_polyphase_2d_t<9, 9>(Q);
            else if (Q.winx == 5 && Q.winy == 5) k_polyphase_2d_t<5, 5>(Q);
            else k_polyphase_2d_t<0, 0>(Q);
        });
        return;
    }
    std::vector<float> tmp((size_t)dw * sh * 4);
    PolyphaseParams X{src, tmp.data(), sw, sh, dw, sh, 4, tx.window, tx.weights.data(), tx.left.data(), wrap};
    size_t total = (size_t)dw * sh * 4;
    emu::launch(dim3((unsigned)((total + 255) / 256)), dim3(256), 0, [&] { k_polyphase_x(X); });
    PolyphaseParams Y{tmp.data(), dst, dw, sh, dw, dh, 4, ty.window, ty.weights.data(), ty.left.data(), wrap};
    total = (size_t)dw * dh * 4;
    emu::launch(dim3((unsigned)((total + 255) / 256)), dim3(256), 0, [&] { k_polyphase_y(Y); });
}

void emu_normalize(float *data, size_t pixels, int expand_pack) {
    NormalizeParams P{data, pixels, expand_pack};
    emu::launch(dim3((unsigned)((pixels + 255) / 256)), dim3(256), 0, [&] { k_normalize(P); });
}

void emu_grey_scale(float *data, size_t pixels, float r, float g, float b, float a) {
    const float sum = r + g + b + a;
    GreyScaleParams P{data, pixels, {r / sum, g / sum, b / sum, a / sum}};
    emu::launch(dim3((unsigned)((pixels + 255) / 256)), dim3(256), 0, [&] { k_grey_scale(P); });
}

void emu_to_normal_map(const float *src, float *dst, int w, int h, int wrap, const float *fw) {
    NormalMapParams P;
    P.src = src; P.dst = dst; P.w = w; P.h = h; P.wrap = wrap;
    build_blended_sobel(fw, P.kdu);
    size_t pixels = (size_t)w * h;
    emu::launch(dim3((unsigned)((pixels + 255) / 256)), dim3(256), 0, [&] { k_to_normal_map(P); });
}

void emu_scale_bias(float *data, size_t pixels, float scale, float bias) {
    ScaleBiasParams P{data, 3 * pixels, scale, bias};
    emu::launch(dim3((unsigned)((P.count + 255) / 256)), dim3(256), 0, [&] { k_scale_bias(P); });
}

// Format_RGBA writer.  path: 0 = k_pixel_format_x4 (mode = bytes per pixel, 8 = RGBA16F, 16 = RGBA32F), 1 = k_pixel_format, 2 = k_pixel_format_rows.
// The layout (kind, sizes, shifts, bit count, pitch) is derived by the caller exactly as capi.cu: pixel_layout() does.
void emu_pixel_format(const float *planar, int w, int h, unsigned char *out, unsigned pitch, unsigned bitCount, int kind, const unsigned *size,
                      const unsigned *shift, int path, int mode) {
    PixelFormatParams P;
    P.lv.data = planar; P.lv.plane = (size_t)w * h; P.lv.w = w; P.lv.h = h; P.lv.bw = (w + 3) / 4; P.lv.bh = (h + 3) / 4;
    P.lv.to_gamma_table = nullptr;
    P.out = out; P.pitch = pitch; P.bitCount = bitCount; P.kind = kind;
    for (int i = 0; i < 4; i++) { P.size[i] = size[i]; P.shift[i] = shift[i]; }
    if (path == 0) emu::launch(dim3((w / 4 + 255) / 256, h), dim3(256), 0, [&] { k_pixel_format_x4(P, mode); });
    else if (path == 1) emu::launch(dim3((w + 255) / 256, h), dim3(256), 0, [&] { k_pixel_format(P); });
    else emu::launch(dim3((h + 63) / 64), dim3(64), 0, [&] { k_pixel_format_rows(P); });
}
}
