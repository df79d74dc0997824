// C-ABI implementation (include/nvtt_b200.h): context, device surfaces, per-level encode dispatch and the whole
// InputOptions pipeline, all on one CUDA stream per context.  Host logic mirrors
// Compressor::Private::compress (src/nvtt/Context.cpp:217-346, 486-516) and chooseCpuCompressor (:1038-1163).
// There is no CPU code path: without a CUDA device nvttb_context_create fails and nothing else can be called.
#include "../../include/nvtt_b200.h"
#include "host_tables.h"
#include "dds_reader.h"
#include "kernels/bc_alpha.cuh"
#include "kernels/bc3_color.cuh"
#include "kernels/bc1_icbc.cuh"
#include "kernels/bc1_quick.cuh"
#include "kernels/bc6h.cuh"
#include "kernels/bc7.cuh"
#include "kernels/bc7_search.cuh"
#include "kernels/bc6h_search.cuh"
#include "kernels/image_ops.cuh"
#include "kernels/bc_decode.cuh"
#include "kernels/pixel_format.cuh"
#include "kernels/shard_exchange.cuh"
#include "kernels/polyphase_tma.cuh"

#include <cuda_runtime.h>
#include <cuda.h>
#include <map>
#include <string>
#include <thread>
#include <tuple>
#include <vector>
#include <string.h>
#include <stdio.h>
#include <stdlib.h>
#include <ctype.h>
#if defined(__linux__)
#include <pthread.h>
#include <sched.h>
#endif

using namespace nvb;

// nvtt enums used here (src/nvtt/nvtt.h:80-277)
enum { F_RGB = 0, F_DXT1 = 1, F_DXT1a = 2, F_DXT3 = 3, F_DXT5 = 4, F_DXT5n = 5, F_BC4 = 6, F_BC5 = 7, F_BC6 = 10, F_BC7 = 11, F_BC3_RGBM = 12 };
enum { Q_Fastest = 0, Q_Normal = 1, Q_Production = 2, Q_Highest = 3 };
enum { AM_None = 0, AM_Transparency = 1, AM_Premultiplied = 2 };
enum { MF_Box = 0, MF_Triangle = 1, MF_Kaiser = 2 };
enum { RF_Box = 0, RF_Triangle = 1, RF_Kaiser = 2, RF_Mitchell = 3 };

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct PolyDev {
    int window = 0, length = 0;
    int max_extent = 0;  // most source samples any NVB_PF_TILE-wide run of outputs touches (for the fused 2-D kernel)
    float *weights = nullptr;
    int *left = nullptr;
    // exact 2:1 tables: every output has the same weights and left[i] = 2 i + left0 (checked on the host table)
    bool uniform2 = false;
    int left0 = 0;
    float w0[NVB_PT_MAXW] = {};
};

enum { K_ALPHA = 0, K_ALPHA_OPT, K_ALPHA_DXT3, K_BC3_COLOR, K_BC1A_COLOR, K_BC1, K_DXT1_QUICK, K_BC6_ROUGH, K_BC6_TILES, K_BC6_SETUP, K_BC6_ORDER, K_BC6_SEARCH, K_BC6_FINISH, K_BC6_SELECT, K_BC7_ROUGH, K_BC7_TILES, K_BC7_SETUP, K_BC7_ORDER, K_BC7_SEARCH, K_BC7_FINISH, K_BC7_SELECT, K_SET_IMAGE, K_GAMMA, K_BOX_DOWN, K_POLY_X, K_POLY_Y, K_POLY_2D, K_NORMALIZE, K_SCALE_BIAS, K_GREY_SCALE, K_NORMAL_MAP, K_DECODE, K_ERROR_METRIC, K_BINARIZE, K_QUANTIZE, K_PIXEL_FORMAT, K_XCHG, K_COUNT };
static const char *const kKernelNames[K_COUNT] = {"k_alpha_blocks", "k_alpha_optimal", "k_alpha_dxt3", "k_bc3_color", "k_bc1a_color", "k_bc1_icbc", "k_dxt1_quick", "k_bc6_rough", "k_bc6_tiles", "k_bc6_setup", "k_bc6_order", "k_bc6_search", "k_bc6_finish", "k_bc6_select", "k_bc7_rough", "k_bc7_tiles", "k_bc7_setup", "k_bc7_order", "k_bc7_search", "k_bc7_finish", "k_bc7_select", "k_set_image", "k_gamma", "k_box_down",
                                                  "k_polyphase_x", "k_polyphase_y", "k_polyphase_2d", "k_normalize", "k_scale_bias", "k_grey_scale", "k_to_normal_map", "k_decode_blocks", "k_error_metric", "k_binarize", "k_quantize", "k_pixel_format", "k_xchg"};
struct ProfRec {
    int kid;
    cudaEvent_t a, b;
    double units;  // pixels (or texels) the launch processed
};

struct NvttbContext {
    int device = 0;
    bool profiling = false;
    std::vector<ProfRec> prof;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    cudaStream_t stream = nullptr;
    // copy engines run beside the kernels: host->device bands are uploaded on h2d_stream while earlier bands are
    // converted and encoded on `stream`; finished level-0 bands go back on d2h_stream while the mip chain is computed
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    enum { MAX_BANDS = 9 };
    cudaEvent_t ev_up[MAX_BANDS] = {}, ev_enc[MAX_BANDS] = {}, ev_stage_free = nullptr, ev_tail = nullptr;
    // BC7: the eight mode searches of a level are independent until the final select, so each runs on its own stream
    // (small levels cannot fill 148 SMs with one mode's candidates; together they do)
    cudaStream_t mode_stream[8] = {};
    cudaEvent_t ev_fork[2] = {}, ev_join[2][8] = {};  // [parity of the chunk]
    std::string err;
    uint64_t launches = 0;
    float *d_to_gamma = nullptr, *d_to_linear = nullptr;
    unsigned short *d_cand = nullptr;
    int *d_cand_off = nullptr;
    unsigned char *d_om5 = nullptr, *d_om6 = nullptr;
    unsigned char *d_om5a = nullptr, *d_om6a = nullptr;  // OMatchAlpha5/6
    unsigned *d_cand_idx = nullptr;
    unsigned short *d_cand3 = nullptr;
    int *d_cand3_off = nullptr;
    // ICBC tables: [four splits | three splits] u16, [four_total | three_total] int, [mid5 | mid6] float, [match5 | match6] u8
    unsigned short *d_icbc_splits = nullptr;
    int *d_icbc_totals = nullptr;
    float *d_icbc_mid = nullptr;
    unsigned char *d_icbc_match = nullptr;
    int icbc_four_count = 0;
    DevBuf in_stage, tmp_filter, tmp_level, out_dev, lvlA, lvlB, enc_scratch;
    // block-row sharding of one image: band 0 runs the tail levels on tail_stream (own encoder scratch, own level buffers)
    // while its share of the big levels is still being encoded on `stream`
    cudaStream_t tail_stream = nullptr;
    // The mip levels of an image are built and encoded on side_stream (highest priority) while level 0 is still being encoded
    // on `stream`: a small level cannot fill 148 SMs and costs at least the latency of one block (tens of microseconds for the
    // cluster-fit encoders) - one after the other on one stream that is 1 % of a step on one GPU and 5 % when the image is
    // spread over eight.  Side by side with level 0 their blocks simply take free slots.
    cudaStream_t side_stream = nullptr;
    cudaEvent_t ev_side_go = nullptr, ev_side_done = nullptr, ev_cv[MAX_BANDS] = {};
    // host input, level 0 band by band: odd bands are encoded on alt_stream, so that a band starts while the last blocks of the
    // one before it are still running (a cluster-fit block takes ~100 us; eight drains are 1 % of an 8192² image)
    cudaStream_t alt_stream = nullptr;
    cudaEvent_t ev_alt_done = nullptr;
    enum { MAX_LEVELS = 32 };
    cudaEvent_t ev_lvl[MAX_LEVELS] = {}, ev_tail_done = nullptr;
    DevBuf tail_scratch, tail_lvl, exchange;
    unsigned *h_fault = nullptr;  // host-mapped: set by a device-side wait that timed out
    unsigned multi_seq = 0;       // nvttb_process_multi: image counter of the exchange buffer this context owns
    void *h_out = nullptr;  // pinned
    size_t h_out_cap = 0;
    std::map<std::tuple<int, unsigned, unsigned, unsigned, int, int>, PolyDev> poly_cache;
};

struct NvttbSurface {
    NvttbContext *ctx = nullptr;
    DevBuf buf;
    int w = 0, h = 0;
    int wrapMode = 2;  // nvtt default WrapMode_Mirror (Surface.cpp / InputOptions.cpp:98)
    int alphaMode = 0;
    int isNormalMap = 0;
};

static int fail(NvttbContext *c, int code, const char *what, cudaError_t e = cudaSuccess) {
    if (c) {
        c->err = what;
        if (e != cudaSuccess) {
            c->err += ": ";
            c->err += cudaGetErrorString(e);
        }
    }
    return code;
}
#define CK(call)                                                        \
    do {                                                                \
        cudaError_t _e = (call);                                        \
        if (_e != cudaSuccess) return fail(ctx, NVTTB_ERR_CUDA, #call, _e); \
    } while (0)

// Every kernel launch goes through this: counts it and, while profiling, brackets it with CUDA events on the stream.
#define NVB_LAUNCH_ON(ctx, strm_, kid_, units_, KFN, grid, block, ...)                 \
    do {                                                                               \
        ProfRec _r;                                                                    \
        if ((ctx)->profiling) {                                                        \
            _r.kid = (kid_);                                                           \
            _r.units = (double)(units_);                                               \
            cudaEventCreate(&_r.a);                                                    \
            cudaEventCreate(&_r.b);                                                    \
            cudaEventRecord(_r.a, (strm_));                                            \
        }                                                                              \
        KFN<<<(grid), (block), 0, (strm_)>>>(__VA_ARGS__);                             \
        if ((ctx)->profiling) {                                                        \
            cudaEventRecord(_r.b, (strm_));                                            \
            (ctx)->prof.push_back(_r);                                                 \
        }                                                                              \
        (ctx)->launches++;                                                             \
    } while (0)
#define NVB_LAUNCH(ctx, kid_, units_, KFN, grid, block, ...) NVB_LAUNCH_ON(ctx, (ctx)->stream, kid_, units_, KFN, grid, block, __VA_ARGS__)

static int ensure(NvttbContext *ctx, DevBuf &b, size_t bytes) {
    if (b.cap >= bytes && b.p) return NVTTB_OK;
    if (b.p) CK(cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    CK(cudaMalloc(&b.p, bytes));
    b.cap = bytes;
    return NVTTB_OK;
}
static int ensure_pinned(NvttbContext *ctx, size_t bytes) {
    if (ctx->h_out_cap >= bytes && ctx->h_out) return NVTTB_OK;
    if (ctx->h_out) CK(cudaFreeHost(ctx->h_out));
    ctx->h_out = nullptr;
    ctx->h_out_cap = 0;
    CK(cudaHostAlloc(&ctx->h_out, bytes, cudaHostAllocPortable));
    ctx->h_out_cap = bytes;
    return NVTTB_OK;
}

// experiment overrides read from the environment: positive integers only (they are used as divisors)
static int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    const int x = v ? atoi(v) : dflt;
    return x >= 1 ? x : dflt;
}

static inline unsigned grid_for(size_t items, int per_cta) {
    size_t g = (items + per_cta - 1) / per_cta;
    const size_t cap = 148u * 32u;  // grid-stride kernels: a few waves of the 148 SMs
    if (g > cap) g = cap;
    if (g == 0) g = 1;
    return (unsigned)g;
}

static int block_bytes(int format) {
    switch (format) {
    case F_DXT1: case F_DXT1a: case F_BC4: return 8;
    case F_DXT3: case F_DXT5: case F_DXT5n: case F_BC5: case F_BC6: case F_BC7: case F_BC3_RGBM: return 16;
    default: return 0;
    }
}

extern "C" {

int nvttb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int nvttb_context_create(int device, NvttbContext **out) {
    if (!out) return NVTTB_ERR_INVALID_INPUT;
    *out = nullptr;
    // Streams are multiplexed onto CUDA_DEVICE_MAX_CONNECTIONS hardware queues (default 8); the copy / tail / mode streams of a
    // context overlap best when they do not share one.  Only effective before the process creates its CUDA context.
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0 || device < 0 || device >= n) {
        fprintf(stderr, "nvtt_b200: no usable CUDA device (%s); this library has no CPU fallback\n",
                e != cudaSuccess ? cudaGetErrorString(e) : "device index out of range");
        return NVTTB_ERR_CUDA;
    }
    NvttbContext *ctx = new NvttbContext();
    ctx->device = device;
    auto bail = [&](const char *what, cudaError_t err) {
        fprintf(stderr, "nvtt_b200: %s: %s\n", what, cudaGetErrorString(err));
        delete ctx;
        return NVTTB_ERR_CUDA;
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    if ((e = cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    if ((e = cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    for (int i = 0; i < NvttbContext::MAX_BANDS; i++) {
        if ((e = cudaEventCreateWithFlags(&ctx->ev_up[i], cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
        if ((e = cudaEventCreateWithFlags(&ctx->ev_enc[i], cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    }
    for (int i = 0; i < 8; i++) {
        for (int p = 0; p < 2; p++)
            if ((e = cudaEventCreateWithFlags(&ctx->ev_join[p][i], cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    }
    for (int p = 0; p < 2; p++)
        if ((e = cudaEventCreateWithFlags(&ctx->ev_fork[p], cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&ctx->ev_stage_free, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&ctx->ev_tail, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&ctx->ev_tail_done, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    {
        // highest priority: their few blocks are scheduled ahead of the pending blocks of a big level on `stream`
        int prio_lo = 0, prio_hi = 0;
        if ((e = cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi)) != cudaSuccess) return bail("cudaDeviceGetStreamPriorityRange", e);
        if ((e = cudaStreamCreateWithPriority(&ctx->tail_stream, cudaStreamNonBlocking, prio_hi)) != cudaSuccess) return bail("cudaStreamCreate", e);
        if ((e = cudaStreamCreateWithPriority(&ctx->side_stream, cudaStreamNonBlocking, prio_hi)) != cudaSuccess) return bail("cudaStreamCreate", e);
    }
    if ((e = cudaStreamCreateWithFlags(&ctx->alt_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    if ((e = cudaEventCreateWithFlags(&ctx->ev_alt_done, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&ctx->ev_side_go, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&ctx->ev_side_done, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    for (int i = 0; i < NvttbContext::MAX_BANDS; i++)
        if ((e = cudaEventCreateWithFlags(&ctx->ev_cv[i], cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    for (int i = 0; i < NvttbContext::MAX_LEVELS; i++)
        if ((e = cudaEventCreateWithFlags(&ctx->ev_lvl[i], cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaHostAlloc((void **)&ctx->h_fault, sizeof(unsigned), cudaHostAllocMapped | cudaHostAllocPortable)) != cudaSuccess) return bail("cudaHostAlloc", e);
    *ctx->h_fault = 0;
    float tg[512], tl[512];
    build_gamma_tables(tg, tl);
    std::vector<uint16_t> cand;
    int off[18];
    build_squish_splits(cand, off);
    uint8_t om5[512], om6[512];
    build_omatch(om5, 32);
    build_omatch(om6, 64);
    {
        std::vector<uint32_t> cidx;
        build_squish_split_indices(cidx);
        if ((e = cudaMalloc(&ctx->d_cand_idx, cidx.size() * 4)) != cudaSuccess) return bail("cudaMalloc", e);
        cudaMemcpy(ctx->d_cand_idx, cidx.data(), cidx.size() * 4, cudaMemcpyHostToDevice);
        std::vector<uint16_t> cand3;
        int off3[18];
        build_squish_splits3(cand3, off3);
        uint8_t om5a[512], om6a[512];
        build_omatch(om5a, 32, true);
        build_omatch(om6a, 64, true);
        if ((e = cudaMalloc(&ctx->d_cand3, cand3.size() * 2)) != cudaSuccess) return bail("cudaMalloc", e);
        if ((e = cudaMalloc(&ctx->d_cand3_off, sizeof(off3))) != cudaSuccess) return bail("cudaMalloc", e);
        if ((e = cudaMalloc(&ctx->d_om5a, 512)) != cudaSuccess) return bail("cudaMalloc", e);
        if ((e = cudaMalloc(&ctx->d_om6a, 512)) != cudaSuccess) return bail("cudaMalloc", e);
        cudaMemcpy(ctx->d_cand3, cand3.data(), cand3.size() * 2, cudaMemcpyHostToDevice);
        cudaMemcpy(ctx->d_cand3_off, off3, sizeof(off3), cudaMemcpyHostToDevice);
        cudaMemcpy(ctx->d_om5a, om5a, 512, cudaMemcpyHostToDevice);
        cudaMemcpy(ctx->d_om6a, om6a, 512, cudaMemcpyHostToDevice);
    }
    if ((e = cudaMalloc(&ctx->d_to_gamma, sizeof(tg))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc(&ctx->d_to_linear, sizeof(tl))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc(&ctx->d_cand, cand.size() * 2)) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc(&ctx->d_cand_off, sizeof(off))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc(&ctx->d_om5, 512)) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc(&ctx->d_om6, 512)) != cudaSuccess) return bail("cudaMalloc", e);
    {
        std::vector<uint16_t> four, three;
        int totals[32];
        float mid[96];
        uint8_t match[1024];
        build_icbc_splits(four, totals, three, totals + 16);
        build_icbc_midpoints(mid, mid + 32);
        build_icbc_match(match, 32);
        build_icbc_match(match + 512, 64);
        ctx->icbc_four_count = (int)four.size();
        std::vector<uint16_t> all(four);
        all.insert(all.end(), three.begin(), three.end());
        if ((e = cudaMalloc(&ctx->d_icbc_splits, all.size() * 2)) != cudaSuccess) return bail("cudaMalloc", e);
        if ((e = cudaMalloc(&ctx->d_icbc_totals, sizeof(totals))) != cudaSuccess) return bail("cudaMalloc", e);
        if ((e = cudaMalloc(&ctx->d_icbc_mid, sizeof(mid))) != cudaSuccess) return bail("cudaMalloc", e);
        if ((e = cudaMalloc(&ctx->d_icbc_match, sizeof(match))) != cudaSuccess) return bail("cudaMalloc", e);
        cudaMemcpy(ctx->d_icbc_splits, all.data(), all.size() * 2, cudaMemcpyHostToDevice);
        cudaMemcpy(ctx->d_icbc_totals, totals, sizeof(totals), cudaMemcpyHostToDevice);
        cudaMemcpy(ctx->d_icbc_mid, mid, sizeof(mid), cudaMemcpyHostToDevice);
        cudaMemcpy(ctx->d_icbc_match, match, sizeof(match), cudaMemcpyHostToDevice);
    }
    cudaMemcpy(ctx->d_to_gamma, tg, sizeof(tg), cudaMemcpyHostToDevice);
    cudaMemcpy(ctx->d_to_linear, tl, sizeof(tl), cudaMemcpyHostToDevice);
    cudaMemcpy(ctx->d_cand, cand.data(), cand.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(ctx->d_cand_off, off, sizeof(off), cudaMemcpyHostToDevice);
    cudaMemcpy(ctx->d_om5, om5, 512, cudaMemcpyHostToDevice);
    if ((e = cudaMemcpy(ctx->d_om6, om6, 512, cudaMemcpyHostToDevice)) != cudaSuccess) return bail("cudaMemcpy", e);
    *out = ctx;
    return NVTTB_OK;
}

void nvttb_context_destroy(NvttbContext *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->h2d_stream);
    cudaStreamSynchronize(ctx->d2h_stream);
    cudaFree(ctx->d_to_gamma);
    cudaFree(ctx->d_to_linear);
    cudaFree(ctx->d_cand);
    cudaFree(ctx->d_cand_off);
    cudaFree(ctx->d_om5);
    cudaFree(ctx->d_om6);
    cudaFree(ctx->d_om5a);
    cudaFree(ctx->d_om6a);
    cudaFree(ctx->d_cand3);
    cudaFree(ctx->d_cand_idx);
    cudaFree(ctx->d_cand3_off);
    cudaFree(ctx->d_icbc_splits);
    cudaFree(ctx->d_icbc_totals);
    cudaFree(ctx->d_icbc_mid);
    cudaFree(ctx->d_icbc_match);
    cudaFree(ctx->in_stage.p);
    cudaFree(ctx->tmp_filter.p);
    cudaFree(ctx->tmp_level.p);
    cudaFree(ctx->out_dev.p);
    cudaFree(ctx->lvlA.p);
    cudaFree(ctx->lvlB.p);
    cudaFree(ctx->enc_scratch.p);
    cudaStreamSynchronize(ctx->tail_stream);
    cudaFree(ctx->tail_scratch.p);
    cudaFree(ctx->tail_lvl.p);
    cudaFree(ctx->exchange.p);
    if (ctx->h_fault) cudaFreeHost(ctx->h_fault);
    for (int i = 0; i < NvttbContext::MAX_LEVELS; i++)
        if (ctx->ev_lvl[i]) cudaEventDestroy(ctx->ev_lvl[i]);
    if (ctx->ev_tail_done) cudaEventDestroy(ctx->ev_tail_done);
    cudaStreamDestroy(ctx->tail_stream);
    if (ctx->side_stream) {
        cudaStreamSynchronize(ctx->side_stream);
        cudaStreamDestroy(ctx->side_stream);
    }
    if (ctx->alt_stream) {
        cudaStreamSynchronize(ctx->alt_stream);
        cudaStreamDestroy(ctx->alt_stream);
    }
    if (ctx->ev_alt_done) cudaEventDestroy(ctx->ev_alt_done);
    if (ctx->ev_side_go) cudaEventDestroy(ctx->ev_side_go);
    if (ctx->ev_side_done) cudaEventDestroy(ctx->ev_side_done);
    for (int i = 0; i < NvttbContext::MAX_BANDS; i++)
        if (ctx->ev_cv[i]) cudaEventDestroy(ctx->ev_cv[i]);
    if (ctx->h_out) cudaFreeHost(ctx->h_out);
    for (auto &kv : ctx->poly_cache) {
        cudaFree(kv.second.weights);
        cudaFree(kv.second.left);
    }
    for (int i = 0; i < NvttbContext::MAX_BANDS; i++) {
        if (ctx->ev_up[i]) cudaEventDestroy(ctx->ev_up[i]);
        if (ctx->ev_enc[i]) cudaEventDestroy(ctx->ev_enc[i]);
    }
    for (int i = 0; i < 8; i++) {
        if (ctx->mode_stream[i]) { cudaStreamSynchronize(ctx->mode_stream[i]); cudaStreamDestroy(ctx->mode_stream[i]); }
        for (int p = 0; p < 2; p++)
            if (ctx->ev_join[p][i]) cudaEventDestroy(ctx->ev_join[p][i]);
    }
    for (int p = 0; p < 2; p++)
        if (ctx->ev_fork[p]) cudaEventDestroy(ctx->ev_fork[p]);
    if (ctx->ev_stage_free) cudaEventDestroy(ctx->ev_stage_free);
    if (ctx->ev_tail) cudaEventDestroy(ctx->ev_tail);
    cudaStreamDestroy(ctx->h2d_stream);
    cudaStreamDestroy(ctx->d2h_stream);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *nvttb_last_error(const NvttbContext *ctx) { return ctx ? ctx->err.c_str() : "null context"; }
uint64_t nvttb_launch_count(const NvttbContext *ctx) { return ctx ? ctx->launches : 0; }
const char *nvttb_build_variant(void) {
#ifdef NVB_FASTMATH
    return "fastmath";
#else
    return "strict";
#endif
}
void *nvttb_stream(NvttbContext *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
int nvttb_synchronize(NvttbContext *ctx) {
    if (!ctx) return NVTTB_ERR_INVALID_INPUT;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->h_fault && *ctx->h_fault) {
        *ctx->h_fault = 0;
        return fail(ctx, NVTTB_ERR_CUDA, "sharded image: a band never delivered its rows (device-side wait timed out)");
    }
    return NVTTB_OK;
}

int nvttb_timer_start(NvttbContext *ctx) {
    if (!ctx) return NVTTB_ERR_INVALID_INPUT;
    if (!ctx->t0) {
        CK(cudaEventCreate(&ctx->t0));
        CK(cudaEventCreate(&ctx->t1));
    }
    CK(cudaEventRecord(ctx->t0, ctx->stream));
    return NVTTB_OK;
}
int nvttb_timer_stop(NvttbContext *ctx, float *ms) {
    if (!ctx || !ms || !ctx->t0) return NVTTB_ERR_INVALID_INPUT;
    CK(cudaEventRecord(ctx->t1, ctx->stream));
    CK(cudaEventSynchronize(ctx->t1));
    CK(cudaEventElapsedTime(ms, ctx->t0, ctx->t1));
    return NVTTB_OK;
}
int nvttb_profile_begin(NvttbContext *ctx) {
    if (!ctx) return NVTTB_ERR_INVALID_INPUT;
    CK(cudaStreamSynchronize(ctx->stream));
    for (auto &r : ctx->prof) {
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    ctx->prof.clear();
    ctx->profiling = true;
    return NVTTB_OK;
}
int nvttb_profile_end(NvttbContext *ctx, NvttbKernelStat *stats, int max_stats, int *count) {
    if (!ctx || !stats || !count) return NVTTB_ERR_INVALID_INPUT;
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->profiling = false;
    NvttbKernelStat acc[K_COUNT];
    for (int k = 0; k < K_COUNT; k++) {
        acc[k].name = kKernelNames[k];
        acc[k].launches = 0;
        acc[k].total_ms = 0.0;
        acc[k].max_ms = 0.0;
        acc[k].max_units = 0.0;
        acc[k].total_units = 0.0;
    }
    for (auto &r : ctx->prof) {
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        NvttbKernelStat &a = acc[r.kid];
        a.launches++;
        a.total_ms += ms;
        a.total_units += r.units;
        if (ms > a.max_ms) {
            a.max_ms = ms;
            a.max_units = r.units;
        }
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    ctx->prof.clear();
    int n = 0;
    for (int k = 0; k < K_COUNT && n < max_stats; k++)
        if (acc[k].launches) stats[n++] = acc[k];
    *count = n;
    return NVTTB_OK;
}

size_t nvttb_level_size(int format, int w, int h) {
    if (w <= 0 || h <= 0) return 0;
    return (size_t)((w + 3) / 4) * ((h + 3) / 4) * block_bytes(format);
}

int nvttb_format_supported(int format, int quality) {
    switch (format) {
    case F_BC4:
    case F_BC5:
        return quality >= Q_Fastest && quality <= Q_Highest;
    case F_DXT1:
        return quality >= Q_Fastest && quality <= Q_Highest;
    case F_DXT5:
    case F_DXT3:
        return quality >= Q_Fastest && quality <= Q_Highest;
    case F_DXT5n:
        return quality >= Q_Fastest && quality <= Q_Highest;
    case F_DXT1a:
        return quality >= Q_Fastest && quality <= Q_Highest;
    case F_BC6:
    case F_BC7:
    case F_BC3_RGBM:
        return 1;  // quality is ignored for BC6 / BC7 (CompressorDX11.cpp:42-102) and BC3-RGBM (BlockCompressor.cpp:235-238)
    default:
        return 0;
    }
}

}  // extern "C"

// The side streams of the BC6H / BC7 encoders are created on first use: streams share a small number of hardware queues
// (CUDA_DEVICE_MAX_CONNECTIONS), and work queued behind a waiting kernel of another stream in the same queue cannot start.
static int ensure_mode_streams(NvttbContext *ctx) {
    for (int i = 0; i < 8; i++)
        if (!ctx->mode_stream[i]) CK(cudaStreamCreateWithFlags(&ctx->mode_stream[i], cudaStreamNonBlocking));
    return NVTTB_OK;
}

// BC7, one mode of one chunk of blocks, on the mode's own stream: rough shape ranking (modes 0,1,2,3,7), start endpoints
// per candidate, the searcher state machine, finish + reduction over the block's candidates.
#ifndef NVB_BC7_CHUNK
#define NVB_BC7_CHUNK 65536  // blocks per chunk (~12 KB of searcher records each): 8k 1210, 16k 1177, 32k 1146, 64k 1133 ms for a 2048² chain
#endif
template <int M, int NCAND> struct Bc7ModeBytes {
    using X = Bc7X<M>;
    static constexpr size_t setup = (size_t)NCAND * X::NR * 16, idx = (size_t)NCAND * 16, res = (size_t)NCAND * X::NR * X::NLSB * 16;
    static constexpr size_t perm = X::NR > 1 ? (size_t)NCAND * X::NR * 4 : 0;
    static constexpr size_t per_block = setup + idx + res + perm;
};
static constexpr size_t kBc7CounterBytes = 8 * 256;  // one 256-byte counter record per mode
static constexpr size_t kBc7ChunkBytesPerBlock = 256 + Bc7ModeBytes<0, 4>::per_block + Bc7ModeBytes<1, 16>::per_block + Bc7ModeBytes<2, 16>::per_block +
                                                 Bc7ModeBytes<3, 16>::per_block + Bc7ModeBytes<4, 8>::per_block + Bc7ModeBytes<5, 4>::per_block +
                                                 Bc7ModeBytes<6, 1>::per_block + Bc7ModeBytes<7, 16>::per_block;

template <int M, int NCAND> static int launch_bc7_mode(NvttbContext *ctx, Bc7SearchParams S, unsigned char *&arena, unsigned *counters, int parity, double units) {
    using X = Bc7X<M>;
    using B = Bc7ModeBytes<M, NCAND>;
    const int n = S.nblk;
    S.setup = (uint4 *)arena;
    S.setup_idx = (uint4 *)(arena + B::setup * NVB_BC7_CHUNK);
    S.res = (uint4 *)(arena + (B::setup + B::idx) * NVB_BC7_CHUNK);
    S.perm = B::perm ? (unsigned *)(arena + (B::setup + B::idx + B::res) * NVB_BC7_CHUNK) : nullptr;
    S.counters = counters + M * 64;
    arena += B::per_block * NVB_BC7_CHUNK;
    cudaStream_t st = ctx->mode_stream[M];
    CK(cudaStreamWaitEvent(st, ctx->ev_fork[parity], 0));
    CK(cudaMemsetAsync(S.counters, 0, NVB_BX_COUNTERS * sizeof(unsigned), st));
    if constexpr (M == 0 || M == 1 || M == 2 || M == 3 || M == 7)
        NVB_LAUNCH_ON(ctx, st, K_BC7_ROUGH, units, k_bc7_rough<M>, grid_for(n, NVB_BC7_ROUGH_WARPS), NVB_BC7_ROUGH_WARPS * 32, S.P, S.blk0, S.blk0 + n);
    const unsigned cgrid = (unsigned)(((size_t)n * NCAND + 127) / 128);
    NVB_LAUNCH_ON(ctx, st, K_BC7_SETUP, units, (k_bc7_setup<M, NCAND>), cgrid, 128, S);
    if constexpr (X::NR > 1) {
        const unsigned ogrid = (unsigned)(((size_t)n * NCAND * X::NR + 255) / 256);
        NVB_LAUNCH_ON(ctx, st, K_BC7_ORDER, units, (k_bc7_order<M, NCAND, 0>), ogrid, 256, S);
        NVB_LAUNCH_ON(ctx, st, K_BC7_ORDER, units, (k_bc7_order<M, NCAND, 1>), ogrid, 256, S);
    }
    // searchers are handed out dynamically; ~2 per thread on small levels (tuned on B200), at most what the GPU can hold
    const size_t searchers = (size_t)n * (X::SPLIT ? 4 : NCAND * X::NR * X::NLSB);
    static const int spt = env_int("NVB_BC7_SPT", 2);
    size_t sgrid = (searchers / spt + 127) / 128;
    if (sgrid > 148u * 6u) sgrid = 148u * 6u;
    if (sgrid < 1) sgrid = 1;
    // two trials per loop trip where the palettes fit in registers (not mode 6: 16 entries x 4 channels; not the split modes)
    if constexpr (M == 0 || M == 1 || M == 2 || M == 3 || M == 7) NVB_LAUNCH_ON(ctx, st, K_BC7_SEARCH, units, (k_bc7_search2<M>), (unsigned)sgrid, 128, S);
    else NVB_LAUNCH_ON(ctx, st, K_BC7_SEARCH, units, (k_bc7_search<M, 0>), (unsigned)sgrid, 128, S);
    if constexpr (M == 4) {
        CK(cudaMemsetAsync(S.counters, 0, sizeof(unsigned), st));
        NVB_LAUNCH_ON(ctx, st, K_BC7_SEARCH, units, (k_bc7_search<M, 1>), (unsigned)sgrid, 128, S);
    }
    NVB_LAUNCH_ON(ctx, st, K_BC7_FINISH, units, (k_bc7_finish<M, NCAND>), cgrid, 128, S);
    CK(cudaEventRecord(ctx->ev_join[parity][M], st));
    return NVTTB_OK;
}

// encoder scratch of one level (BC6H / BC7 only): what encode_device sizes ctx->enc_scratch to
static constexpr size_t kBc6ScratchPerBlock = 80 + 32 + 8 + 256 + 32 + 96 + 16 + 96 + 8;
static size_t encoder_scratch_bytes(int format, int w, int h) {
    const size_t nb = (size_t)((w + 3) / 4) * ((h + 3) / 4);
    if (format == F_BC6) return nb * kBc6ScratchPerBlock + 256 + 64;
    if (format == F_BC7) return nb * (80 + 128 + 32) + 16 + kBc7CounterBytes + (size_t)NVB_BC7_CHUNK * (kBc7ChunkBytesPerBlock + 256);
    return 0;
}

// ---- level encode on device buffers (async on ctx->stream) --------------------------------------------------
// d_rgba points at row 0 of the rows to encode (h of them); plane = floats between the planes of the level they belong to
struct CycView { int rpc, n, i; };  // LevelView::cyc_*: the rows are one GPU's concatenated chunks of a sharded level
static int encode_device(NvttbContext *ctx, const NvttbEncodeDesc *d, const float *d_rgba, int w, int h, unsigned char *d_out, size_t plane = 0, const CycView *cyc = nullptr) {
    if (!nvttb_format_supported(d->format, d->quality)) return fail(ctx, NVTTB_ERR_UNSUPPORTED_FEATURE, "format/quality not implemented");
    LevelView lv;
    lv.data = d_rgba;
    lv.plane = plane ? plane : (size_t)w * h;
    lv.w = w;
    lv.h = h;
    lv.bw = (w + 3) / 4;
    lv.bh = (h + 3) / 4;
    lv.to_gamma_table = d->applyToGamma ? ctx->d_to_gamma : nullptr;
    if (cyc) {
        lv.cyc_rpc = cyc->rpc;
        lv.cyc_n = cyc->n;
        lv.cyc_i = cyc->i;
    }
    const int nb = lv.bw * lv.bh;
    auto alpha = [&](int channel, int stride, int offset, bool optimal) {
        AlphaBlocksParams P;
        P.lv = lv;
        P.channel = channel;
        P.out = d_out;
        P.out_stride = stride;
        P.out_offset = offset;
        P.mode = optimal ? 1 : 0;
        if (optimal) {
            // OptimalCompress::compressDXT5A: one warp per channel-block, 4 warps per CTA
            NVB_LAUNCH(ctx, K_ALPHA_OPT, (double)w * h, k_alpha_optimal, grid_for(nb, 4), 128, P);
        } else {
            NVB_LAUNCH(ctx, K_ALPHA, (double)w * h, k_alpha_blocks, grid_for(nb, 128), 128, P);
        }
    };
    if (d->format == F_DXT1) {
        // CompressorDXT1 -> ICBC: Fastest -> Level 1, Production -> Level 9, Normal and Highest -> Level 8 (BlockCompressor.cpp:211-217)
        Bc1Params P;
        P.lv = lv;
        P.out = d_out;
        P.out_stride = 8;
        P.out_offset = 0;
        P.level = (d->quality == Q_Fastest) ? 1 : (d->quality == Q_Production) ? 9 : 8;
        P.transparency = (d->alphaMode == AM_Transparency);
        P.cw[0] = d->colorWeights[0];
        P.cw[1] = d->colorWeights[1];
        P.cw[2] = d->colorWeights[2];
        P.four = ctx->d_icbc_splits;
        P.three = ctx->d_icbc_splits + ctx->icbc_four_count;
        P.four_total = ctx->d_icbc_totals;
        P.three_total = ctx->d_icbc_totals + 16;
        P.midpoints5 = ctx->d_icbc_mid;
        P.midpoints6 = ctx->d_icbc_mid + 32;
        P.match5 = ctx->d_icbc_match;
        P.match6 = ctx->d_icbc_match + 512;
#define NVB_BC1_GO(KFN) NVB_LAUNCH(ctx, K_BC1, (double)w * h, KFN, (nb + NVB_BC1_GROUPS - 1) / NVB_BC1_GROUPS, NVB_BC1_GROUPS * 16, P)
        NVB_BC1_DISPATCH(P, NVB_BC1_GO);
#undef NVB_BC1_GO
    } else if (d->format == F_BC3_RGBM) {
        // CompressorBC3_RGBM -> compress_dxt5_rgbm (BlockCompressor.cpp:235-238, CompressorDXT5_RGBM.cpp:54-118):
        // colour block = ICBC Quality_Default (Level 8) on (R,G,B)/M with weights w*M, no 3-colour mode, colour weights 1;
        // alpha block = weighted brute-force fit of the multiplier that compensates the colour block's error
        Bc1Params P;
        P.lv = lv;
        P.out = d_out;
        P.out_stride = 16;
        P.out_offset = 8;
        P.level = 8;
        P.transparency = (d->alphaMode == AM_Transparency);
        P.cw[0] = P.cw[1] = P.cw[2] = 1.0f;
        P.four = ctx->d_icbc_splits;
        P.three = ctx->d_icbc_splits + ctx->icbc_four_count;
        P.four_total = ctx->d_icbc_totals;
        P.three_total = ctx->d_icbc_totals + 16;
        P.midpoints5 = ctx->d_icbc_mid;
        P.midpoints6 = ctx->d_icbc_mid + 32;
        P.match5 = ctx->d_icbc_match;
        P.match6 = ctx->d_icbc_match + 512;
        P.rgbm = 1;
        P.rgbm_min = d->rgbmThreshold;
#define NVB_BC1_GO(KFN) NVB_LAUNCH(ctx, K_BC1, (double)w * h, KFN, (nb + NVB_BC1_GROUPS - 1) / NVB_BC1_GROUPS, NVB_BC1_GROUPS * 16, P)
        NVB_BC1_DISPATCH(P, NVB_BC1_GO);
#undef NVB_BC1_GO
        RgbmAlphaParams A;
        A.lv = lv;
        A.out = d_out;
        A.min_m = d->rgbmThreshold;
        A.transparency = P.transparency;
        NVB_LAUNCH(ctx, K_ALPHA_OPT, (double)w * h, k_rgbm_alpha, grid_for(nb, 4), 128, A);
    } else if (d->format == F_BC4) {
        // Fastest/Normal -> QuickCompress, Production/Highest -> OptimalCompress (Context.cpp:1098-1115)
        alpha(0, 8, 0, d->quality >= Q_Production);
    } else if (d->format == F_BC5) {
        alpha(0, 16, 0, d->quality >= Q_Production);
        alpha(1, 16, 8, d->quality >= Q_Production);
    } else if (d->format == F_DXT1a && d->quality != Q_Fastest) {
        // CompressorDXT1a (CompressorDX9.cpp:83-111): squish weighted cluster fit in DXT1 mode (3-colour, then 4-colour when opaque)
        Bc3ColorParams P;
        P.lv = lv;
        P.out = d_out;
        P.out_stride = 8;
        P.out_offset = 0;
        P.dxt5n = 0;
        P.metric[0] = d->colorWeights[0];
        P.metric[1] = d->colorWeights[1];
        P.metric[2] = d->colorWeights[2];
        P.weight_by_alpha = (d->alphaMode == AM_Transparency);
        P.cand = ctx->d_cand;
        P.cand_idx = ctx->d_cand_idx;
        P.cand_off = ctx->d_cand_off;
        P.omatch5 = ctx->d_om5;
        P.omatch6 = ctx->d_om6;
        P.cand3 = ctx->d_cand3;
        P.cand3_off = ctx->d_cand3_off;
        P.omatch5a = ctx->d_om5a;
        P.omatch6a = ctx->d_om6a;
        NVB_LAUNCH(ctx, K_BC1A_COLOR, (double)w * h, k_bc1a_color, (nb + NVB_BC3_GROUPS - 1) / NVB_BC3_GROUPS, NVB_BC3_GROUPS * 16, P);
    } else if (d->format == F_DXT1a || (d->quality == Q_Fastest && (d->format == F_DXT5 || d->format == F_DXT3 || d->format == F_DXT5n))) {
        // FastCompressorDXT1a / DXT3 / DXT5 / DXT5n (CompressorDX9.cpp:55-81): QuickCompress colour block (+ alpha block)
        const bool wide = d->format != F_DXT1a;
        if (d->format == F_DXT5) alpha(3, 16, 0, false);
        if (d->format == F_DXT5n) alpha(0, 16, 0, false);  // swizzle(4,1,5,0): alpha block = red channel
        if (d->format == F_DXT3) {
            AlphaBlocksParams A;
            A.lv = lv; A.channel = 3; A.out = d_out; A.out_stride = 16; A.out_offset = 0; A.mode = 0;
            NVB_LAUNCH(ctx, K_ALPHA_DXT3, (double)w * h, k_alpha_dxt3, grid_for(nb, 128), 128, A);
        }
        Dxt1QuickParams Q;
        Q.lv = lv;
        Q.out = d_out;
        Q.out_stride = wide ? 16 : 8;
        Q.out_offset = wide ? 8 : 0;
        Q.dxt1a = (d->format == F_DXT1a);
        Q.dxt5n = (d->format == F_DXT5n);
        Q.omatch5 = ctx->d_om5;
        Q.omatch6 = ctx->d_om6;
        NVB_LAUNCH(ctx, K_DXT1_QUICK, (double)w * h, k_dxt1_quick, grid_for(nb, 128), 128, Q);
    } else if (d->format == F_DXT5 || d->format == F_DXT3 || d->format == F_DXT5n) {
        if (d->format == F_DXT5) {
            alpha(3, 16, 0, d->quality == Q_Highest);  // CompressorDX9.cpp:149-157
        } else if (d->format == F_DXT5n) {
            // "rgba.swizzle(4,1,5,0)": alpha block = red channel; QuickCompress, or OptimalCompress at Highest (CompressorDX9.cpp:212-221)
            alpha(0, 16, 0, d->quality == Q_Highest);
        } else {
            AlphaBlocksParams A;  // CompressorDXT3: OptimalCompress::compressDXT3A on the alpha channel (CompressorDX9.cpp:119-124)
            A.lv = lv;
            A.channel = 3;
            A.out = d_out;
            A.out_stride = 16;
            A.out_offset = 0;
            A.mode = 0;
            NVB_LAUNCH(ctx, K_ALPHA_DXT3, (double)w * h, k_alpha_dxt3, grid_for(nb, 128), 128, A);
        }
        if (d->format == F_DXT5n && d->quality == Q_Highest) {
            // CompressorDXT5n at Quality_Highest: brute-force green block (CompressorDX9.cpp:184-187)
            AlphaBlocksParams G;
            G.lv = lv; G.channel = 1; G.out = d_out; G.out_stride = 16; G.out_offset = 8; G.mode = 1; G.omatch6 = ctx->d_om6;
            NVB_LAUNCH(ctx, K_ALPHA_OPT, (double)w * h, k_dxt1g_optimal, grid_for(nb, 4), 128, G);
            CK(cudaGetLastError());
            return NVTTB_OK;
        }
        Bc3ColorParams P;
        P.lv = lv;
        P.out = d_out;
        P.out_stride = 16;
        P.out_offset = 8;
        P.dxt5n = (d->format == F_DXT5n);
        P.metric[0] = P.dxt5n ? 0.0f : d->colorWeights[0];
        P.metric[1] = P.dxt5n ? 1.0f : d->colorWeights[1];
        P.metric[2] = P.dxt5n ? 0.0f : d->colorWeights[2];
        P.weight_by_alpha = (d->alphaMode == AM_Transparency);
        P.cand = ctx->d_cand;
        P.cand_idx = ctx->d_cand_idx;
        P.cand_off = ctx->d_cand_off;
        P.omatch5 = ctx->d_om5;
        P.omatch6 = ctx->d_om6;
        P.cand3 = nullptr;
        P.cand3_off = nullptr;
        P.omatch5a = nullptr;
        P.omatch6a = nullptr;
        NVB_LAUNCH(ctx, K_BC3_COLOR, (double)w * h, k_bc3_color, (nb + NVB_BC3_GROUPS - 1) / NVB_BC3_GROUPS, NVB_BC3_GROUPS * 16, P);
    }
    else if (d->format == F_BC6) {
        // scratch per block: 20 floats of rough endpoints, 2 candidate blocks, 2 errors; then the searcher records:
        // texel tile (256), meta (2 x 16), start endpoints (3 x 32), start indices (2 x 8), results (3 x 32), order (2 x 4)
        int rc = ensure(ctx, ctx->enc_scratch, encoder_scratch_bytes(F_BC6, w, h));
        if (rc != NVTTB_OK) return rc;
        if ((rc = ensure_mode_streams(ctx)) != NVTTB_OK) return rc;
        Bc6Params P;
        P.lv = lv;
        P.out = d_out;
        // ZOH::Utils::FORMAT: unsigned for PixelType_UnsignedFloat / UnsignedNorm / UnsignedInt (CompressorDX11.cpp:47-56)
        P.is_signed = !(d->pixelType == 5 || d->pixelType == 0 || d->pixelType == 2);
        P.transparency = (d->alphaMode == AM_Transparency);
        unsigned char *base = (unsigned char *)ctx->enc_scratch.p;
        Bc6SearchParams S;
        float4 *tiles = (float4 *)base;               base += (size_t)nb * 256;
        S.tiles = tiles;
        S.meta = (int4 *)base;                        base += (size_t)nb * 32;
        S.setup = (int4 *)base;                       base += (size_t)nb * 96;
        S.res = (int4 *)base;                         base += (size_t)nb * 96;
        S.setup_idx = (uint2 *)base;                  base += (size_t)nb * 16;
        P.cand = base;                                base += (size_t)nb * 32;
        P.rough = (float *)base;                      base += (size_t)nb * 80;
        P.cand_err = (float *)base;                   base += (size_t)nb * 8;
        S.perm = (unsigned *)base;                    base += (size_t)nb * 8;
        base += (16 - ((size_t)base & 15)) & 15;
        S.counters = (unsigned *)base;
        S.P = P;
        const double units = (double)w * h;
        CK(cudaMemsetAsync(S.counters, 0, NVB_BC6_COUNTERS * sizeof(unsigned), ctx->stream));
        NVB_LAUNCH(ctx, K_BC6_TILES, units, k_bc6_tiles, (unsigned)(((size_t)nb * 16 + 255) / 256), 256, S, tiles);
        const int padded = (nb + 127) / 128 * 128;
        // searchers are handed out dynamically, at most what the GPU holds (searchers per thread tuned on B200:
        // profiles/r1e_summary.md; NVB_BC6_SPT1/2 override for experiments).
        static const int spt1 = env_int("NVB_BC6_SPT1", 1), spt2 = env_int("NVB_BC6_SPT2", 2);
        size_t g1 = ((size_t)nb / spt1 + 127) / 128, g2 = ((size_t)nb * 2 / spt2 + 127) / 128;
        if (g1 > 148u * 5u) g1 = 148u * 5u;
        if (g2 > 148u * 6u) g2 = 148u * 6u;
        if (g1 < 1) g1 = 1;
        if (g2 < 1) g2 = 1;
        // Two independent chains after the tiles: the one-region encoding (line fit, setup, the long 16 x 16 searches) on a
        // second stream, the two-region one (32-shape ranking, setup, order, searches) on the context's stream; finish joins.
        CK(cudaEventRecord(ctx->ev_fork[0], ctx->stream));
        cudaStream_t s1 = ctx->mode_stream[0];
        CK(cudaStreamWaitEvent(s1, ctx->ev_fork[0], 0));
        NVB_LAUNCH_ON(ctx, s1, K_BC6_ROUGH, units, k_bc6_rough_one, (unsigned)((nb + 127) / 128), 128, S);
        NVB_LAUNCH_ON(ctx, s1, K_BC6_SETUP, units, k_bc6_setup, (unsigned)((nb + 127) / 128), 128, S, padded, 0);
        NVB_LAUNCH_ON(ctx, s1, K_BC6_SEARCH, units, k_bc6_search<1>, (unsigned)g1, 128, S);
        CK(cudaEventRecord(ctx->ev_join[0][0], s1));
        NVB_LAUNCH(ctx, K_BC6_ROUGH, units, k_bc6_rough, grid_for(nb, NVB_BC6_ROUGH_WARPS), NVB_BC6_ROUGH_WARPS * 32, P, 2);
        NVB_LAUNCH(ctx, K_BC6_SETUP, units, k_bc6_setup, (unsigned)((nb + 127) / 128), 128, S, padded, 1);
        NVB_LAUNCH(ctx, K_BC6_ORDER, units, k_bc6_order<0>, (unsigned)(((size_t)nb * 2 + 255) / 256), 256, S);
        NVB_LAUNCH(ctx, K_BC6_ORDER, units, k_bc6_order<1>, (unsigned)(((size_t)nb * 2 + 255) / 256), 256, S);
        NVB_LAUNCH(ctx, K_BC6_SEARCH, units, k_bc6_search<2>, (unsigned)g2, 128, S);
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join[0][0], 0));
        NVB_LAUNCH(ctx, K_BC6_FINISH, units, k_bc6_finish, 2 * padded / 128, 128, S, padded);
        NVB_LAUNCH(ctx, K_BC6_SELECT, units, k_bc6_select, grid_for(nb, 256), 256, P);
    }
    else if (d->format == F_BC7) {
        // scratch per block of the level: 5 x 16 shape bytes, 8 x 16 candidate bytes, 8 errors; per block of a chunk: the
        // texel tile and the setup / result records of every searcher (Bc7ModeBytes)
        int rc = ensure(ctx, ctx->enc_scratch, encoder_scratch_bytes(F_BC7, w, h));
        if (rc != NVTTB_OK) return rc;
        if ((rc = ensure_mode_streams(ctx)) != NVTTB_OK) return rc;
        Bc7Params P;
        P.lv = lv;
        P.out = d_out;
        P.shapes = (unsigned char *)ctx->enc_scratch.p;
        P.cand = P.shapes + (size_t)nb * 80;
        P.cand_err = (float *)(P.shapes + (size_t)nb * 208);
        unsigned char *chunk_base = P.shapes + (size_t)nb * 240;
        chunk_base += (16 - ((size_t)chunk_base & 15)) & 15;
        unsigned *counters = (unsigned *)chunk_base;
        chunk_base += kBc7CounterBytes;
        const double units = (double)w * h;
        // Chunks of the level flow through the eight mode streams back to back.  A mode's records are only touched by its own
        // stream; the texel tiles are shared by all modes, so there are two tile buffers and chunk c+2 waits for chunk c.
        int nchunks = 0;
        for (int blk0 = 0; blk0 < nb; blk0 += NVB_BC7_CHUNK, ++nchunks) {
            const int parity = nchunks & 1;
            Bc7SearchParams S;
            S.P = P;
            S.blk0 = blk0;
            S.nblk = nb - blk0 < NVB_BC7_CHUNK ? nb - blk0 : NVB_BC7_CHUNK;
            float4 *tiles = (float4 *)(chunk_base + (size_t)parity * NVB_BC7_CHUNK * 256);
            S.tiles = tiles;
            S.setup = S.setup_idx = S.res = nullptr;
            S.perm = S.counters = nullptr;
            unsigned char *arena = chunk_base + (size_t)2 * NVB_BC7_CHUNK * 256;
            const double cu = units * S.nblk / nb;
            if (nchunks >= 2)
                for (int m = 0; m < 8; m++) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join[parity][m], 0));
            NVB_LAUNCH(ctx, K_BC7_TILES, cu, k_bc7_tiles, (unsigned)((S.nblk * 16 + 255) / 256), 256, S, tiles);
            CK(cudaEventRecord(ctx->ev_fork[parity], ctx->stream));  // level, scratch and tiles are ready at this point of ctx->stream
            if ((rc = launch_bc7_mode<6, 1>(ctx, S, arena, counters, parity, cu)) != NVTTB_OK) return rc;  // few, very long searches: start them first
            if ((rc = launch_bc7_mode<3, 16>(ctx, S, arena, counters, parity, cu)) != NVTTB_OK) return rc;
            if ((rc = launch_bc7_mode<7, 16>(ctx, S, arena, counters, parity, cu)) != NVTTB_OK) return rc;
            if ((rc = launch_bc7_mode<1, 16>(ctx, S, arena, counters, parity, cu)) != NVTTB_OK) return rc;
            if ((rc = launch_bc7_mode<0, 4>(ctx, S, arena, counters, parity, cu)) != NVTTB_OK) return rc;
            if ((rc = launch_bc7_mode<2, 16>(ctx, S, arena, counters, parity, cu)) != NVTTB_OK) return rc;
            if ((rc = launch_bc7_mode<4, 8>(ctx, S, arena, counters, parity, cu)) != NVTTB_OK) return rc;
            if ((rc = launch_bc7_mode<5, 4>(ctx, S, arena, counters, parity, cu)) != NVTTB_OK) return rc;
        }
        // join: the last chunk of every mode stream (stream order covers the earlier ones)
        for (int m = 0; m < 8; m++) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join[(nchunks - 1) & 1][m], 0));
        NVB_LAUNCH(ctx, K_BC7_SELECT, units, k_bc7_select, grid_for(nb, 256), 256, P);
    }
    CK(cudaGetLastError());
    return NVTTB_OK;
}

// ---- polyphase tables, cached per (filter, params, src, dst) ------------------------------------------------
static int get_poly(NvttbContext *ctx, const FilterDesc &f, int src, int dst, PolyDev *out) {
    unsigned uw, u0, u1;
    memcpy(&uw, &f.width, 4);
    memcpy(&u0, &f.p0, 4);
    memcpy(&u1, &f.p1, 4);
    auto key = std::make_tuple(f.kind, uw, u0, u1, src, dst);
    auto it = ctx->poly_cache.find(key);
    if (it != ctx->poly_cache.end()) {
        *out = it->second;
        return NVTTB_OK;
    }
    PolyphaseTable t;
    build_polyphase(f, (unsigned)src, (unsigned)dst, t);
    PolyDev pd;
    pd.window = t.window;
    pd.length = t.length;
    for (int i0 = 0; i0 < t.length; i0 += NVB_PF_TILE) {
        const int i1 = (i0 + NVB_PF_TILE - 1 < t.length) ? i0 + NVB_PF_TILE - 1 : t.length - 1;
        const int ext = t.left[i1] + t.window - t.left[i0];
        if (ext > pd.max_extent) pd.max_extent = ext;
    }
    if (dst * 2 == src && t.window <= NVB_PT_MAXW && t.length > 0) {
        pd.uniform2 = true;
        pd.left0 = t.left[0];
        for (int i = 0; i < t.length && pd.uniform2; i++) {
            if (t.left[i] != 2 * i + pd.left0) pd.uniform2 = false;
            if (memcmp(&t.weights[(size_t)i * t.window], &t.weights[0], t.window * sizeof(float)) != 0) pd.uniform2 = false;
        }
        for (int j = 0; j < t.window; j++) pd.w0[j] = t.weights[j];
    }
    CK(cudaMalloc(&pd.weights, t.weights.size() * sizeof(float)));
    CK(cudaMalloc(&pd.left, t.left.size() * sizeof(int)));
    CK(cudaMemcpy(pd.weights, t.weights.data(), t.weights.size() * sizeof(float), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(pd.left, t.left.data(), t.left.size() * sizeof(int), cudaMemcpyHostToDevice));
    ctx->poly_cache[key] = pd;
    *out = pd;
    return NVTTB_OK;
}

// src (sw x sh) -> dst (dw x dh), X pass into ctx->tmp_filter then Y pass (FloatImage::resize, FloatImage.cpp:761-808)
// cuTensorMapEncodeTiled through the runtime (no link against libcuda)
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                    const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled tensor_map_encoder() {
    static PFN_encodeTiled fn = []() -> PFN_encodeTiled {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
        return (PFN_encodeTiled)p;
    }();
    return fn;
}

// Optional block encode of the new level inside the filter kernel (k_polyphase_tma<W, true>): BC4 / BC5 quick alpha blocks.
struct MipFuse {
    unsigned char *out;     // the level's blocks
    int channels, stride;   // 1 / 8 (BC4) or 2 / 16 (BC5)
    const float *to_gamma;  // fused toGamma(2.2) table or null
    bool done;              // set when the level was encoded by the filter kernel
};

template <int W> static int launch_polyphase_tma(NvttbContext *ctx, const PolyDev &px, const PolyDev &py, int wrap, const float *src, int sw, int sh,
                                                 float *dst, int dw, int dh, bool normalize, bool *done, MipFuse *fuse = nullptr) {
    using G = PtGeom<W>;
    PFN_encodeTiled enc = tensor_map_encoder();
    if (!enc) return NVTTB_OK;  // *done stays false: the caller takes the plain kernel
    CUtensorMap map;
    const cuuint64_t dims[3] = {(cuuint64_t)sw, (cuuint64_t)sh, 4};
    const cuuint64_t strides[2] = {(cuuint64_t)sw * 4, (cuuint64_t)sw * sh * 4};
    const cuuint32_t box[3] = {(cuuint32_t)G::BOX_W, (cuuint32_t)G::IN_H, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)src, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return NVTTB_OK;
    // a function attribute belongs to the device that is current when it is set: once per device (each device has its own
    // host thread in nvttb_process_multi, so the slots are not shared)
    static bool attr_set[64] = {};
    if (ctx->device < 0 || ctx->device >= 64 || !attr_set[ctx->device]) {
        CK(cudaFuncSetAttribute(k_polyphase_tma<W, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES));
        CK(cudaFuncSetAttribute(k_polyphase_tma<W, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES));
        if (ctx->device >= 0 && ctx->device < 64) attr_set[ctx->device] = true;
    }
    PolyTmaParams Q;
    Q.src = src; Q.dst = dst; Q.sw = sw; Q.sh = sh; Q.dw = dw; Q.dh = dh; Q.wrap = wrap;
    Q.left0x = px.left0; Q.left0y = py.left0;
    for (int j = 0; j < NVB_PT_MAXW; j++) { Q.wx[j] = px.w0[j]; Q.wy[j] = py.w0[j]; }
    Q.tiles_x = (dw + NVB_PT_TW - 1) / NVB_PT_TW;
    Q.tiles_y = (dh + NVB_PT_TH - 1) / NVB_PT_TH;
    Q.normalize = normalize ? 1 : 0;
    if (fuse) {
        Q.enc_out = fuse->out;
        Q.enc_channels = fuse->channels;
        Q.enc_stride = fuse->stride;
        Q.enc_to_gamma = fuse->to_gamma;
    }
    const int ntiles = Q.tiles_x * Q.tiles_y;
    // persistent: three CTAs per SM (74 KB of shared memory each), every CTA the same number of tiles (+-1)
    const int per_cta = (ntiles + 148 * 3 - 1) / (148 * 3);
    const int grid = (ntiles + per_cta - 1) / per_cta;
    ProfRec r_;
    if (ctx->profiling) {
        r_.kid = K_POLY_2D; r_.units = (double)dw * dh;
        cudaEventCreate(&r_.a); cudaEventCreate(&r_.b); cudaEventRecord(r_.a, ctx->stream);
    }
    if (fuse) k_polyphase_tma<W, true><<<grid, NVB_PT_THREADS, G::SMEM_BYTES, ctx->stream>>>(map, Q);
    else k_polyphase_tma<W, false><<<grid, NVB_PT_THREADS, G::SMEM_BYTES, ctx->stream>>>(map, Q);
    if (ctx->profiling) { cudaEventRecord(r_.b, ctx->stream); ctx->prof.push_back(r_); }
    ctx->launches++;
    CK(cudaGetLastError());
    *done = true;
    if (fuse) fuse->done = true;
    return NVTTB_OK;
}

// normalize: also apply the normal-map renormalisation of a mip (expand -> normalise -> pack) to planes 0..2 of dst
static int resize_device(NvttbContext *ctx, const FilterDesc &f, int wrap, const float *src, int sw, int sh, float *dst, int dw, int dh, bool normalize = false,
                         MipFuse *fuse = nullptr) {
    PolyDev px, py;
    int rc;
    if ((rc = get_poly(ctx, f, sw, dw, &px)) != NVTTB_OK) return rc;
    if ((rc = get_poly(ctx, f, sh, dh, &py)) != NVTTB_OK) return rc;
    auto post = [&]() -> int {
        if (normalize) {
            NormalizeParams P{dst, (size_t)dw * dh, 1};
            NVB_LAUNCH(ctx, K_NORMALIZE, (double)P.pixels, k_normalize, grid_for(P.pixels, 256), 256, P);
            CK(cudaGetLastError());
        }
        return NVTTB_OK;
    };
    static const bool no_tma = getenv("NVB_NO_TMA") != nullptr;
    if (!no_tma && px.uniform2 && py.uniform2 && px.window == py.window && px.left0 == -(px.window - 3) / 2 && py.left0 == px.left0 && (sw & 3) == 0 && ((size_t)src & 15) == 0 && dw >= 2 * NVB_PT_TW && dh >= 2 * NVB_PT_TH &&
        (size_t)dw * dh >= (size_t)128 * NVB_PT_TW * NVB_PT_TH) {  // small levels: more (tile, plane) CTAs beat a persistent loop
        // TMA-fed persistent kernel for the 2:1 mip filters (+ fused renormalisation)
        bool done = false;
        if (px.window == 13) rc = launch_polyphase_tma<13>(ctx, px, py, wrap, src, sw, sh, dst, dw, dh, normalize, &done, fuse);
        else if (px.window == 9) rc = launch_polyphase_tma<9>(ctx, px, py, wrap, src, sw, sh, dst, dw, dh, normalize, &done, fuse);
        else if (px.window == 5) rc = launch_polyphase_tma<5>(ctx, px, py, wrap, src, sw, sh, dst, dw, dh, normalize, &done, fuse);
        if (rc != NVTTB_OK) return rc;
        if (done) return NVTTB_OK;
    }
    if (px.max_extent <= NVB_PF_EXT && py.max_extent <= NVB_PF_EXT && px.window <= NVB_PF_MAXWIN && py.window <= NVB_PF_MAXWIN) {
        // fused X+Y in shared memory: the dw x sh intermediate never reaches HBM
        Polyphase2DParams Q{src, dst, sw, sh, dw, dh, px.window, py.window, px.weights, px.left, py.weights, py.left, wrap};
        dim3 grid((dw + NVB_PF_TILE - 1) / NVB_PF_TILE, (dh + NVB_PF_TILE - 1) / NVB_PF_TILE, 4);
        // compile-time windows for the 2:1 mip filters (Kaiser 13, Mitchell 9, Triangle 5 taps), run-time windows otherwise
        if (Q.winx == 13 && Q.winy == 13) NVB_LAUNCH(ctx, K_POLY_2D, (double)dw * dh, (k_polyphase_2d_t<13, 13>), grid, 256, Q);
        else if (Q.winx == 9 && Q.winy == 9) NVB_LAUNCH(ctx, K_POLY_2D, (double)dw * dh, (k_polyphase_2d_t<9, 9>), grid, 256, Q);
        else if (Q.winx == 5 && Q.winy == 5) NVB_LAUNCH(ctx, K_POLY_2D, (double)dw * dh, (k_polyphase_2d_t<5, 5>), grid, 256, Q);
        else NVB_LAUNCH(ctx, K_POLY_2D, (double)dw * dh, (k_polyphase_2d_t<0, 0>), grid, 256, Q);
        CK(cudaGetLastError());
        return post();
    }
    if ((rc = ensure(ctx, ctx->tmp_filter, (size_t)dw * sh * 4 * sizeof(float))) != NVTTB_OK) return rc;
    float *tmp = (float *)ctx->tmp_filter.p;
    PolyphaseParams X{src, tmp, sw, sh, dw, sh, 4, px.window, px.weights, px.left, wrap};
    NVB_LAUNCH(ctx, K_POLY_X, (double)dw * sh, k_polyphase_x, grid_for((size_t)dw * sh * 4, 256), 256, X);
    PolyphaseParams Y{tmp, dst, dw, sh, dw, dh, 4, py.window, py.weights, py.left, wrap};
    NVB_LAUNCH(ctx, K_POLY_Y, (double)dw * dh, k_polyphase_y, grid_for((size_t)dw * dh * 4, 256), 256, Y);
    CK(cudaGetLastError());
    return post();
}

// Surface::buildNextMipmap semantics on raw device buffers.  Returns via *dw,*dh the new size.
static int next_mip_device(NvttbContext *ctx, int mipmapFilter, float filterWidth, float p0, float p1, int alphaMode, int wrap,
                           const float *src, int sw, int sh, float *dst, int dw, int dh, bool normalize = false, MipFuse *fuse = nullptr) {
    if (mipmapFilter == MF_Box && filterWidth == 0.5f && alphaMode != AM_Transparency) {
        BoxDownParams P{src, dst, sw, sh, dw, dh, 4};
        if ((sw & 3) == 0 && (sh & 1) == 0 && sh >= 2 && dh <= 65535 && ((size_t)src & 15) == 0 && ((size_t)dst & 7) == 0) {
            const dim3 grid(grid_for((size_t)dw / 2, 256), dh, 4);
            NVB_LAUNCH(ctx, K_BOX_DOWN, (double)dw * dh, k_box_down_even4, grid, 256, P);
        } else {
            NVB_LAUNCH(ctx, K_BOX_DOWN, (double)dw * dh, k_box_down, grid_for((size_t)dw * dh * 4, 256), 256, P);
        }
        if (normalize) {
            NormalizeParams N{dst, (size_t)dw * dh, 1};
            NVB_LAUNCH(ctx, K_NORMALIZE, (double)N.pixels, k_normalize, grid_for(N.pixels, 256), 256, N);
        }
        CK(cudaGetLastError());
        return NVTTB_OK;
    }
    FilterDesc f;
    f.kind = (mipmapFilter == MF_Box) ? Filter_Box : (mipmapFilter == MF_Triangle) ? Filter_Triangle : Filter_Kaiser;
    f.width = filterWidth;
    f.p0 = (f.kind == Filter_Kaiser) ? p0 : 0.0f;
    f.p1 = (f.kind == Filter_Kaiser) ? p1 : 0.0f;
    return resize_device(ctx, f, wrap, src, sw, sh, dst, dw, dh, normalize, fuse);
}

static void default_filter(int filter, float *width, float params[2]) {
    params[0] = params[1] = 0.0f;
    if (filter == RF_Box) *width = 0.5f;
    else if (filter == RF_Triangle) *width = 1.0f;
    else if (filter == RF_Kaiser) {
        *width = 3.0f;
        params[0] = 4.0f;
        params[1] = 1.0f;
    } else {
        *width = 2.0f;
        params[0] = 1.0f / 3.0f;
        params[1] = 1.0f / 3.0f;
    }
}

static bool nv_equal(float f0, float f1) {  // nv::equal, src/nvmath/nvmath.h:139-143
    const float eps = 0.0001f;
    float m = 1.0f;
    if (fabsf(f0) > m) m = fabsf(f0);
    if (fabsf(f1) > m) m = fabsf(f1);
    return fabs(f0 - f1) <= eps * m;
}

static int gamma_device(NvttbContext *ctx, float *data, size_t pixels, bool toLinear, float gamma) {
    if (nv_equal(gamma, 1.0f)) return NVTTB_OK;
    GammaParams P;
    P.data = data;
    P.count = 3 * pixels;
    if (gamma == 2.2f) {
        P.mode = toLinear ? 0 : 1;
        P.table = toLinear ? ctx->d_to_linear : ctx->d_to_gamma;
        P.power = 0.0f;
    } else {
        P.mode = 2;
        P.table = nullptr;
        P.power = toLinear ? gamma : 1.0f / gamma;
    }
    NVB_LAUNCH(ctx, K_GAMMA, (double)pixels, k_gamma, grid_for(P.count, 256), 256, P);
    CK(cudaGetLastError());
    return NVTTB_OK;
}

// Surface::toGreyScale: the four scales are divided by their sum first (Surface.cpp:1738-1742)
static int grey_scale_device(NvttbContext *ctx, float *data, size_t pixels, float r, float g, float b, float a) {
    const float sum = r + g + b + a;
    GreyScaleParams P;
    P.data = data;
    P.pixels = pixels;
    P.scale[0] = r / sum;
    P.scale[1] = g / sum;
    P.scale[2] = b / sum;
    P.scale[3] = a / sum;
    NVB_LAUNCH(ctx, K_GREY_SCALE, (double)pixels, k_grey_scale, grid_for(pixels, 256), 256, P);
    CK(cudaGetLastError());
    return NVTTB_OK;
}

// Surface::toNormalMap: src -> dst (different buffers), 9x9 blended Sobel on the alpha plane
static int normal_map_device(NvttbContext *ctx, const float *src, float *dst, int w, int h, int wrap, const float filterWeights[4]) {
    NormalMapParams P;
    P.src = src;
    P.dst = dst;
    P.w = w;
    P.h = h;
    P.wrap = wrap;
    build_blended_sobel(filterWeights, P.kdu);
    NVB_LAUNCH(ctx, K_NORMAL_MAP, (double)w * h, k_to_normal_map, grid_for((size_t)w * h, 256), 256, P);
    CK(cudaGetLastError());
    return NVTTB_OK;
}

static size_t input_bpp(int inputFormat) {
    switch (inputFormat) {
    case 0: return 4;
    case 1: return 8;
    case 2: return 16;
    case 3: return 4;
    default: return 0;
    }
}

// convert `pixels` interleaved texels at d_src to planar fp32 at dst (plane stride `plane`), optionally fusing toLinear(2.2)
static int convert_device(NvttbContext *ctx, int inputFormat, const void *d_src, size_t pixels, float *dst, size_t plane, bool fuseToLinear) {
    SetImageParams P{d_src, dst, (int)pixels, inputFormat, fuseToLinear ? ctx->d_to_linear : nullptr, plane};
    if (inputFormat == 0 && pixels >= 4096 && (pixels & 3) == 0 && (plane & 3) == 0 && ((size_t)d_src & 15) == 0 && ((size_t)dst & 15) == 0)
        NVB_LAUNCH(ctx, K_SET_IMAGE, (double)pixels, k_set_image_bgra8_x4, grid_for(pixels / 4, 256), 256, P);
    else
        NVB_LAUNCH(ctx, K_SET_IMAGE, (double)pixels, k_set_image, grid_for(pixels, 256), 256, P);
    CK(cudaGetLastError());
    return NVTTB_OK;
}

// upload (if needed) + convert to planar fp32, optionally fusing toLinear(2.2)
static int set_image_device(NvttbContext *ctx, int inputFormat, int w, int h, const void *data, int location, float *dst, bool fuseToLinear) {
    const size_t bpp = input_bpp(inputFormat);
    if (bpp == 0 || w <= 0 || h <= 0 || !data) return fail(ctx, NVTTB_ERR_INVALID_INPUT, "bad input image");
    const size_t bytes = (size_t)w * h * bpp;
    const void *d_src = data;
    if (location == NVTTB_HOST) {
        int rc = ensure(ctx, ctx->in_stage, bytes);
        if (rc != NVTTB_OK) return rc;
        CK(cudaMemcpyAsync(ctx->in_stage.p, data, bytes, cudaMemcpyHostToDevice, ctx->stream));
        d_src = ctx->in_stage.p;
    }
    return convert_device(ctx, inputFormat, d_src, (size_t)w * h, dst, (size_t)w * h, fuseToLinear);
}

extern "C" {

int nvttb_encode_level(NvttbContext *ctx, const NvttbEncodeDesc *desc, const float *rgba, int rgba_location, void *out,
                       int out_location, size_t out_capacity) {
    if (!ctx || !desc || !rgba || !out) return fail(ctx, NVTTB_ERR_INVALID_INPUT, "null argument");
    CK(cudaSetDevice(ctx->device));
    const int w = desc->width, h = desc->height;
    const size_t size = nvttb_level_size(desc->format, w, h);
    if (size == 0) return fail(ctx, NVTTB_ERR_UNSUPPORTED_FEATURE, "unsupported format");
    if (out_capacity < size) return fail(ctx, NVTTB_ERR_INVALID_INPUT, "output buffer too small");
    int rc;
    const float *d_rgba = rgba;
    if (rgba_location == NVTTB_HOST) {
        const size_t bytes = (size_t)w * h * 4 * sizeof(float);
        if ((rc = ensure(ctx, ctx->tmp_level, bytes)) != NVTTB_OK) return rc;
        CK(cudaMemcpyAsync(ctx->tmp_level.p, rgba, bytes, cudaMemcpyHostToDevice, ctx->stream));
        d_rgba = (const float *)ctx->tmp_level.p;
    }
    unsigned char *d_out = (unsigned char *)out;
    if (out_location == NVTTB_HOST) {
        if ((rc = ensure(ctx, ctx->out_dev, size)) != NVTTB_OK) return rc;
        d_out = (unsigned char *)ctx->out_dev.p;
    }
    if ((rc = encode_device(ctx, desc, d_rgba, w, h, d_out)) != NVTTB_OK) return rc;
    if (out_location == NVTTB_HOST) {
        CK(cudaMemcpyAsync(out, d_out, size, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    } else if (rgba_location == NVTTB_HOST) {
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return NVTTB_OK;
}

// ---- Format_RGB / Format_RGBA: PixelFormatConverter::compress (src/nvtt/CompressorRGB.cpp:410-575) --------------------------
namespace {
struct PixelLayout {
    unsigned bitCount, pitch;
    int kind, aligned;
    unsigned size[4], shift[4];
};
void mask_shift_size(unsigned mask, unsigned *shift, unsigned *size) {  // PixelFormat::maskShiftAndSize
    *shift = 0;
    *size = 0;
    if (!mask) return;
    while ((mask & 1) == 0) { ++*shift; mask >>= 1; }
    while ((mask & 1) == 1) { ++*size; mask >>= 1; }
}
// 0 when the description is one the reference asserts on (more than 32 bits of fixed point, bad float sizes are zero-padded as there)
bool pixel_layout(const NvttbPixelFormatDesc *d, PixelLayout *L) {
    if (d->width <= 0 || d->height <= 0 || d->pitchAlignment <= 0 || (d->pitchAlignment & (d->pitchAlignment - 1))) return false;
    memset(L, 0, sizeof *L);
    const unsigned sz[4] = {d->rsize, d->gsize, d->bsize, d->asize};
    if (d->pixelType == 4 /* PixelType_Float */) {
        for (int i = 0; i < 4; i++) { if (sz[i] > 32) return false; L->size[i] = sz[i]; }
        L->bitCount = sz[0] + sz[1] + sz[2] + sz[3];
        L->kind = 2;
        L->aligned = 1;
        for (int i = 0; i < 4; i++) if (sz[i] != 0 && sz[i] != 16 && sz[i] != 32) L->aligned = 0;
    } else {
        if (d->bitcount != 0) {
            L->bitCount = d->bitcount;
            const unsigned mask[4] = {d->rmask, d->gmask, d->bmask, d->amask};
            for (int i = 0; i < 4; i++) mask_shift_size(mask[i], &L->shift[i], &L->size[i]);
        } else {
            for (int i = 0; i < 4; i++) { if (sz[i] > 32) return false; L->size[i] = sz[i]; }
            L->bitCount = sz[0] + sz[1] + sz[2] + sz[3];
            L->shift[3] = 0;
            L->shift[2] = L->shift[3] + sz[3];
            L->shift[1] = L->shift[2] + sz[2];
            L->shift[0] = L->shift[1] + sz[1];
        }
        if (L->bitCount > 32) return false;  // nvCheck(bitCount <= 32)
        // UnsignedNorm = 0, UnsignedInt = 2; SignedNorm / SignedInt write zero components; SharedExp writes zeros unless 9/9/9/5
        if (d->pixelType == 0) L->kind = 0;
        else if (d->pixelType == 2) L->kind = 1;
        else if (d->pixelType == 6 /* PixelType_SharedExp */ && L->size[0] == 9 && L->size[1] == 9 && L->size[2] == 9 && L->size[3] == 5) {
            // R9G9B9E5 (toFloat3SE): always one 32-bit word per pixel - putBits(v.v, 32) whatever bitCount says - so only the
            // layouts whose bitCount is 32 keep the scanline pitch consistent with what is written
            if (L->bitCount != 32) return false;
            L->kind = 4;
        }
        else L->kind = 3;
        L->aligned = (L->bitCount % 8) == 0;
    }
    if (L->bitCount == 0) return false;
    const unsigned alignBits = 8u * (unsigned)d->pitchAlignment;  // computeBitPitch / computeBytePitch (nvimage.h:11-24)
    const unsigned bitPitch = (((unsigned)d->width * L->bitCount + alignBits - 1) / alignBits) * alignBits;
    L->pitch = (bitPitch + 7) / 8;
    return true;
}
}  // namespace

size_t nvttb_pixel_format_level_size(const NvttbPixelFormatDesc *desc) {
    PixelLayout L;
    if (!desc || !pixel_layout(desc, &L)) return 0;
    return (size_t)L.pitch * desc->height;
}

static int convert_device(NvttbContext *ctx, const NvttbPixelFormatDesc *d, const PixelLayout &L, const float *d_rgba, unsigned char *d_out) {
    PixelFormatParams P;
    P.lv.data = d_rgba;
    P.lv.plane = (size_t)d->width * d->height;
    P.lv.w = d->width;
    P.lv.h = d->height;
    P.lv.bw = (d->width + 3) / 4;
    P.lv.bh = (d->height + 3) / 4;
    P.lv.to_gamma_table = nullptr;
    P.out = d_out;
    P.pitch = L.pitch;
    P.bitCount = L.bitCount;
    P.kind = L.kind;
    for (int i = 0; i < 4; i++) { P.size[i] = L.size[i]; P.shift[i] = L.shift[i]; }
    // four pixels per thread when every row start and every 4-pixel group stays 16-byte aligned in both images
    int mode = 0;
    const unsigned bytes = L.bitCount / 8;
    if (L.aligned && L.kind <= 1 && (bytes == 1 || bytes == 2 || bytes == 4)) mode = (int)bytes;
    else if (L.kind == 2 && L.size[0] == 16 && L.size[1] == 16 && L.size[2] == 16 && L.size[3] == 16) mode = 8;
    else if (L.kind == 2 && L.size[0] == 32 && L.size[1] == 32 && L.size[2] == 32 && L.size[3] == 32) mode = 16;
    if (mode && d->width % 4 == 0 && L.pitch == (unsigned)d->width * bytes && ((size_t)d_rgba & 15) == 0 && ((size_t)d_out & 15) == 0 &&
        (((size_t)d->width * d->height) & 3) == 0) {
        const dim3 grid(grid_for(d->width / 4, 256), d->height);
        NVB_LAUNCH(ctx, K_PIXEL_FORMAT, (double)d->width * d->height, k_pixel_format_x4, grid, 256, P, mode);
    } else if (L.aligned) {
        const dim3 grid(grid_for(d->width, 256), d->height);
        NVB_LAUNCH(ctx, K_PIXEL_FORMAT, (double)d->width * d->height, k_pixel_format, grid, 256, P);
    } else {
        NVB_LAUNCH(ctx, K_PIXEL_FORMAT, (double)d->width * d->height, k_pixel_format_rows, grid_for(d->height, 64), 64, P);
    }
    return NVTTB_OK;
}

int nvttb_convert_level(NvttbContext *ctx, const NvttbPixelFormatDesc *desc, const float *rgba, int rgba_location, void *out,
                        int out_location, size_t out_capacity) {
    if (!ctx || !desc || !rgba || !out) return fail(ctx, NVTTB_ERR_INVALID_INPUT, "null argument");
    PixelLayout L;
    if (!pixel_layout(desc, &L)) return fail(ctx, NVTTB_ERR_UNSUPPORTED_FEATURE, "unsupported pixel format");
    if (desc->height > 65535) return fail(ctx, NVTTB_ERR_UNSUPPORTED_FEATURE, "level too tall");
    CK(cudaSetDevice(ctx->device));
    const int w = desc->width, h = desc->height;
    const size_t size = (size_t)L.pitch * h;
    if (out_capacity < size) return fail(ctx, NVTTB_ERR_INVALID_INPUT, "output buffer too small");
    int rc;
    const float *d_rgba = rgba;
    if (rgba_location == NVTTB_HOST) {
        const size_t bytes = (size_t)w * h * 4 * sizeof(float);
        if ((rc = ensure(ctx, ctx->tmp_level, bytes)) != NVTTB_OK) return rc;
        CK(cudaMemcpyAsync(ctx->tmp_level.p, rgba, bytes, cudaMemcpyHostToDevice, ctx->stream));
        d_rgba = (const float *)ctx->tmp_level.p;
    }
    unsigned char *d_out = (unsigned char *)out;
    if (out_location == NVTTB_HOST) {
        if ((rc = ensure(ctx, ctx->out_dev, size)) != NVTTB_OK) return rc;
        d_out = (unsigned char *)ctx->out_dev.p;
    }
    if ((rc = convert_device(ctx, desc, L, d_rgba, d_out)) != NVTTB_OK) return rc;
    if (out_location == NVTTB_HOST) {
        CK(cudaMemcpyAsync(out, d_out, size, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    } else if (rgba_location == NVTTB_HOST) {
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return NVTTB_OK;
}

// ---- surfaces ---------------------------------------------------------------------------------------------
int nvttb_surface_create(NvttbContext *ctx, NvttbSurface **out) {
    if (!ctx || !out) return NVTTB_ERR_INVALID_INPUT;
    NvttbSurface *s = new NvttbSurface();
    s->ctx = ctx;
    *out = s;
    return NVTTB_OK;
}
void nvttb_surface_destroy(NvttbSurface *s) {
    if (!s) return;
    if (s->buf.p) {
        cudaSetDevice(s->ctx->device);
        cudaStreamSynchronize(s->ctx->stream);
        cudaFree(s->buf.p);
    }
    delete s;
}
int nvttb_surface_clone(const NvttbSurface *s, NvttbSurface **out) {
    if (!s || !out) return NVTTB_ERR_INVALID_INPUT;
    NvttbContext *ctx = s->ctx;
    NvttbSurface *c = new NvttbSurface();
    c->ctx = ctx;
    c->w = s->w;
    c->h = s->h;
    c->wrapMode = s->wrapMode;
    c->alphaMode = s->alphaMode;
    c->isNormalMap = s->isNormalMap;
    if (s->buf.p && s->w > 0) {
        const size_t bytes = (size_t)s->w * s->h * 16;
        int rc = ensure(ctx, c->buf, bytes);
        if (rc != NVTTB_OK) { delete c; return rc; }
        cudaError_t e = cudaMemcpyAsync(c->buf.p, s->buf.p, bytes, cudaMemcpyDeviceToDevice, ctx->stream);
        if (e != cudaSuccess) { cudaFree(c->buf.p); delete c; return fail(ctx, NVTTB_ERR_CUDA, "clone", e); }
    }
    *out = c;
    return NVTTB_OK;
}
void nvttb_surface_set_wrap_mode(NvttbSurface *s, int m) { if (s) s->wrapMode = m; }
void nvttb_surface_set_alpha_mode(NvttbSurface *s, int m) { if (s) s->alphaMode = m; }
void nvttb_surface_set_normal_map(NvttbSurface *s, int b) { if (s) s->isNormalMap = b; }
int nvttb_surface_width(const NvttbSurface *s) { return s ? s->w : 0; }
int nvttb_surface_height(const NvttbSurface *s) { return s ? s->h : 0; }
const float *nvttb_surface_device_data(const NvttbSurface *s) { return s ? (const float *)s->buf.p : nullptr; }

int nvttb_surface_set_image(NvttbSurface *s, int inputFormat, int w, int h, const void *data, int location) {
    if (!s) return NVTTB_ERR_INVALID_INPUT;
    NvttbContext *ctx = s->ctx;
    CK(cudaSetDevice(ctx->device));
    int rc = ensure(ctx, s->buf, (size_t)w * h * 16);
    if (rc != NVTTB_OK) return rc;
    if ((rc = set_image_device(ctx, inputFormat, w, h, data, location, (float *)s->buf.p, false)) != NVTTB_OK) return rc;
    s->w = w;
    s->h = h;
    if (location == NVTTB_HOST) CK(cudaStreamSynchronize(ctx->stream));
    return NVTTB_OK;
}

int nvttb_surface_set_image_2d(NvttbSurface *s, int format, int decoder, int w, int h, const void *data, int location, int bc6Signed) {
    if (!s || !data || w <= 0 || h <= 0) return NVTTB_ERR_INVALID_INPUT;
    NvttbContext *ctx = s->ctx;
    const bool ok = format == F_DXT1 || format == F_DXT3 || format == F_DXT5 || format == F_DXT5n || format == 12 /*BC3_RGBM*/ ||
                    format == F_BC4 || format == F_BC5 || format == F_BC6 || format == F_BC7;
    if (!ok || decoder < 0 || decoder > 2) return fail(ctx, NVTTB_ERR_INVALID_INPUT, "setImage2D: format / decoder not decodable");
    CK(cudaSetDevice(ctx->device));
    const int bw = (w + 3) / 4, bh = (h + 3) / 4;
    const size_t bytes = (size_t)bw * bh * ((format == F_DXT1 || format == F_BC4) ? 8 : 16);
    int rc = ensure(ctx, s->buf, (size_t)w * h * 16);
    if (rc != NVTTB_OK) return rc;
    const unsigned char *d_blocks = (const unsigned char *)data;
    if (location == NVTTB_HOST) {
        if ((rc = ensure(ctx, ctx->in_stage, bytes)) != NVTTB_OK) return rc;
        CK(cudaMemcpyAsync(ctx->in_stage.p, data, bytes, cudaMemcpyHostToDevice, ctx->stream));
        d_blocks = (const unsigned char *)ctx->in_stage.p;
    }
    DecodeParams P;
    P.blocks = d_blocks;
    P.out = (float *)s->buf.p;
    P.w = w; P.h = h; P.bw = bw; P.bh = bh;
    P.format = format; P.decoder = decoder; P.bc6_signed = bc6Signed;
    if (format == F_BC6 || format == F_BC7) NVB_LAUNCH(ctx, K_DECODE, (double)w * h, k_decode_blocks, grid_for((size_t)bw * bh, 128), 128, P);
    else NVB_LAUNCH(ctx, K_DECODE, (double)w * h, k_decode_dxt, grid_for((size_t)bw * bh, 128), 128, P);
    CK(cudaGetLastError());
    s->w = w;
    s->h = h;
    if (location == NVTTB_HOST) CK(cudaStreamSynchronize(ctx->stream));
    return NVTTB_OK;
}

static int error_metric(const NvttbSurface *ref, const NvttbSurface *img, int mode, float *out) {
    if (!ref || !img || !out) return NVTTB_ERR_INVALID_INPUT;
    NvttbContext *ctx = ref->ctx;
    if (!ref->buf.p || !img->buf.p || ref->w != img->w || ref->h != img->h) {  // !sameLayout
        *out = FLT_MAX;
        return NVTTB_OK;
    }
    CK(cudaSetDevice(ctx->device));
    const size_t count = (size_t)ref->w * ref->h;
    const unsigned grid = grid_for(count, 256);
    int rc = ensure(ctx, ctx->tmp_filter, (size_t)grid * sizeof(double));
    if (rc != NVTTB_OK) return rc;
    ErrorParams P;
    P.ref = (const float *)ref->buf.p;
    P.img = (const float *)img->buf.p;
    P.count = count;
    P.mode = mode;
    P.partial = (double *)ctx->tmp_filter.p;
    NVB_LAUNCH(ctx, K_ERROR_METRIC, (double)count, k_error_metric, grid, 256, P);
    CK(cudaGetLastError());
    std::vector<double> part(grid);
    CK(cudaMemcpyAsync(part.data(), P.partial, grid * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    double mse = 0;
    for (unsigned i = 0; i < grid; i++) mse += part[i];
    *out = (mode == 4) ? (float)(mse / (double)(unsigned)count) : (float)sqrt(mse / (double)(unsigned)count);
    return NVTTB_OK;
}
int nvttb_rms_error(const NvttbSurface *reference, const NvttbSurface *img, float *out) {
    return error_metric(reference, img, (reference && reference->alphaMode == AM_Transparency) ? 1 : 0, out);
}
int nvttb_rms_alpha_error(const NvttbSurface *reference, const NvttbSurface *img, float *out) { return error_metric(reference, img, 2, out); }
int nvttb_angular_error(const NvttbSurface *reference, const NvttbSurface *img, float *out) { return error_metric(reference, img, 3, out); }
int nvttb_cielab_error(const NvttbSurface *reference, const NvttbSurface *img, float *out) { return error_metric(reference, img, 4, out); }

int nvttb_surface_to_linear(NvttbSurface *s, float gamma) {
    if (!s) return NVTTB_ERR_INVALID_INPUT;
    if (!s->buf.p || s->w == 0) return NVTTB_OK;
    cudaSetDevice(s->ctx->device);
    return gamma_device(s->ctx, (float *)s->buf.p, (size_t)s->w * s->h, true, gamma);
}
int nvttb_surface_to_gamma(NvttbSurface *s, float gamma) {
    if (!s) return NVTTB_ERR_INVALID_INPUT;
    if (!s->buf.p || s->w == 0) return NVTTB_OK;
    cudaSetDevice(s->ctx->device);
    return gamma_device(s->ctx, (float *)s->buf.p, (size_t)s->w * s->h, false, gamma);
}

int nvttb_surface_build_next_mipmap(NvttbSurface *s, int mipmapFilter, int useParams, float filterWidth, const float *params, int *built) {
    if (!s) return NVTTB_ERR_INVALID_INPUT;
    NvttbContext *ctx = s->ctx;
    if (built) *built = 0;
    if (!s->buf.p || (s->w == 1 && s->h == 1)) return NVTTB_OK;
    CK(cudaSetDevice(ctx->device));
    float fw, pr[2];
    default_filter(mipmapFilter, &fw, pr);
    if (useParams) {
        fw = filterWidth;
        if (params) { pr[0] = params[0]; pr[1] = params[1]; }
    }
    const int dw = s->w / 2 > 1 ? s->w / 2 : 1, dh = s->h / 2 > 1 ? s->h / 2 : 1;
    DevBuf nb;
    int rc = ensure(ctx, nb, (size_t)dw * dh * 16);
    if (rc != NVTTB_OK) return rc;
    rc = next_mip_device(ctx, mipmapFilter, fw, pr[0], pr[1], s->alphaMode, s->wrapMode, (const float *)s->buf.p, s->w, s->h, (float *)nb.p, dw, dh);
    if (rc != NVTTB_OK) { cudaFree(nb.p); return rc; }
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(s->buf.p);
    s->buf = nb;
    s->w = dw;
    s->h = dh;
    if (built) *built = 1;
    return NVTTB_OK;
}

int nvttb_surface_resize(NvttbSurface *s, int w, int h, int resizeFilter, int useParams, float filterWidth, const float *params) {
    if (!s) return NVTTB_ERR_INVALID_INPUT;
    NvttbContext *ctx = s->ctx;
    if (!s->buf.p || (w == s->w && h == s->h)) return NVTTB_OK;
    if (w <= 0 || h <= 0) return fail(ctx, NVTTB_ERR_INVALID_INPUT, "bad resize extent");
    CK(cudaSetDevice(ctx->device));
    float fw, pr[2];
    default_filter(resizeFilter, &fw, pr);
    if (useParams) {
        fw = filterWidth;
        if (params) { pr[0] = params[0]; pr[1] = params[1]; }
    }
    FilterDesc f;
    f.kind = resizeFilter;
    f.width = (resizeFilter == RF_Mitchell) ? 2.0f : fw;  // MitchellFilter() ignores filterWidth (Surface.cpp:1212-1216)
    f.p0 = (resizeFilter == RF_Kaiser || resizeFilter == RF_Mitchell) ? pr[0] : 0.0f;
    f.p1 = (resizeFilter == RF_Kaiser || resizeFilter == RF_Mitchell) ? pr[1] : 0.0f;
    DevBuf nb;
    int rc = ensure(ctx, nb, (size_t)w * h * 16);
    if (rc != NVTTB_OK) return rc;
    rc = resize_device(ctx, f, s->wrapMode, (const float *)s->buf.p, s->w, s->h, (float *)nb.p, w, h);
    if (rc != NVTTB_OK) { cudaFree(nb.p); return rc; }
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(s->buf.p);
    s->buf = nb;
    s->w = w;
    s->h = h;
    return NVTTB_OK;
}

static int scale_bias(NvttbSurface *s, float scale, float bias) {
    NvttbContext *ctx = s->ctx;
    if (!s->buf.p) return NVTTB_OK;
    CK(cudaSetDevice(ctx->device));
    ScaleBiasParams P{(float *)s->buf.p, (size_t)3 * s->w * s->h, scale, bias};
    NVB_LAUNCH(ctx, K_SCALE_BIAS, (double)s->w * s->h, k_scale_bias, grid_for(P.count, 256), 256, P);
    CK(cudaGetLastError());
    return NVTTB_OK;
}
int nvttb_surface_expand_normals(NvttbSurface *s) { return s ? scale_bias(s, 2.0f, -1.0f) : NVTTB_ERR_INVALID_INPUT; }
int nvttb_surface_pack_normals(NvttbSurface *s) { return s ? scale_bias(s, 0.5f, 0.5f) : NVTTB_ERR_INVALID_INPUT; }
int nvttb_surface_scale_bias(NvttbSurface *s, int channel, int count, float scale, float bias) {
    if (!s || channel < 0 || count < 1 || channel + count > 4) return NVTTB_ERR_INVALID_INPUT;
    NvttbContext *ctx = s->ctx;
    if (!s->buf.p) return NVTTB_OK;
    CK(cudaSetDevice(ctx->device));
    const size_t n = (size_t)s->w * s->h;
    ScaleBiasParams P{(float *)s->buf.p + (size_t)channel * n, (size_t)count * n, scale, bias};
    NVB_LAUNCH(ctx, K_SCALE_BIAS, (double)n, k_scale_bias, grid_for(P.count, 256), 256, P);
    CK(cudaGetLastError());
    return NVTTB_OK;
}
int nvttb_surface_clamp(NvttbSurface *s, int channel, float low, float high) {
    if (!s || channel < 0 || channel > 3) return NVTTB_ERR_INVALID_INPUT;
    NvttbContext *ctx = s->ctx;
    if (!s->buf.p) return NVTTB_OK;
    CK(cudaSetDevice(ctx->device));
    const size_t n = (size_t)s->w * s->h;
    ClampParams P{(float *)s->buf.p + (size_t)channel * n, n, low, high};
    NVB_LAUNCH(ctx, K_SCALE_BIAS, (double)n, k_clamp, grid_for(n, 256), 256, P);
    CK(cudaGetLastError());
    return NVTTB_OK;
}
int nvttb_surface_range(const NvttbSurface *s, int channel, int alpha_channel, float alpha_ref, float *range_min, float *range_max) {
    if (!s || channel < 0 || channel > 3 || alpha_channel > 3) return NVTTB_ERR_INVALID_INPUT;
    NvttbContext *ctx = s->ctx;
    float lo = FLT_MAX, hi = -FLT_MAX;
    if (s->buf.p) {
        CK(cudaSetDevice(ctx->device));
        const size_t n = (size_t)s->w * s->h;
        const unsigned grid = grid_for(n, 256);
        int rc = ensure(ctx, ctx->tmp_filter, (size_t)grid * sizeof(float2));
        if (rc != NVTTB_OK) return rc;
        RangeParams P{(const float *)s->buf.p + (size_t)channel * n, alpha_channel >= 0 ? (const float *)s->buf.p + (size_t)alpha_channel * n : nullptr,
                      n, alpha_ref, (float2 *)ctx->tmp_filter.p};
        NVB_LAUNCH(ctx, K_ERROR_METRIC, (double)n, k_range, grid, 256, P);
        CK(cudaGetLastError());
        std::vector<float2> part(grid);
        CK(cudaMemcpyAsync(part.data(), P.partial, grid * sizeof(float2), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        for (unsigned i = 0; i < grid; i++) {
            if (part[i].x < lo) lo = part[i].x;
            if (part[i].y > hi) hi = part[i].y;
        }
    }
    if (range_min) *range_min = lo;
    if (range_max) *range_max = hi;
    return NVTTB_OK;
}
int nvttb_surface_tone_map(NvttbSurface *s, int toneMapper) {
    if (!s || toneMapper < 0 || toneMapper > 3) return NVTTB_ERR_INVALID_INPUT;
    NvttbContext *ctx = s->ctx;
    if (!s->buf.p) return NVTTB_OK;
    CK(cudaSetDevice(ctx->device));
    ToneMapParams P{(float *)s->buf.p, (size_t)s->w * s->h, toneMapper};
    NVB_LAUNCH(ctx, K_SCALE_BIAS, (double)P.pixels, k_tone_map, grid_for(P.pixels, 256), 256, P);
    CK(cudaGetLastError());
    return NVTTB_OK;
}
int nvttb_surface_to_rgbm(NvttbSurface *s, float range, float threshold) {
    if (!s) return NVTTB_ERR_INVALID_INPUT;
    (void)range;  // shadowed by a local in the reference (Surface.cpp:1905): no effect there either
    NvttbContext *ctx = s->ctx;
    if (!s->buf.p) return NVTTB_OK;
    CK(cudaSetDevice(ctx->device));
    threshold = threshold < 1e-6f ? 1e-6f : (threshold > 1.0f ? 1.0f : threshold);
    ToRgbmParams P{(float *)s->buf.p, (size_t)s->w * s->h, threshold};
    NVB_LAUNCH(ctx, K_SCALE_BIAS, (double)P.pixels, k_to_rgbm, grid_for(P.pixels, 256), 256, P);
    CK(cudaGetLastError());
    return NVTTB_OK;
}
static int quantize_channel(NvttbSurface *s, int channel, QuantizeParams P) {
    NvttbContext *ctx = s->ctx;
    if (channel < 0 || channel > 3) return NVTTB_ERR_INVALID_INPUT;
    if (!s->buf.p) return NVTTB_OK;
    CK(cudaSetDevice(ctx->device));
    const size_t n = (size_t)s->w * s->h;
    P.data = (float *)s->buf.p + (size_t)channel * n;
    P.w = s->w;
    P.h = s->h;
    P.carry = nullptr;
    if (P.dither) {
        // Floyd-Steinberg: one CTA walks the plane as a wavefront (k_quantize_dither); two rows of carried errors
        int rc = ensure(ctx, ctx->tmp_filter, (size_t)2 * s->w * sizeof(float));
        if (rc != NVTTB_OK) return rc;
        P.carry = (float *)ctx->tmp_filter.p;
        NVB_LAUNCH(ctx, K_QUANTIZE, (double)n, k_quantize_dither, 1, NVB_FS_ROWS, P);
    } else {
        NVB_LAUNCH(ctx, K_QUANTIZE, (double)n, k_quantize, grid_for(n, 256), 256, P);
    }
    CK(cudaGetLastError());
    return NVTTB_OK;
}
int nvttb_surface_binarize(NvttbSurface *s, int channel, float threshold, int dither) {
    if (!s) return NVTTB_ERR_INVALID_INPUT;
    QuantizeParams P = {};
    P.binarize = 1;
    P.threshold = threshold;
    P.dither = dither ? 1 : 0;
    return quantize_channel(s, channel, P);
}
int nvttb_surface_quantize(NvttbSurface *s, int channel, int bits, int exactEndPoints, int dither) {
    if (!s || bits < 1 || bits > 24) return NVTTB_ERR_INVALID_INPUT;
    QuantizeParams P = {};
    P.binarize = 0;
    P.dither = dither ? 1 : 0;
    if (exactEndPoints) {  // floor(x * (range - 1) + 0.5) / (range - 1)
        P.scale = (float)((1 << bits) - 1);
        P.offset0 = 0.5f;
        P.offset1 = 0.0f;
    } else {  // (floor(x * range) + 0.5) / range
        P.scale = (float)(1 << bits);
        P.offset0 = 0.0f;
        P.offset1 = 0.5f;
    }
    return quantize_channel(s, channel, P);
}
int nvttb_surface_normalize_normal_map(NvttbSurface *s) {
    if (!s) return NVTTB_ERR_INVALID_INPUT;
    NvttbContext *ctx = s->ctx;
    if (!s->buf.p || !s->isNormalMap) return NVTTB_OK;  // Surface::normalizeNormalMap is a no-op unless flagged (Surface.cpp:2810-2817)
    CK(cudaSetDevice(ctx->device));
    NormalizeParams P{(float *)s->buf.p, (size_t)s->w * s->h, 0};
    NVB_LAUNCH(ctx, K_NORMALIZE, (double)P.pixels, k_normalize, grid_for(P.pixels, 256), 256, P);
    CK(cudaGetLastError());
    return NVTTB_OK;
}
int nvttb_surface_to_grey_scale(NvttbSurface *s, float r, float g, float b, float a) {
    if (!s) return NVTTB_ERR_INVALID_INPUT;
    NvttbContext *ctx = s->ctx;
    if (!s->buf.p) return NVTTB_OK;
    CK(cudaSetDevice(ctx->device));
    return grey_scale_device(ctx, (float *)s->buf.p, (size_t)s->w * s->h, r, g, b, a);
}
int nvttb_surface_to_normal_map(NvttbSurface *s, float sm, float medium, float big, float large) {
    if (!s) return NVTTB_ERR_INVALID_INPUT;
    NvttbContext *ctx = s->ctx;
    if (!s->buf.p) return NVTTB_OK;
    CK(cudaSetDevice(ctx->device));
    DevBuf nb;
    int rc = ensure(ctx, nb, (size_t)s->w * s->h * 16);
    if (rc != NVTTB_OK) return rc;
    const float fw[4] = {sm, medium, big, large};
    rc = normal_map_device(ctx, (const float *)s->buf.p, (float *)nb.p, s->w, s->h, s->wrapMode, fw);
    if (rc != NVTTB_OK) { cudaFree(nb.p); return rc; }
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(s->buf.p);
    s->buf = nb;
    s->isNormalMap = 1;  // Surface::toNormalMap sets the flag (Surface.cpp:2807)
    return NVTTB_OK;
}

int nvttb_surface_download(const NvttbSurface *s, float *out) {
    if (!s || !out) return NVTTB_ERR_INVALID_INPUT;
    NvttbContext *ctx = s->ctx;
    if (!s->buf.p) return fail(ctx, NVTTB_ERR_INVALID_INPUT, "null surface");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(out, s->buf.p, (size_t)s->w * s->h * 16, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return NVTTB_OK;
}

int nvttb_surface_encode(NvttbSurface *s, const NvttbEncodeDesc *desc, void *out, int out_location, size_t out_capacity) {
    if (!s || !desc || !out) return NVTTB_ERR_INVALID_INPUT;
    if (!s->buf.p) return fail(s->ctx, NVTTB_ERR_INVALID_INPUT, "null surface");
    NvttbEncodeDesc d = *desc;
    d.width = s->w;
    d.height = s->h;
    d.alphaMode = s->alphaMode;
    return nvttb_encode_level(s->ctx, &d, (const float *)s->buf.p, NVTTB_DEVICE, out, out_location, out_capacity);
}

// ---- whole pipeline -----------------------------------------------------------------------------------------
int nvttb_process_mip_count(const NvttbProcessDesc *d) {
    if (!d || d->width <= 0 || d->height <= 0) return 0;
    int count = 1;
    if (d->generateMipmaps) {
        unsigned w = d->width, h = d->height;
        while (w != 1 || h != 1) {
            w = w / 2 > 1 ? w / 2 : 1;
            h = h / 2 > 1 ? h / 2 : 1;
            count++;
        }
        if (d->maxLevel > 0 && d->maxLevel < count) count = d->maxLevel;
    }
    return count;
}

// ---- block-row sharding of one image (NvttbProcessDesc.band*) ---------------------------------------------------------------
struct ShardGeom {
    int N = 1, b = 0;  // bands, this band
    int C = 0;         // level-0 texel rows per chunk; 0 = nothing can be distributed (band 0 encodes every level)
    int K = 0;         // chunks per band
};
// false: not sharded
static bool shard_geom(const NvttbProcessDesc *d, ShardGeom *g) {
    *g = ShardGeom();
    if (d->bandCount <= 1) return false;
    g->N = d->bandCount;
    g->b = d->bandIndex;
    const int H = d->height;
    const int C = d->bandChunkRows > 0 ? d->bandChunkRows : (H % d->bandCount == 0 ? H / d->bandCount : 0);
    if (C <= 0 || (C & 3) != 0 || H % C != 0 || (H / C) % d->bandCount != 0) return true;
    g->C = C;
    g->K = H / C / d->bandCount;
    return true;
}
// block rows of one chunk at level m (whose height is h); 0 = the level is not distributed
static int level_rpc(const NvttbProcessDesc *d, const ShardGeom &g, int m, int h) {
    if (g.C == 0 || m > 24) return 0;
    const int q = 4 << m;
    if (g.C % q != 0 || ((long long)h << m) != (long long)d->height) return 0;
    return g.C / q;
}
static void level_extent(const NvttbProcessDesc *d, int level, int *w, int *h) {
    int lw = d->width, lh = d->height;
    for (int m = 0; m < level; m++) {
        lw = lw / 2 > 1 ? lw / 2 : 1;
        lh = lh / 2 > 1 ? lh / 2 : 1;
    }
    *w = lw;
    *h = lh;
}

extern "C" int nvttb_process_band_slices(const NvttbProcessDesc *d, int level, size_t *offset, size_t *bytes, size_t *pitch, int *count) {
    if (!d || !offset || !bytes || !pitch || !count || level < 0 || level >= nvttb_process_mip_count(d)) return NVTTB_ERR_INVALID_INPUT;
    int w, h;
    level_extent(d, level, &w, &h);
    ShardGeom g;
    *offset = 0;
    *pitch = 0;
    if (!shard_geom(d, &g)) {
        *bytes = nvttb_level_size(d->encode.format, w, h);
        *count = 1;
        return NVTTB_OK;
    }
    const int rpc = level_rpc(d, g, level, h);
    if (rpc == 0) {  // tail level: band 0 encodes all of it
        *bytes = g.b == 0 ? nvttb_level_size(d->encode.format, w, h) : 0;
        *count = g.b == 0 ? 1 : 0;
        return NVTTB_OK;
    }
    const size_t row_bytes = (size_t)((w + 3) / 4) * block_bytes(d->encode.format);
    *bytes = (size_t)rpc * row_bytes;
    *pitch = (size_t)g.N * rpc * row_bytes;
    *offset = (size_t)g.b * rpc * row_bytes;
    *count = g.K;
    return NVTTB_OK;
}

static size_t face_bytes(const NvttbProcessDesc *d) {
    const int mips = nvttb_process_mip_count(d);
    size_t total = 0;
    for (int m = 0; m < mips; m++) {
        size_t off, bytes, pitch;
        int count;
        nvttb_process_band_slices(d, m, &off, &bytes, &pitch, &count);
        total += bytes * (size_t)count;
    }
    return total;
}

static size_t whole_face_bytes(const NvttbProcessDesc *d) {
    NvttbProcessDesc t = *d;
    t.bandIndex = 0;
    t.bandCount = 0;
    return face_bytes(&t);
}

size_t nvttb_process_output_size(const NvttbProcessDesc *d) {
    if (!d) return 0;
    return face_bytes(d) * (size_t)(d->faceCount > 0 ? d->faceCount : 1);
}

// Runs faces [f0,f1) and leaves their encoded chains in d_out (face-major, mip-minor).
// h_out (optional, pinned host memory of the same layout as d_out): finished pieces are copied back on d2h_stream while
// the rest of the chain is still being computed; the caller synchronises d2h_stream.
// in_place (sharded images only): d_out is the whole chain of the processed faces and the band's slices go to their final offsets.
// While alive (and `on`), everything the context launches goes to side_stream.  Per-launch event timing (nvttb_profile_*)
// describes the main stream; launches that overlap it would be counted twice.
struct SideScope {
    NvttbContext *c;
    cudaStream_t *other;
    bool on, prof;
    SideScope(NvttbContext *ctx, bool enable, cudaStream_t *which = nullptr) : c(ctx), other(which ? which : &ctx->side_stream), on(enable), prof(ctx->profiling) {
        if (on) {
            std::swap(c->stream, *other);
            c->profiling = false;
        }
    }
    ~SideScope() {
        if (on) {
            std::swap(c->stream, *other);
            c->profiling = prof;
        }
    }
};

static int process_faces(NvttbContext *ctx, const NvttbProcessDesc *d, const void *const *images, int loc, int f0, int f1,
                         unsigned char *d_out, unsigned char *h_out = nullptr, bool in_place = false) {
    const int mips = nvttb_process_mip_count(d);
    const size_t fbytes = face_bytes(d);
    const int W = d->width, H = d->height;
    int rc;
    // level buffers: A holds level 0, B the levels 1.. one after the other (every level stays until the face is done, so the
    // encode of a level need not finish before the next ones are built)
    DevBuf &A = ctx->lvlA, &B = ctx->lvlB;  // persistent scratch: no cudaMalloc/cudaFree per call
    if ((rc = ensure(ctx, A, (size_t)W * H * 16)) != NVTTB_OK) return rc;
    if (mips > NvttbContext::MAX_LEVELS) return fail(ctx, NVTTB_ERR_INVALID_INPUT, "too many levels");
    size_t mip_off[NvttbContext::MAX_LEVELS + 1];  // floats
    mip_off[0] = mip_off[1] = 0;
    {
        int w = W, h = H;
        for (int m = 1; m < mips; m++) {
            w = w / 2 > 1 ? w / 2 : 1;
            h = h / 2 > 1 ? h / 2 : 1;
            mip_off[m + 1] = mip_off[m] + (((size_t)4 * w * h + 3) & ~(size_t)3);  // 16-byte aligned levels
        }
    }
    if (mips > 1 && (rc = ensure(ctx, B, mip_off[mips] * sizeof(float))) != NVTTB_OK) return rc;
    auto cleanup = [&]() {
        cudaStreamSynchronize(ctx->h2d_stream);
        cudaStreamSynchronize(ctx->stream);
        cudaStreamSynchronize(ctx->side_stream);
        cudaStreamSynchronize(ctx->alt_stream);
        cudaStreamSynchronize(ctx->d2h_stream);
    };
    const bool toNormal = d->convertToNormalMap != 0;
    const bool isNormal = d->isNormalMap || toNormal;  // toNormalMap flags the surface as a normal map (Surface.cpp:2807)
    const bool colour = !isNormal;
    const bool linFast = colour && d->inputGamma == 2.2f;
    const bool gamFast = colour && d->outputGamma == 2.2f;
    const bool gamSlow = colour && !gamFast && !nv_equal(d->outputGamma, 1.0f);
    // Row-band pipeline for host input: level 0 is uploaded, converted and encoded band by band, so the H2D copy of band
    // b+1 (h2d_stream) runs under the encode of band b (stream) and the D2H copy of band b's blocks (d2h_stream) under
    // everything that follows.  Needs the per-texel-only prologue (fused toLinear or none).
    const size_t bpp = input_bpp(d->inputFormat);
    ShardGeom geom;
    const bool sharded = shard_geom(d, &geom);
    // (not for BC6H: a level is a chain of seven kernels around a dynamically scheduled search, and one chain per band costs
    // far more in tails - 348 instead of 220 ms for a 6 x 2048² cube - than the 0.6 ms per face the upload would hide)
    const bool banded = !sharded && loc == NVTTB_HOST && bpp != 0 && !toNormal && !(colour && !linFast) && !gamSlow && H >= 64 && (size_t)W * H >= (1u << 14) &&
                        d->encode.format != F_BC6;
    // mips on side_stream, beside the level-0 encode (not for the encoders that share per-context scratch, BC6H / BC7: they are
    // long enough for the latency of a small level not to matter)
    const bool use_side = mips > 1 && !sharded && !gamSlow && encoder_scratch_bytes(d->encode.format, W, H) == 0;
    // opt-in: the 2:1 filter kernel block-encodes the level it builds (BC4 / BC5 quick alpha blocks; polyphase_tma.cuh)
    static const bool fused_env = getenv("NVTT_B200_FUSED_MIP_ENCODE") != nullptr;
    const bool can_fuse = fused_env && !sharded && !gamSlow && (d->encode.format == F_BC4 || d->encode.format == F_BC5) && d->encode.quality < Q_Production;
    const int bhTotal = (H + 3) / 4;
    // Band boundaries in texel rows.  Large images: eight bands, the first one cut again at 1/8 of its rows - nothing can be
    // encoded before the first band has arrived, and uploading is ~8x faster than encoding (BC1 Production), so a first band of
    // 1/64 of the image (4 MB of an 8192² BGRA8 image: 0.08 ms instead of 0.64 ms) still hides the upload of the next one.
    int band_y[NvttbContext::MAX_BANDS + 1];
    int nbands = 1;
    band_y[0] = 0;
    band_y[1] = H;
    if (banded) {
        const int even = bhTotal >= 64 ? 8 : 2;
        const int bandBlockRows = (bhTotal + even - 1) / even;
        nbands = 0;
        if (even == 8 && bandBlockRows >= 16) band_y[++nbands] = (bandBlockRows / 8) * 4;
        for (int b = 1; b <= even; b++) {
            const int y = b * bandBlockRows * 4 < H ? b * bandBlockRows * 4 : H;
            if (y > band_y[nbands]) band_y[++nbands] = y;
        }
    }
    const size_t bs = (size_t)block_bytes(d->encode.format), bw0 = (size_t)((W + 3) / 4);
    for (int f = f0; f < f1; f++) {
        unsigned char *out = d_out + (size_t)(f - f0) * fbytes;
        unsigned char *hout = h_out ? h_out + (size_t)(f - f0) * fbytes : nullptr;
        // in_place: d_out is the whole chain of the processed faces
        unsigned char *whole_face = d_out + (size_t)(f - f0) * whole_face_bytes(d);
        size_t level_off = 0;
        bool level0_done = false;
        if (banded) {
            if (!images[f]) return fail(ctx, NVTTB_ERR_INVALID_INPUT, "bad input image");
            if ((rc = ensure(ctx, ctx->in_stage, (size_t)W * H * bpp)) != NVTTB_OK) { cleanup(); return rc; }
            // the staging buffer may still be read by conversion kernels of the previous face or of an earlier call (an event
            // that was never recorded is complete, so the wait is free the first time)
            CK(cudaStreamWaitEvent(ctx->h2d_stream, ctx->ev_stage_free, 0));
            for (int b = 0; b < nbands; b++) {
                const int y0 = band_y[b], y1 = band_y[b + 1];
                const size_t off = (size_t)y0 * W * bpp;
                CK(cudaMemcpyAsync((char *)ctx->in_stage.p + off, (const char *)images[f] + off, (size_t)(y1 - y0) * W * bpp, cudaMemcpyHostToDevice, ctx->h2d_stream));
                CK(cudaEventRecord(ctx->ev_up[b], ctx->h2d_stream));
            }
            NvttbEncodeDesc e = d->encode;
            e.width = W;
            e.alphaMode = d->alphaMode;
            e.applyToGamma = gamFast ? 1 : 0;
            if (use_side) {
                // the level buffer may still be read on `stream` by the previous face / call: the side stream starts behind it
                CK(cudaEventRecord(ctx->ev_side_go, ctx->stream));
                CK(cudaStreamWaitEvent(ctx->side_stream, ctx->ev_side_go, 0));
            }
            for (int b = 0; b < nbands; b++) {
                const int y0 = band_y[b], y1 = band_y[b + 1];
                float *rows = (float *)A.p + (size_t)y0 * W;
                {
                    // conversion as soon as the band has arrived: on the (high priority) side stream when there is one, so that it
                    // does not queue behind the encode of an earlier band
                    SideScope side(ctx, use_side);
                    CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_up[b], 0));
                    if ((rc = convert_device(ctx, d->inputFormat, (const char *)ctx->in_stage.p + (size_t)y0 * W * bpp, (size_t)(y1 - y0) * W, rows, (size_t)W * H, linFast)) != NVTTB_OK) { cleanup(); return rc; }
                    if (use_side) CK(cudaEventRecord(ctx->ev_cv[b], ctx->stream));
                    if (y1 == H) CK(cudaEventRecord(ctx->ev_stage_free, ctx->stream));  // level 0 is complete (the mips follow on this stream)
                }
                e.height = y1 - y0;
                const size_t ooff = (size_t)(y0 / 4) * bw0 * bs, obytes = (size_t)((y1 - y0 + 3) / 4) * bw0 * bs;
                {
                    SideScope alt(ctx, use_side && (b & 1), &ctx->alt_stream);  // odd bands on the second encode stream
                    if (use_side) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_cv[b], 0));
                    if (use_side && b == 1) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_side_go, 0));  // alt_stream too starts behind the previous image
                    if ((rc = encode_device(ctx, &e, rows, W, y1 - y0, out + ooff, (size_t)W * H)) != NVTTB_OK) { cleanup(); return rc; }
                    if (hout) {
                        CK(cudaEventRecord(ctx->ev_enc[b], ctx->stream));
                        CK(cudaStreamWaitEvent(ctx->d2h_stream, ctx->ev_enc[b], 0));
                        CK(cudaMemcpyAsync(hout + ooff, out + ooff, obytes, cudaMemcpyDeviceToHost, ctx->d2h_stream));
                    }
                    if (use_side && (b & 1)) CK(cudaEventRecord(ctx->ev_alt_done, ctx->stream));
                }
            }
            if (use_side && nbands > 1) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_alt_done, 0));
            level0_done = true;
        } else {
            // setImage (+ toLinear).  The staging buffer is shared with the banded pipeline's copy stream.
            if (loc == NVTTB_HOST) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_stage_free, 0));
            if ((rc = set_image_device(ctx, d->inputFormat, W, H, images[f], loc, (float *)A.p, linFast)) != NVTTB_OK) { cleanup(); return rc; }
            if (colour && !linFast) {
                if ((rc = gamma_device(ctx, (float *)A.p, (size_t)W * H, true, d->inputGamma)) != NVTTB_OK) { cleanup(); return rc; }
            }
        }
        float *cur = (float *)A.p;
        if (toNormal) {
            // img.toGreyScale(heightFactors); img.toNormalMap(bumpFrequencyScale)  (Context.cpp:269-272)
            if ((rc = ensure(ctx, ctx->tmp_level, (size_t)W * H * 16)) != NVTTB_OK) { cleanup(); return rc; }
            if ((rc = grey_scale_device(ctx, cur, (size_t)W * H, d->heightFactors[0], d->heightFactors[1], d->heightFactors[2], d->heightFactors[3])) != NVTTB_OK) { cleanup(); return rc; }
            if ((rc = normal_map_device(ctx, cur, (float *)ctx->tmp_level.p, W, H, d->wrapMode, d->bumpFrequencyScale)) != NVTTB_OK) { cleanup(); return rc; }
            CK(cudaMemcpyAsync(cur, ctx->tmp_level.p, (size_t)W * H * 16, cudaMemcpyDeviceToDevice, ctx->stream));
        }
        if (use_side && !banded) {
            CK(cudaEventRecord(ctx->ev_side_go, ctx->stream));  // level 0 is complete
            CK(cudaStreamWaitEvent(ctx->side_stream, ctx->ev_side_go, 0));
        }  // banded: level 0 was converted on the side stream itself
        int w = W, h = H;
        for (int m = 0; m < mips; m++) {
            SideScope side(ctx, use_side && m > 0);
            MipFuse fuse{nullptr, 0, 0, nullptr, false};
            if (m > 0) {
                const int dw = w / 2 > 1 ? w / 2 : 1, dh = h / 2 > 1 ? h / 2 : 1;
                float fw, pr[2];
                default_filter(d->mipmapFilter, &fw, pr);
                if (d->mipmapFilter == MF_Kaiser) {
                    fw = d->kaiserWidth;
                    pr[0] = d->kaiserAlpha;
                    pr[1] = d->kaiserStretch;
                }
                // normal maps: the mip is renormalised (expand -> normalise -> pack) before it is encoded and down-sampled again
                float *nxt = (float *)B.p + mip_off[m];
                fuse = MipFuse{out, d->encode.format == F_BC5 ? 2 : 1, d->encode.format == F_BC5 ? 16 : 8, gamFast ? ctx->d_to_gamma : nullptr, false};
                if ((rc = next_mip_device(ctx, d->mipmapFilter, fw, pr[0], pr[1], d->alphaMode, d->wrapMode, cur, w, h, nxt, dw, dh, isNormal && d->normalizeMipmaps,
                                          can_fuse ? &fuse : nullptr)) != NVTTB_OK) { cleanup(); return rc; }
                cur = nxt;
                w = dw;
                h = dh;
            }
            // tmp = img; tmp.toGamma(outputGamma); compress(tmp)
            NvttbEncodeDesc e = d->encode;
            e.width = w;
            e.height = h;
            e.alphaMode = d->alphaMode;
            e.applyToGamma = gamFast ? 1 : 0;
            const float *src = cur;
            if (gamSlow) {
                if ((rc = ensure(ctx, ctx->tmp_level, (size_t)w * h * 16)) != NVTTB_OK) { cleanup(); return rc; }
                cudaMemcpyAsync(ctx->tmp_level.p, cur, (size_t)w * h * 16, cudaMemcpyDeviceToDevice, ctx->stream);
                if ((rc = gamma_device(ctx, (float *)ctx->tmp_level.p, (size_t)w * h, false, d->outputGamma)) != NVTTB_OK) { cleanup(); return rc; }
                src = (const float *)ctx->tmp_level.p;
            }
            if (sharded) {
                // every band has the whole fp32 level (replicated front end); each encodes the rows of its own chunks
                size_t soff, sbytes, spitch;
                int scount;
                nvttb_process_band_slices(d, m, &soff, &sbytes, &spitch, &scount);
                const int rpc = level_rpc(d, geom, m, h);
                for (int j = 0; j < scount; j++) {
                    const int y0 = rpc ? (j * geom.N + geom.b) * rpc * 4 : 0, rows = rpc ? rpc * 4 : h;
                    e.height = rows;
                    unsigned char *dst = in_place ? whole_face + level_off + soff + (size_t)j * spitch : out;
                    if ((rc = encode_device(ctx, &e, src + (size_t)y0 * w, w, rows, dst, (size_t)w * h)) != NVTTB_OK) { cleanup(); return rc; }
                    out += sbytes;
                }
                level_off += nvttb_level_size(e.format, w, h);
                continue;
            }
            if (!(m == 0 && level0_done) && !fuse.done) {  // fuse.done: the filter kernel has encoded the level it built
                if ((rc = encode_device(ctx, &e, src, w, h, out)) != NVTTB_OK) { cleanup(); return rc; }
            }
            out += nvttb_level_size(e.format, w, h);
        }
        if (use_side) CK(cudaEventRecord(ctx->ev_side_done, ctx->side_stream));
        if (hout && !in_place) {
            // whatever of this face has not been sent yet: the mip tail (banded) or the whole chain
            const size_t done = level0_done ? nvttb_level_size(d->encode.format, W, H) : 0;
            if (fbytes > done) {
                if (use_side && level0_done) {
                    // the mips were encoded on the side stream: they go home as soon as they are done, under the level-0 encode
                    CK(cudaStreamWaitEvent(ctx->d2h_stream, ctx->ev_side_done, 0));
                } else {
                    if (use_side) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_side_done, 0));
                    CK(cudaEventRecord(ctx->ev_tail, ctx->stream));
                    CK(cudaStreamWaitEvent(ctx->d2h_stream, ctx->ev_tail, 0));
                }
                CK(cudaMemcpyAsync(hout + done, d_out + (size_t)(f - f0) * fbytes + done, fbytes - done, cudaMemcpyDeviceToHost, ctx->d2h_stream));
            }
        }
        // the face is done when both streams are; the level buffers are reused by the next face
        if (use_side) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_side_done, 0));
        // the level buffers are reused by the next face: same stream, so ordering is implicit
    }
    CK(cudaGetLastError());
    return NVTTB_OK;
}

// ---- band-local front end of a sharded image ------------------------------------------------------------------------------------
struct LocalPlan {
    ShardGeom g;
    int mips = 0;
    int k = -1;  // last distributed level
    int wk = 0, hk = 0;
};
// true when a band can build everything it encodes from its own chunks of the source image: 2x2 Box mips (a chunk of
// 4 << k rows down-samples to whole block rows of levels 0..k without looking at its neighbours) and per-texel colour ops only
static bool shard_local_plan(const NvttbProcessDesc *d, LocalPlan *p) {
    *p = LocalPlan();
    if (!shard_geom(d, &p->g) || p->g.C == 0) return false;
    const bool toNormal = d->convertToNormalMap != 0;
    const bool colour = !(d->isNormalMap || toNormal);
    if (toNormal || d->mipmapFilter != MF_Box || d->alphaMode == AM_Transparency || input_bpp(d->inputFormat) == 0) return false;
    if (colour && !(d->inputGamma == 2.2f || nv_equal(d->inputGamma, 1.0f))) return false;
    if (colour && !(d->outputGamma == 2.2f || nv_equal(d->outputGamma, 1.0f))) return false;
    p->mips = nvttb_process_mip_count(d);
    int w = d->width, h = d->height;
    for (int m = 0; m < p->mips; m++) {
        if (level_rpc(d, p->g, m, h) == 0) break;
        if (((long long)w << m) != (long long)d->width || (m > 0 && (w & 1) && w != 1)) return false;  // widths must halve exactly too
        p->k = m;
        p->wk = w;
        p->hk = h;
        w = w / 2 > 1 ? w / 2 : 1;
        h = h / 2 > 1 ? h / 2 : 1;
    }
    return p->k >= 0;
}

extern "C" size_t nvttb_process_exchange_size(const NvttbProcessDesc *d) {
    LocalPlan p;
    if (!d || !shard_local_plan(d, &p)) return 0;
    return (size_t)NVB_XCHG_HEADER + (size_t)2 * 4 * p.wk * p.hk * sizeof(float);
}

// Every device buffer one band-local shard call needs, sized up front.  A cudaMalloc / cudaFree may wait for the device to go
// idle, and band 0 keeps a waiting kernel resident until every band has delivered its rows - so nothing may allocate once
// the first launch of the image is out (matters when several bands share one GPU, as in the tests; and for band 0 itself).
static int shard_local_prepare(NvttbContext *ctx, const NvttbProcessDesc *d, const LocalPlan &pl, bool host_in, bool own_out) {
    const ShardGeom &g = pl.g;
    const int W = d->width, HL = g.K * g.C;
    int rc;
    size_t tot = 0;
    for (int m = 0; m <= pl.k; m++) tot += (size_t)4 * (W >> m) * (HL >> m);
    if ((rc = ensure(ctx, ctx->lvlA, tot * sizeof(float))) != NVTTB_OK) return rc;
    if (host_in && (rc = ensure(ctx, ctx->in_stage, (size_t)HL * W * input_bpp(d->inputFormat))) != NVTTB_OK) return rc;
    if (own_out && (rc = ensure(ctx, ctx->out_dev, whole_face_bytes(d))) != NVTTB_OK) return rc;
    const size_t es = encoder_scratch_bytes(d->encode.format, W, HL);
    if (es && (rc = ensure(ctx, ctx->enc_scratch, es)) != NVTTB_OK) return rc;
    if (es && (rc = ensure_mode_streams(ctx)) != NVTTB_OK) return rc;
    if (pl.k < pl.mips - 1 && g.b == 0) {
        const int w1 = pl.wk / 2 > 1 ? pl.wk / 2 : 1, h1 = pl.hk / 2 > 1 ? pl.hk / 2 : 1;
        if ((rc = ensure(ctx, ctx->tail_lvl, (size_t)2 * 4 * w1 * h1 * sizeof(float))) != NVTTB_OK) return rc;
        const size_t ts = encoder_scratch_bytes(d->encode.format, w1, h1);
        if (ts && (rc = ensure(ctx, ctx->tail_scratch, ts)) != NVTTB_OK) return rc;
    }
    return NVTTB_OK;
}

extern "C" int nvttb_process_prepare(NvttbContext *ctx, const NvttbProcessDesc *d, int images_location, int own_output) {
    if (!ctx || !d) return NVTTB_ERR_INVALID_INPUT;
    LocalPlan pl;
    if (d->bandCount <= 1 || !d->bandExchange || !shard_local_plan(d, &pl)) return NVTTB_OK;  // nothing waits on the device in the other paths
    CK(cudaSetDevice(ctx->device));
    return shard_local_prepare(ctx, d, pl, images_location == NVTTB_HOST, own_output != 0);
}

// One face of a sharded image on band g.b: upload / convert / down-sample the band's own chunks, hand the rows of level k to
// band 0, encode the band's rows of levels 0..k; band 0 also runs the tail (levels k+1..) on tail_stream.
// out: whole-face layout on the device (possibly a peer's memory); h_out: whole-face layout on the host, or null.
static int process_shard_local(NvttbContext *ctx, const NvttbProcessDesc *d, const LocalPlan &pl, const void *image, int loc,
                               unsigned char *out, unsigned char *h_out) {
    const ShardGeom &g = pl.g;
    const int N = g.N, b = g.b, C = g.C, K = g.K, k = pl.k, mips = pl.mips;
    const int W = d->width;
    const size_t bpp = input_bpp(d->inputFormat);
    const int HL = K * C;  // rows of the band's level 0 (its chunks concatenated)
    if (!image) return fail(ctx, NVTTB_ERR_INVALID_INPUT, "bad input image");
    if (mips > NvttbContext::MAX_LEVELS) return fail(ctx, NVTTB_ERR_INVALID_INPUT, "too many levels");
    int rc;
    size_t loff[NvttbContext::MAX_LEVELS], lvl_off[NvttbContext::MAX_LEVELS + 1], tot = 0;
    for (int m = 0; m <= k; m++) {
        loff[m] = tot;
        tot += (size_t)4 * (W >> m) * (HL >> m);
    }
    lvl_off[0] = 0;
    {
        int w = W, h = d->height;
        for (int m = 0; m < mips; m++) {
            lvl_off[m + 1] = lvl_off[m] + nvttb_level_size(d->encode.format, w, h);
            w = w / 2 > 1 ? w / 2 : 1;
            h = h / 2 > 1 ? h / 2 : 1;
        }
    }
    if ((rc = shard_local_prepare(ctx, d, pl, loc == NVTTB_HOST, false)) != NVTTB_OK) return rc;
    float *const chain = (float *)ctx->lvlA.p;
    const bool isNormal = d->isNormalMap != 0;
    const bool colour = !isNormal;
    const bool linFast = colour && d->inputGamma == 2.2f;
    const bool gamFast = colour && d->outputGamma == 2.2f;
    const size_t bs = (size_t)block_bytes(d->encode.format);
    const bool tail = k < mips - 1;
    auto cleanup = [&]() {
        cudaStreamSynchronize(ctx->h2d_stream);
        cudaStreamSynchronize(ctx->stream);
        cudaStreamSynchronize(ctx->side_stream);
        cudaStreamSynchronize(ctx->alt_stream);
        cudaStreamSynchronize(ctx->tail_stream);
        cudaStreamSynchronize(ctx->d2h_stream);
    };
    NvttbEncodeDesc e = d->encode;
    e.alphaMode = d->alphaMode;
    e.applyToGamma = gamFast ? 1 : 0;
    const bool host_in = loc == NVTTB_HOST;
    const bool per_chunk_events = K <= NvttbContext::MAX_BANDS;
    const size_t chunk_in = (size_t)C * W * bpp;
    // Everything but the level-0 encode runs on side_stream (process_faces explains why): conversion of the chunks as they
    // arrive, the band's mip chain, the export to band 0 and the encodes of levels 1..k.  Not for the encoders that share
    // per-context scratch (BC6H / BC7).
    const bool use_side = encoder_scratch_bytes(d->encode.format, W, HL) == 0;
    const bool side_conv = use_side && per_chunk_events;
    if (use_side) {
        // the previous image may still be read on `stream`: the side stream starts behind it
        CK(cudaEventRecord(ctx->ev_side_go, ctx->stream));
        CK(cudaStreamWaitEvent(ctx->side_stream, ctx->ev_side_go, 0));
    }
    // 1. the band's chunks: upload (copy stream), convert and - for host input - encode level 0 piece by piece, so that
    //    the copies of the following pieces hide under the encode.  A piece is a chunk, except that the first chunk is cut at
    //    1/8 of its rows: nothing can be encoded before the first piece has arrived, and all the GPUs of the box pull their
    //    first piece from the same host memory at the same moment.
    struct Piece { int j, r0, nr, ev; };  // chunk, first row inside it, rows, index of its events
    Piece pieces[NvttbContext::MAX_BANDS + 1];
    int npieces = 0;
    {
        const bool split = host_in && per_chunk_events && K + 1 <= NvttbContext::MAX_BANDS && C >= 64 && (C / 8) % 4 == 0;
        for (int j = 0; j < K; j++) {
            if (j == 0 && split) {
                pieces[npieces++] = Piece{0, 0, C / 8, K};
                pieces[npieces++] = Piece{0, C / 8, C - C / 8, 0};
            } else {
                pieces[npieces++] = Piece{j, 0, C, per_chunk_events ? j : 0};
            }
        }
    }
    const size_t row_in = (size_t)W * bpp;
    if (host_in) {
        CK(cudaStreamWaitEvent(ctx->h2d_stream, ctx->ev_stage_free, 0));
        for (int q = 0; q < npieces; q++) {
            const Piece &pc = pieces[q];
            const size_t c = (size_t)pc.j * N + b;
            CK(cudaMemcpyAsync((char *)ctx->in_stage.p + pc.j * chunk_in + pc.r0 * row_in, (const char *)image + c * chunk_in + pc.r0 * row_in,
                               (size_t)pc.nr * row_in, cudaMemcpyHostToDevice, ctx->h2d_stream));
            if (per_chunk_events) CK(cudaEventRecord(ctx->ev_up[pc.ev], ctx->h2d_stream));
        }
        if (!per_chunk_events) CK(cudaEventRecord(ctx->ev_up[0], ctx->h2d_stream));
    }
    const int rpc0 = C / 4;
    const size_t row_bytes0 = (size_t)((W + 3) / 4) * bs;
    for (int q = 0; q < npieces; q++) {
        const Piece &pc = pieces[q];
        const size_t c = (size_t)pc.j * N + b;
        float *rows = chain + ((size_t)pc.j * C + pc.r0) * W;
        {
            SideScope side(ctx, side_conv);
            if (host_in && (per_chunk_events || q == 0)) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_up[pc.ev], 0));
            const char *src = host_in ? (const char *)ctx->in_stage.p + pc.j * chunk_in + pc.r0 * row_in : (const char *)image + c * chunk_in + pc.r0 * row_in;
            if ((rc = convert_device(ctx, d->inputFormat, src, (size_t)pc.nr * W, rows, (size_t)W * HL, linFast)) != NVTTB_OK) { cleanup(); return rc; }
            if (side_conv) CK(cudaEventRecord(ctx->ev_cv[pc.ev], ctx->stream));
            if (host_in && q == npieces - 1) CK(cudaEventRecord(ctx->ev_stage_free, ctx->stream));
        }
        if (host_in) {
            // odd pieces on the second encode stream: a piece starts while the last blocks of the one before it are still running
            const bool on_alt = side_conv && (q & 1);
            SideScope alt(ctx, on_alt, &ctx->alt_stream);
            if (side_conv) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_cv[pc.ev], 0));
            if (on_alt && q == 1) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_side_go, 0));  // alt_stream too starts behind the previous image
            e.width = W;
            e.height = pc.nr;
            const size_t ooff = lvl_off[0] + (c * rpc0 + pc.r0 / 4) * row_bytes0;
            if ((rc = encode_device(ctx, &e, rows, W, pc.nr, out + ooff, (size_t)W * HL)) != NVTTB_OK) { cleanup(); return rc; }
            if (h_out) {
                cudaEvent_t ev = ctx->ev_enc[pc.ev];
                CK(cudaEventRecord(ev, ctx->stream));
                CK(cudaStreamWaitEvent(ctx->d2h_stream, ev, 0));
                CK(cudaMemcpyAsync(h_out + ooff, out + ooff, (size_t)(pc.nr / 4) * row_bytes0, cudaMemcpyDeviceToHost, ctx->d2h_stream));
            }
            if (on_alt) CK(cudaEventRecord(ctx->ev_alt_done, ctx->stream));
        }
    }
    if (host_in && side_conv && npieces > 1) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_alt_done, 0));
    if (use_side && !side_conv) {
        // the chunks were converted on `stream`: the side stream continues from there
        CK(cudaEventRecord(ctx->ev_side_go, ctx->stream));
        CK(cudaStreamWaitEvent(ctx->side_stream, ctx->ev_side_go, 0));
    }
    // 2. the band's rows of levels 1..k
    {
        SideScope side(ctx, use_side);
        for (int m = 1; m <= k; m++) {
            float *lm = chain + loff[m];
            if ((rc = next_mip_device(ctx, MF_Box, 0.5f, 0.0f, 0.0f, d->alphaMode, d->wrapMode, chain + loff[m - 1], W >> (m - 1), HL >> (m - 1), lm, W >> m, HL >> m)) != NVTTB_OK) { cleanup(); return rc; }
            if (isNormal && d->normalizeMipmaps) {
                NormalizeParams P{lm, (size_t)(W >> m) * (HL >> m), 1};
                NVB_LAUNCH(ctx, K_NORMALIZE, (double)P.pixels, k_normalize, grid_for(P.pixels, 256), 256, P);
            }
        }
    }
    // 3. rows of level k -> band 0 (peer stores), then the arrival flag
    unsigned *xhdr = (unsigned *)d->bandExchange;
    const unsigned seq = d->bandSequence;
    float *ximg = tail ? (float *)((char *)d->bandExchange + NVB_XCHG_HEADER) + (size_t)(seq & 1u) * 4 * pl.wk * pl.hk : nullptr;
    if (tail) {
        SideScope side(ctx, use_side);
        // the buffer of this parity was last used by image seq - 2: band 0 must have finished reading it
        if (seq >= 3) NVB_LAUNCH(ctx, K_XCHG, 0.0, k_xchg_wait, 1, 32, xhdr + NVB_XCHG_ACK, 1, seq - 2, ctx->h_fault);
        ExportRowsParams X{chain + loff[k], ximg, pl.wk, HL >> k, pl.hk, C >> k, N, b};
        NVB_LAUNCH(ctx, K_XCHG, (double)X.w * X.hl, k_export_rows, grid_for((size_t)4 * X.w * X.hl, 256), 256, X);
        NVB_LAUNCH(ctx, K_XCHG, 0.0, k_xchg_signal, 1, 1, xhdr + b, seq);
    }
    // 4. band 0: the tail levels from the exchanged level, on the second stream with its own scratch
    if (tail && b == 0) {
        const int w1 = pl.wk / 2 > 1 ? pl.wk / 2 : 1, h1 = pl.hk / 2 > 1 ? pl.hk / 2 : 1;
        const size_t lvl1 = (size_t)4 * w1 * h1;
        std::swap(ctx->stream, ctx->tail_stream);
        std::swap(ctx->enc_scratch, ctx->tail_scratch);
        // per-launch event timing (nvttb_profile_*) describes the main stream; launches that overlap it would be counted twice
        const bool was_profiling = ctx->profiling;
        ctx->profiling = false;
        auto unswap = [&]() {
            std::swap(ctx->stream, ctx->tail_stream);
            std::swap(ctx->enc_scratch, ctx->tail_scratch);
            ctx->profiling = was_profiling;
        };
        NVB_LAUNCH(ctx, K_XCHG, 0.0, k_xchg_wait, 1, NVB_XCHG_FLAGS, xhdr, N, seq, ctx->h_fault);
        const float *cur = ximg;
        int w = pl.wk, h = pl.hk;
        for (int m = k + 1; m < mips; m++) {
            const int dw = w / 2 > 1 ? w / 2 : 1, dh = h / 2 > 1 ? h / 2 : 1;
            float *nxt = (float *)ctx->tail_lvl.p + (size_t)((m - k - 1) & 1) * lvl1;
            if ((rc = next_mip_device(ctx, MF_Box, 0.5f, 0.0f, 0.0f, d->alphaMode, d->wrapMode, cur, w, h, nxt, dw, dh)) != NVTTB_OK) { unswap(); cleanup(); return rc; }
            if (m == k + 1) NVB_LAUNCH(ctx, K_XCHG, 0.0, k_xchg_signal, 1, 1, xhdr + NVB_XCHG_ACK, seq);  // the exchanged level has been read
            cur = nxt;
            w = dw;
            h = dh;
            if (isNormal && d->normalizeMipmaps) {
                NormalizeParams P{nxt, (size_t)w * h, 1};
                NVB_LAUNCH(ctx, K_NORMALIZE, (double)P.pixels, k_normalize, grid_for(P.pixels, 256), 256, P);
            }
            e.width = w;
            e.height = h;
            if ((rc = encode_device(ctx, &e, cur, w, h, out + lvl_off[m])) != NVTTB_OK) { unswap(); cleanup(); return rc; }
        }
        if (h_out) {
            cudaError_t ce = cudaEventRecord(ctx->ev_tail, ctx->stream);
            if (ce == cudaSuccess) ce = cudaStreamWaitEvent(ctx->d2h_stream, ctx->ev_tail, 0);
            if (ce == cudaSuccess) ce = cudaMemcpyAsync(h_out + lvl_off[k + 1], out + lvl_off[k + 1], lvl_off[mips] - lvl_off[k + 1], cudaMemcpyDeviceToHost, ctx->d2h_stream);
            if (ce != cudaSuccess) { unswap(); cleanup(); return fail(ctx, NVTTB_ERR_CUDA, "tail copy", ce); }
        }
        cudaError_t ce = cudaEventRecord(ctx->ev_tail_done, ctx->stream);
        unswap();
        if (ce != cudaSuccess) { cleanup(); return fail(ctx, NVTTB_ERR_CUDA, "cudaEventRecord", ce); }
    }
    // 5. the band's rows of the distributed levels: one launch per level over the concatenated chunks
    for (int m = host_in ? 1 : 0; m <= k; m++) {
        SideScope side(ctx, use_side && m > 0);
        if (m == 0 && side_conv) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_cv[pieces[npieces - 1].ev], 0));  // level 0 was converted on the side stream
        const int w = W >> m, hl = HL >> m;
        const int rpc = C / (4 << m);
        const size_t row_bytes = (size_t)((w + 3) / 4) * bs;
        e.width = w;
        e.height = hl;
        CycView cyc{rpc, N, b};
        if ((rc = encode_device(ctx, &e, chain + loff[m], w, hl, out + lvl_off[m], 0, &cyc)) != NVTTB_OK) { cleanup(); return rc; }
        if (h_out) {
            CK(cudaEventRecord(ctx->ev_lvl[m], ctx->stream));
            CK(cudaStreamWaitEvent(ctx->d2h_stream, ctx->ev_lvl[m], 0));
            const size_t first = lvl_off[m] + (size_t)b * rpc * row_bytes, pitch = (size_t)N * rpc * row_bytes;
            CK(cudaMemcpy2DAsync(h_out + first, pitch, out + first, pitch, (size_t)rpc * row_bytes, K, cudaMemcpyDeviceToHost, ctx->d2h_stream));
        }
    }
    if (use_side) {
        CK(cudaEventRecord(ctx->ev_side_done, ctx->side_stream));
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_side_done, 0));
    }
    if (tail && b == 0) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_tail_done, 0));
    CK(cudaGetLastError());
    return NVTTB_OK;
}

static int check_process(NvttbContext *ctx, const NvttbProcessDesc *d, const void *const *images, int *f0, int *f1) {
    if (!ctx || !d || !images) return fail(ctx, NVTTB_ERR_INVALID_INPUT, "null argument");
    if (d->width <= 0 || d->height <= 0 || d->faceCount <= 0) return fail(ctx, NVTTB_ERR_INVALID_INPUT, "bad extent");
    if (d->bandCount > 1 && (d->bandIndex < 0 || d->bandIndex >= d->bandCount)) return fail(ctx, NVTTB_ERR_INVALID_INPUT, "bad band index");
    if (d->bandCount > 1 && d->bandChunkRows < 0) return fail(ctx, NVTTB_ERR_INVALID_INPUT, "bad band chunk rows");
    if (d->firstFace < 0 || d->lastFace < 0 || (d->lastFace != 0 && d->lastFace < d->firstFace)) return fail(ctx, NVTTB_ERR_INVALID_INPUT, "bad face range");
    if (!nvttb_format_supported(d->encode.format, d->encode.quality)) return fail(ctx, NVTTB_ERR_UNSUPPORTED_FEATURE, "format/quality not implemented");
    *f0 = 0;
    *f1 = d->faceCount;
    if (d->lastFace > d->firstFace) {
        *f0 = d->firstFace;
        *f1 = d->lastFace < d->faceCount ? d->lastFace : d->faceCount;
    }
    for (int f = *f0; f < *f1; f++)
        if (!images[f]) return fail(ctx, NVTTB_ERR_INVALID_INPUT, "missing face image");
    return NVTTB_OK;
}

size_t nvttb_process_whole_output_size(const NvttbProcessDesc *d) {
    if (!d) return 0;
    int f0 = d->firstFace, f1 = d->lastFace;
    if (f0 == 0 && f1 == 0) f1 = d->faceCount;
    return whole_face_bytes(d) * (size_t)(f1 > f0 ? f1 - f0 : 0);
}
int nvttb_device_alloc(NvttbContext *ctx, size_t bytes, void **device_ptr) {
    if (!ctx || !device_ptr) return NVTTB_ERR_INVALID_INPUT;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMalloc(device_ptr, bytes));
    CK(cudaMemset(*device_ptr, 0, bytes));  // exchange buffers start with all flags down
    return NVTTB_OK;
}
int nvttb_device_free(NvttbContext *ctx, void *device_ptr) {
    if (!ctx) return NVTTB_ERR_INVALID_INPUT;
    CK(cudaSetDevice(ctx->device));
    CK(cudaFree(device_ptr));
    return NVTTB_OK;
}
int nvttb_ipc_export(NvttbContext *ctx, void *device_ptr, unsigned char handle[64]) {
    if (!ctx || !device_ptr || !handle) return NVTTB_ERR_INVALID_INPUT;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    CK(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, device_ptr));
    memcpy(handle, &h, 64);
    return NVTTB_OK;
}
int nvttb_ipc_open(NvttbContext *ctx, const unsigned char handle[64], void **device_ptr) {
    if (!ctx || !device_ptr || !handle) return NVTTB_ERR_INVALID_INPUT;
    CK(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CK(cudaIpcOpenMemHandle(device_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return NVTTB_OK;
}
int nvttb_ipc_close(NvttbContext *ctx, void *device_ptr) {
    if (!ctx) return NVTTB_ERR_INVALID_INPUT;
    CK(cudaSetDevice(ctx->device));
    CK(cudaIpcCloseMemHandle(device_ptr));
    return NVTTB_OK;
}
int nvttb_host_register(NvttbContext *ctx, void *host_ptr, size_t bytes) {
    if (!ctx || !host_ptr || !bytes) return NVTTB_ERR_INVALID_INPUT;
    CK(cudaSetDevice(ctx->device));
    CK(cudaHostRegister(host_ptr, bytes, cudaHostRegisterPortable));
    return NVTTB_OK;
}
int nvttb_host_unregister(NvttbContext *ctx, void *host_ptr) {
    if (!ctx || !host_ptr) return NVTTB_ERR_INVALID_INPUT;
    CK(cudaSetDevice(ctx->device));
    CK(cudaHostUnregister(host_ptr));
    return NVTTB_OK;
}

// Asynchronous on the context's stream when the images are on the device.  With HOST images the call returns once the
// texels have been consumed (the caller may free them) - the encoded chain is complete after nvttb_synchronize.
int nvttb_process_to_device(NvttbContext *ctx, const NvttbProcessDesc *d, const void *const *images, int loc, void *out_device,
                            size_t out_capacity, size_t *written) {
    int f0, f1, rc;
    if ((rc = check_process(ctx, d, images, &f0, &f1)) != NVTTB_OK) return rc;
    CK(cudaSetDevice(ctx->device));
    const size_t total = face_bytes(d) * (size_t)(f1 - f0);
    const bool in_place = d->bandCount > 1 && d->bandOutputInPlace;
    if (total == 0 && !in_place) {
        if (written) *written = 0;
        return NVTTB_OK;
    }
    const size_t need = in_place ? whole_face_bytes(d) * (size_t)(f1 - f0) : total;
    if (!out_device || out_capacity < need) return fail(ctx, NVTTB_ERR_INVALID_INPUT, "output buffer too small");
    LocalPlan pl;
    if (in_place && d->bandExchange && f1 - f0 == 1 && shard_local_plan(d, &pl)) {
        if (d->bandSequence == 0) return fail(ctx, NVTTB_ERR_INVALID_INPUT, "bandSequence must start at 1");
        rc = process_shard_local(ctx, d, pl, images[f0], loc, (unsigned char *)out_device, nullptr);
    } else {
        rc = process_faces(ctx, d, images, loc, f0, f1, (unsigned char *)out_device, nullptr, in_place);
    }
    if (rc != NVTTB_OK) return rc;
    if (loc == NVTTB_HOST) {
        CK(cudaStreamSynchronize(ctx->h2d_stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    if (written) *written = total;
    return NVTTB_OK;
}

int nvttb_process_shard(NvttbContext *ctx, const NvttbProcessDesc *d, const void *const *images, int loc, void *out_device, void *out_host) {
    int f0, f1, rc;
    if ((rc = check_process(ctx, d, images, &f0, &f1)) != NVTTB_OK) return rc;
    if (d->bandCount <= 1) return fail(ctx, NVTTB_ERR_INVALID_INPUT, "nvttb_process_shard needs bandCount > 1");
    CK(cudaSetDevice(ctx->device));
    const size_t wbytes = whole_face_bytes(d), need = wbytes * (size_t)(f1 - f0);
    unsigned char *out = (unsigned char *)out_device;
    if (!out) {
        if ((rc = ensure(ctx, ctx->out_dev, need)) != NVTTB_OK) return rc;
        out = (unsigned char *)ctx->out_dev.p;
    }
    LocalPlan pl;
    if (d->bandExchange && f1 - f0 == 1 && shard_local_plan(d, &pl)) {
        if (d->bandSequence == 0) return fail(ctx, NVTTB_ERR_INVALID_INPUT, "bandSequence must start at 1");
        if ((rc = process_shard_local(ctx, d, pl, images[f0], loc, out, (unsigned char *)out_host)) != NVTTB_OK) return rc;
    } else {
        if ((rc = process_faces(ctx, d, images, loc, f0, f1, out, nullptr, true)) != NVTTB_OK) return rc;
        if (out_host) {
            const int mips = nvttb_process_mip_count(d);
            for (int f = f0; f < f1; f++) {
                size_t lo = (size_t)(f - f0) * wbytes;
                int w = d->width, h = d->height;
                for (int m = 0; m < mips; m++) {
                    size_t off, bytes, pitch;
                    int count;
                    nvttb_process_band_slices(d, m, &off, &bytes, &pitch, &count);
                    if (count == 1) CK(cudaMemcpyAsync((char *)out_host + lo + off, out + lo + off, bytes, cudaMemcpyDeviceToHost, ctx->stream));
                    else if (count > 1) CK(cudaMemcpy2DAsync((char *)out_host + lo + off, pitch, out + lo + off, pitch, bytes, count, cudaMemcpyDeviceToHost, ctx->stream));
                    lo += nvttb_level_size(d->encode.format, w, h);
                    w = w / 2 > 1 ? w / 2 : 1;
                    h = h / 2 > 1 ? h / 2 : 1;
                }
            }
        }
    }
    if (out_host || loc == NVTTB_HOST) {
        CK(cudaStreamSynchronize(ctx->h2d_stream));
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaStreamSynchronize(ctx->d2h_stream));
        if (*ctx->h_fault) {
            *ctx->h_fault = 0;
            return fail(ctx, NVTTB_ERR_CUDA, "sharded image: a band never delivered its rows (device-side wait timed out)");
        }
    }
    return NVTTB_OK;
}

static int emit_chain(NvttbContext *ctx, const NvttbProcessDesc *d, int f0, int f1, const unsigned char *p, NvttbEmitFn emit, void *user) {
    const int mips = nvttb_process_mip_count(d);
    for (int f = f0; f < f1; f++) {
        int w = d->width, h = d->height;
        for (int m = 0; m < mips; m++) {
            size_t off, bytes, pitch;
            int count;
            nvttb_process_band_slices(d, m, &off, &bytes, &pitch, &count);
            const size_t sz = bytes * (size_t)count;
            if (sz != 0 && !emit(user, f, m, w, h, 1, p, sz)) return fail(ctx, NVTTB_ERR_FILE_WRITE, "emit callback asked to stop");
            p += sz;
            w = w / 2 > 1 ? w / 2 : 1;
            h = h / 2 > 1 ? h / 2 : 1;
        }
    }
    return NVTTB_OK;
}

int nvttb_process(NvttbContext *ctx, const NvttbProcessDesc *d, const void *const *images, int loc, NvttbEmitFn emit, void *user) {
    int f0, f1, rc;
    if ((rc = check_process(ctx, d, images, &f0, &f1)) != NVTTB_OK) return rc;
    if (!emit) return fail(ctx, NVTTB_ERR_FILE_OPEN, "no output handler");
    // the emit path owns a buffer of this band's bytes only: in-place (whole-chain) output is nvttb_process_to_device / _shard
    if (d->bandCount > 1 && d->bandOutputInPlace) return fail(ctx, NVTTB_ERR_INVALID_INPUT, "bandOutputInPlace needs nvttb_process_to_device or nvttb_process_shard");
    CK(cudaSetDevice(ctx->device));
    const size_t fbytes = face_bytes(d);
    const size_t total = fbytes * (size_t)(f1 - f0);
    if (total == 0) return NVTTB_OK;  // a band that owns no block row of this (small) image emits nothing
    if ((rc = ensure(ctx, ctx->out_dev, total)) != NVTTB_OK) return rc;
    if ((rc = ensure_pinned(ctx, total)) != NVTTB_OK) return rc;
    if ((rc = process_faces(ctx, d, images, loc, f0, f1, (unsigned char *)ctx->out_dev.p, (unsigned char *)ctx->h_out)) != NVTTB_OK) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaStreamSynchronize(ctx->d2h_stream));
    return emit_chain(ctx, d, f0, f1, (const unsigned char *)ctx->h_out, emit, user);
}

// ---- several GPUs, one process ------------------------------------------------------------------------------------------------
// level-0 rows per chunk for n bands: 4 * 2^j with whole chunks, the same number per band, and about four chunks per band
static int auto_chunk_rows(int H, int n) {
    int best = 0;
    for (int C = 4; C <= H; C <<= 1) {
        if (H % C != 0 || (H / C) % n != 0) continue;
        best = C;
        if (H / C / n <= 4) break;
    }
    return best;
}

// Host threads of nvttb_process_multi: keep the thread that feeds GPU `device` (pageable -> pinned staging copies, launches)
// on the CPUs of the NUMA node that GPU hangs off, so that its staging traffic stays on the local memory controller and its
// PCIe root.  Linux sysfs only; silently does nothing when the topology is not exposed (containers, single-node hosts) or
// NVTT_B200_NO_AFFINITY is set.
static void pin_thread_near_device(int device) {
#if defined(__linux__)
    static const bool off = getenv("NVTT_B200_NO_AFFINITY") != nullptr;
    if (off) return;
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) { cudaGetLastError(); return; }
    for (char *c = bus; *c; c++) *c = (char)tolower(*c);
    char path[128];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/local_cpulist", bus);
    FILE *f = fopen(path, "r");
    if (!f) return;
    char list[1024] = {0};
    const bool ok = fgets(list, sizeof list, f) != nullptr;
    fclose(f);
    if (!ok) return;
    cpu_set_t set;
    CPU_ZERO(&set);
    int count = 0;
    char *save = nullptr;  // strtok_r: this runs in one host thread per GPU at the same time
    for (char *tok = strtok_r(list, ",\n", &save); tok; tok = strtok_r(nullptr, ",\n", &save)) {  // "0-31,64-95"
        int a = 0, b = 0;
        const int got = sscanf(tok, "%d-%d", &a, &b);
        if (got == 1) b = a;
        if (got < 1 || a < 0 || b < a) continue;
        for (int c = a; c <= b && c < CPU_SETSIZE; c++) { CPU_SET(c, &set); count++; }
    }
    if (count > 0) pthread_setaffinity_np(pthread_self(), sizeof set, &set);  // failure (cgroup limits): stay where we are
#else
    (void)device;
#endif
}

int nvttb_bind_thread_to_device(NvttbContext *ctx) {
    if (!ctx) return NVTTB_ERR_INVALID_INPUT;
    pin_thread_near_device(ctx->device);
    return NVTTB_OK;
}

int nvttb_process_multi(NvttbContext *const *ctxs, int n, const NvttbProcessDesc *d, const void *const *images, NvttbEmitFn emit, void *user) {
    if (!ctxs || n <= 0 || !ctxs[0]) return NVTTB_ERR_INVALID_INPUT;
    NvttbContext *ctx = ctxs[0];
    for (int i = 1; i < n; i++)
        if (!ctxs[i]) return fail(ctx, NVTTB_ERR_INVALID_INPUT, "null context");
    if (n == 1 || !d || d->bandCount > 1) return nvttb_process(ctx, d, images, NVTTB_HOST, emit, user);
    int f0, f1, rc;
    if ((rc = check_process(ctx, d, images, &f0, &f1)) != NVTTB_OK) return rc;
    if (!emit) return fail(ctx, NVTTB_ERR_FILE_OPEN, "no output handler");
    const int faces = f1 - f0;
    const size_t wbytes = whole_face_bytes(d), total = wbytes * (size_t)faces;
    CK(cudaSetDevice(ctx->device));
    if ((rc = ensure_pinned(ctx, total)) != NVTTB_OK) return rc;
    unsigned char *const h_out = (unsigned char *)ctx->h_out;
    std::vector<int> rcs(n, NVTTB_OK);
    if (faces >= n || (size_t)d->width * d->height < ((size_t)1 << 20)) {
        // independent faces / array slices (or images too small to cut): deal them out, each GPU copies its chains home
        const int users = faces < n ? faces : n;
        std::vector<std::thread> th;
        for (int t = 0; t < users; t++) {
            th.emplace_back([&, t]() {
                NvttbContext *c = ctxs[t];
                const int lo = f0 + (int)((long long)faces * t / users), hi = f0 + (int)((long long)faces * (t + 1) / users);
                if (hi <= lo) return;
                pin_thread_near_device(c->device);
                int r = cudaSetDevice(c->device) == cudaSuccess ? NVTTB_OK : NVTTB_ERR_CUDA;
                if (r == NVTTB_OK) r = ensure(c, c->out_dev, wbytes * (size_t)(hi - lo));
                if (r == NVTTB_OK) r = process_faces(c, d, images, NVTTB_HOST, lo, hi, (unsigned char *)c->out_dev.p, h_out + (size_t)(lo - f0) * wbytes);
                if (r == NVTTB_OK && (cudaStreamSynchronize(c->stream) != cudaSuccess || cudaStreamSynchronize(c->d2h_stream) != cudaSuccess)) r = NVTTB_ERR_CUDA;
                rcs[t] = r;
            });
        }
        for (auto &x : th) x.join();
    } else {
        // one large image at a time, block-row sharded over the largest number of GPUs that divides it evenly
        int bands = n, C = 0;
        for (; bands > 1; bands--)
            if ((C = auto_chunk_rows(d->height, bands)) != 0) break;
        if (bands <= 1) return nvttb_process(ctx, d, images, NVTTB_HOST, emit, user);
        NvttbProcessDesc base = *d;
        base.bandCount = bands;
        base.bandChunkRows = C;
        base.bandOutputInPlace = 1;
        base.bandExchange = nullptr;
        // band-local front end: needs the exchange buffer on GPU 0 and peer access to it from the others
        bool peers = true;
        for (int t = 1; t < bands && peers; t++) {
            int can = 0;
            // Two bands on ONE GPU: the caller's images are usually pageable, a pageable copy is synchronous in the driver and
            // (measured) waits while band 0's device-side wait is resident - the band would never deliver.  Replicated front end.
            for (int u = 0; u < t; u++)
                if (ctxs[u]->device == ctxs[t]->device) peers = false;
            if (!peers) break;
            if (cudaDeviceCanAccessPeer(&can, ctxs[t]->device, ctx->device) != cudaSuccess || !can) peers = false;
            if (peers) {
                cudaSetDevice(ctxs[t]->device);
                const cudaError_t pe = cudaDeviceEnablePeerAccess(ctx->device, 0);
                if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) peers = false;
                cudaGetLastError();
            }
        }
        CK(cudaSetDevice(ctx->device));
        const size_t xbytes = nvttb_process_exchange_size(&base);
        if (peers && xbytes) {
            if (ctx->exchange.cap < xbytes) {
                if ((rc = ensure(ctx, ctx->exchange, xbytes)) != NVTTB_OK) return rc;
                CK(cudaMemset(ctx->exchange.p, 0, xbytes));
                ctx->multi_seq = 0;
            }
            base.bandExchange = ctx->exchange.p;
        }
        if (base.bandExchange) {
            for (int t = 0; t < bands; t++) {
                NvttbProcessDesc mine = base;
                mine.bandIndex = t;
                if ((rc = nvttb_process_prepare(ctxs[t], &mine, NVTTB_HOST, 1)) != NVTTB_OK) return rc;
            }
        }
        for (int f = f0; f < f1; f++) {
            base.firstFace = f;
            base.lastFace = f + 1;
            base.bandSequence = ++ctx->multi_seq;
            std::vector<std::thread> th;
            for (int t = 0; t < bands; t++) {
                th.emplace_back([&, t]() {
                    NvttbProcessDesc mine = base;
                    mine.bandIndex = t;
                    pin_thread_near_device(ctxs[t]->device);
                    rcs[t] = nvttb_process_shard(ctxs[t], &mine, images, NVTTB_HOST, nullptr, h_out + (size_t)(f - f0) * wbytes);
                });
            }
            for (auto &x : th) x.join();
            for (int t = 0; t < bands; t++)
                if (rcs[t] != NVTTB_OK) break;
        }
    }
    for (int t = 0; t < n; t++)
        if (rcs[t] != NVTTB_OK) {
            if (t != 0) ctx->err = std::string("GPU ") + std::to_string(ctxs[t]->device) + ": " + ctxs[t]->err;
            return rcs[t];
        }
    return emit_chain(ctx, d, f0, f1, h_out, emit, user);
}

}  // extern "C"
