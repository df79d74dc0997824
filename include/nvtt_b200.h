/*
 * nvtt_b200 — C ABI of the B200-native (sm_100a) block-compression + mip-generation path.
 *
 * This is the drop-in boundary: plain C, plain pointers and sizes.  Each entry point names the interface of the
 * reference (castano/nvidia-texture-tools 2.1.2) that it replaces; enum values are the reference's own
 * (nvtt::Format, nvtt::Quality, ... src/nvtt/nvtt.h:80-277) so a caller can pass them through unchanged.
 * The C++ mirror of nvtt::Compressor / CompressionOptions / InputOptions / OutputOptions lives in
 * nvidia-texture-tools_b200/host/ and is implemented on top of these functions only.
 *
 * There is NO CPU fallback: every function that computes needs a CUDA device and fails with
 * NVTTB_ERR_CUDA (nvtt::Error_CudaError) otherwise.
 */
#ifndef NVTT_B200_H
#define NVTT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define NVTTB_API __declspec(dllexport)
#else
#define NVTTB_API __attribute__((visibility("default")))
#endif

/* Return codes: 0 = success, otherwise 1 + nvtt::Error (src/nvtt/nvtt.h:345-356). */
enum {
    NVTTB_OK = 0,
    NVTTB_ERR_UNKNOWN = 1,
    NVTTB_ERR_INVALID_INPUT = 2,
    NVTTB_ERR_UNSUPPORTED_FEATURE = 3,
    NVTTB_ERR_CUDA = 4,
    NVTTB_ERR_FILE_OPEN = 5,
    NVTTB_ERR_FILE_WRITE = 6,
    NVTTB_ERR_UNSUPPORTED_OUTPUT_FORMAT = 7
};

/* Where a buffer lives. */
enum { NVTTB_HOST = 0, NVTTB_DEVICE = 1 };

typedef struct NvttbContext NvttbContext; /* one per GPU: stream, lookup tables, scratch (replaces nvtt::Compressor::Private state, src/nvtt/Context.h) */
typedef struct NvttbSurface NvttbSurface; /* device-resident planar fp32 RGBA image (replaces nvtt::Surface, src/nvtt/Surface.h:37-75) */

/* ---- context -------------------------------------------------------------------------------------------- */
/* Number of CUDA devices visible (0 if none / no driver).  Replaces nv::cuda::isHardwarePresent (src/nvtt/cuda/CudaUtils.cpp). */
NVTTB_API int nvttb_device_count(void);
/* Which arithmetic contract this build of the library follows: "strict" (libnvtt_b200.so: nvcc -fmad=false, every fp32
 * operation rounded like the reference's -ffp-contract=off build - the parity contract) or "fastmath"
 * (libnvtt_b200_fastmath.so: the same sources with FMA contraction allowed - faster, NOT bit-identical to the reference;
 * SURVEY.md 7.2 item 7).  Same symbols in both; a caller picks one at link / dlopen time. */
NVTTB_API const char *nvttb_build_variant(void);
/* Create a context on `device`.  Replaces Compressor::Compressor + enableCudaAcceleration (src/nvtt/Context.cpp:61-98). */
NVTTB_API int nvttb_context_create(int device, NvttbContext **out);
NVTTB_API void nvttb_context_destroy(NvttbContext *ctx);
/* Human-readable description of the last failure on this context (never NULL). */
NVTTB_API const char *nvttb_last_error(const NvttbContext *ctx);
/* Kernel launches issued by this context so far (for bench.py's gpu_launches). */
NVTTB_API uint64_t nvttb_launch_count(const NvttbContext *ctx);
/* Block until all work queued on the context's stream is done. */
NVTTB_API int nvttb_synchronize(NvttbContext *ctx);
/* The context's cudaStream_t (as void*), so a caller can record CUDA events on the launching stream. */
NVTTB_API void *nvttb_stream(NvttbContext *ctx);

/* Device-side timing on the context's stream (CUDA events), for bench.py: start / stop+elapsed in ms. */
NVTTB_API int nvttb_timer_start(NvttbContext *ctx);
NVTTB_API int nvttb_timer_stop(NvttbContext *ctx, float *elapsed_ms);
/* Per-kernel CUDA-event timing of everything launched between begin and end (bench.py's roofline leg).
 * units = texels the launches covered.  Replaces nothing in the reference (it only has nv::Timer, src/nvcore/Timer.cpp). */
typedef struct NvttbKernelStat {
    const char *name;
    int launches;
    double total_ms, max_ms; /* summed / longest single launch */
    double total_units, max_units; /* texels covered by all launches / by the longest one */
} NvttbKernelStat;
NVTTB_API int nvttb_profile_begin(NvttbContext *ctx);
NVTTB_API int nvttb_profile_end(NvttbContext *ctx, NvttbKernelStat *stats, int max_stats, int *count);

/* ---- one mip level: the nv::CompressorInterface::compress seam ------------------------------------------ */
/* Replaces Compressor::Private::compress(AlphaMode,w,h,d,face,mip,const float*,...) + chooseCpuCompressor
 * (src/nvtt/Context.cpp:486-516,1038-1163; src/nvtt/Compressor.h:34-38). */
typedef struct NvttbEncodeDesc {
    int format;            /* nvtt::Format */
    int quality;           /* nvtt::Quality */
    int alphaMode;         /* nvtt::AlphaMode */
    int pixelType;         /* nvtt::PixelType (BC6: UnsignedFloat / Float) */
    float colorWeights[4]; /* CompressionOptions::setColorWeights */
    int width, height;     /* texels; depth is 1 */
    int applyToGamma;      /* 1: fuse Surface::toGamma(2.2) on R,G,B into the block gather (pipeline use) */
    float rgbmThreshold;   /* CompressionOptions::setRGBMThreshold (Format_BC3_RGBM; the reference's default is 0.15) */
} NvttbEncodeDesc;

/* Bytes of one encoded level = blocks * block size (nv::computeImageSize, src/nvtt/Surface.cpp:210-218); 0 if unsupported. */
NVTTB_API size_t nvttb_level_size(int format, int width, int height);
/* 1 if (format, quality) is implemented by this library. */
NVTTB_API int nvttb_format_supported(int format, int quality);
/* rgba: planar fp32 [4][h][w] (FloatImage layout) on host or device; out: BCn bytes, host or device.
 * Asynchronous when both buffers are on the device; synchronous otherwise. */
NVTTB_API int nvttb_encode_level(NvttbContext *ctx, const NvttbEncodeDesc *desc, const float *rgba, int rgba_location,
                                 void *out, int out_location, size_t out_capacity);

/* ---- Format_RGB / Format_RGBA: uncompressed pixel formats -------------------------------------------------
 * Replaces PixelFormatConverter::compress (src/nvtt/CompressorRGB.cpp:410-575) and, for the size, nv::computeImageSize's
 * Format_RGBA branch (src/nvtt/Surface.cpp:210-214).  The fields are CompressionOptions::Private's after
 * setPixelFormat(bitcount, masks) / setPixelFormat(rsize, gsize, bsize, asize) / setPixelType / setPitchAlignment
 * (src/nvtt/CompressionOptions.cpp:113-173): bitcount != 0 selects the mask form.  PixelType_SharedExp 9/9/9/5 (RGB9E5)
 * is not implemented (size 0 / NVTTB_ERR_UNSUPPORTED_FEATURE); the signed types write zeros, as the reference does. */
typedef struct NvttbPixelFormatDesc {
    int pixelType;                        /* nvtt::PixelType */
    unsigned bitcount;                    /* mask form: bits per pixel (<= 32); 0 = size form */
    unsigned rmask, gmask, bmask, amask;  /* mask form */
    unsigned rsize, gsize, bsize, asize;  /* size form (and the float channel widths: 0, 10, 11, 16 or 32) */
    int pitchAlignment;                   /* bytes, power of two (default 1) */
    int width, height;                    /* texels; depth is 1 */
} NvttbPixelFormatDesc;
/* Bytes of one level = height * computeBytePitch(width, bits, pitchAlignment); 0 if the description is unsupported. */
NVTTB_API size_t nvttb_pixel_format_level_size(const NvttbPixelFormatDesc *desc);
/* rgba: planar fp32 [4][h][w] on host or device; out: height scanlines of the pitch above, host or device. */
NVTTB_API int nvttb_convert_level(NvttbContext *ctx, const NvttbPixelFormatDesc *desc, const float *rgba, int rgba_location,
                                  void *out, int out_location, size_t out_capacity);

/* ---- Surface ops on the device (the image-op seam called from src/nvtt/Context.cpp:267-343) ------------- */
NVTTB_API int nvttb_surface_create(NvttbContext *ctx, NvttbSurface **out);
NVTTB_API void nvttb_surface_destroy(NvttbSurface *s);
NVTTB_API int nvttb_surface_clone(const NvttbSurface *s, NvttbSurface **out);       /* Surface copy (COW in the reference) */
NVTTB_API void nvttb_surface_set_wrap_mode(NvttbSurface *s, int wrapMode);          /* Surface::setWrapMode */
NVTTB_API void nvttb_surface_set_alpha_mode(NvttbSurface *s, int alphaMode);        /* Surface::setAlphaMode */
NVTTB_API void nvttb_surface_set_normal_map(NvttbSurface *s, int isNormalMap);      /* Surface::setNormalMap */
NVTTB_API int nvttb_surface_width(const NvttbSurface *s);
NVTTB_API int nvttb_surface_height(const NvttbSurface *s);
/* Surface::setImage(InputFormat,w,h,1,data)  src/nvtt/Surface.cpp:728-815 */
NVTTB_API int nvttb_surface_set_image(NvttbSurface *s, int inputFormat, int w, int h, const void *data, int location);
/* Surface::toLinear / toGamma  src/nvtt/Surface.cpp:1470-1488 */
NVTTB_API int nvttb_surface_to_linear(NvttbSurface *s, float gamma);
NVTTB_API int nvttb_surface_to_gamma(NvttbSurface *s, float gamma);
/* Surface::buildNextMipmap(filter[,filterWidth,params])  src/nvtt/Surface.cpp:1344-1406.
 * params may be NULL (defaults: Box 0.5, Triangle 1.0, Kaiser 3.0/alpha 4/stretch 1).  *built = 0 when the surface is already 1x1. */
NVTTB_API int nvttb_surface_build_next_mipmap(NvttbSurface *s, int mipmapFilter, int useParams, float filterWidth,
                                              const float *params, int *built);
/* Surface::resize(w,h,1,filter[,filterWidth,params])  src/nvtt/Surface.cpp:1152-1219 */
NVTTB_API int nvttb_surface_resize(NvttbSurface *s, int w, int h, int resizeFilter, int useParams, float filterWidth,
                                   const float *params);
/* Surface::expandNormals / normalizeNormalMap / packNormals  src/nvtt/Surface.cpp:2810-2817,2952-2963 */
NVTTB_API int nvttb_surface_expand_normals(NvttbSurface *s);
NVTTB_API int nvttb_surface_normalize_normal_map(NvttbSurface *s);
NVTTB_API int nvttb_surface_pack_normals(NvttbSurface *s);
/* Surface::scaleBias(channel, scale, bias) for `count` consecutive channels -> FloatImage::scaleBias (src/nvtt/Surface.cpp:1669-1677,
 * src/nvimage/FloatImage.cpp:231-242); packNormals / expandNormals with caller-chosen values are scaleBias(0, 3, ...). */
NVTTB_API int nvttb_surface_scale_bias(NvttbSurface *s, int channel, int count, float scale, float bias);
/* Surface::clamp(channel, low, high)  src/nvtt/Surface.cpp:1679-1686 */
NVTTB_API int nvttb_surface_clamp(NvttbSurface *s, int channel, float low, float high);
/* Surface::range(channel, &min, &max, alpha_channel, alpha_ref)  src/nvtt/Surface.cpp:526-566 (alpha_channel < 0: no alpha test) */
NVTTB_API int nvttb_surface_range(const NvttbSurface *s, int channel, int alpha_channel, float alpha_ref, float *range_min, float *range_max);
/* Surface::toneMap(ToneMapper, params)  src/nvtt/Surface.cpp:2444-2494 (0 Linear, 1 Reindhart, 2 Halo, 3 Lightmap) */
NVTTB_API int nvttb_surface_tone_map(NvttbSurface *s, int toneMapper);
/* Surface::toRGBM(range, threshold)  src/nvtt/Surface.cpp:1862-1946 */
NVTTB_API int nvttb_surface_to_rgbm(NvttbSurface *s, float range, float threshold);
/* Surface::binarize(channel, threshold, dither) and Surface::quantize(channel, bits, exactEndPoints, dither)
 * src/nvtt/Surface.cpp:2656-2775.  dither != 0 = the reference's Floyd-Steinberg scan (bit-identical: a skewed wavefront on
 * one SM per plane, so it is slow next to everything else here - use it the way the reference does, for final output). */
NVTTB_API int nvttb_surface_binarize(NvttbSurface *s, int channel, float threshold, int dither);
NVTTB_API int nvttb_surface_quantize(NvttbSurface *s, int channel, int bits, int exactEndPoints, int dither);
/* Surface::toGreyScale / toNormalMap  src/nvtt/Surface.cpp:1732-1756,2794-2808 */
NVTTB_API int nvttb_surface_to_grey_scale(NvttbSurface *s, float r, float g, float b, float a);
NVTTB_API int nvttb_surface_to_normal_map(NvttbSurface *s, float sm, float medium, float big, float large);
/* Surface::setImage2D(format, decoder, w, h, data): decode a BCn level into the surface  src/nvtt/Surface.cpp:908-1118.
 * format: BC1, BC2, BC3, BC3n, BC3_RGBM, BC4, BC5, BC6, BC7 (nvtt::Format values); decoder: nvtt::Decoder (D3D10 / D3D9 / NV5x).
 * bc6Signed: state of the reference's global ZOH::Utils::FORMAT at decode time (0 = unsigned, its value unless a signed
 * BC6 encode ran before). */
NVTTB_API int nvttb_surface_set_image_2d(NvttbSurface *s, int format, int decoder, int w, int h, const void *data, int location, int bc6Signed);
/* nvtt::rmsError(reference, img) / nvtt::rmsAlphaError  src/nvtt/Surface.cpp:3270-3279 -> nv::rmsColorError / rmsAlphaError
 * src/nvimage/ErrorMetric.cpp:13-73.  The colour error is alpha-weighted when the reference surface's alpha mode is
 * AlphaMode_Transparency.  *out = FLT_MAX when the layouts differ (as in the reference). */
NVTTB_API int nvttb_rms_error(const NvttbSurface *reference, const NvttbSurface *img, float *out);
NVTTB_API int nvttb_rms_alpha_error(const NvttbSurface *reference, const NvttbSurface *img, float *out);
/* nvtt::angularError = nv::rmsAngularError (src/nvtt/Surface.cpp:3287-3291, src/nvimage/ErrorMetric.cpp:475-511): RMS angle
 * between the unpacked, normalised normals, in radians (acosf: CUDA vs glibc, 1e-5 relative). */
NVTTB_API int nvttb_angular_error(const NvttbSurface *reference, const NvttbSurface *img, float *out);
/* nvtt::cieLabError = nv::cieLabError (src/nvtt/Surface.cpp:3282-3285, src/nvimage/ErrorMetric.cpp:192-342): mean length of the
 * CIE-Lab difference; built on powf, so CUDA and glibc agree to ~1e-4 relative, not bit for bit. */
NVTTB_API int nvttb_cielab_error(const NvttbSurface *reference, const NvttbSurface *img, float *out);
/* Surface::data(): copy planar fp32 RGBA (4*w*h floats) to the host. */
NVTTB_API int nvttb_surface_download(const NvttbSurface *s, float *out);
/* Device pointer of the planar fp32 data (valid until the next op on the surface). */
NVTTB_API const float *nvttb_surface_device_data(const NvttbSurface *s);
/* Compressor::compress(Surface, face, mip, ...)  src/nvtt/Context.cpp:146-149,477-484.  desc->width/height are ignored. */
NVTTB_API int nvttb_surface_encode(NvttbSurface *s, const NvttbEncodeDesc *desc, void *out, int out_location, size_t out_capacity);

/* ---- the whole InputOptions pipeline on the device ------------------------------------------------------ */
/* Replaces Compressor::process -> Compressor::Private::compress(InputOptions...) (src/nvtt/Context.cpp:117-120,217-346):
 * per face: setImage -> [toGreyScale+toNormalMap] -> toLinear -> level 0 -> { buildNextMipmap -> [renormalise] ->
 * toGamma -> compress } per level.  The DDS/KTX header is written by the C++ host layer, not here. */
typedef struct NvttbProcessDesc {
    int inputFormat;   /* nvtt::InputFormat */
    int width, height; /* level-0 extent of every face */
    int faceCount;     /* 1 (2D), 6 (cube) or array size */
    int wrapMode;      /* nvtt::WrapMode */
    int mipmapFilter;  /* nvtt::MipmapFilter */
    int generateMipmaps;
    int maxLevel;      /* <= 0: full chain */
    float kaiserWidth, kaiserAlpha, kaiserStretch;
    float inputGamma, outputGamma;
    int isNormalMap, convertToNormalMap, normalizeMipmaps;
    float heightFactors[4];      /* InputOptions::setHeightEvaluation */
    float bumpFrequencyScale[4]; /* InputOptions::setNormalFilter */
    int alphaMode;     /* nvtt::AlphaMode */
    NvttbEncodeDesc encode; /* format, quality, colour weights, pixel type (width/height/applyToGamma ignored) */
    int firstFace, lastFace;   /* process faces [firstFace, lastFace); 0,0 = all (multi-GPU sharding by face) */
    /* Block-row sharding of ONE image over bandCount GPUs (bandCount <= 1: off).  Level 0 is cut into chunks of
     * bandChunkRows texel rows (a multiple of 4; 0 = height / bandCount, i.e. one contiguous band per GPU) and chunk c belongs
     * to band c % bandCount (cyclic: content of different cost is spread over the GPUs).  A mip level is distributed while a
     * chunk still covers whole block rows of it (bandChunkRows % (4 << level) == 0); the remaining small levels ("the tail")
     * are encoded by band 0 alone.  nvttb_process_band_slices is the layout contract.  The call produces, per face, the
     * concatenation of this band's slices level by level; emit is called once per level with them. */
    int bandIndex, bandCount;
    /* nvttb_process_to_device / nvttb_process_shard only, with bandCount > 1: out_device is the WHOLE encoded chain of the
     * processed faces (as one GPU would produce it) and this band stores its slices at their final offsets.  The buffer may live
     * on another GPU (peer memory, see nvttb_ipc_*): the encoder's stores then go over NVLink and no gather is needed. */
    int bandOutputInPlace;
    int bandChunkRows;
    /* Band-local front end (Box mip filter, 2.2 / 1.0 gammas, one face per call): with bandExchange != NULL a band uploads,
     * converts and down-samples ONLY its own chunks; the rows of the last distributed level travel to band 0 through this
     * buffer (device memory on band 0's GPU, nvttb_process_exchange_size bytes, zero-filled once; the other bands map it as
     * peer memory) with release/acquire flags over NVLink, and band 0 runs the tail from it on a second stream.
     * bandSequence: 1, 2, 3, ... - the same number on every band for one image, one more for the next image that uses the
     * same exchange buffer.  With bandExchange == NULL every band builds the whole fp32 chain (any filter, no exchange). */
    void *bandExchange;
    unsigned bandSequence;
} NvttbProcessDesc;

/* Called once per (face, mip) in the reference's order (face-major, mip-minor); data is host memory owned by the
 * library and valid only during the call — exactly like OutputHandler::writeData (src/nvtt/nvtt.h:330-342).
 * Return 0 to stop (maps to Error_FileWrite). */
typedef int (*NvttbEmitFn)(void *user, int face, int mip, int width, int height, int depth, const void *data, size_t size);

/* images[f] = level-0 texels of face f in inputFormat, on host or device. */
NVTTB_API int nvttb_process(NvttbContext *ctx, const NvttbProcessDesc *desc, const void *const *images, int images_location,
                            NvttbEmitFn emit, void *user);
/* Same pipeline, but the encoded chain of every processed face stays in one caller-provided DEVICE buffer
 * (face-major, mip-minor, tightly packed); nothing is copied to the host.  *written = bytes produced. */
NVTTB_API int nvttb_process_to_device(NvttbContext *ctx, const NvttbProcessDesc *desc, const void *const *images,
                                      int images_location, void *out_device, size_t out_capacity, size_t *written);
/* Total bytes nvttb_process emits for this description (Compressor::estimateSize, src/nvtt/Context.cpp:122-137). */
NVTTB_API size_t nvttb_process_output_size(const NvttbProcessDesc *desc);
/* Where band `desc->bandIndex` of `desc->bandCount` sits inside mip level `level`: `*count` slices of `*bytes` bytes each, the
 * first at byte `*offset` of the level and the following ones `*pitch` bytes apart (*count = 0: this band emits nothing for the
 * level).  With bandCount <= 1 the one slice is the level.  Replaces nothing in the reference (it has no multi-device path);
 * it is the contract between the GPUs of a box for assembling one image. */
NVTTB_API int nvttb_process_band_slices(const NvttbProcessDesc *desc, int level, size_t *offset, size_t *bytes, size_t *pitch, int *count);
/* Bytes of the exchange buffer of the band-local front end (0: the description does not qualify and bandExchange is ignored). */
NVTTB_API size_t nvttb_process_exchange_size(const NvttbProcessDesc *desc);
/* One band's share of ONE block-row sharded image (desc->bandCount > 1).  out_device: whole-chain layout (this GPU's or a
 * peer's memory; NULL = an internal buffer); out_host (optional): pinned / registered host memory in the whole-chain layout -
 * the band's slices are also copied there (each GPU writes its slice of the output over its own PCIe link).  Host images may
 * be pageable or pinned - but when several bands share ONE GPU (tests) use pinned memory: a pageable copy is synchronous inside
 * the driver and was measured to wait while band 0's device-side wait kernel is resident, so the band would never deliver.
 * Synchronous when out_host or a host image is given, asynchronous on the context's stream otherwise. */
NVTTB_API int nvttb_process_shard(NvttbContext *ctx, const NvttbProcessDesc *desc, const void *const *images, int images_location,
                                  void *out_device, void *out_host);
/* Sizes every device buffer a band-local nvttb_process_shard / nvttb_process_to_device call with this description will use, so
 * that the call itself allocates nothing.  Needed when several bands share one GPU (band 0 keeps a waiting kernel resident
 * until all bands have delivered, and a cudaMalloc on that GPU may wait for it); harmless otherwise.  own_output: the call will
 * be given out_device = NULL. */
NVTTB_API int nvttb_process_prepare(NvttbContext *ctx, const NvttbProcessDesc *desc, int images_location, int own_output);
/* Restrict the CALLING host thread to the CPUs of the NUMA node the context's GPU is attached to (sysfs local_cpulist), so
 * that the staging copies and launches of a one-thread-per-GPU / one-process-per-GPU caller stay on the local memory controller
 * and PCIe root.  nvttb_process_multi does this for its own threads.  No-op where the topology is not exposed or when
 * NVTT_B200_NO_AFFINITY is set.  The reference has no counterpart (its ThreadPool, src/nvthread/ThreadPool.cpp, is not NUMA aware). */
NVTTB_API int nvttb_bind_thread_to_device(NvttbContext *ctx);
/* The whole pipeline for host images on SEVERAL GPUs of one process (one host thread per context): large single images are
 * block-row sharded (band-local front end where it applies), cube faces / array slices are dealt out face by face.  emit sees
 * exactly what nvttb_process would produce on one GPU.  contexts[0] owns the pinned output buffer. */
NVTTB_API int nvttb_process_multi(NvttbContext *const *contexts, int context_count, const NvttbProcessDesc *desc,
                                  const void *const *images, NvttbEmitFn emit, void *user);
/* cudaHostRegister / cudaHostUnregister (portable): lets several processes (one per GPU) DMA into one shared host buffer. */
NVTTB_API int nvttb_host_register(NvttbContext *ctx, void *host_ptr, size_t bytes);
NVTTB_API int nvttb_host_unregister(NvttbContext *ctx, void *host_ptr);
/* ---- one output buffer shared by the GPUs of a box (block-row sharding of one image) ------------------------------------
 * The owner allocates the chain buffer and exports it; the other ranks (processes) open it and pass the pointer to
 * nvttb_process_to_device with bandOutputInPlace = 1.  Thin wrappers over cudaMalloc / cudaIpcGetMemHandle /
 * cudaIpcOpenMemHandle(lazy peer access) / cudaIpcCloseMemHandle; the 64 handle bytes travel over any transport.
 * Replaces nothing in the reference (it has no multi-device path). */
NVTTB_API int nvttb_device_alloc(NvttbContext *ctx, size_t bytes, void **device_ptr);
NVTTB_API int nvttb_device_free(NvttbContext *ctx, void *device_ptr);
NVTTB_API int nvttb_ipc_export(NvttbContext *ctx, void *device_ptr, unsigned char handle[64]);
NVTTB_API int nvttb_ipc_open(NvttbContext *ctx, const unsigned char handle[64], void **device_ptr);
NVTTB_API int nvttb_ipc_close(NvttbContext *ctx, void *device_ptr);
/* Bytes of the whole chain of the processed faces, ignoring the band fields (the size of the shared buffer above). */
NVTTB_API size_t nvttb_process_whole_output_size(const NvttbProcessDesc *desc);

/* ---- DDS / DDS10 container reader (host only): pre-made mip chains as pipeline input ------------------------------------------
 * Replaces nv::DirectDrawSurface::load / isValid / isSupported / mipmapCount / surfaceSize / offset (src/nvimage/DirectDrawSurface.cpp
 * :983-1100,1143-1322) for a file image in memory.  blockFormat: nvtt::Format of a BCn file (decode it with
 * nvttb_surface_set_image_2d), else -1; inputFormat: nvtt::InputFormat when the surfaces can be handed to nvttb_process /
 * InputOptions::setMipmapData as they are (B8G8R8A8, RGBA16F, RGBA32F, R32F), else -1. */
typedef struct NvttbDdsInfo {
    int width, height, depth;
    int mipCount, faceCount, arraySize;  /* faceCount: 6 for a cube map, else 1 */
    int textureType;                     /* nvtt::TextureType */
    int blockFormat, inputFormat;
    unsigned bitsPerPixel, blockBytes, headerBytes;
    unsigned dxgiFormat, fourcc;
    int hasAlpha, isNormalMap;
} NvttbDdsInfo;
/* NVTTB_ERR_INVALID_INPUT: not a DDS file / truncated; NVTTB_ERR_UNSUPPORTED_FEATURE: a layout the reference's reader rejects too */
NVTTB_API int nvttb_dds_describe(const void *file, size_t bytes, NvttbDdsInfo *info);
/* Byte offset and size of surface (face, mip) inside the file (faces of an array: face = slice * faceCount + cube face) */
NVTTB_API int nvttb_dds_surface(const NvttbDdsInfo *info, int face, int mip, size_t *offset, size_t *bytes, int *width, int *height, int *depth);

/* Number of mip levels the pipeline produces (nv::countMipmaps, src/nvtt/Surface.cpp:181-193, capped by maxLevel). */
NVTTB_API int nvttb_process_mip_count(const NvttbProcessDesc *desc);

#ifdef __cplusplus
}
#endif
#endif /* NVTT_B200_H */
