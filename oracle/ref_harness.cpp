// TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// C-ABI harness around the UNMODIFIED reference library (NVTT 2.1.2, /root/reference), compiled by
// oracle/build_ref.sh into oracle/_ref/libnvtt_ref.so.  It only *calls* the reference's public API
// (src/nvtt/nvtt.h) — no reference source is copied here.  The same file, compiled unchanged against the drop-in header
// nvidia-texture-tools_b200/host/nvtt/nvtt.h (tests/build_host_harness.sh), drives OUR library
// through the identical nvtt:: calls — that is the source-compatibility test of the boundary.  The tests use it (1) to pin the plain-C
// restatement in oracle/*.c and (2) as the checker for the CUDA path; bench.py uses it as the
// `--impl reference` arm / cpu_baseline ("kind": "reference").
#include <nvtt/nvtt.h>
#include <string.h>
#include <stdlib.h>
#include <vector>

using namespace nvtt;

#ifdef REF_ZERO_NEW
// Pinned parity build only: OptimalCompress::compressDXT5A reads the not-yet-written output block
// (OptimalCompressDXT.cpp:546; the buffer is a bare `new uint8[]`, BlockCompressor.cpp:106), so the reference's BC4/BC5
// Production output depends on heap garbage.  Replacing the global allocation functions *of this .so* (linked
// -Bsymbolic) with zero-filling ones pins that read to "zero-filled output buffer" without touching reference sources.
#include <new>
void *operator new(size_t n) { void *p = calloc(n ? n : 1, 1); if (!p) throw std::bad_alloc(); return p; }
void *operator new[](size_t n) { void *p = calloc(n ? n : 1, 1); if (!p) throw std::bad_alloc(); return p; }
void operator delete(void *p) noexcept { free(p); }
void operator delete[](void *p) noexcept { free(p); }
void operator delete(void *p, size_t) noexcept { free(p); }
void operator delete[](void *p, size_t) noexcept { free(p); }
#endif

namespace {
struct MemHandler : public OutputHandler {
    std::vector<unsigned char> buf;
    int images = 0;
    virtual void beginImage(int, int, int, int, int, int) { images++; }
    virtual bool writeData(const void *data, int size) {
        const unsigned char *p = (const unsigned char *)data;
        buf.insert(buf.end(), p, p + size);
        return true;
    }
    virtual void endImage() {}
};
struct ErrCount : public ErrorHandler {
    int n = 0, last = -1;
    virtual void error(Error e) { n++; last = (int)e; }
};

struct SeqDispatcher : public TaskDispatcher {
    virtual void dispatch(Task *task, void *context, int count) {
        for (int i = 0; i < count; i++) task(context, i);
    }
};

void setup_co(CompressionOptions &co, int format, int quality, const float *cw, int pixelType) {
    co.setFormat((Format)format);
    co.setQuality((Quality)quality);
    if (cw) co.setColorWeights(cw[0], cw[1], cw[2], cw[3]);
    co.setPixelType((PixelType)pixelType);
}
}  // namespace

extern "C" {

// One mip level through Compressor::compress(w,h,d,face,mip,rgba,...) (src/nvtt/Context.cpp:187-190,486-516).
// planar fp32 RGBA [c][y][x].  threads: 0 = reference default (nvthread pool), 1 = sequential dispatcher.
// alphaMode != None goes through the Surface API so that the alpha mode reaches the compressor.
// Returns bytes written (<= out_cap) or -1.
long ref_compress_level(int format, int quality, int alphaMode, int w, int h, const float *rgba,
                        const float *colorWeights4, int pixelType, int threads, unsigned char *out, long out_cap) {
    Compressor ctx;
    ctx.enableCudaAcceleration(false);
    SeqDispatcher seq;
    if (threads == 1) ctx.setTaskDispatcher(&seq);
    CompressionOptions co;
    setup_co(co, format, quality, colorWeights4, pixelType);
    OutputOptions oo;
    MemHandler mh;
    ErrCount eh;
    oo.setOutputHandler(&mh);
    oo.setErrorHandler(&eh);
    oo.setOutputHeader(false);
    bool ok;
    if (alphaMode == 0) {
        ok = ctx.compress(w, h, 1, 0, 0, rgba, co, oo);
    } else {
        Surface s;
        s.setAlphaMode((AlphaMode)alphaMode);
        // RGBA_32F input is interleaved; re-interleave from planar.
        std::vector<float> il((size_t)w * h * 4);
        for (size_t i = 0; i < (size_t)w * h; i++)
            for (int c = 0; c < 4; c++) il[i * 4 + c] = rgba[(size_t)c * w * h + i];
        s.setImage(InputFormat_RGBA_32F, w, h, 1, il.data());
        ok = ctx.compress(s, 0, 0, co, oo);
    }
    if (!ok || eh.n) return -1;
    if ((long)mh.buf.size() > out_cap) return -2;
    memcpy(out, mh.buf.data(), mh.buf.size());
    return (long)mh.buf.size();
}

struct RefProcessDesc {
    int inputFormat;     // nvtt::InputFormat
    int textureType;     // nvtt::TextureType
    int width, height, faces;
    int wrapMode;        // nvtt::WrapMode
    int mipmapFilter;    // nvtt::MipmapFilter
    int generateMipmaps; // bool
    int maxLevel;        // -1 = all
    float kaiserWidth, kaiserAlpha, kaiserStretch;
    float inputGamma, outputGamma;
    int isNormalMap, convertToNormalMap, normalizeMipmaps;
    int alphaMode;
    int format, quality, pixelType;
    float colorWeights[4];
    int outputHeader;    // bool
    int container;       // nvtt::Container
    int threads;         // 0 default pool, 1 sequential
    int quantization;    // CompressionOptions::setQuantization: bit 0 colour dithering, bit 1 alpha dithering, bit 2 binary alpha
    int alphaThreshold;  // 0..255 (127 = the default)
    // Format_RGB / Format_RGBA only.  pixelFormatMode: 0 = CompressionOptions defaults, 1 = setPixelFormat(bitcount, masks),
    // 2 = setPixelFormat(rsize, gsize, bsize, asize); pitchAlignment 0 = leave the default (1)
    int pixelFormatMode;
    unsigned bitcount, rmask, gmask, bmask, amask;
    int rsize, gsize, bsize, asize;
    int pitchAlignment;
    // InputOptions::setMaxExtents / setRoundMode (0 / RoundMode_None = leave the defaults), and caller-supplied mip levels:
    // mipImages[mip * faces + f] for mip >= 1 (userMips = number of levels incl. level 0 present in the array; NULL entries are
    // levels the caller leaves to the filter)
    int maxExtent, roundMode;
    int userMips;
    const void *const *mipImages;
};

// Whole InputOptions pipeline: Compressor::process (src/nvtt/Context.cpp:117-120,217-346).
// images[f] = level-0 data of face f in `inputFormat`.  Returns total bytes (header + every level).
long ref_process(const RefProcessDesc *d, const void *const *images, unsigned char *out, long out_cap) {
    InputOptions io;
    io.setTextureLayout((TextureType)d->textureType, d->width, d->height, 1, d->textureType == TextureType_Array ? d->faces : 1);
    io.setFormat((InputFormat)d->inputFormat);
    for (int f = 0; f < d->faces; f++) io.setMipmapData(images[f], d->width, d->height, 1, f, 0);
    if (d->userMips > 1 && d->mipImages) {
        int w = d->width, h = d->height;
        for (int m = 1; m < d->userMips; m++) {
            w = w / 2 > 1 ? w / 2 : 1;
            h = h / 2 > 1 ? h / 2 : 1;
            for (int f = 0; f < d->faces; f++)
                if (d->mipImages[m * d->faces + f]) io.setMipmapData(d->mipImages[m * d->faces + f], w, h, 1, f, m);
        }
    }
    if (d->maxExtent > 0) io.setMaxExtents(d->maxExtent);
    if (d->roundMode != 0) io.setRoundMode((RoundMode)d->roundMode);
    io.setWrapMode((WrapMode)d->wrapMode);
    io.setMipmapFilter((MipmapFilter)d->mipmapFilter);
    io.setMipmapGeneration(d->generateMipmaps != 0, d->maxLevel);
    io.setKaiserParameters(d->kaiserWidth, d->kaiserAlpha, d->kaiserStretch);
    io.setGamma(d->inputGamma, d->outputGamma);
    io.setNormalMap(d->isNormalMap != 0);
    io.setConvertToNormalMap(d->convertToNormalMap != 0);
    io.setNormalizeMipmaps(d->normalizeMipmaps != 0);
    io.setAlphaMode((AlphaMode)d->alphaMode);
    CompressionOptions co;
    setup_co(co, d->format, d->quality, d->colorWeights, d->pixelType);
    if (d->quantization) co.setQuantization((d->quantization & 1) != 0, (d->quantization & 2) != 0, (d->quantization & 4) != 0, d->alphaThreshold);
    if (d->pixelFormatMode == 1) co.setPixelFormat(d->bitcount, d->rmask, d->gmask, d->bmask, d->amask);
    else if (d->pixelFormatMode == 2) co.setPixelFormat((unsigned char)d->rsize, (unsigned char)d->gsize, (unsigned char)d->bsize, (unsigned char)d->asize);
    if (d->pitchAlignment > 0) co.setPitchAlignment(d->pitchAlignment);
    OutputOptions oo;
    MemHandler mh;
    ErrCount eh;
    oo.setOutputHandler(&mh);
    oo.setErrorHandler(&eh);
    oo.setOutputHeader(d->outputHeader != 0);
    oo.setContainer((Container)d->container);
    Compressor ctx;
    ctx.enableCudaAcceleration(false);
    SeqDispatcher seq;
    if (d->threads == 1) ctx.setTaskDispatcher(&seq);
    if (!ctx.process(io, co, oo)) return -1;
    if (eh.n) return -1;
    if (out == NULL) return (long)mh.buf.size();
    if ((long)mh.buf.size() > out_cap) return -2;
    memcpy(out, mh.buf.data(), mh.buf.size());
    return (long)mh.buf.size();
}

// Header only: Compressor::outputHeader(type,w,h,d,arraySize,mipmapCount,isNormalMap,...) (src/nvtt/Context.cpp:604-869)
long ref_output_header(int textureType, int w, int h, int arraySize, int mipmapCount, int isNormalMap, int format, int container,
                       unsigned char *out, long out_cap) {
    Compressor ctx;
    CompressionOptions co;
    co.setFormat((Format)format);
    OutputOptions oo;
    MemHandler mh;
    oo.setOutputHandler(&mh);
    oo.setContainer((Container)container);
    if (!ctx.outputHeader((TextureType)textureType, w, h, 1, arraySize, mipmapCount, isNormalMap != 0, co, oo)) return -1;
    if ((long)mh.buf.size() > out_cap) return -2;
    memcpy(out, mh.buf.data(), mh.buf.size());
    return (long)mh.buf.size();
}

// ---- Surface ("imperative") API, handle based (src/nvtt/Surface.cpp) ----
void *ref_surf_create(int wrapMode, int alphaMode, int isNormalMap) {
    Surface *s = new Surface();
    s->setWrapMode((WrapMode)wrapMode);
    s->setAlphaMode((AlphaMode)alphaMode);
    s->setNormalMap(isNormalMap != 0);
    return s;
}
void ref_surf_destroy(void *h) { delete (Surface *)h; }
int ref_surf_set_image(void *h, int inputFormat, int w, int ht, const void *data) {
    return ((Surface *)h)->setImage((InputFormat)inputFormat, w, ht, 1, data) ? 1 : 0;
}
int ref_surf_width(void *h) { return ((Surface *)h)->width(); }
int ref_surf_height(void *h) { return ((Surface *)h)->height(); }
// copies planar fp32 RGBA (4*w*h floats)
void ref_surf_get(void *h, float *out) {
    Surface *s = (Surface *)h;
    memcpy(out, s->data(), sizeof(float) * 4 * (size_t)s->width() * s->height() * s->depth());
}
void ref_surf_to_linear(void *h, float g) { ((Surface *)h)->toLinear(g); }
void ref_surf_to_gamma(void *h, float g) { ((Surface *)h)->toGamma(g); }
int ref_surf_build_next_mipmap(void *h, int filter, int useParams, float filterWidth, float p0, float p1) {
    Surface *s = (Surface *)h;
    if (useParams) {
        float params[2] = {p0, p1};
        return s->buildNextMipmap((MipmapFilter)filter, filterWidth, params) ? 1 : 0;
    }
    return s->buildNextMipmap((MipmapFilter)filter) ? 1 : 0;
}
void ref_surf_resize(void *h, int w, int ht, int filter, int useParams, float filterWidth, float p0, float p1) {
    Surface *s = (Surface *)h;
    if (useParams) {
        float params[2] = {p0, p1};
        s->resize(w, ht, 1, (ResizeFilter)filter, filterWidth, params);
    } else {
        s->resize(w, ht, 1, (ResizeFilter)filter);
    }
}
void ref_surf_expand_normals(void *h) { ((Surface *)h)->expandNormals(); }
void ref_surf_pack_normals(void *h) { ((Surface *)h)->packNormals(); }
void ref_surf_normalize_normal_map(void *h) { ((Surface *)h)->normalizeNormalMap(); }
void ref_surf_to_grey_scale(void *h, float r, float g, float b, float a) { ((Surface *)h)->toGreyScale(r, g, b, a); }
void ref_surf_to_normal_map(void *h, float sm, float md, float bg, float lg) { ((Surface *)h)->toNormalMap(sm, md, bg, lg); }
void ref_surf_quantize(void *h, int channel, int bits, int exactEndPoints, int dither) { ((Surface *)h)->quantize(channel, bits, exactEndPoints != 0, dither != 0); }
void ref_surf_binarize(void *h, int channel, float threshold, int dither) { ((Surface *)h)->binarize(channel, threshold, dither != 0); }

// Decode a BCn level with the library's decoder (Surface::setImage2D, src/nvtt/Surface.cpp:908-1118) -> planar fp32.
int ref_decode_ex(int format, int decoder, int w, int h, const void *data, float *out) {
    Surface s;
    if (!s.setImage2D((Format)format, (Decoder)decoder, w, h, data)) return 0;
    memcpy(out, s.data(), sizeof(float) * 4 * (size_t)w * h);
    return 1;
}
int ref_decode(int format, int w, int h, const void *data, float *out) { return ref_decode_ex(format, 0, w, h, data, out); }

// nvtt::rmsError / rmsAlphaError between an RGBA32F image (reference, with the given alpha mode) and a decoded BCn level
int ref_rms_error(int format, int w, int h, const void *blocks, const float *rgba32f, int alphaMode, float *rms, float *rmsAlpha) {
    Surface ref, img;
    ref.setAlphaMode((AlphaMode)alphaMode);
    if (!ref.setImage(InputFormat_RGBA_32F, w, h, 1, rgba32f)) return 0;
    if (!img.setImage2D((Format)format, Decoder_D3D10, w, h, blocks)) return 0;
    *rms = nvtt::rmsError(ref, img);
    *rmsAlpha = nvtt::rmsAlphaError(ref, img);
    return 1;
}
// nvtt::cieLabError between an RGBA32F image and its decoded BCn level
int ref_cielab_error(int format, int w, int h, const void *blocks, const float *rgba32f, float *out) {
    Surface ref, img;
    if (!ref.setImage(InputFormat_RGBA_32F, w, h, 1, rgba32f)) return 0;
    if (!img.setImage2D((Format)format, Decoder_D3D10, w, h, blocks)) return 0;
    *out = nvtt::cieLabError(ref, img);
    return 1;
}
// nvtt::angularError between a packed normal map (RGBA32F) and its decoded BCn level
int ref_angular_error(int format, int w, int h, const void *blocks, const float *rgba32f, float *out) {
    Surface ref, img;
    if (!ref.setImage(InputFormat_RGBA_32F, w, h, 1, rgba32f)) return 0;
    if (!img.setImage2D((Format)format, Decoder_D3D10, w, h, blocks)) return 0;
    *out = nvtt::angularError(ref, img);
    return 1;
}

int ref_version() { return (int)nvtt::version(); }
}
