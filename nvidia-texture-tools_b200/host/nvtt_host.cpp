// C++ host layer: the nvtt:: classes of host/nvtt/nvtt.h implemented on the C ABI (include/nvtt_b200.h) only.
// Mirrors the control flow and error behaviour of the reference's src/nvtt/Context.cpp:61-516 (process, compress,
// outputHeader, estimateSize), InputOptions.cpp:96-330, CompressionOptions.cpp:48-200, OutputOptions.cpp:44-180 and the
// DDS header writer nvimage/DirectDrawSurface.cpp:574-830.  No image is ever processed on the CPU here.
#include "nvtt/nvtt.h"
#include "../../include/nvtt_b200.h"

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <math.h>
#include <string>
#include <vector>

using namespace nvtt;

namespace {
inline int imax(int a, int b) { return a > b ? a : b; }
inline int imin(int a, int b) { return a < b ? a : b; }

// The GPU contexts of this process.  Surfaces and single-level calls live on the first one (device NVTT_B200_DEVICE,
// default 0); Compressor::process spreads large jobs over all of them (nvttb_process_multi): NVTT_B200_DEVICES = "all"
// (default), a count, or a comma-separated device list.
struct Gpu {
    NvttbContext *ctx = nullptr;
    std::vector<NvttbContext *> all;
    bool tried = false, triedAll = false;
    NvttbContext *get() {
        if (!tried) {
            tried = true;
            int dev = 0;
            if (const char *e = getenv("NVTT_B200_DEVICE")) dev = atoi(e);
            if (nvttb_context_create(dev, &ctx) != NVTTB_OK) ctx = nullptr;
        }
        return ctx;
    }
    const std::vector<NvttbContext *> &pool() {
        if (!triedAll) {
            triedAll = true;
            NvttbContext *first = get();
            if (first) {
                all.push_back(first);
                int firstDev = 0;
                if (const char *e = getenv("NVTT_B200_DEVICE")) firstDev = atoi(e);
                const int n = nvttb_device_count();
                std::vector<int> devs;
                const char *e = getenv("NVTT_B200_DEVICES");
                if (e && strchr(e, ',')) {
                    for (const char *p = e; *p;) {
                        devs.push_back(atoi(p));
                        p = strchr(p, ',');
                        if (!p) break;
                        p++;
                    }
                } else {
                    int want = (e && strcmp(e, "all") != 0) ? atoi(e) : n;
                    for (int d = 0; d < n && (int)devs.size() < want; d++) devs.push_back(d);
                }
                for (int d : devs) {
                    if (d == firstDev || d < 0 || d >= n) continue;
                    NvttbContext *c = nullptr;
                    if (nvttb_context_create(d, &c) == NVTTB_OK) all.push_back(c);
                }
            }
        }
        return all;
    }
};
Gpu g_gpu;
SurfaceLoader g_surfaceLoader = nullptr;

unsigned previousPowerOfTwo(unsigned v) {
    unsigned p = 1;
    while (p * 2 <= v && p * 2 != 0) p *= 2;
    return p;
}
unsigned nextPowerOfTwo(unsigned v) {
    unsigned p = 1;
    while (p < v) p *= 2;
    return p;
}
unsigned nearestPowerOfTwo(unsigned v) {
    const unsigned np2 = nextPowerOfTwo(v), pp2 = previousPowerOfTwo(v);
    return (np2 - v <= v - pp2) ? np2 : pp2;
}
int countMipmaps(int w, int h, int d) {
    int m = 0;
    while (w != 1 || h != 1 || d != 1) {
        w = imax(1, w / 2);
        h = imax(1, h / 2);
        d = imax(1, d / 2);
        m++;
    }
    return m + 1;
}
int blockSize(Format f) {
    switch (f) {
    case Format_DXT1: case Format_DXT1a: case Format_DXT1n: case Format_BC4: case Format_CTX1: return 8;
    case Format_DXT3: case Format_DXT5: case Format_DXT5n: case Format_BC3_RGBM: case Format_BC5: case Format_BC6: case Format_BC7: return 16;
    default: return 0;
    }
}
// nv::getTargetExtent (src/nvtt/Surface.cpp:220-327), including its dead "nearest multiple of four" branch.
void getTargetExtent(int *width, int *height, int *depth, int maxExtent, RoundMode roundMode, TextureType type) {
    int w = *width, h = *height, d = *depth;
    if (roundMode != RoundMode_None && maxExtent > 0) maxExtent = (int)previousPowerOfTwo((unsigned)maxExtent);
    const int m = imax(imax(w, h), d);
    if (maxExtent > 0 && m > maxExtent) {
        w = imax((w * maxExtent) / m, 1);
        h = imax((h * maxExtent) / m, 1);
        d = imax((d * maxExtent) / m, 1);
    }
    if (type == TextureType_2D) d = 1;
    else if (type == TextureType_Cube) { w = h = (w + h) / 2; d = 1; }
    if (roundMode == RoundMode_ToNextPowerOfTwo) { w = nextPowerOfTwo(w); h = nextPowerOfTwo(h); d = nextPowerOfTwo(d); }
    else if (roundMode == RoundMode_ToNearestPowerOfTwo) { w = nearestPowerOfTwo(w); h = nearestPowerOfTwo(h); d = nearestPowerOfTwo(d); }
    else if (roundMode == RoundMode_ToPreviousPowerOfTwo) { w = previousPowerOfTwo(w); h = previousPowerOfTwo(h); d = previousPowerOfTwo(d); }
    else if (roundMode == RoundMode_ToNextMultipleOfFour) { w = (w + 3) & ~3; h = (h + 3) & ~3; d = (d + 3) & ~3; }
    else if (roundMode == RoundMode_ToPreviousMultipleOfFour) { w = imax(w & ~3, 4); h = imax(h & ~3, 4); d = imax(d & ~3, 4); }
    if (type == TextureType_2D || type == TextureType_Cube) d = 1;
    *width = w; *height = h; *depth = d;
}
}  // namespace

// ---- option objects ---------------------------------------------------------------------------------------
struct CompressionOptions::Private {
    Format format;
    Quality quality;
    float colorWeight[4];
    PixelType pixelType;
    Decoder decoder;
    bool enableColorDithering, enableAlphaDithering, binaryAlpha;
    int alphaThreshold;
    float rgbmThreshold;
    // Format_RGB / Format_RGBA (CompressionOptions.h:40-75)
    unsigned bitcount, rmask, gmask, bmask, amask;
    unsigned char rsize, gsize, bsize, asize;
    int pitchAlignment;
    unsigned getBitCount() const {
        if (format == Format_RGBA) return bitcount != 0 ? bitcount : (unsigned)rsize + gsize + bsize + asize;
        return 0;
    }
};
CompressionOptions::CompressionOptions() : m(*new Private()) { reset(); }
CompressionOptions::~CompressionOptions() { delete &m; }
void CompressionOptions::reset() {
    m.format = Format_DXT1;
    m.quality = Quality_Normal;
    m.colorWeight[0] = m.colorWeight[1] = m.colorWeight[2] = m.colorWeight[3] = 1.0f;
    m.pixelType = PixelType_UnsignedNorm;
    m.decoder = Decoder_D3D10;
    m.enableColorDithering = m.enableAlphaDithering = m.binaryAlpha = false;
    m.alphaThreshold = 127;
    m.rgbmThreshold = 0.15f;  // CompressionOptions.cpp:53
    m.bitcount = 32;
    m.bmask = 0x000000FF; m.gmask = 0x0000FF00; m.rmask = 0x00FF0000; m.amask = 0xFF000000;
    m.rsize = m.gsize = m.bsize = m.asize = 8;
    m.pitchAlignment = 1;
}
// CompressionOptions.cpp:113-173
void CompressionOptions::setPixelFormat(unsigned int bitCount, unsigned int rmask, unsigned int gmask, unsigned int bmask, unsigned int amask) {
    m.bitcount = bitCount;
    m.rmask = rmask; m.gmask = gmask; m.bmask = bmask; m.amask = amask;
    m.rsize = m.gsize = m.bsize = m.asize = 0;
}
void CompressionOptions::setPixelFormat(unsigned char rsize, unsigned char gsize, unsigned char bsize, unsigned char asize) {
    m.bitcount = 0;
    m.rmask = m.gmask = m.bmask = m.amask = 0;
    m.rsize = rsize; m.gsize = gsize; m.bsize = bsize; m.asize = asize;
}
void CompressionOptions::setPitchAlignment(int pitchAlignment) { m.pitchAlignment = pitchAlignment; }
void CompressionOptions::setRGBMThreshold(float min_m) { m.rgbmThreshold = min_m; }
void CompressionOptions::setFormat(Format f) { m.format = f; }
void CompressionOptions::setQuality(Quality q) { m.quality = q; }
void CompressionOptions::setColorWeights(float r, float g, float b, float a) {
    m.colorWeight[0] = r; m.colorWeight[1] = g; m.colorWeight[2] = b; m.colorWeight[3] = a;
}
void CompressionOptions::setPixelType(PixelType t) { m.pixelType = t; }
void CompressionOptions::setQuantization(bool c, bool a, bool b, int t) {
    m.enableColorDithering = c; m.enableAlphaDithering = a; m.binaryAlpha = b; m.alphaThreshold = t;
}
void CompressionOptions::setTargetDecoder(Decoder d) { m.decoder = d; }
void CompressionOptions::setExternalCompressor(const char *) {}  // CompressionOptions.cpp:177; no external codec is ever compiled in
Format CompressionOptions::format() const { return m.format; }

struct InputOptions::Private {
    WrapMode wrapMode;
    TextureType textureType;
    InputFormat inputFormat;
    AlphaMode alphaMode;
    int width, height, depth, faceCount, mipmapCount, imageCount;
    std::vector<std::vector<unsigned char>> images;  // [mip * faceCount + face], empty = NULL
    float inputGamma, outputGamma;
    bool generateMipmaps;
    int maxLevel;
    MipmapFilter mipmapFilter;
    float kaiserWidth, kaiserAlpha, kaiserStretch;
    bool isNormalMap, normalizeMipmaps, convertToNormalMap;
    float heightFactors[4], bumpFrequencyScale[4];
    int maxExtent;
    RoundMode roundMode;
};
InputOptions::InputOptions() : m(*new Private()) {
    m.width = m.height = m.depth = m.faceCount = m.mipmapCount = m.imageCount = 0;
    reset();
}
InputOptions::~InputOptions() { delete &m; }
void InputOptions::reset() {
    m.wrapMode = WrapMode_Mirror;
    m.textureType = TextureType_2D;
    m.inputFormat = InputFormat_BGRA_8UB;
    m.alphaMode = AlphaMode_None;
    m.inputGamma = m.outputGamma = 2.2f;
    m.generateMipmaps = true;
    m.maxLevel = -1;
    m.mipmapFilter = MipmapFilter_Box;
    m.kaiserWidth = 3;
    m.kaiserAlpha = 4.0f;
    m.kaiserStretch = 1.0f;
    m.isNormalMap = false;
    m.normalizeMipmaps = true;
    m.convertToNormalMap = false;
    m.heightFactors[0] = m.heightFactors[1] = m.heightFactors[2] = 0.0f;
    m.heightFactors[3] = 1.0f;
    const float s = 1.0f + 0.5f + 0.25f + 0.125f;
    m.bumpFrequencyScale[0] = 1.0f / s; m.bumpFrequencyScale[1] = 0.5f / s; m.bumpFrequencyScale[2] = 0.25f / s; m.bumpFrequencyScale[3] = 0.125f / s;
    m.maxExtent = 0;
    m.roundMode = RoundMode_None;
}
void InputOptions::setTextureLayout(TextureType type, int w, int h, int d, int arraySize) {
    resetTextureLayout();
    m.textureType = type;
    m.width = w; m.height = h; m.depth = d;
    m.faceCount = (type == TextureType_Cube) ? 6 : (type == TextureType_Array ? arraySize : 1);
    m.mipmapCount = countMipmaps(w, h, d);
    m.imageCount = m.mipmapCount * m.faceCount;
    m.images.assign(m.imageCount, std::vector<unsigned char>());
}
void InputOptions::resetTextureLayout() {
    m.images.clear();
    m.width = m.height = m.depth = m.faceCount = m.mipmapCount = m.imageCount = 0;
}
bool InputOptions::setMipmapData(const void *data, int width, int height, int depth, int face, int mipLevel) {
    if ((unsigned)face >= (unsigned)m.faceCount) return false;
    if ((unsigned)mipLevel >= (unsigned)m.mipmapCount) return false;
    const int idx = mipLevel * m.faceCount + face;
    if (idx >= m.imageCount) return false;
    int w = m.width, h = m.height, d = m.depth;
    for (int i = 0; i < mipLevel; i++) { w = imax(1, w / 2); h = imax(1, h / 2); d = imax(1, d / 2); }
    if (w != width || h != height || d != depth) return false;
    size_t bpp;
    switch (m.inputFormat) {
    case InputFormat_BGRA_8UB: bpp = 4; break;
    case InputFormat_RGBA_16F: bpp = 8; break;
    case InputFormat_RGBA_32F: bpp = 16; break;
    case InputFormat_R_32F: bpp = 4; break;
    default: return false;
    }
    const size_t bytes = (size_t)width * height * depth * bpp;
    m.images[idx].assign((const unsigned char *)data, (const unsigned char *)data + bytes);
    return true;
}
void InputOptions::setFormat(InputFormat f) { m.inputFormat = f; }
void InputOptions::setAlphaMode(AlphaMode a) { m.alphaMode = a; }
void InputOptions::setGamma(float i, float o) { m.inputGamma = i; m.outputGamma = o; }
void InputOptions::setWrapMode(WrapMode w) { m.wrapMode = w; }
void InputOptions::setMipmapFilter(MipmapFilter f) { m.mipmapFilter = f; }
void InputOptions::setMipmapGeneration(bool e, int maxLevel) { m.generateMipmaps = e; m.maxLevel = maxLevel; }
void InputOptions::setKaiserParameters(float w, float a, float s) { m.kaiserWidth = w; m.kaiserAlpha = a; m.kaiserStretch = s; }
void InputOptions::setNormalMap(bool b) { m.isNormalMap = b; }
void InputOptions::setConvertToNormalMap(bool c) { m.convertToNormalMap = c; }
void InputOptions::setHeightEvaluation(float r, float g, float b, float a) { m.heightFactors[0] = r; m.heightFactors[1] = g; m.heightFactors[2] = b; m.heightFactors[3] = a; }
void InputOptions::setNormalFilter(float s, float md, float bg, float lg) {
    const float total = s + md + bg + lg;  // InputOptions.cpp: normalised to sum 1
    m.bumpFrequencyScale[0] = s / total; m.bumpFrequencyScale[1] = md / total; m.bumpFrequencyScale[2] = bg / total; m.bumpFrequencyScale[3] = lg / total;
}
void InputOptions::setNormalizeMipmaps(bool b) { m.normalizeMipmaps = b; }
void InputOptions::setMaxExtents(int d) { m.maxExtent = d; }
void InputOptions::setRoundMode(RoundMode r) { m.roundMode = r; }

namespace {
struct FileOutputHandler : public OutputHandler {
    FILE *fp;
    bool own;
    FileOutputHandler(const char *name) : fp(fopen(name, "wb")), own(true) {}
    FileOutputHandler(FILE *f) : fp(f), own(false) {}
    ~FileOutputHandler() { if (fp && own) fclose(fp); }
    void beginImage(int, int, int, int, int, int) {}
    void endImage() {}
    bool writeData(const void *data, int size) { return fp && fwrite(data, 1, (size_t)size, fp) == (size_t)size; }
};
}  // namespace

struct OutputOptions::Private {
    std::string fileName;
    OutputHandler *outputHandler;
    ErrorHandler *errorHandler;
    bool outputHeader;
    Container container;
    int version;
    bool srgb;
    bool deleteOutputHandler;
    void error(Error e) const { if (errorHandler) errorHandler->error(e); }
    bool writeData(const void *d, int n) const { return outputHandler ? outputHandler->writeData(d, n) : true; }
};
OutputOptions::OutputOptions() : m(*new Private()) {
    m.outputHandler = nullptr;
    m.deleteOutputHandler = false;
    reset();
}
OutputOptions::~OutputOptions() {
    if (m.deleteOutputHandler) delete m.outputHandler;
    delete &m;
}
void OutputOptions::reset() {
    if (m.deleteOutputHandler) delete m.outputHandler;
    m.fileName.clear();
    m.outputHandler = nullptr;
    m.errorHandler = nullptr;
    m.outputHeader = true;
    m.container = Container_DDS;
    m.version = 0;
    m.srgb = false;
    m.deleteOutputHandler = false;
}
void OutputOptions::setFileName(const char *fileName) {
    if (m.deleteOutputHandler) delete m.outputHandler;
    m.fileName = fileName;
    m.outputHandler = nullptr;
    m.deleteOutputHandler = false;
    FileOutputHandler *oh = new FileOutputHandler(fileName);
    if (!oh->fp) { delete oh; return; }
    m.outputHandler = oh;
    m.deleteOutputHandler = true;
}
void OutputOptions::setFileHandle(void *fp) {
    if (m.deleteOutputHandler) delete m.outputHandler;
    m.fileName.clear();
    m.outputHandler = new FileOutputHandler((FILE *)fp);
    m.deleteOutputHandler = true;
}
void OutputOptions::setOutputHandler(OutputHandler *oh) {
    if (m.deleteOutputHandler) delete m.outputHandler;
    m.fileName.clear();
    m.outputHandler = oh;
    m.deleteOutputHandler = false;
}
void OutputOptions::setErrorHandler(ErrorHandler *eh) { m.errorHandler = eh; }
void OutputOptions::setOutputHeader(bool b) { m.outputHeader = b; }
void OutputOptions::setContainer(Container c) { m.container = c; }
void OutputOptions::setUserVersion(int v) { m.version = v; }
void OutputOptions::setSrgbFlag(bool b) { m.srgb = b; }

// ---- Surface ------------------------------------------------------------------------------------------------
struct Surface::Private {
    NvttbSurface *s = nullptr;
    WrapMode wrapMode = WrapMode_Mirror;
    AlphaMode alphaMode = AlphaMode_None;
    bool isNormalMap = false;
    mutable std::vector<float> hostCopy;
    mutable bool hostValid = false;
};
namespace {
void surf_sync_flags(Surface::Private *m) {
    if (!m->s) return;
    nvttb_surface_set_wrap_mode(m->s, m->wrapMode);
    nvttb_surface_set_alpha_mode(m->s, m->alphaMode);
    nvttb_surface_set_normal_map(m->s, m->isNormalMap ? 1 : 0);
}
}  // namespace
Surface::Surface() : m(new Private()) {}
Surface::Surface(const Surface &o) : m(new Private()) { *this = o; }
Surface::~Surface() {
    if (m->s) nvttb_surface_destroy(m->s);
    delete m;
}
void Surface::operator=(const Surface &o) {
    if (this == &o) return;
    if (m->s) { nvttb_surface_destroy(m->s); m->s = nullptr; }
    m->wrapMode = o.m->wrapMode;
    m->alphaMode = o.m->alphaMode;
    m->isNormalMap = o.m->isNormalMap;
    m->hostValid = false;
    if (o.m->s) nvttb_surface_clone(o.m->s, &m->s);  // deep device copy (the reference's copy-on-write, eager)
}
void Surface::setWrapMode(WrapMode w) { m->wrapMode = w; surf_sync_flags(m); }
void Surface::setAlphaMode(AlphaMode a) { m->alphaMode = a; surf_sync_flags(m); }
void Surface::setNormalMap(bool b) { m->isNormalMap = b; surf_sync_flags(m); }
bool Surface::isNull() const { return m->s == nullptr || nvttb_surface_width(m->s) == 0; }
int Surface::width() const { return m->s ? nvttb_surface_width(m->s) : 0; }
int Surface::height() const { return m->s ? nvttb_surface_height(m->s) : 0; }
int Surface::depth() const { return isNull() ? 0 : 1; }
TextureType Surface::type() const { return TextureType_2D; }
WrapMode Surface::wrapMode() const { return m->wrapMode; }
AlphaMode Surface::alphaMode() const { return m->alphaMode; }
bool Surface::isNormalMap() const { return m->isNormalMap; }
int Surface::countMipmaps() const { return isNull() ? 0 : ::countMipmaps(width(), height(), 1); }
const float *Surface::data() const {
    if (isNull()) return nullptr;
    if (!m->hostValid) {
        m->hostCopy.resize((size_t)4 * width() * height());
        if (nvttb_surface_download(m->s, m->hostCopy.data()) != NVTTB_OK) return nullptr;
        m->hostValid = true;
    }
    return m->hostCopy.data();
}
bool Surface::setImage(InputFormat format, int w, int h, int d, const void *data) {
    if (d != 1) return false;  // 3D textures are outside the hot path
    NvttbContext *ctx = g_gpu.get();
    if (!ctx) return false;
    if (!m->s) {
        if (nvttb_surface_create(ctx, &m->s) != NVTTB_OK) return false;
        surf_sync_flags(m);
    }
    m->hostValid = false;
    return nvttb_surface_set_image(m->s, format, w, h, data, NVTTB_HOST) == NVTTB_OK;
}
// Surface::setImage2D (Surface.cpp:908-1118): BCn blocks -> planar fp32, on the device
bool Surface::setImage2D(Format format, Decoder decoder, int w, int h, const void *data) {
    NvttbContext *ctx = g_gpu.get();
    if (!ctx) return false;
    if (!m->s) {
        if (nvttb_surface_create(ctx, &m->s) != NVTTB_OK) return false;
        surf_sync_flags(m);
    }
    m->hostValid = false;
    return nvttb_surface_set_image_2d(m->s, format, decoder, w, h, data, NVTTB_HOST, 0) == NVTTB_OK;
}
void Surface::resize(int w, int h, int d, ResizeFilter filter) {
    if (isNull() || d != 1) return;
    m->hostValid = false;
    nvttb_surface_resize(m->s, w, h, filter, 0, 0.0f, nullptr);
}
void Surface::resize(int w, int h, int d, ResizeFilter filter, float filterWidth, const float *params) {
    if (isNull() || d != 1) return;
    m->hostValid = false;
    nvttb_surface_resize(m->s, w, h, filter, 1, filterWidth, params);
}
bool Surface::canMakeNextMipmap(int min_size) {
    if (isNull()) return false;
    const int w = width(), h = height();
    if (min_size == 1) return !(w == 1 && h == 1);
    return !(w <= min_size || h <= min_size);
}
bool Surface::buildNextMipmap(MipmapFilter filter, int min_size) {
    if (!canMakeNextMipmap(min_size)) return false;
    int built = 0;
    m->hostValid = false;
    return nvttb_surface_build_next_mipmap(m->s, filter, 0, 0.0f, nullptr, &built) == NVTTB_OK && built;
}
bool Surface::buildNextMipmap(MipmapFilter filter, float filterWidth, const float *params, int min_size) {
    if (!canMakeNextMipmap(min_size)) return false;
    int built = 0;
    m->hostValid = false;
    // params == NULL keeps the filter's own defaults (Kaiser: alpha 4, stretch 1)
    float def[2] = {4.0f, 1.0f};
    return nvttb_surface_build_next_mipmap(m->s, filter, 1, filterWidth, params ? params : def, &built) == NVTTB_OK && built;
}
void Surface::toLinear(float g) { if (!isNull()) { m->hostValid = false; nvttb_surface_to_linear(m->s, g); } }
void Surface::toGamma(float g) { if (!isNull()) { m->hostValid = false; nvttb_surface_to_gamma(m->s, g); } }
void Surface::toGreyScale(float r, float g, float b, float a) { if (!isNull()) { m->hostValid = false; nvttb_surface_to_grey_scale(m->s, r, g, b, a); } }
void Surface::toNormalMap(float sm, float md, float bg, float lg) {
    if (isNull()) return;
    m->hostValid = false;
    if (nvttb_surface_to_normal_map(m->s, sm, md, bg, lg) == NVTTB_OK) { m->isNormalMap = true; surf_sync_flags(m); }
}
void Surface::normalizeNormalMap() { if (!isNull() && m->isNormalMap) { m->hostValid = false; nvttb_surface_normalize_normal_map(m->s); } }
void Surface::binarize(int channel, float threshold, bool dither) {
    if (isNull()) return;
    m->hostValid = false;
    nvttb_surface_binarize(m->s, channel, threshold, dither ? 1 : 0);
}
void Surface::quantize(int channel, int bits, bool exactEndPoints, bool dither) {
    if (isNull()) return;
    m->hostValid = false;
    nvttb_surface_quantize(m->s, channel, bits, exactEndPoints ? 1 : 0, dither ? 1 : 0);
}
// Surface::packNormals / expandNormals = scaleBias(0, 3, scale, bias) for any values (Surface.cpp:2952-2963)
void Surface::packNormals(float scale, float bias) {
    if (isNull()) return;
    m->hostValid = false;
    nvttb_surface_scale_bias(m->s, 0, 3, scale, bias);
}
void Surface::expandNormals(float scale, float bias) {
    if (isNull()) return;
    m->hostValid = false;
    nvttb_surface_scale_bias(m->s, 0, 3, scale, bias);
}
namespace {
bool nv_equal_f(float a, float b) {  // nv::equal (nvmath.h:139-143)
    float mx = 1.0f;
    if (fabsf(a) > mx) mx = fabsf(a);
    if (fabsf(b) > mx) mx = fabsf(b);
    return fabsf(a - b) <= 0.0001f * mx;
}
}  // namespace
void Surface::scaleBias(int channel, float scale, float bias) {
    if (isNull()) return;
    if (nv_equal_f(scale, 1.0f) && nv_equal_f(bias, 0.0f)) return;  // Surface.cpp:1672
    m->hostValid = false;
    nvttb_surface_scale_bias(m->s, channel, 1, scale, bias);
}
void Surface::clamp(int channel, float low, float high) {
    if (isNull()) return;
    m->hostValid = false;
    nvttb_surface_clamp(m->s, channel, low, high);
}
void Surface::toRGBM(float range, float threshold) {
    if (isNull()) return;
    m->hostValid = false;
    nvttb_surface_to_rgbm(m->s, range, threshold);
}
void Surface::toneMap(ToneMapper tm, float *) {
    if (isNull()) return;
    m->hostValid = false;
    nvttb_surface_tone_map(m->s, (int)tm);
}
void Surface::range(int channel, float *rangeMin, float *rangeMax, int alpha_channel, float alpha_ref) const {
    float lo = FLT_MAX, hi = -FLT_MAX;
    if (m->s) nvttb_surface_range(m->s, channel, alpha_channel, alpha_ref, &lo, &hi);
    if (rangeMin) *rangeMin = lo;
    if (rangeMax) *rangeMax = hi;
}
bool Surface::load(const char *fileName, bool *hasAlpha) {
    if (!g_surfaceLoader || !fileName) return false;
    return g_surfaceLoader(*this, fileName, hasAlpha);
}
void nvtt::setSurfaceLoader(SurfaceLoader loader) { g_surfaceLoader = loader; }
// planar source channels (Surface.cpp:817-901): rearranged into the interleaved layout of the same InputFormat on the host
// (a pure copy), then the regular device conversion
bool Surface::setImage(InputFormat format, int w, int h, int d, const void *r, const void *g, const void *b, const void *a) {
    if (d != 1 || w <= 0 || h <= 0) return false;
    const size_t n = (size_t)w * h;
    if (format == InputFormat_R_32F) return setImage(format, w, h, d, r);
    if (format == InputFormat_BGRA_8UB) {
        std::vector<unsigned char> t(n * 4);
        const unsigned char *R = (const unsigned char *)r, *G = (const unsigned char *)g, *B = (const unsigned char *)b, *A = (const unsigned char *)a;
        for (size_t i = 0; i < n; i++) { t[4 * i] = B[i]; t[4 * i + 1] = G[i]; t[4 * i + 2] = R[i]; t[4 * i + 3] = A[i]; }
        return setImage(format, w, h, d, t.data());
    }
    if (format == InputFormat_RGBA_16F) {
        std::vector<unsigned short> t(n * 4);
        const unsigned short *R = (const unsigned short *)r, *G = (const unsigned short *)g, *B = (const unsigned short *)b, *A = (const unsigned short *)a;
        for (size_t i = 0; i < n; i++) { t[4 * i] = R[i]; t[4 * i + 1] = G[i]; t[4 * i + 2] = B[i]; t[4 * i + 3] = A[i]; }
        return setImage(format, w, h, d, t.data());
    }
    if (format == InputFormat_RGBA_32F) {
        std::vector<float> t(n * 4);
        const float *R = (const float *)r, *G = (const float *)g, *B = (const float *)b, *A = (const float *)a;
        for (size_t i = 0; i < n; i++) { t[4 * i] = R[i]; t[4 * i + 1] = G[i]; t[4 * i + 2] = B[i]; t[4 * i + 3] = A[i]; }
        return setImage(format, w, h, d, t.data());
    }
    return false;
}

// ---- Compressor ---------------------------------------------------------------------------------------------
struct Compressor::Private {
    bool cudaSupported, cudaEnabled;
    TaskDispatcher *dispatcher;
};
Compressor::Compressor() : m(*new Private()) {
    m.cudaSupported = g_gpu.get() != nullptr;
    m.cudaEnabled = m.cudaSupported;
    m.dispatcher = nullptr;
    if (!m.cudaSupported) fprintf(stderr, "nvtt (B200): no CUDA device available — compression calls will fail with Error_CudaError\n");
}
Compressor::~Compressor() { delete &m; }
// The GPU path is the only implementation: the request is accepted for source compatibility (nvcompress -nocuda) but there is
// no CPU encoder to switch to, so it does not change what runs.
void Compressor::enableCudaAcceleration(bool) { m.cudaEnabled = m.cudaSupported; }
bool Compressor::isCudaAccelerationEnabled() const { return m.cudaEnabled; }
void Compressor::setTaskDispatcher(TaskDispatcher *disp) { m.dispatcher = disp; }

namespace {
void fill_encode(NvttbEncodeDesc *e, const CompressionOptions::Private &co, AlphaMode am, int w, int h) {
    e->format = co.format;
    e->quality = co.quality;
    e->alphaMode = am;
    e->pixelType = co.pixelType;
    for (int i = 0; i < 4; i++) e->colorWeights[i] = co.colorWeight[i];
    e->width = w;
    e->height = h;
    e->applyToGamma = 0;
    e->rgbmThreshold = co.rgbmThreshold;
}
void fill_pixel_format(NvttbPixelFormatDesc *p, const CompressionOptions::Private &co, int w, int h) {
    p->pixelType = co.pixelType;
    p->bitcount = co.bitcount;
    p->rmask = co.rmask; p->gmask = co.gmask; p->bmask = co.bmask; p->amask = co.amask;
    p->rsize = co.rsize; p->gsize = co.gsize; p->bsize = co.bsize; p->asize = co.asize;
    p->pitchAlignment = co.pitchAlignment;
    p->width = w;
    p->height = h;
}
unsigned byte_pitch(unsigned w, unsigned bitsize, unsigned alignmentInBytes) {  // nv::computeBytePitch (nvimage.h:11-24)
    const unsigned a = 8 * alignmentInBytes;
    return (((w * bitsize + a - 1) / a) * a + 7) / 8;
}
// nv::computeImageSize (Surface.cpp:210-218)
int image_size(int w, int h, int d, const CompressionOptions::Private &co) {
    if (co.format == Format_RGBA) return d * h * (int)byte_pitch(w, co.getBitCount(), co.pitchAlignment);
    return ((w + 3) / 4) * ((h + 3) / 4) * blockSize(co.format) * d;
}
// Compressor::Private::estimateSize for one level (Context.cpp:1225-1245), the KTX imageSize field: unlike the public
// estimateSize it reads `bitcount` directly, so a Format_RGBA layout given as channel sizes (bitcount == 0) reports 0 bytes
// (and gets no mip padding).  Kept.
int ktx_level_size(int w, int h, const CompressionOptions::Private &co) {
    if (co.format == Format_RGBA) return h * (int)byte_pitch(w, co.bitcount, co.pitchAlignment);
    return image_size(w, h, 1, co);
}
// findDXGIFormat (src/nvimage/DirectDrawSurface.cpp:483-546): the mask layouts that have a DXGI equivalent
uint32_t find_dxgi_format(unsigned bitcount, unsigned r, unsigned g, unsigned b, unsigned a) {
    static const struct { uint32_t dxgi, bits, r, g, b, a; } t[] = {
        {87, 32, 0xFF0000, 0xFF00, 0xFF, 0xFF000000}, {88, 32, 0xFF0000, 0xFF00, 0xFF, 0},  {85, 16, 0xF800, 0x7E0, 0x1F, 0},
        {86, 16, 0x7C00, 0x3E0, 0x1F, 0x8000},        {65, 8, 0, 0, 0, 8},                  {24, 32, 0x3FF, 0xFFC00, 0x3FF00000, 0xC0000000},
        {28, 32, 0xFF, 0xFF00, 0xFF0000, 0xFF000000}, {35, 32, 0xFFFF, 0xFFFF0000, 0, 0},   {61, 8, 0xFF, 0, 0, 0},
        {56, 16, 0xFFFF, 0, 0, 0},                    {49, 16, 0xFF, 0xFF00, 0, 0},
    };
    // (the reference's table also lists layouts without a DXGI code; they resolve to DXGI_FORMAT_UNKNOWN = unsupported, like no match)
    for (const auto &e : t)
        if (e.bits == bitcount && e.r == r && e.g == g && e.b == b && e.a == a) return e.dxgi;
    return 0;
}

struct DDSHeaderBytes {  // nvimage/DirectDrawSurface.h:283-430 layout, little endian
    uint32_t fourcc, size, flags, height, width, pitch, depth, mipmapcount, reserved[11];
    uint32_t pf_size, pf_flags, pf_fourcc, pf_bitcount, pf_rmask, pf_gmask, pf_bmask, pf_amask;
    uint32_t caps1, caps2, caps3, caps4, notused;
    uint32_t dxgiFormat, resourceDimension, miscFlag, arraySize, reserved10;
};
inline uint32_t fourcc(char a, char b, char c, char d) { return (uint32_t)(uint8_t)a | ((uint32_t)(uint8_t)b << 8) | ((uint32_t)(uint8_t)c << 16) | ((uint32_t)(uint8_t)d << 24); }

struct EmitCtx {
    const OutputOptions::Private *oo;
};
int emit_cb(void *user, int face, int mip, int w, int h, int d, const void *data, size_t size) {
    const OutputOptions::Private *oo = ((EmitCtx *)user)->oo;
    if (oo->outputHandler) {
        oo->outputHandler->beginImage((int)size, w, h, d, face, mip);
        oo->outputHandler->writeData(data, (int)size);
        oo->outputHandler->endImage();
    }
    return 1;
}
}  // namespace

bool Compressor::outputHeader(TextureType textureType, int w, int h, int d, int arraySize, int mipmapCount, bool isNormalMap,
                              const CompressionOptions &compressionOptions, const OutputOptions &outputOptions) const {
    const CompressionOptions::Private &co = compressionOptions.m;
    const OutputOptions::Private &oo = outputOptions.m;
    if (w <= 0 || h <= 0 || d <= 0 || arraySize <= 0 || mipmapCount <= 0) {
        oo.error(Error_InvalidInput);
        return false;
    }
    if (!oo.outputHeader) return true;
    if (oo.container == Container_KTX) {
        // KtxHeader (src/nvimage/KtxFile.h:113-131, KtxFile.cpp:17-34) as filled in by Context.cpp:870-1035
        struct KtxHeaderBytes {
            uint8_t identifier[12];
            uint32_t endianness, glType, glTypeSize, glFormat, glInternalFormat, glBaseInternalFormat;
            uint32_t pixelWidth, pixelHeight, pixelDepth, numberOfArrayElements, numberOfFaces, numberOfMipmapLevels, bytesOfKeyValueData;
        } kh;
        static_assert(sizeof(KtxHeaderBytes) == 64, "KTX header is 64 bytes");
        static const uint8_t ident[12] = {0xAB, 0x4B, 0x54, 0x58, 0x20, 0x31, 0x31, 0xBB, 0x0D, 0x0A, 0x1A, 0x0A};
        memset(&kh, 0, sizeof kh);
        memcpy(kh.identifier, ident, 12);
        kh.endianness = 0x04030201;
        kh.glType = 0;
        kh.glTypeSize = 1;
        kh.glFormat = 0;
        kh.numberOfFaces = 1;
        if (textureType == TextureType_Cube) kh.numberOfFaces = 6;
        else if (textureType == TextureType_3D) kh.pixelDepth = d;
        else if (textureType == TextureType_Array) kh.numberOfArrayElements = arraySize;
        kh.pixelWidth = w;
        kh.pixelHeight = h;
        kh.numberOfMipmapLevels = mipmapCount;
        const Format f = co.format;
        bool supported = true;
        if (f == Format_RGBA) {  // Context.cpp:904-951
            const unsigned bitcount = co.getBitCount();
            auto gl = [&](uint32_t type, uint32_t typeSize, uint32_t format, uint32_t internal) {
                kh.glType = type; kh.glTypeSize = typeSize; kh.glFormat = format; kh.glInternalFormat = internal; kh.glBaseInternalFormat = format;
            };
            if (co.pixelType == PixelType_Float) {
                if (co.rsize == 16 && co.gsize == 16 && co.bsize == 16 && co.asize == 16) gl(0x140B, 2, 0x1908, 0x881A);      // half, RGBA16F
                else if (co.rsize == 11 && co.gsize == 11 && co.bsize == 10 && co.asize == 0) gl(0x8C3B, 4, 0x1907, 0x8C3A);  // R11F_G11F_B10F
                else supported = false;
            } else if (bitcount == 16 && co.rsize == 16) {
                gl(0x1403, 2, 0x1903, 0x822A);  // unsigned short, R16
            } else if (co.bitcount == 24 && co.rmask == 0xFF0000 && co.gmask == 0xFF00 && co.bmask == 0xFF && co.amask == 0) {
                gl(0x1401, 1, 0x80E0, 0x8051);  // BGR, RGB8
            } else if (co.bitcount == 32 && co.rmask == 0xFF0000 && co.gmask == 0xFF00 && co.bmask == 0xFF && co.amask == 0xFF000000) {
                gl(0x1401, 1, 0x80E1, 0x8058);  // BGRA, RGBA8
            } else if (co.bitcount == 32 && co.rmask == 0xFF && co.gmask == 0xFF00 && co.bmask == 0xFF0000 && co.amask == 0xFF000000) {
                gl(0x1401, 1, 0x1908, 0x8058);  // RGBA, RGBA8
            } else {
                supported = false;
            }
        }
        else if (f == Format_DXT1 || f == Format_DXT1n) { kh.glInternalFormat = oo.srgb ? 0x8C4C : 0x83F0; kh.glBaseInternalFormat = 0x1907; }
        else if (f == Format_DXT1a) { kh.glInternalFormat = oo.srgb ? 0x8C4D : 0x83F1; kh.glBaseInternalFormat = 0x1908; }
        else if (f == Format_DXT3) { kh.glInternalFormat = oo.srgb ? 0x8C4E : 0x83F2; kh.glBaseInternalFormat = 0x1908; }
        else if (f == Format_DXT5 || f == Format_DXT5n || f == Format_BC3_RGBM) { kh.glInternalFormat = oo.srgb ? 0x8C4F : 0x83F3; kh.glBaseInternalFormat = 0x1908; }
        else if (f == Format_BC4) { kh.glInternalFormat = 0x8DBB; kh.glBaseInternalFormat = 0x1903; }
        else if (f == Format_BC5) { kh.glInternalFormat = 0x8DBD; kh.glBaseInternalFormat = 0x8227; }
        else if (f == Format_BC6) { kh.glInternalFormat = (co.pixelType == PixelType_Float) ? 0x8E8E : 0x8E8F; kh.glBaseInternalFormat = 0x1907; }
        else if (f == Format_BC7) { kh.glInternalFormat = oo.srgb ? 0x8E8D : 0x8E8C; kh.glBaseInternalFormat = 0x1908; }
        else supported = false;  // ETC / PVR: out of scope of this library
        if (!supported) {
            oo.error(Error_UnsupportedOutputFormat);
            return false;
        }
        const bool ok = oo.writeData(&kh, 64);
        if (!ok) oo.error(Error_FileWrite);
        return ok;
    }
    DDSHeaderBytes hd;
    memset(&hd, 0, sizeof hd);
    hd.fourcc = fourcc('D', 'D', 'S', ' ');
    hd.size = 124;
    hd.flags = 0x1 | 0x1000;  // CAPS | PIXELFORMAT
    hd.reserved[9] = fourcc('N', 'V', 'T', 'T');
    hd.reserved[10] = (2 << 16) | (1 << 8) | 2;
    hd.pf_size = 32;
    hd.caps1 = 0x1000;  // TEXTURE
    hd.reserved[7] = fourcc('U', 'V', 'E', 'R');
    hd.reserved[8] = (uint32_t)oo.version;
    if (textureType == TextureType_2D) { hd.resourceDimension = 3; hd.miscFlag = 0; hd.arraySize = 1; }
    else if (textureType == TextureType_Cube) { hd.caps1 |= 0x8; hd.caps2 = 0x200 | 0xFC00; hd.resourceDimension = 3; hd.miscFlag = 0x4; hd.arraySize = 1; }
    else if (textureType == TextureType_3D) { hd.caps2 = 0x200000; hd.resourceDimension = 4; hd.arraySize = 1; hd.flags |= 0x800000; hd.depth = d; }
    else { hd.resourceDimension = 3; hd.arraySize = arraySize; }
    hd.flags |= 0x4; hd.width = w;
    hd.flags |= 0x2; hd.height = h;
    if (mipmapCount == 0 || mipmapCount == 1) {
        hd.flags &= ~0x20000u;
        hd.mipmapcount = 1;
        hd.caps1 = (hd.caps2 == 0) ? 0x1000 : (0x1000 | 0x8);
    } else {
        hd.flags |= 0x20000;
        hd.mipmapcount = mipmapCount;
        hd.caps1 |= 0x8 | 0x400000;
    }
    bool supported = true;
    const Format f = co.format;
    if (oo.container == Container_DDS10) {
        hd.pf_flags = 0x4;
        hd.pf_fourcc = fourcc('D', 'X', '1', '0');
        if (f == Format_RGBA) {  // Context.cpp:649-682
            const unsigned bitcount = co.getBitCount();
            if (co.pixelType == PixelType_Float) {
                if (co.rsize == 16 && co.gsize == 16 && co.bsize == 16 && co.asize == 16) hd.dxgiFormat = 10;  // R16G16B16A16_FLOAT
                else if (co.rsize == 11 && co.gsize == 11 && co.bsize == 10 && co.asize == 0) hd.dxgiFormat = 26;  // R11G11B10_FLOAT
                else supported = false;
            } else if (bitcount == 16 && co.rsize == 16) {
                hd.dxgiFormat = 56;  // R16_UNORM
            } else {
                hd.dxgiFormat = find_dxgi_format(co.bitcount, co.rmask, co.gmask, co.bmask, co.amask);
                if (hd.dxgiFormat == 0) supported = false;
            }
        } else if (f == Format_DXT1 || f == Format_DXT1a || f == Format_DXT1n) {
            hd.dxgiFormat = oo.srgb ? 72 : 71;
            if (f == Format_DXT1a) hd.pf_flags |= 0x1;
            if (isNormalMap) hd.pf_flags |= 0x80000000u;
        } else if (f == Format_DXT3) hd.dxgiFormat = oo.srgb ? 75 : 74;
        else if (f == Format_DXT5 || f == Format_BC3_RGBM) hd.dxgiFormat = oo.srgb ? 78 : 77;
        else if (f == Format_DXT5n) { hd.dxgiFormat = 77; if (isNormalMap) hd.pf_flags |= 0x80000000u; }
        else if (f == Format_BC4) hd.dxgiFormat = 80;
        else if (f == Format_BC5) { hd.dxgiFormat = 83; if (isNormalMap) hd.pf_flags |= 0x80000000u; }
        else if (f == Format_BC6) hd.dxgiFormat = 95;  // always UF16 (Context.cpp:709-710)
        else if (f == Format_BC7) { hd.dxgiFormat = oo.srgb ? 99 : 98; if (isNormalMap) hd.pf_flags |= 0x80000000u; }
        else supported = false;
    } else {
        hd.flags &= ~0x8u;
        hd.flags |= 0x80000;  // LINEARSIZE
        hd.pitch = (uint32_t)image_size(w, h, d, co);
        hd.pf_flags = 0x4;
        if (f == Format_RGBA) {  // Context.cpp:726-793
            hd.flags &= ~0x80000u;
            hd.flags |= 0x8;  // PITCH
            hd.pitch = byte_pitch(w, co.getBitCount(), co.pitchAlignment);
            if (co.pixelType == PixelType_Float) {
                const int rs = co.rsize, gs = co.gsize, bs = co.bsize, as = co.asize;
                if (rs == 16 && gs == 0 && bs == 0 && as == 0) hd.pf_fourcc = 111;        // D3DFMT_R16F
                else if (rs == 16 && gs == 16 && bs == 0 && as == 0) hd.pf_fourcc = 112;  // D3DFMT_G16R16F
                else if (rs == 16 && gs == 16 && bs == 16 && as == 16) hd.pf_fourcc = 113;
                else if (rs == 32 && gs == 0 && bs == 0 && as == 0) hd.pf_fourcc = 114;
                else if (rs == 32 && gs == 32 && bs == 0 && as == 0) hd.pf_fourcc = 115;
                else if (rs == 32 && gs == 32 && bs == 32 && as == 32) hd.pf_fourcc = 116;
                else supported = false;
            } else {
                unsigned bitcount = co.getBitCount(), rmask = co.rmask, gmask = co.gmask, bmask = co.bmask, amask = co.amask;
                if (co.bitcount == 0 && bitcount <= 32) {
                    const unsigned ashift = 0, bshift = ashift + co.asize, gshift = bshift + co.bsize, rshift = gshift + co.gsize;
                    // the reference shifts ints by up to 32 here; x86 takes the count modulo 32, made explicit
                    rmask = ((1u << (co.rsize & 31)) - 1) << (rshift & 31);
                    gmask = ((1u << (co.gsize & 31)) - 1) << (gshift & 31);
                    bmask = ((1u << (co.bsize & 31)) - 1) << (bshift & 31);
                    amask = ((1u << (co.asize & 31)) - 1) << (ashift & 31);
                } else if (co.bitcount == 0) {
                    supported = false;
                }
                if (supported) {  // DDSHeader::setPixelFormat (DirectDrawSurface.cpp:730-783)
                    hd.pf_flags = 0;
                    if (rmask != 0 || gmask != 0 || bmask != 0) {
                        hd.pf_flags = (gmask == 0 && bmask == 0) ? 0x20000u : 0x40u;  // LUMINANCE : RGB
                        if (amask != 0) hd.pf_flags |= 0x1;                            // ALPHAPIXELS
                    } else if (amask != 0) {
                        hd.pf_flags |= 0x2;  // ALPHA
                    }
                    if (bitcount == 0) {
                        unsigned total = rmask | gmask | bmask | amask;
                        while (total != 0) { bitcount++; total >>= 1; }
                    }
                    hd.pf_fourcc = 0;
                    hd.pf_bitcount = bitcount;
                    hd.pf_rmask = rmask; hd.pf_gmask = gmask; hd.pf_bmask = bmask; hd.pf_amask = amask;
                }
            }
        }
        else if (f == Format_DXT1 || f == Format_DXT1a || f == Format_DXT1n) { hd.pf_fourcc = fourcc('D', 'X', 'T', '1'); if (isNormalMap) hd.pf_flags |= 0x80000000u; }
        else if (f == Format_DXT3) hd.pf_fourcc = fourcc('D', 'X', 'T', '3');
        else if (f == Format_DXT5 || f == Format_BC3_RGBM) hd.pf_fourcc = fourcc('D', 'X', 'T', '5');
        else if (f == Format_DXT5n) { hd.pf_fourcc = fourcc('D', 'X', 'T', '5'); if (isNormalMap) { hd.pf_flags |= 0x80000000u; hd.pf_bitcount = fourcc('A', '2', 'D', '5'); } }
        else if (f == Format_BC4) hd.pf_fourcc = fourcc('A', 'T', 'I', '1');
        else if (f == Format_BC5) { hd.pf_fourcc = fourcc('A', 'T', 'I', '2'); if (isNormalMap) { hd.pf_flags |= 0x80000000u; hd.pf_bitcount = fourcc('A', '2', 'X', 'Y'); } }
        else supported = false;  // BC6/BC7 need the DX10 header (Context.cpp:826-834)
        if (oo.srgb) hd.pf_flags |= 0x40000000u;
    }
    if (!supported) {
        oo.error(Error_UnsupportedOutputFormat);
        return false;
    }
    const int headerSize = (hd.pf_fourcc == fourcc('D', 'X', '1', '0')) ? 148 : 128;
    const bool ok = oo.writeData(&hd, headerSize);
    if (!ok) oo.error(Error_FileWrite);
    return ok;
}

bool Compressor::outputHeader(const Surface &img, int mipmapCount, const CompressionOptions &co, const OutputOptions &oo) const {
    return outputHeader(img.type(), img.width(), img.height(), img.depth(), 1, mipmapCount, img.isNormalMap(), co, oo);
}

int Compressor::estimateSize(int w, int h, int d, int mipmapCount, const CompressionOptions &co) const {
    int size = 0;
    for (int i = 0; i < mipmapCount; i++) {
        size += image_size(w, h, d, co.m);
        w = imax(1, w / 2); h = imax(1, h / 2); d = imax(1, d / 2);
    }
    return size;
}
int Compressor::estimateSize(const Surface &img, int mipmapCount, const CompressionOptions &co) const {
    return estimateSize(img.width(), img.height(), img.depth(), mipmapCount, co);
}
int Compressor::estimateSize(const InputOptions &io, const CompressionOptions &co) const {
    int w = io.m.width, h = io.m.height, d = io.m.depth;
    getTargetExtent(&w, &h, &d, io.m.maxExtent, io.m.roundMode, io.m.textureType);
    int mipmapCount = 1;
    if (io.m.generateMipmaps) {
        mipmapCount = countMipmaps(w, h, d);
        if (io.m.maxLevel > 0) mipmapCount = imin(mipmapCount, io.m.maxLevel);
    }
    return io.m.faceCount * estimateSize(w, h, d, mipmapCount, co);
}

// Compressor::Private::compress(AlphaMode,w,h,d,face,mip,rgba,...)  (Context.cpp:486-516)
static bool compress_level(const Compressor::Private &m, AlphaMode am, int w, int h, int d, int face, int mip, const float *rgba, int loc,
                           const CompressionOptions::Private &co, const OutputOptions::Private &oo) {
    const int size = image_size(w, h, d, co);
    if (oo.outputHandler) oo.outputHandler->beginImage(size, w, h, d, face, mip);
    bool ok = true;
    NvttbContext *ctx = g_gpu.get();
    if (!ctx || !m.cudaEnabled) {
        oo.error(Error_CudaError);
        ok = false;
    } else if (d == 1 && co.format == Format_RGBA) {
        NvttbPixelFormatDesc p;
        fill_pixel_format(&p, co, w, h);
        std::vector<unsigned char> out((size_t)size);
        const int rc = (size > 0 && (size_t)size == nvttb_pixel_format_level_size(&p))
                           ? nvttb_convert_level(ctx, &p, rgba, loc, out.data(), NVTTB_HOST, out.size())
                           : NVTTB_ERR_UNSUPPORTED_FEATURE;
        if (rc != NVTTB_OK) {
            oo.error((Error)(rc - 1));
            ok = false;
        } else {
            oo.writeData(out.data(), size);
        }
    } else if (d != 1 || !nvttb_format_supported(co.format, co.quality)) {
        oo.error(Error_UnsupportedFeature);  // reference: signals the error and still returns true (Context.cpp:504-515)
    } else {
        NvttbEncodeDesc e;
        fill_encode(&e, co, am, w, h);
        std::vector<unsigned char> out((size_t)size);
        const int rc = nvttb_encode_level(ctx, &e, rgba, loc, out.data(), NVTTB_HOST, out.size());
        if (rc != NVTTB_OK) {
            oo.error((Error)(rc - 1));
            ok = false;
        } else {
            oo.writeData(out.data(), size);
        }
    }
    if (oo.outputHandler) oo.outputHandler->endImage();
    return ok;
}

bool Compressor::compress(int w, int h, int d, int face, int mipmap, const float *rgba, const CompressionOptions &co, const OutputOptions &oo) const {
    return compress_level(m, AlphaMode_None, w, h, d, face, mipmap, rgba, NVTTB_HOST, co.m, oo.m);
}
bool Compressor::compress(const Surface &img, int face, int mipmap, const CompressionOptions &co, const OutputOptions &oo) const {
    if (img.isNull()) return false;
    return compress_level(m, img.alphaMode(), img.width(), img.height(), 1, face, mipmap, nvttb_surface_device_data(img.m->s), NVTTB_DEVICE, co.m, oo.m);
}

// Compressor::Private::quantize (Context.cpp:519-541)
static void quantize_surface(Surface &img, const CompressionOptions::Private &co) {
    if (co.enableColorDithering && co.format >= Format_BC1 && co.format <= Format_BC3) {
        img.quantize(0, 5, true, true);
        img.quantize(1, 6, true, true);
        img.quantize(2, 5, true, true);
    }
    if (co.enableColorDithering && co.format == Format_RGB) {
        img.quantize(0, co.rsize, true, true);
        img.quantize(1, co.gsize, true, true);
        img.quantize(2, co.bsize, true, true);
    }
    if (co.enableAlphaDithering && co.format == Format_RGB) img.quantize(3, co.asize, true, true);
    if (!co.enableAlphaDithering && co.binaryAlpha) img.binarize(3, float(co.alphaThreshold) / 255.0f, co.enableAlphaDithering);
}

// Compressor::Private::compress(InputOptions...)  (Context.cpp:217-346), DDS order.
bool Compressor::process(const InputOptions &inputOptions, const CompressionOptions &compressionOptions, const OutputOptions &outputOptions) const {
    const InputOptions::Private &io = inputOptions.m;
    const CompressionOptions::Private &co = compressionOptions.m;
    const OutputOptions::Private &oo = outputOptions.m;
    if (oo.outputHandler == nullptr) {  // hasValidOutputHandler
        oo.error(Error_FileOpen);
        return false;
    }
    const int faceCount = io.faceCount;
    int width = io.width, height = io.height, depth = io.depth;
    const int arraySize = io.textureType == TextureType_Array ? faceCount : 1;
    getTargetExtent(&width, &height, &depth, io.maxExtent, io.roundMode, io.textureType);
    const bool canUseSourceImages = (io.width == width && io.height == height && io.depth == depth);
    int mipmapCount = 1;
    if (io.generateMipmaps) {
        mipmapCount = countMipmaps(width, height, depth);
        if (io.maxLevel > 0) mipmapCount = imin(mipmapCount, io.maxLevel);
    }
    if (!outputHeader(io.textureType, width, height, depth, arraySize, mipmapCount, io.isNormalMap, compressionOptions, outputOptions)) return false;

    NvttbContext *ctx = g_gpu.get();
    if (!ctx || !m.cudaEnabled) {
        oo.error(Error_CudaError);
        return false;
    }
    // Compressor::Private::quantize (Context.cpp:519-541): colour dithering acts on BC1..BC3 (5/6/5 bits, Floyd-Steinberg);
    // alpha dithering only on Format_RGB, i.e. never here; binary alpha = non-dithered binarize, and only when alpha dithering
    // is off (so nvcompress's settings for -bc1a / -bc2 are both no-ops, as in the reference).
    const bool isRGB = co.format == Format_RGB;
    const bool colorDither = co.enableColorDithering && ((co.format >= Format_BC1 && co.format <= Format_BC3) || isRGB);
    const bool binarizeAlpha = !co.enableAlphaDithering && co.binaryAlpha;
    const bool quantizeStep = colorDither || binarizeAlpha || (isRGB && co.enableAlphaDithering);
    if (depth != 1 || !(isRGB || nvttb_format_supported(co.format, co.quality))) {
        oo.error(Error_UnsupportedFeature);
        return false;
    }
    for (int f = 0; f < faceCount; f++)
        if (io.images[f].empty()) { oo.error(Error_InvalidInput); return false; }
    bool userMips = false;
    for (int i = faceCount; i < io.imageCount; i++) userMips = userMips || !io.images[i].empty();

    if (oo.container == Container_KTX) {
        // KTX stores the faces of one mip level together: mip-major order, every level prefixed with its byte size
        // (Context.cpp:347-472).  Kept quirks: the Kaiser parameters are passed as {stretch, alpha}, and normal-map mips
        // are renormalised without the expand / pack pair of the DDS loop.
        std::vector<Surface> images;
        int w = width, h = height;
        uint32_t imageSize = (uint32_t)ktx_level_size(w, h, co) * faceCount;
        oo.writeData(&imageSize, 4);
        for (int f = 0; f < faceCount; f++) {
            Surface s;
            s.setWrapMode(io.wrapMode);
            s.setAlphaMode(io.alphaMode);
            s.setNormalMap(io.isNormalMap);
            if (!s.setImage(io.inputFormat, io.width, io.height, 1, io.images[f].data())) { oo.error(Error_CudaError); return false; }
            if (io.convertToNormalMap) {
                s.toGreyScale(io.heightFactors[0], io.heightFactors[1], io.heightFactors[2], io.heightFactors[3]);
                s.toNormalMap(io.bumpFrequencyScale[0], io.bumpFrequencyScale[1], io.bumpFrequencyScale[2], io.bumpFrequencyScale[3]);
            }
            if (!s.isNormalMap()) s.toLinear(io.inputGamma);
            s.resize(w, h, 1, ResizeFilter_Box);
            Surface tmp = s;
            if (!s.isNormalMap()) tmp.toGamma(io.outputGamma);
            if (quantizeStep) quantize_surface(tmp, co);
            if (!compress(tmp, f, 0, compressionOptions, outputOptions)) return false;
            images.push_back(s);
        }
        static const unsigned char padding[3] = {0, 0, 0};
        for (int mip = 1; mip < mipmapCount; mip++) {
            w = imax(1, w / 2);
            h = imax(1, h / 2);
            imageSize = (uint32_t)ktx_level_size(w, h, co) * faceCount;
            oo.writeData(&imageSize, 4);
            for (int f = 0; f < faceCount; f++) {
                Surface &img = images[f];
                const int idx = mip * faceCount + f;
                // mipChainBroken[] is never set in the reference: a user-supplied level is used whenever it exists
                if (idx < io.imageCount && !io.images[idx].empty()) {
                    img.setImage(io.inputFormat, w, h, 1, io.images[idx].data());
                    if (!img.isNormalMap()) img.toLinear(io.inputGamma);
                } else if (io.mipmapFilter == MipmapFilter_Kaiser) {
                    const float params[2] = {io.kaiserStretch, io.kaiserAlpha};
                    img.buildNextMipmap(MipmapFilter_Kaiser, io.kaiserWidth, params);
                } else {
                    img.buildNextMipmap(io.mipmapFilter);
                }
                Surface tmp;
                if (img.isNormalMap()) {
                    if (io.normalizeMipmaps) img.normalizeNormalMap();
                    tmp = img;
                } else {
                    tmp = img;
                    tmp.toGamma(io.outputGamma);
                }
                if (quantizeStep) quantize_surface(tmp, co);
                if (!compress(tmp, f, mip, compressionOptions, outputOptions)) return false;
            }
            const int mipPadding = 3 - ((imageSize + 3) % 4);
            if (mipPadding != 0) oo.writeData(padding, mipPadding);
        }
        return true;
    }

    if (canUseSourceImages && !userMips && !quantizeStep && !isRGB) {
        // fused device pipeline
        NvttbProcessDesc d;
        memset(&d, 0, sizeof d);
        d.inputFormat = io.inputFormat;
        d.width = width; d.height = height; d.faceCount = faceCount;
        d.wrapMode = io.wrapMode;
        d.mipmapFilter = io.mipmapFilter;
        d.generateMipmaps = io.generateMipmaps ? 1 : 0;
        d.maxLevel = io.maxLevel;
        d.kaiserWidth = io.kaiserWidth; d.kaiserAlpha = io.kaiserAlpha; d.kaiserStretch = io.kaiserStretch;
        d.inputGamma = io.inputGamma; d.outputGamma = io.outputGamma;
        d.isNormalMap = io.isNormalMap; d.convertToNormalMap = io.convertToNormalMap; d.normalizeMipmaps = io.normalizeMipmaps;
        for (int i = 0; i < 4; i++) { d.heightFactors[i] = io.heightFactors[i]; d.bumpFrequencyScale[i] = io.bumpFrequencyScale[i]; }
        d.alphaMode = io.alphaMode;
        fill_encode(&d.encode, co, io.alphaMode, width, height);
        std::vector<const void *> ptrs(faceCount);
        for (int f = 0; f < faceCount; f++) ptrs[f] = io.images[f].data();
        EmitCtx ec{&oo};
        // every visible GPU works on the job: one large image is block-row sharded, faces / array slices are dealt out
        const std::vector<NvttbContext *> &pool = g_gpu.pool();
        const int rc = pool.size() > 1 ? nvttb_process_multi(pool.data(), (int)pool.size(), &d, ptrs.data(), emit_cb, &ec)
                                       : nvttb_process(ctx, &d, ptrs.data(), NVTTB_HOST, emit_cb, &ec);
        if (rc != NVTTB_OK) {
            oo.error((Error)(rc - 1));
            return false;
        }
        return true;
    }

    // general path (input resize and/or user-supplied mip levels): the reference's loop, one Surface op at a time
    Surface img;
    img.setWrapMode(io.wrapMode);
    img.setAlphaMode(io.alphaMode);
    img.setNormalMap(io.isNormalMap);
    for (int f = 0; f < faceCount; f++) {
        int w = width, h = height;
        bool canUseSourceImagesForThisFace = canUseSourceImages;
        if (!img.setImage(io.inputFormat, io.width, io.height, 1, io.images[f].data())) { oo.error(Error_CudaError); return false; }
        if (io.convertToNormalMap) {
            img.toGreyScale(io.heightFactors[0], io.heightFactors[1], io.heightFactors[2], io.heightFactors[3]);
            img.toNormalMap(io.bumpFrequencyScale[0], io.bumpFrequencyScale[1], io.bumpFrequencyScale[2], io.bumpFrequencyScale[3]);
        }
        if (!img.isNormalMap()) img.toLinear(io.inputGamma);
        img.resize(w, h, 1, ResizeFilter_Box);
        {
            Surface tmp = img;
            if (!img.isNormalMap()) tmp.toGamma(io.outputGamma);
            if (quantizeStep) quantize_surface(tmp, co);
            if (!compress(tmp, f, 0, compressionOptions, outputOptions)) return false;
        }
        for (int mip = 1; mip < mipmapCount; mip++) {
            w = imax(1, w / 2);
            h = imax(1, h / 2);
            const int idx = mip * faceCount + f;
            bool useSourceImages = false;
            if (canUseSourceImagesForThisFace) {
                if (io.images[idx].empty()) canUseSourceImagesForThisFace = false;
                else useSourceImages = true;
            }
            if (useSourceImages) {
                img.setImage(io.inputFormat, w, h, 1, io.images[idx].data());
                if (!img.isNormalMap()) img.toLinear(io.inputGamma);
            } else if (io.mipmapFilter == MipmapFilter_Kaiser) {
                const float params[2] = {io.kaiserAlpha, io.kaiserStretch};
                img.buildNextMipmap(MipmapFilter_Kaiser, io.kaiserWidth, params);
            } else {
                img.buildNextMipmap(io.mipmapFilter);
            }
            Surface tmp;
            if (img.isNormalMap()) {
                if (io.normalizeMipmaps) {
                    img.expandNormals();
                    img.normalizeNormalMap();
                    img.packNormals();
                }
                tmp = img;
            } else {
                tmp = img;
                tmp.toGamma(io.outputGamma);
            }
            if (quantizeStep) quantize_surface(tmp, co);
            if (!compress(tmp, f, mip, compressionOptions, outputOptions)) return false;
        }
    }
    return true;
}

float nvtt::rmsError(const Surface &reference, const Surface &img) {
    float v = FLT_MAX;
    if (reference.m->s && img.m->s) nvttb_rms_error(reference.m->s, img.m->s, &v);
    return v;
}
float nvtt::cieLabError(const Surface &reference, const Surface &img) {
    float v = FLT_MAX;
    if (reference.m->s && img.m->s) nvttb_cielab_error(reference.m->s, img.m->s, &v);
    return v;
}
float nvtt::angularError(const Surface &reference, const Surface &img) {
    float v = FLT_MAX;
    if (reference.m->s && img.m->s) nvttb_angular_error(reference.m->s, img.m->s, &v);
    return v;
}
float nvtt::rmsAlphaError(const Surface &reference, const Surface &img) {
    float v = FLT_MAX;
    if (reference.m->s && img.m->s) nvttb_rms_alpha_error(reference.m->s, img.m->s, &v);
    return v;
}

unsigned int nvtt::version() { return NVTT_VERSION; }
const char *nvtt::errorString(Error e) {
    static const char *const names[] = {"Unknown error", "Invalid input", "Unsupported feature", "CUDA error", "Error opening file", "Error writing through output handler", "Unsupported output format"};
    return (unsigned)e < 7 ? names[e] : "Invalid error";
}
