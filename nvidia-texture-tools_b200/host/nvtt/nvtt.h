// nvtt.h — B200-native drop-in for the hot-path subset of NVTT 2.1.2's public C++ API.
//
// Same namespace, class names, method names, argument meaning, enum values and error behaviour as the reference's
// src/nvtt/nvtt.h:80-447 (Format, Quality, CompressionOptions, InputOptions, OutputHandler, ErrorHandler,
// OutputOptions, TaskDispatcher, Compressor) and the Surface methods that lie on the BCn + mip path
// (src/nvtt/nvtt.h:472-605).  Everything is implemented on top of the C ABI in include/nvtt_b200.h; there is no
// CPU code path.  Source compatible: a caller written against the reference's header (e.g. oracle/ref_harness.cpp,
// nvcompress' use of Compressor::process) compiles unchanged against this one.
#ifndef NVTT_B200_NVTT_H
#define NVTT_B200_NVTT_H

#define NVTT_VERSION 20102
#define NVTT_API __attribute__((visibility("default")))

namespace nvtt {
struct Surface;

enum Format {
    Format_RGB, Format_RGBA = Format_RGB,
    Format_DXT1, Format_DXT1a, Format_DXT3, Format_DXT5, Format_DXT5n,
    Format_BC1 = Format_DXT1, Format_BC1a = Format_DXT1a, Format_BC2 = Format_DXT3, Format_BC3 = Format_DXT5, Format_BC3n = Format_DXT5n,
    Format_BC4, Format_BC5, Format_DXT1n, Format_CTX1, Format_BC6, Format_BC7, Format_BC3_RGBM,
    Format_ETC1, Format_ETC2_R, Format_ETC2_RG, Format_ETC2_RGB, Format_ETC2_RGBA, Format_ETC2_RGB_A1, Format_ETC2_RGBM,
    Format_PVR_2BPP_RGB, Format_PVR_4BPP_RGB, Format_PVR_2BPP_RGBA, Format_PVR_4BPP_RGBA,
    Format_Count
};
enum PixelType { PixelType_UnsignedNorm = 0, PixelType_SignedNorm = 1, PixelType_UnsignedInt = 2, PixelType_SignedInt = 3, PixelType_Float = 4, PixelType_UnsignedFloat = 5, PixelType_SharedExp = 6 };
enum Quality { Quality_Fastest, Quality_Normal, Quality_Production, Quality_Highest };
enum Decoder { Decoder_D3D10, Decoder_D3D9, Decoder_NV5x };
enum WrapMode { WrapMode_Clamp, WrapMode_Repeat, WrapMode_Mirror };
enum TextureType { TextureType_2D, TextureType_Cube, TextureType_3D, TextureType_Array };
enum InputFormat { InputFormat_BGRA_8UB, InputFormat_RGBA_16F, InputFormat_RGBA_32F, InputFormat_R_32F };
enum MipmapFilter { MipmapFilter_Box, MipmapFilter_Triangle, MipmapFilter_Kaiser };
enum ResizeFilter { ResizeFilter_Box, ResizeFilter_Triangle, ResizeFilter_Kaiser, ResizeFilter_Mitchell };
enum RoundMode { RoundMode_None, RoundMode_ToNextPowerOfTwo, RoundMode_ToNearestPowerOfTwo, RoundMode_ToPreviousPowerOfTwo, RoundMode_ToNextMultipleOfFour, RoundMode_ToNearestMultipleOfFour, RoundMode_ToPreviousMultipleOfFour };
enum AlphaMode { AlphaMode_None, AlphaMode_Transparency, AlphaMode_Premultiplied };
enum Error { Error_Unknown, Error_InvalidInput, Error_UnsupportedFeature, Error_CudaError, Error_FileOpen, Error_FileWrite, Error_UnsupportedOutputFormat, Error_Count };
enum Container { Container_DDS, Container_DDS10, Container_KTX };
enum ToneMapper { ToneMapper_Linear, ToneMapper_Reindhart, ToneMapper_Halo, ToneMapper_Lightmap };

struct CompressionOptions {
    NVTT_API CompressionOptions();
    NVTT_API ~CompressionOptions();
    NVTT_API void reset();
    NVTT_API void setFormat(Format format);
    NVTT_API void setQuality(Quality quality);
    NVTT_API void setColorWeights(float red, float green, float blue, float alpha = 1.0f);
    NVTT_API void setPixelFormat(unsigned int bitcount, unsigned int rmask, unsigned int gmask, unsigned int bmask, unsigned int amask);
    NVTT_API void setPixelFormat(unsigned char rsize, unsigned char gsize, unsigned char bsize, unsigned char asize);
    NVTT_API void setPixelType(PixelType pixelType);
    NVTT_API void setPitchAlignment(int pitchAlignment);
    NVTT_API void setQuantization(bool colorDithering, bool alphaDithering, bool binaryAlpha, int alphaThreshold = 127);
    NVTT_API void setRGBMThreshold(float min_m);
    NVTT_API void setTargetDecoder(Decoder decoder);
    NVTT_API void setExternalCompressor(const char *name);  // stored and ignored, as in a stock reference build (no HAVE_* codec)
    NVTT_API Format format() const;
    struct Private;
    Private &m;
private:
    CompressionOptions(const CompressionOptions &);
    void operator=(const CompressionOptions &);
};

struct InputOptions {
    NVTT_API InputOptions();
    NVTT_API ~InputOptions();
    NVTT_API void reset();
    NVTT_API void setTextureLayout(TextureType type, int w, int h, int d = 1, int arraySize = 1);
    NVTT_API void resetTextureLayout();
    NVTT_API bool setMipmapData(const void *data, int w, int h, int d = 1, int face = 0, int mipmap = 0);  // copies
    NVTT_API void setFormat(InputFormat format);
    NVTT_API void setAlphaMode(AlphaMode alphaMode);
    NVTT_API void setGamma(float inputGamma, float outputGamma);
    NVTT_API void setWrapMode(WrapMode mode);
    NVTT_API void setMipmapFilter(MipmapFilter filter);
    NVTT_API void setMipmapGeneration(bool enabled, int maxLevel = -1);
    NVTT_API void setKaiserParameters(float width, float alpha, float stretch);
    NVTT_API void setNormalMap(bool b);
    NVTT_API void setConvertToNormalMap(bool convert);
    NVTT_API void setHeightEvaluation(float redScale, float greenScale, float blueScale, float alphaScale);
    NVTT_API void setNormalFilter(float sm, float medium, float big, float large);
    NVTT_API void setNormalizeMipmaps(bool b);
    NVTT_API void setMaxExtents(int d);
    NVTT_API void setRoundMode(RoundMode mode);
    struct Private;
    Private &m;
private:
    InputOptions(const InputOptions &);
    void operator=(const InputOptions &);
};

struct OutputHandler {
    virtual ~OutputHandler() {}
    virtual void beginImage(int size, int width, int height, int depth, int face, int miplevel) = 0;
    virtual bool writeData(const void *data, int size) = 0;
    virtual void endImage() = 0;
};
struct ErrorHandler {
    virtual ~ErrorHandler() {}
    virtual void error(Error e) = 0;
};

struct OutputOptions {
    NVTT_API OutputOptions();
    NVTT_API ~OutputOptions();
    NVTT_API void reset();
    NVTT_API void setFileName(const char *fileName);
    NVTT_API void setFileHandle(void *fp);
    NVTT_API void setOutputHandler(OutputHandler *outputHandler);
    NVTT_API void setErrorHandler(ErrorHandler *errorHandler);
    NVTT_API void setOutputHeader(bool outputHeader);
    NVTT_API void setContainer(Container container);
    NVTT_API void setUserVersion(int version);
    NVTT_API void setSrgbFlag(bool b);
    struct Private;
    Private &m;
private:
    OutputOptions(const OutputOptions &);
    void operator=(const OutputOptions &);
};

typedef void Task(void *context, int id);
struct TaskDispatcher {
    virtual ~TaskDispatcher() {}
    virtual void dispatch(Task *task, void *context, int count) = 0;
};

struct Compressor {
    NVTT_API Compressor();
    NVTT_API ~Compressor();
    // The GPU is the only implementation: enableCudaAcceleration(false) is accepted but cannot select a CPU path; without a
    // CUDA device every compress call reports Error_CudaError and returns false.
    NVTT_API void enableCudaAcceleration(bool enable);
    NVTT_API bool isCudaAccelerationEnabled() const;
    NVTT_API void setTaskDispatcher(TaskDispatcher *disp);  // accepted and ignored: blocks are scheduled by the GPU
    NVTT_API bool process(const InputOptions &inputOptions, const CompressionOptions &compressionOptions, const OutputOptions &outputOptions) const;
    NVTT_API int estimateSize(const InputOptions &inputOptions, const CompressionOptions &compressionOptions) const;
    NVTT_API bool outputHeader(const Surface &img, int mipmapCount, const CompressionOptions &compressionOptions, const OutputOptions &outputOptions) const;
    NVTT_API bool compress(const Surface &img, int face, int mipmap, const CompressionOptions &compressionOptions, const OutputOptions &outputOptions) const;
    NVTT_API int estimateSize(const Surface &img, int mipmapCount, const CompressionOptions &compressionOptions) const;
    NVTT_API bool outputHeader(TextureType type, int w, int h, int d, int arraySize, int mipmapCount, bool isNormalMap, const CompressionOptions &compressionOptions, const OutputOptions &outputOptions) const;
    NVTT_API bool compress(int w, int h, int d, int face, int mipmap, const float *rgba, const CompressionOptions &compressionOptions, const OutputOptions &outputOptions) const;
    NVTT_API int estimateSize(int w, int h, int d, int mipmapCount, const CompressionOptions &compressionOptions) const;
    struct Private;
    Private &m;
private:
    Compressor(const Compressor &);
    void operator=(const Compressor &);
};
typedef Compressor Context;

// Device-resident surface: the methods of nvtt::Surface that are on the BCn + mip path.
struct Surface {
    NVTT_API Surface();
    NVTT_API Surface(const Surface &img);
    NVTT_API ~Surface();
    NVTT_API void operator=(const Surface &img);
    NVTT_API void setWrapMode(WrapMode mode);
    NVTT_API void setAlphaMode(AlphaMode alphaMode);
    NVTT_API void setNormalMap(bool isNormalMap);
    NVTT_API bool isNull() const;
    NVTT_API int width() const;
    NVTT_API int height() const;
    NVTT_API int depth() const;
    NVTT_API TextureType type() const;
    NVTT_API WrapMode wrapMode() const;
    NVTT_API AlphaMode alphaMode() const;
    NVTT_API bool isNormalMap() const;
    NVTT_API int countMipmaps() const;
    NVTT_API const float *data() const;  // host copy of the planar fp32 RGBA data, refreshed on demand
    NVTT_API bool load(const char *fileName, bool *hasAlpha = 0);  // decodes through the installed SurfaceLoader (see below)
    NVTT_API void range(int channel, float *rangeMin, float *rangeMax, int alpha_channel = -1, float alpha_ref = 0.f) const;
    NVTT_API bool setImage(InputFormat format, int w, int h, int d, const void *data);
    NVTT_API bool setImage(InputFormat format, int w, int h, int d, const void *r, const void *g, const void *b, const void *a);
    NVTT_API bool setImage2D(Format format, Decoder decoder, int w, int h, const void *data);
    NVTT_API void resize(int w, int h, int d, ResizeFilter filter);
    NVTT_API void resize(int w, int h, int d, ResizeFilter filter, float filterWidth, const float *params = 0);
    NVTT_API bool buildNextMipmap(MipmapFilter filter, int min_size = 1);
    NVTT_API bool buildNextMipmap(MipmapFilter filter, float filterWidth, const float *params = 0, int min_size = 1);
    NVTT_API bool canMakeNextMipmap(int min_size = 1);
    NVTT_API void toLinear(float gamma);
    NVTT_API void toGamma(float gamma);
    NVTT_API void toGreyScale(float redScale, float greenScale, float blueScale, float alphaScale);
    NVTT_API void toNormalMap(float sm, float medium, float big, float large);
    NVTT_API void binarize(int channel, float threshold, bool dither);
    NVTT_API void quantize(int channel, int bits, bool exactEndPoints, bool dither);
    NVTT_API void scaleBias(int channel, float scale, float bias);
    NVTT_API void clamp(int channel, float low = 0.0f, float high = 1.0f);
    NVTT_API void toRGBM(float range = 1.0f, float threshold = 0.25f);
    NVTT_API void toneMap(ToneMapper tm, float *parameters);
    NVTT_API void normalizeNormalMap();
    NVTT_API void packNormals(float scale = 0.5f, float bias = 0.5f);
    NVTT_API void expandNormals(float scale = 2.0f, float bias = -1.0f);
    struct Private;
    Private *m;
};

// src/nvtt/nvtt.h:699-700
NVTT_API float rmsError(const Surface &reference, const Surface &img);
NVTT_API float rmsAlphaError(const Surface &reference, const Surface &img);
NVTT_API float angularError(const Surface &reference, const Surface &img);
NVTT_API float cieLabError(const Surface &reference, const Surface &img);

// Image FILE decoding is not part of the GPU path and this library links no image codecs: Surface::load hands the file name
// to the loader installed here (an application - or the nvcompress build recipe - provides one on top of its own image reader;
// it fills the surface with setImage / setImage2D).  Without a loader Surface::load returns false.  Not in the reference.
typedef bool (*SurfaceLoader)(Surface &dst, const char *fileName, bool *hasAlpha);
NVTT_API void setSurfaceLoader(SurfaceLoader loader);

NVTT_API unsigned int version();
NVTT_API const char *errorString(Error e);
}  // namespace nvtt
#endif
