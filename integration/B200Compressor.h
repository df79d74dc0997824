// B200Compressor.h - the file a maintainer of the reference adds to src/nvtt/ to route Compressor::Private::compress
// (src/nvtt/Context.cpp:486-516) to the B200 library through its C ABI (include/nvtt_b200.h).  It is written against the
// REFERENCE's internal headers (Compressor.h, CompressionOptions.h, OutputOptions.h) and is compiled and run by
// tests/build_integration.sh + tests/test_gpu_integration.py (linked with the unmodified reference objects and libnvtt_b200.so).
//
//   Compressor::Compressor():            nvttb_context_create(0, &m.b200);            (next to icbc::init_dxt1(), Context.cpp:72)
//   Compressor::Private::compress(...):  instead of the "#if defined HAVE_CUDA ... chooseGpuCompressor" block (:492-498):
//       if (m.b200 && (compressionOptions.format == Format_RGBA || nvttb_format_supported(compressionOptions.format, compressionOptions.quality)))
//           compressor = new nv::B200Compressor(m.b200);
#ifndef NVTT_B200COMPRESSOR_H
#define NVTT_B200COMPRESSOR_H

#include "Compressor.h"
#include "CompressionOptions.h"
#include "OutputOptions.h"
#include "nvcore/Array.inl"

#include <nvtt_b200.h>

namespace nv {

struct B200Compressor : public CompressorInterface {
    NvttbContext * ctx;
    explicit B200Compressor(NvttbContext * c) : ctx(c) {}

    virtual void compress(nvtt::AlphaMode alphaMode, uint w, uint h, uint d, const float * rgba, nvtt::TaskDispatcher *,
                          const nvtt::CompressionOptions::Private & co, const nvtt::OutputOptions::Private & oo)
    {
        if (d != 1) { oo.error(nvtt::Error_UnsupportedFeature); return; }
        int rc;
        Array<uint8> mem;
        if (co.format == nvtt::Format_RGBA) {
            // chooseCpuCompressor -> PixelFormatConverter (Context.cpp:1040-1043, CompressorRGB.cpp:410-575)
            NvttbPixelFormatDesc p;
            p.pixelType = co.pixelType;
            p.bitcount = co.bitcount;
            p.rmask = co.rmask;  p.gmask = co.gmask;  p.bmask = co.bmask;  p.amask = co.amask;
            p.rsize = co.rsize;  p.gsize = co.gsize;  p.bsize = co.bsize;  p.asize = co.asize;
            p.pitchAlignment = co.pitchAlignment;
            p.width = int(w);  p.height = int(h);
            const size_t size = nvttb_pixel_format_level_size(&p);
            if (size == 0) { oo.error(nvtt::Error_UnsupportedFeature); return; }
            mem.resize(uint(size));
            rc = nvttb_convert_level(ctx, &p, rgba, NVTTB_HOST, mem.buffer(), NVTTB_HOST, size);
        }
        else {
            NvttbEncodeDesc e;
            e.format = co.format;  e.quality = co.quality;  e.alphaMode = alphaMode;  e.pixelType = co.pixelType;
            e.colorWeights[0] = co.colorWeight.x;  e.colorWeights[1] = co.colorWeight.y;
            e.colorWeights[2] = co.colorWeight.z;  e.colorWeights[3] = co.colorWeight.w;
            e.width = int(w);  e.height = int(h);
            e.applyToGamma = 0;
            e.rgbmThreshold = co.rgbmThreshold;
            const size_t size = nvttb_level_size(co.format, int(w), int(h));
            if (size == 0) { oo.error(nvtt::Error_UnsupportedFeature); return; }
            mem.resize(uint(size));
            rc = nvttb_encode_level(ctx, &e, rgba, NVTTB_HOST, mem.buffer(), NVTTB_HOST, size);
        }
        if (rc != NVTTB_OK) { oo.error(nvtt::Error(rc - 1)); return; }
        oo.writeData(mem.buffer(), int(mem.count()));  // one writeData per mip level, like BlockCompressor.cpp:110,202
    }
};

} // nv namespace

#endif // NVTT_B200COMPRESSOR_H
