// TEST INFRASTRUCTURE ONLY.  Compiles integration/B200Compressor.h - the binding INTEGRATION.md section B asks a maintainer of
// the reference to add - against the REFERENCE's own internal headers and drives it the way Compressor::Private::compress
// does (Context.cpp:486-516): options objects of the reference, its OutputOptions::Private callbacks, one level in.
#include <nvtt/nvtt.h>               // the reference's public header (src/nvtt on the include path, not host/)
#include <nvtt/B200Compressor.h>     // integration/B200Compressor.h, found through -I integration/..
#include <string.h>
#include <vector>

namespace {
struct MemHandler : public nvtt::OutputHandler {
    std::vector<unsigned char> buf;
    virtual void beginImage(int, int, int, int, int, int) {}
    virtual bool writeData(const void * data, int size) { buf.insert(buf.end(), (const unsigned char *)data, (const unsigned char *)data + size); return true; }
    virtual void endImage() {}
};
struct ErrCount : public nvtt::ErrorHandler {
    int n; ErrCount() : n(0) {}
    virtual void error(nvtt::Error) { n++; }
};
}

// pixelFormat: 0 = none (BCn), 1 = setPixelFormat(32, 0xFF0000, 0xFF00, 0xFF, 0xFF000000), 2 = setPixelFormat(5, 6, 5, 0)
extern "C" long integ_compress_level(int format, int quality, int alphaMode, int w, int h, const float * rgba, int pixelFormat, unsigned char * out, long cap)
{
    static NvttbContext * ctx = NULL;
    if (!ctx && nvttb_context_create(0, &ctx) != NVTTB_OK) return -10;
    nvtt::CompressionOptions co;
    co.setFormat((nvtt::Format)format);
    co.setQuality((nvtt::Quality)quality);
    if (pixelFormat == 1) co.setPixelFormat(32, 0xFF0000, 0xFF00, 0xFF, 0xFF000000);
    if (pixelFormat == 2) co.setPixelFormat(5, 6, 5, 0);
    nvtt::OutputOptions oo;
    MemHandler mh;
    ErrCount eh;
    oo.setOutputHandler(&mh);
    oo.setErrorHandler(&eh);
    nv::B200Compressor c(ctx);
    oo.m.beginImage(0, w, h, 1, 0, 0);
    c.compress((nvtt::AlphaMode)alphaMode, (uint)w, (uint)h, 1, rgba, NULL, co.m, oo.m);
    oo.m.endImage();
    if (eh.n) return -1;
    if ((long)mh.buf.size() > cap) return -2;
    memcpy(out, mh.buf.data(), mh.buf.size());
    return (long)mh.buf.size();
}
