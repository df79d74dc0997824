// TEST INFRASTRUCTURE ONLY.  The image-FILE side of nvcompress when it is built against the B200 drop-in library: the
// reference's nvcompress reads its input with the reference's own nvimage reader (ImageIO / DirectDrawSurface, linked in
// "for file IO only"), and nvtt::Surface::load - which in the reference calls the same reader (src/nvtt/Surface.cpp:568-659) -
// is routed to it through nvtt::setSurfaceLoader.  Nothing here computes: decoded texels go straight into
// Surface::setImage / setImage2D of the drop-in, i.e. onto the GPU.
#include <nvtt/nvtt.h>  // the drop-in header (host/nvtt/nvtt.h comes first on the include path)

#include <nvcore/Ptr.h>
#include <nvcore/StrLib.h>
#include <nvimage/DirectDrawSurface.h>
#include <nvimage/FloatImage.h>
#include <nvimage/Image.h>
#include <nvimage/ImageIO.h>

#include <stdlib.h>
#include <string.h>
#include <vector>

namespace {
bool load_surface(nvtt::Surface &dst, const char *fileName, bool *hasAlpha) {
    nv::AutoPtr<nv::FloatImage> img(nv::ImageIO::loadFloat(fileName));
    if (img == NULL) {
        if (!nv::strEqual(nv::Path::extension(fileName), ".dds")) return false;
        nv::DirectDrawSurface dds;
        if (!dds.load(fileName)) return false;
        if (dds.header.isBlockFormat()) {
            const int w = dds.surfaceWidth(0), h = dds.surfaceHeight(0);
            const unsigned size = dds.surfaceSize(0);
            std::vector<unsigned char> data(size);
            dds.readSurface(0, 0, data.data(), size);
            nvtt::Format f;
            if (dds.header.hasDX10Header()) {
                const unsigned fmt = dds.header.header10.dxgiFormat;
                if (fmt == nv::DXGI_FORMAT_BC1_UNORM || fmt == nv::DXGI_FORMAT_BC1_UNORM_SRGB) f = nvtt::Format_BC1;
                else if (fmt == nv::DXGI_FORMAT_BC2_UNORM || fmt == nv::DXGI_FORMAT_BC2_UNORM_SRGB) f = nvtt::Format_BC2;
                else if (fmt == nv::DXGI_FORMAT_BC3_UNORM || fmt == nv::DXGI_FORMAT_BC3_UNORM_SRGB) f = nvtt::Format_BC3;
                else if (fmt == nv::DXGI_FORMAT_BC6H_UF16) f = nvtt::Format_BC6;
                else if (fmt == nv::DXGI_FORMAT_BC7_UNORM || fmt == nv::DXGI_FORMAT_BC7_UNORM_SRGB) f = nvtt::Format_BC7;
                else return false;
            } else {
                const unsigned fourcc = dds.header.pf.fourcc;
                if (fourcc == nv::FOURCC_DXT1) f = nvtt::Format_BC1;
                else if (fourcc == nv::FOURCC_DXT3) f = nvtt::Format_BC2;
                else if (fourcc == nv::FOURCC_DXT5) f = nvtt::Format_BC3;
                else return false;
            }
            return dst.setImage2D(f, nvtt::Decoder_D3D10, w, h, data.data());
        }
        nv::Image im;
        nv::imageFromDDS(&im, dds, /*face=*/0, /*mipmap=*/0);
        return dst.setImage(nvtt::InputFormat_BGRA_8UB, im.width, im.height, im.depth, im.pixels());
    }
    if (hasAlpha != NULL) *hasAlpha = (img->componentCount() == 4);
    img->resizeChannelCount(4);  // "Block compressors expect a 4 channel texture"
    return dst.setImage(nvtt::InputFormat_RGBA_32F, img->width(), img->height(), img->depth(), img->channel(0), img->channel(1),
                        img->channel(2), img->channel(3));
}
struct Install {
    Install() { nvtt::setSurfaceLoader(load_surface); }
} g_install;
}  // namespace
