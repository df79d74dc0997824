// DDS / DDS10 container READER (host code, no GPU work): validates a .dds file image in memory and locates every
// (face, mip) surface in it, so that pre-made mip chains can be fed to the pipeline (InputOptions::setMipmapData for levels
// > 0) or decoded (BCn) without going through an image library.  Mirrors the reference's nv::DirectDrawSurface:
// isValid / isSupported (src/nvimage/DirectDrawSurface.cpp:1002-1100), mipmapCount / width / height / depth / arraySize
// (:1143-1177), isTexture3D / isTextureCube / isTextureArray (:1202-1225), surfaceSize / faceSize / offset (:1253-1322),
// DDSHeader::blockSize / pixelSize (:902-963).  The matching writer is host/nvtt_host.cpp (Compressor::outputHeader).
#pragma once
#include "../../include/nvtt_b200.h"
#include <string.h>

namespace nvb {
namespace dds {
enum : unsigned {
    FOURCC_DDS = 0x20534444u, FOURCC_DX10 = 0x30315844u, FOURCC_NVTT = 0x5454564Eu,
    FOURCC_DXT1 = 0x31545844u, FOURCC_DXT2 = 0x32545844u, FOURCC_DXT3 = 0x33545844u, FOURCC_DXT4 = 0x34545844u, FOURCC_DXT5 = 0x35545844u,
    FOURCC_RXGB = 0x42475852u, FOURCC_ATI1 = 0x31495441u, FOURCC_ATI2 = 0x32495441u,
    DDSD_HEIGHT = 0x2u, DDSD_WIDTH = 0x4u, DDSD_MIPMAPCOUNT = 0x20000u, DDSD_DEPTH = 0x800000u,
    DDPF_ALPHAPIXELS = 0x1u, DDPF_FOURCC = 0x4u, DDPF_RGB = 0x40u, DDPF_LUMINANCE = 0x20000u, DDPF_NORMAL = 0x80000000u,
    DDSCAPS_TEXTURE = 0x1000u, DDSCAPS2_CUBEMAP = 0x200u, DDSCAPS2_CUBEMAP_ALL_FACES = 0xFC00u, DDSCAPS2_VOLUME = 0x200000u,
};
inline unsigned rd32(const unsigned char *p) { return (unsigned)p[0] | ((unsigned)p[1] << 8) | ((unsigned)p[2] << 16) | ((unsigned)p[3] << 24); }

// bits per pixel of the D3DFORMAT codes that appear as a FourCC, and of the DXGI formats the reference's table knows
inline unsigned d3d_pixel_bits(unsigned f) {
    switch (f) {
    case 36: return 64;   // A16B16G16R16
    case 111: return 16;  // R16F
    case 112: return 32;  // G16R16F
    case 113: return 64;  // A16B16G16R16F
    case 114: return 32;  // R32F
    case 115: return 64;  // G32R32F
    case 116: return 128; // A32B32G32R32F
    default: return 0;
    }
}
inline unsigned dxgi_pixel_bits(unsigned f) {
    switch (f) {
    case 2: return 128;                                    // R32G32B32A32_FLOAT
    case 6: return 96;                                     // R32G32B32_FLOAT
    case 10: case 11: case 16: return 64;                  // R16G16B16A16_FLOAT / UNORM, R32G32_FLOAT
    case 24: case 26: case 28: case 29: case 34: case 35: case 41: case 87: case 88: case 91: case 93: return 32;
    case 49: case 54: case 56: case 85: case 86: return 16; // R8G8_UNORM, R16_FLOAT, R16_UNORM, B5G6R5, B5G5R5A1
    case 61: case 65: return 8;                            // R8_UNORM, A8_UNORM
    default: return 0;
    }
}
inline unsigned block_bytes(unsigned fourcc, unsigned dxgi, int *format) {
    *format = -1;
    switch (fourcc) {
    case FOURCC_DXT1: *format = 1; return 8;
    case FOURCC_ATI1: *format = 6; return 8;
    case FOURCC_DXT2: case FOURCC_DXT3: *format = 3; return 16;
    case FOURCC_DXT4: case FOURCC_DXT5: *format = 4; return 16;
    case FOURCC_RXGB: *format = 5; return 16;
    case FOURCC_ATI2: *format = 7; return 16;
    case FOURCC_DX10:
        if (dxgi >= 70 && dxgi <= 72) { *format = 1; return 8; }
        if (dxgi >= 79 && dxgi <= 81) { *format = 6; return 8; }
        if (dxgi >= 73 && dxgi <= 75) { *format = 3; return 16; }
        if (dxgi >= 76 && dxgi <= 78) { *format = 4; return 16; }
        if (dxgi >= 82 && dxgi <= 84) { *format = 7; return 16; }
        if (dxgi >= 94 && dxgi <= 96) { *format = 10; return 16; }
        if (dxgi >= 97 && dxgi <= 99) { *format = 11; return 16; }
        return 0;
    default: return 0;
    }
}
}  // namespace dds
}  // namespace nvb

extern "C" int nvttb_dds_describe(const void *file, size_t bytes, NvttbDdsInfo *o) {
    using namespace nvb::dds;
    if (!file || !o || bytes < 128) return NVTTB_ERR_INVALID_INPUT;
    memset(o, 0, sizeof *o);
    const unsigned char *h = (const unsigned char *)file;
    // isValid
    if (rd32(h) != FOURCC_DDS || rd32(h + 4) != 124) return NVTTB_ERR_INVALID_INPUT;
    const unsigned flags = rd32(h + 8);
    if ((flags & (DDSD_WIDTH | DDSD_HEIGHT)) != (DDSD_WIDTH | DDSD_HEIGHT)) return NVTTB_ERR_INVALID_INPUT;
    if (rd32(h + 76) != 32) return NVTTB_ERR_INVALID_INPUT;                 // pf.size
    const unsigned caps1 = rd32(h + 108), caps2 = rd32(h + 112);
    if (!(caps1 & DDSCAPS_TEXTURE)) return NVTTB_ERR_INVALID_INPUT;
    const unsigned pfFlags = rd32(h + 80), fourcc = rd32(h + 84), bitcount = rd32(h + 88);
    const unsigned rmask = rd32(h + 92), gmask = rd32(h + 96), bmask = rd32(h + 100), amask = rd32(h + 104);
    const bool dx10 = (pfFlags & DDPF_FOURCC) && fourcc == FOURCC_DX10;
    unsigned dxgi = 0, dim = 0, arraySize = 1;
    if (dx10) {
        if (bytes < 148) return NVTTB_ERR_INVALID_INPUT;
        dxgi = rd32(h + 128);
        dim = rd32(h + 132);
        arraySize = rd32(h + 140);
    }
    o->width = (int)rd32(h + 16);
    o->height = (int)rd32(h + 12);
    o->depth = (flags & DDSD_DEPTH) ? (int)rd32(h + 24) : 1;
    o->mipCount = (flags & DDSD_MIPMAPCOUNT) ? (int)rd32(h + 28) : 1;
    if (o->mipCount < 1) o->mipCount = 1;
    o->arraySize = (int)arraySize;
    o->headerBytes = dx10 ? 148u : 128u;
    o->fourcc = (pfFlags & DDPF_FOURCC) ? fourcc : 0u;
    o->dxgiFormat = dxgi;
    const bool cube = (caps2 & DDSCAPS2_CUBEMAP) != 0;
    const bool volume = dx10 ? dim == 4 : (caps2 & DDSCAPS2_VOLUME) != 0;
    o->textureType = volume ? 2 : cube ? 1 : (dx10 && arraySize > 1) ? 3 : 0;  // nvtt::TextureType
    o->faceCount = cube ? 6 : 1;
    o->isNormalMap = (pfFlags & DDPF_NORMAL) != 0;
    o->blockBytes = block_bytes((pfFlags & DDPF_FOURCC) ? fourcc : 0u, dxgi, &o->blockFormat);
    o->inputFormat = -1;
    // isSupported + pixel size
    if (o->blockBytes == 0) {
        if (dx10) {
            o->bitsPerPixel = dxgi_pixel_bits(dxgi);
            if (dxgi == 87 || dxgi == 88 || dxgi == 91 || dxgi == 93) o->inputFormat = 0;  // B8G8R8A8 / X8 (+ sRGB)
            else if (dxgi == 10) o->inputFormat = 1;
            else if (dxgi == 2) o->inputFormat = 2;
            else if (dxgi == 41) o->inputFormat = 3;
        } else if (pfFlags & DDPF_FOURCC) {
            o->bitsPerPixel = d3d_pixel_bits(fourcc);
            if (fourcc == 113) o->inputFormat = 1;
            else if (fourcc == 116) o->inputFormat = 2;
            else if (fourcc == 114) o->inputFormat = 3;
        } else if ((pfFlags & DDPF_RGB) || (pfFlags & DDPF_LUMINANCE)) {
            o->bitsPerPixel = bitcount;
            if (bitcount == 32 && rmask == 0xFF0000u && gmask == 0xFF00u && bmask == 0xFFu && (amask == 0xFF000000u || amask == 0)) o->inputFormat = 0;
        }
        if (o->bitsPerPixel == 0) return NVTTB_ERR_UNSUPPORTED_FEATURE;  // unknown fourcc / format
    }
    if (!dx10 && cube) {
        if (o->width != o->height) return NVTTB_ERR_UNSUPPORTED_FEATURE;
        if ((caps2 & DDSCAPS2_CUBEMAP_ALL_FACES) != DDSCAPS2_CUBEMAP_ALL_FACES) return NVTTB_ERR_UNSUPPORTED_FEATURE;  // cube maps must contain all faces
    }
    // hasAlpha (DirectDrawSurface.cpp:1088-1118)
    if (rd32(h + 32 + 9 * 4) == FOURCC_NVTT) o->hasAlpha = (pfFlags & DDPF_ALPHAPIXELS) != 0;
    else if (dx10) o->hasAlpha = !(dxgi == 88 || dxgi == 93 || dxgi == 6 || dxgi == 16 || dxgi == 41 || (dxgi >= 70 && dxgi <= 72) || (dxgi >= 79 && dxgi <= 84) || (dxgi >= 94 && dxgi <= 96));
    else if (pfFlags & DDPF_RGB) o->hasAlpha = amask != 0;
    else if (pfFlags & DDPF_FOURCC) o->hasAlpha = !(fourcc == FOURCC_DXT1 || fourcc == FOURCC_ATI1 || fourcc == FOURCC_ATI2 || fourcc == 111 || fourcc == 112 || fourcc == 114 || fourcc == 115);
    // every surface must lie inside the file
    size_t off = 0, sz = 0;
    const int faces = o->faceCount * (o->textureType == 3 ? o->arraySize : 1);
    if (nvttb_dds_surface(o, faces - 1, o->mipCount - 1, &off, &sz, nullptr, nullptr, nullptr) != NVTTB_OK || off + sz > bytes) return NVTTB_ERR_INVALID_INPUT;
    return NVTTB_OK;
}

extern "C" int nvttb_dds_surface(const NvttbDdsInfo *o, int face, int mip, size_t *offset, size_t *bytes, int *w, int *h, int *d) {
    if (!o || !offset || !bytes || mip < 0 || mip >= o->mipCount || face < 0) return NVTTB_ERR_INVALID_INPUT;
    const int faces = o->faceCount * (o->textureType == 3 ? o->arraySize : 1);
    if (face >= faces) return NVTTB_ERR_INVALID_INPUT;
    auto surface_size = [&](int m, int *sw, int *sh, int *sd) -> size_t {
        unsigned x = (unsigned)o->width, y = (unsigned)o->height, z = (unsigned)o->depth;
        for (int i = 0; i < m; i++) {
            x = x / 2 > 1 ? x / 2 : 1;
            y = y / 2 > 1 ? y / 2 : 1;
            z = z / 2 > 1 ? z / 2 : 1;
        }
        if (sw) *sw = (int)x;
        if (sh) *sh = (int)y;
        if (sd) *sd = (int)z;
        if (o->blockBytes == 0) return (size_t)((x * o->bitsPerPixel + 7) / 8) * y * z;  // computeBytePitch(w, bits, 1) * h * d
        return (size_t)o->blockBytes * ((x + 3) / 4) * ((y + 3) / 4) * z;
    };
    size_t faceSize = 0;
    for (int m = 0; m < o->mipCount; m++) faceSize += surface_size(m, nullptr, nullptr, nullptr);
    size_t off = o->headerBytes + (size_t)face * faceSize;
    for (int m = 0; m < mip; m++) off += surface_size(m, nullptr, nullptr, nullptr);
    *offset = off;
    *bytes = surface_size(mip, w, h, d);
    return NVTTB_OK;
}
