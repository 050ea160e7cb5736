// goofy_containers.h -- DDS (DXT1) and KTX 1.1 (ETC1) writers for encoder output; host-side, header-only.
//
// What the reference harness does with its results (saveDds Src/main.cpp:154-183, saveKtx :185-220), written
// from the DDS and KTX 1.1 specifications: one mip level, one face, FourCC 'DXT1' / glInternalFormat
// GL_ETC1_RGB8_OES (0x8D64), linear size / imageSize = width*height/2.  Returns false if the file cannot be written.
#ifndef GOOFY_CONTAINERS_H
#define GOOFY_CONTAINERS_H

#include <cstdint>
#include <cstdio>
#include <cstring>

namespace goofy {
namespace containers {

inline bool writeFile(const char* path, const void* header, size_t headerBytes, const void* payload, size_t payloadBytes)
{
    std::FILE* f = std::fopen(path, "wb");
    if (!f) return false;
    const bool ok = std::fwrite(header, 1, headerBytes, f) == headerBytes && std::fwrite(payload, 1, payloadBytes, f) == payloadBytes;
    return std::fclose(f) == 0 && ok;
}

// 128-byte DDS header: magic, DDS_HEADER (124 bytes) with a FourCC pixel format
inline bool writeDdsDxt1(const char* path, const unsigned char* blocks, uint32_t width, uint32_t height)
{
    uint32_t h[32];
    std::memset(h, 0, sizeof(h));
    const uint32_t bytes = width * height / 2u;
    h[0] = 0x20534444u;                                          // "DDS "
    h[1] = 124u;                                                 // dwSize
    h[2] = 0x1u | 0x2u | 0x4u | 0x1000u | 0x20000u | 0x80000u;   // CAPS HEIGHT WIDTH PIXELFORMAT MIPMAPCOUNT LINEARSIZE
    h[3] = height;
    h[4] = width;
    h[5] = bytes;                                                // dwPitchOrLinearSize
    h[6] = 1u;                                                   // dwDepth
    h[7] = 1u;                                                   // dwMipMapCount
    h[19] = 32u;                                                 // ddspf.dwSize
    h[20] = 0x4u;                                                // DDPF_FOURCC
    h[21] = 0x31545844u;                                         // "DXT1"
    return writeFile(path, h, sizeof(h), blocks, bytes);
}

// 64-byte KTX 1.1 header + imageSize
inline bool writeKtxEtc1(const char* path, const unsigned char* blocks, uint32_t width, uint32_t height)
{
    static const unsigned char id[12] = {0xAB, 0x4B, 0x54, 0x58, 0x20, 0x31, 0x31, 0xBB, 0x0D, 0x0A, 0x1A, 0x0A};
    unsigned char head[68];
    std::memset(head, 0, sizeof(head));
    std::memcpy(head, id, 12);
    const uint32_t bytes = width * height / 2u;
    const uint32_t f[14] = {0x04030201u, 0u, 1u, 0u, 0x8D64u /* GL_ETC1_RGB8_OES */, 0x1907u /* GL_RGB */, width, height,
                            0u, 0u, 1u, 1u, 0u, bytes /* imageSize */};
    std::memcpy(head + 12, f, sizeof(f));
    return writeFile(path, head, sizeof(head), blocks, bytes);
}

}  // namespace containers
}  // namespace goofy

#endif
