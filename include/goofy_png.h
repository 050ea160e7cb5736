// goofy_png.h -- PNG ingest for the benchmark harness: file -> RGBA8, 64-byte aligned, alpha forced to 0xFF.
//
// The loader contract is the reference harness's (loadPngAsRgba8, Src/main.cpp:258-343 of the reference): the image
// comes back as tightly packed RGBA8 in a 64-byte-aligned buffer, alpha := 0xFF whatever the file holds (:328-335),
// and images whose width is not a multiple of 16 or whose height is not a multiple of 4 are rejected (:298-307).
// The reference decodes with lodepng; this is a small decoder written from the specifications instead (PNG 1.2,
// RFC 1950 zlib, RFC 1951 deflate): 8-bit and 16-bit samples, colour types 0 / 2 / 3 / 4 / 6, palette images of
// 1 / 2 / 4 / 8 bits, non-interlaced.  Checksums (CRC-32, Adler-32) are not verified -- the harness reads its own
// test images.  Header-only, host-side, no CUDA.
#ifndef GOOFY_PNG_H
#define GOOFY_PNG_H

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace goofy {
namespace png {

// ------------------------------------------------------------------------------------------ deflate (RFC 1951)
class Inflater {
public:
    Inflater(const uint8_t* data, size_t size) : p_(data), n_(size) {}

    // Decompresses a zlib stream (RFC 1950: 2 header bytes, deflate blocks, Adler-32) into `out`.
    bool run(std::vector<uint8_t>& out)
    {
        if (n_ < 2 || (p_[0] & 0x0F) != 8 || ((p_[0] << 8 | p_[1]) % 31) != 0 || (p_[1] & 0x20)) return false;
        pos_ = 2;
        for (;;) {
            const uint32_t last = bits(1), type = bits(2);
            if (bad_) return false;
            bool ok;
            if (type == 0) ok = stored(out);
            else if (type == 1) ok = fixed(out);
            else if (type == 2) ok = dynamic(out);
            else ok = false;
            if (!ok) return false;
            if (last) return true;
        }
    }

private:
    // A canonical Huffman code as RFC 1951 3.2.2 defines it: codes of one length are consecutive values in symbol
    // order, so (count per length, symbols sorted by length) is all a decoder needs.
    struct Code {
        uint16_t count[16];
        uint16_t symbol[288];
        bool build(const uint8_t* lengths, int n)
        {
            std::memset(count, 0, sizeof(count));
            for (int i = 0; i < n; ++i) ++count[lengths[i]];
            int left = 1;   // over-subscription check
            for (int len = 1; len < 16; ++len) {
                left = left * 2 - count[len];
                if (left < 0) return false;
            }
            uint16_t offset[16];
            offset[1] = 0;
            for (int len = 1; len < 15; ++len) offset[len + 1] = (uint16_t)(offset[len] + count[len]);
            for (int i = 0; i < n; ++i)
                if (lengths[i]) symbol[offset[lengths[i]]++] = (uint16_t)i;
            return true;
        }
    };

    uint32_t bits(int need)
    {
        while (cnt_ < need) {
            if (pos_ >= n_) { bad_ = true; return 0; }
            buf_ |= (uint32_t)p_[pos_++] << cnt_;
            cnt_ += 8;
        }
        const uint32_t v = buf_ & ((1u << need) - 1u);
        buf_ >>= need;
        cnt_ -= need;
        return v;
    }

    // one symbol: walk the code lengths, keeping the first code and the first symbol index of the current length
    int decode(const Code& c)
    {
        int code = 0, first = 0, index = 0;
        for (int len = 1; len < 16; ++len) {
            code |= (int)bits(1);
            if (bad_) return -1;
            const int count = c.count[len];
            if (code - count < first) return c.symbol[index + (code - first)];
            index += count;
            first = (first + count) << 1;
            code <<= 1;
        }
        return -1;
    }

    bool stored(std::vector<uint8_t>& out)
    {
        buf_ = 0;
        cnt_ = 0;   // skip to the next byte boundary
        if (pos_ + 4 > n_) return false;
        const uint32_t len = p_[pos_] | p_[pos_ + 1] << 8, nlen = p_[pos_ + 2] | p_[pos_ + 3] << 8;
        pos_ += 4;
        if ((len ^ 0xFFFFu) != nlen || pos_ + len > n_) return false;
        out.insert(out.end(), p_ + pos_, p_ + pos_ + len);
        pos_ += len;
        return true;
    }

    bool codes(std::vector<uint8_t>& out, const Code& lit, const Code& dist)
    {
        static const uint16_t lenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
        static const uint8_t lenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
        static const uint16_t distBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
        static const uint8_t distExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
        for (;;) {
            const int sym = decode(lit);
            if (sym < 0) return false;
            if (sym < 256) { out.push_back((uint8_t)sym); continue; }
            if (sym == 256) return true;
            if (sym > 285) return false;
            const uint32_t len = lenBase[sym - 257] + bits(lenExtra[sym - 257]);
            const int ds = decode(dist);
            if (ds < 0 || ds > 29) return false;
            const uint32_t d = distBase[ds] + bits(distExtra[ds]);
            if (bad_ || d > out.size()) return false;
            const size_t from = out.size() - d;
            for (uint32_t i = 0; i < len; ++i) out.push_back(out[from + i]);   // may overlap its own output: byte by byte
        }
    }

    bool fixed(std::vector<uint8_t>& out)
    {
        uint8_t lengths[288];
        int i = 0;
        for (; i < 144; ++i) lengths[i] = 8;
        for (; i < 256; ++i) lengths[i] = 9;
        for (; i < 280; ++i) lengths[i] = 7;
        for (; i < 288; ++i) lengths[i] = 8;
        Code lit, dist;
        lit.build(lengths, 288);
        for (i = 0; i < 30; ++i) lengths[i] = 5;
        dist.build(lengths, 30);
        return codes(out, lit, dist);
    }

    bool dynamic(std::vector<uint8_t>& out)
    {
        static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        const int nlen = (int)bits(5) + 257, ndist = (int)bits(5) + 1, ncode = (int)bits(4) + 4;
        if (bad_ || nlen > 286 || ndist > 30) return false;
        uint8_t lengths[320];
        std::memset(lengths, 0, sizeof(lengths));
        for (int i = 0; i < ncode; ++i) lengths[order[i]] = (uint8_t)bits(3);
        Code lenCode;
        if (!lenCode.build(lengths, 19)) return false;
        std::memset(lengths, 0, sizeof(lengths));
        int i = 0;
        while (i < nlen + ndist) {
            const int sym = decode(lenCode);
            if (sym < 0) return false;
            if (sym < 16) { lengths[i++] = (uint8_t)sym; continue; }
            int prev = 0, rep;
            if (sym == 16) {
                if (i == 0) return false;
                prev = lengths[i - 1];
                rep = 3 + (int)bits(2);
            } else if (sym == 17) rep = 3 + (int)bits(3);
            else rep = 11 + (int)bits(7);
            if (bad_ || i + rep > nlen + ndist) return false;
            while (rep--) lengths[i++] = (uint8_t)prev;
        }
        if (lengths[256] == 0) return false;
        Code lit, dist;
        if (!lit.build(lengths, nlen)) return false;
        dist.build(lengths + nlen, ndist);   // an incomplete distance code is legal (a single distance)
        return codes(out, lit, dist);
    }

    const uint8_t* p_;
    size_t n_, pos_ = 0;
    uint32_t buf_ = 0;
    int cnt_ = 0;
    bool bad_ = false;
};

// ------------------------------------------------------------------------------------------ PNG
struct Image {
    unsigned char* rgba = nullptr;   // width * height * 4 bytes, 64-byte aligned; release with freeImage
    unsigned int width = 0, height = 0;
    std::string error;               // empty on success
};

inline void freeImage(Image& im)
{
    std::free(im.rgba);
    im.rgba = nullptr;
}

inline uint32_t be32(const uint8_t* p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }

// requireEncoderShape: apply the reference loader's shape check (width % 16, height % 4).
inline Image load(const char* path, bool requireEncoderShape = true)
{
    Image im;
    std::vector<uint8_t> file;
    {
        FILE* f = std::fopen(path, "rb");
        if (!f) { im.error = "cannot open file"; return im; }
        std::fseek(f, 0, SEEK_END);
        const long size = std::ftell(f);
        std::fseek(f, 0, SEEK_SET);
        file.resize(size > 0 ? (size_t)size : 0);
        const size_t got = file.empty() ? 0 : std::fread(file.data(), 1, file.size(), f);
        std::fclose(f);
        if (got != file.size()) { im.error = "short read"; return im; }
    }
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (file.size() < 8 + 25 || std::memcmp(file.data(), sig, 8) != 0) { im.error = "not a PNG file"; return im; }

    uint32_t width = 0, height = 0;
    int depth = 0, colour = -1, interlace = 0;
    std::vector<uint8_t> idat, palette;
    for (size_t pos = 8; pos + 12 <= file.size();) {
        const uint32_t len = be32(&file[pos]);
        const uint8_t* type = &file[pos + 4];
        const uint8_t* data = &file[pos + 8];
        if (pos + 12 + (size_t)len > file.size()) { im.error = "truncated chunk"; return im; }
        if (!std::memcmp(type, "IHDR", 4) && len >= 13) {
            width = be32(data);
            height = be32(data + 4);
            depth = data[8];
            colour = data[9];
            interlace = data[12];
        } else if (!std::memcmp(type, "PLTE", 4)) palette.assign(data, data + len);
        else if (!std::memcmp(type, "IDAT", 4)) idat.insert(idat.end(), data, data + len);
        else if (!std::memcmp(type, "IEND", 4)) break;
        pos += 12 + (size_t)len;
    }
    if (width == 0 || height == 0 || colour < 0) { im.error = "no IHDR"; return im; }
    if (interlace != 0) { im.error = "interlaced PNG not supported"; return im; }
    int channels;
    switch (colour) {
        case 0: channels = 1; break;
        case 2: channels = 3; break;
        case 3: channels = 1; break;
        case 4: channels = 2; break;
        case 6: channels = 4; break;
        default: im.error = "bad colour type"; return im;
    }
    const bool depthOk = colour == 3 ? (depth == 1 || depth == 2 || depth == 4 || depth == 8)
                                     : colour == 0 ? (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16) : (depth == 8 || depth == 16);
    if (!depthOk) { im.error = "unsupported bit depth"; return im; }
    if (requireEncoderShape) {
        // the reference's messages, Src/main.cpp:298-307
        if (width % 16u != 0u) { im.error = "Incorrect width. Width should be a multiple of 16"; return im; }
        if (height % 4u != 0u) { im.error = "Incorrect height. Height should be a multiple of 4"; return im; }
    }

    std::vector<uint8_t> raw;
    const size_t bitsPerPixel = (size_t)channels * depth, rowBytes = (width * bitsPerPixel + 7) / 8;
    raw.reserve((rowBytes + 1) * height);
    if (!Inflater(idat.data(), idat.size()).run(raw) || raw.size() < (rowBytes + 1) * height) { im.error = "corrupt image data"; return im; }

    // undo the scanline filters in place (PNG 1.2 section 6): bpp = bytes per complete pixel, at least 1
    const size_t bpp = bitsPerPixel >= 8 ? bitsPerPixel / 8 : 1;
    for (uint32_t y = 0; y < height; ++y) {
        uint8_t* row = &raw[(rowBytes + 1) * y + 1];
        const uint8_t* up = y ? row - (rowBytes + 1) : nullptr;
        const int filter = row[-1];
        for (size_t i = 0; i < rowBytes; ++i) {
            const int a = i >= bpp ? row[i - bpp] : 0, b = up ? up[i] : 0, c = (up && i >= bpp) ? up[i - bpp] : 0;
            int pred;
            switch (filter) {
                case 0: pred = 0; break;
                case 1: pred = a; break;
                case 2: pred = b; break;
                case 3: pred = (a + b) >> 1; break;
                case 4: {
                    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
                    pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
                    break;
                }
                default: im.error = "bad filter type"; return im;
            }
            row[i] = (uint8_t)(row[i] + pred);
        }
    }

    const size_t bytes = (size_t)width * height * 4;
    im.rgba = (unsigned char*)std::aligned_alloc(64, (bytes + 63) / 64 * 64);
    if (!im.rgba) { im.error = "out of memory"; return im; }
    for (uint32_t y = 0; y < height; ++y) {
        const uint8_t* row = &raw[(rowBytes + 1) * y + 1];
        unsigned char* dst = im.rgba + (size_t)y * width * 4;
        for (uint32_t x = 0; x < width; ++x, dst += 4) {
            uint8_t r, g, b;
            if (colour == 3 || (colour == 0 && depth < 8)) {
                const uint32_t v = depth == 8 ? row[x] : (row[(size_t)x * depth / 8] >> (8 - depth - (x * depth) % 8)) & ((1u << depth) - 1u);
                if (colour == 3) {
                    if ((size_t)v * 3 + 3 > palette.size()) { r = g = b = 0; }
                    else { r = palette[v * 3]; g = palette[v * 3 + 1]; b = palette[v * 3 + 2]; }
                } else r = g = b = (uint8_t)(v * 255u / ((1u << depth) - 1u));
            } else {
                const size_t step = depth / 8;   // 16-bit samples: keep the high byte
                const uint8_t* s = row + (size_t)x * channels * step;
                if (channels >= 3) { r = s[0]; g = s[step]; b = s[2 * step]; }
                else r = g = b = s[0];
            }
            dst[0] = r;
            dst[1] = g;
            dst[2] = b;
            dst[3] = 0xFF;   // alpha is forced, as the reference loader does (:328-335): the encoders ignore it anyway
        }
    }
    im.width = width;
    im.height = height;
    return im;
}

}  // namespace png
}  // namespace goofy

#endif  // GOOFY_PNG_H
