// png_io.h — minimal PNG reader/writer on top of zlib, replacing the cv::imread / cv::imwrite calls of the
// reference's dataset reader and test-image writer (MON/Core/src/nerf_data.cu:157-205, nerf.cu:255-404) in builds
// without OpenCV.  Non-interlaced 8/16-bit gray, gray+alpha, RGB, RGBA and 8-bit palette images.
#pragma once
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace png_io {

struct Image {
    int width = 0, height = 0, channels = 0, bit_depth = 0;   // channels after decoding: 1 (gray) or 3 (RGB)
    std::vector<uint8_t> u8;    // bit_depth == 8
    std::vector<uint16_t> u16;  // bit_depth == 16 (host byte order)
};

inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

inline bool read(const std::string& path, Image& img, std::string& err) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { err = "cannot open " + path; return false; }
    std::vector<uint8_t> file;
    uint8_t buf[65536];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) file.insert(file.end(), buf, buf + n);
    fclose(f);
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (file.size() < 33 || memcmp(file.data(), sig, 8) != 0) { err = "not a PNG: " + path; return false; }
    size_t pos = 8;
    int color_type = -1, interlace = 0;
    std::vector<uint8_t> idat, palette;
    while (pos + 12 <= file.size()) {
        const uint32_t len = be32(&file[pos]);
        const char* type = reinterpret_cast<const char*>(&file[pos + 4]);
        const uint8_t* data = &file[pos + 8];
        if (pos + 12 + len > file.size()) { err = "truncated PNG: " + path; return false; }
        if (!memcmp(type, "IHDR", 4)) {
            img.width = (int)be32(data); img.height = (int)be32(data + 4);
            img.bit_depth = data[8]; color_type = data[9]; interlace = data[12];
        } else if (!memcmp(type, "PLTE", 4)) {
            palette.assign(data, data + len);
        } else if (!memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), data, data + len);
        } else if (!memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + len;
    }
    if (interlace != 0) { err = "interlaced PNG not supported: " + path; return false; }
    int src_ch;
    switch (color_type) {
        case 0: src_ch = 1; break;
        case 2: src_ch = 3; break;
        case 3: src_ch = 1; break;
        case 4: src_ch = 2; break;
        case 6: src_ch = 4; break;
        default: err = "unsupported PNG colour type: " + path; return false;
    }
    if (!(img.bit_depth == 8 || (img.bit_depth == 16 && color_type != 3))) { err = "unsupported PNG bit depth: " + path; return false; }
    const size_t bpp = (size_t)src_ch * img.bit_depth / 8, stride = bpp * img.width;
    std::vector<uint8_t> raw((stride + 1) * img.height);
    uLongf raw_len = raw.size();
    if (uncompress(raw.data(), &raw_len, idat.data(), idat.size()) != Z_OK || raw_len != raw.size()) { err = "PNG inflate failed: " + path; return false; }
    // undo the per-row filters in place
    std::vector<uint8_t> pix(stride * img.height);
    for (int y = 0; y < img.height; ++y) {
        const uint8_t ft = raw[(stride + 1) * y];
        const uint8_t* in = &raw[(stride + 1) * y + 1];
        uint8_t* out = &pix[stride * y];
        const uint8_t* up = y ? &pix[stride * (y - 1)] : nullptr;
        for (size_t x = 0; x < stride; ++x) {
            const int a = x >= bpp ? out[x - bpp] : 0, b = up ? up[x] : 0, c = (up && x >= bpp) ? up[x - bpp] : 0;
            int pred = 0;
            switch (ft) {
                case 0: pred = 0; break;
                case 1: pred = a; break;
                case 2: pred = b; break;
                case 3: pred = (a + b) >> 1; break;
                case 4: { const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c); pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); break; }
                default: err = "bad PNG filter: " + path; return false;
            }
            out[x] = (uint8_t)(in[x] + pred);
        }
    }
    const size_t npx = (size_t)img.width * img.height;
    if (color_type == 3) {
        img.channels = 3;
        img.u8.resize(npx * 3);
        for (size_t i = 0; i < npx; ++i)
            for (int k = 0; k < 3; ++k) img.u8[i * 3 + k] = (size_t)pix[i] * 3 + k < palette.size() ? palette[(size_t)pix[i] * 3 + k] : 0;
        return true;
    }
    img.channels = (src_ch >= 3) ? 3 : 1;
    if (img.bit_depth == 8) {
        img.u8.resize(npx * img.channels);
        for (size_t i = 0; i < npx; ++i)
            for (int k = 0; k < img.channels; ++k) img.u8[i * img.channels + k] = pix[i * src_ch + k];
    } else {
        img.u16.resize(npx * img.channels);
        for (size_t i = 0; i < npx; ++i)
            for (int k = 0; k < img.channels; ++k) {
                const uint8_t* p = &pix[(i * src_ch + k) * 2];
                img.u16[i * img.channels + k] = (uint16_t)((p[0] << 8) | p[1]);
            }
    }
    return true;
}

inline void put_chunk(std::vector<uint8_t>& out, const char* type, const uint8_t* data, size_t len) {
    const uint32_t l = (uint32_t)len;
    const uint8_t lb[4] = {(uint8_t)(l >> 24), (uint8_t)(l >> 16), (uint8_t)(l >> 8), (uint8_t)l};
    out.insert(out.end(), lb, lb + 4);
    const size_t start = out.size();
    out.insert(out.end(), type, type + 4);
    if (len) out.insert(out.end(), data, data + len);
    const uint32_t crc = (uint32_t)crc32(0L, &out[start], (uInt)(len + 4));
    const uint8_t cb[4] = {(uint8_t)(crc >> 24), (uint8_t)(crc >> 16), (uint8_t)(crc >> 8), (uint8_t)crc};
    out.insert(out.end(), cb, cb + 4);
}

// data: row-major, channels 1 or 3, bit_depth 8 (uint8_t) or 16 (uint16_t, host order)
inline bool write(const std::string& path, int width, int height, int channels, int bit_depth, const void* data) {
    const size_t bpp = (size_t)channels * bit_depth / 8, stride = bpp * width;
    std::vector<uint8_t> raw((stride + 1) * height);
    for (int y = 0; y < height; ++y) {
        raw[(stride + 1) * y] = 0;
        uint8_t* dst = &raw[(stride + 1) * y + 1];
        if (bit_depth == 8) {
            memcpy(dst, static_cast<const uint8_t*>(data) + stride * y, stride);
        } else {
            const uint16_t* src = static_cast<const uint16_t*>(data) + (size_t)width * channels * y;
            for (size_t i = 0; i < (size_t)width * channels; ++i) { dst[2 * i] = (uint8_t)(src[i] >> 8); dst[2 * i + 1] = (uint8_t)src[i]; }
        }
    }
    uLongf clen = compressBound(raw.size());
    std::vector<uint8_t> comp(clen);
    if (compress2(comp.data(), &clen, raw.data(), raw.size(), 6) != Z_OK) return false;
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    uint8_t ihdr[13] = {(uint8_t)(width >> 24), (uint8_t)(width >> 16), (uint8_t)(width >> 8), (uint8_t)width,
                        (uint8_t)(height >> 24), (uint8_t)(height >> 16), (uint8_t)(height >> 8), (uint8_t)height,
                        (uint8_t)bit_depth, (uint8_t)(channels == 3 ? 2 : 0), 0, 0, 0};
    put_chunk(out, "IHDR", ihdr, 13);
    put_chunk(out, "IDAT", comp.data(), clen);
    put_chunk(out, "IEND", nullptr, 0);
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    const bool ok = fwrite(out.data(), 1, out.size(), f) == out.size();
    fclose(f);
    return ok;
}

}  // namespace png_io
