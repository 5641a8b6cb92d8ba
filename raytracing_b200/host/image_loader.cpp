/*
 * image_loader.cpp — 8-bit image textures for the scene loader, replacing the reference's stb_image path
 * (/root/reference/src/loaders/image_loader.cpp:30-63 `LoadSTB`, called from Scene::LoadTexture, scene.cpp:276-323, for
 * ".png", ".tga" and ".jpg").  Output is the reference's texel word: r | g << 8 | b << 16 | a << 24 built from the FILE's
 * channel count the way LoadSTB does it (a channel the file does not have is 0 — a grey image yields (y, 0, 0, 0), grey + alpha
 * yields (y, a, 0, 0)), rows top to bottom.
 *
 * Decoders written here (no third-party code): PNG (zlib inflate from the system library; colour types 0/2/3/4/6, bit depths
 * 1-16, tRNS, Adam7 interlacing), TGA (types 1/2/3/9/10/11: colour-mapped, true-colour 15/16/24/32 bit, grey; RLE; either row
 * order) and JPEG (jpeg_decoder.cpp).  The format is recognised by content, like the reference's loader does.
 */
#include "reference_api.hpp"

#include <zlib.h>

#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace rt_host
{
namespace
{
bool read_file(const char* path, std::vector<unsigned char>& out)
{
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (n < 0) { fclose(f); return false; }
    out.resize((size_t)n);
    bool ok = n == 0 || fread(out.data(), 1, (size_t)n, f) == (size_t)n;
    fclose(f);
    return ok;
}

// pixels: n_channels interleaved bytes per pixel, rows top to bottom -> the reference's packed words
void pack(const std::vector<unsigned char>& px, int w, int h, int nc, TextureImage& out)
{
    out.width = (std::uint32_t)w; out.height = (std::uint32_t)h;
    out.data.resize((size_t)w * h);
    for (size_t i = 0; i < (size_t)w * h; ++i)
    {
        const unsigned char* p = &px[i * nc];
        std::uint32_t r = p[0], g = nc > 1 ? p[1] : 0u, b = nc > 2 ? p[2] : 0u, a = nc > 3 ? p[3] : 0u;
        out.data[i] = (r << 0) | (g << 8) | (b << 16) | (a << 24);       // image_loader.cpp:48-56
    }
}

std::uint32_t be32(const unsigned char* p) { return ((std::uint32_t)p[0] << 24) | ((std::uint32_t)p[1] << 16) | ((std::uint32_t)p[2] << 8) | p[3]; }

int paeth(int a, int b, int c)
{
    int p = a + b - c, pa = p > a ? p - a : a - p, pb = p > b ? p - b : b - p, pc = p > c ? p - c : c - p;
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

bool load_png(const std::vector<unsigned char>& file, TextureImage& out, std::string& err)
{
    static const unsigned char sig[8] = { 137, 80, 78, 71, 13, 10, 26, 10 };
    if (file.size() < 8 || memcmp(file.data(), sig, 8) != 0) { err = "not a PNG file"; return false; }
    std::uint32_t w = 0, h = 0;
    int depth = 0, ctype = -1, interlace = 0;
    std::vector<unsigned char> idat, palette, trns;
    size_t pos = 8;
    bool end = false;
    while (!end && pos + 12 <= file.size())
    {
        std::uint32_t len = be32(&file[pos]);
        const unsigned char* type = &file[pos + 4];
        if (pos + 12 + (size_t)len > file.size()) { err = "truncated PNG chunk"; return false; }
        const unsigned char* data = &file[pos + 8];
        if (!memcmp(type, "IHDR", 4))
        {
            if (len < 13) { err = "bad IHDR"; return false; }
            w = be32(data); h = be32(data + 4); depth = data[8]; ctype = data[9]; interlace = data[12];
            if (data[10] != 0 || data[11] != 0) { err = "unknown PNG compression/filter method"; return false; }
        }
        else if (!memcmp(type, "PLTE", 4)) palette.assign(data, data + len);
        else if (!memcmp(type, "tRNS", 4)) trns.assign(data, data + len);
        else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), data, data + len);
        else if (!memcmp(type, "IEND", 4)) end = true;
        pos += 12 + (size_t)len;
    }
    if (ctype < 0 || w == 0 || h == 0 || w > 32768 || h > 32768) { err = "bad PNG header"; return false; }
    if (interlace > 1) { err = "unknown PNG interlace method"; return false; }
    int src_channels;
    switch (ctype) { case 0: src_channels = 1; break; case 2: src_channels = 3; break; case 3: src_channels = 1; break;
                     case 4: src_channels = 2; break; case 6: src_channels = 4; break; default: err = "bad PNG colour type"; return false; }
    if (!(depth == 8 || depth == 16 || ((ctype == 0 || ctype == 3) && (depth == 1 || depth == 2 || depth == 4)))) { err = "bad PNG bit depth"; return false; }
    if (ctype == 3 && (depth == 16 || palette.size() < 3)) { err = "bad PNG palette"; return false; }
    const size_t bits_pp = (size_t)src_channels * depth, bpp = (bits_pp + 7) / 8;
    // the image is stored as one pass, or as the seven Adam7 passes (each a complete filtered sub-image of the pixels
    // x0 + i * dx, y0 + j * dy)
    struct Pass { std::uint32_t x0, y0, dx, dy; };
    static const Pass adam7[7] = { { 0, 0, 8, 8 }, { 4, 0, 8, 8 }, { 0, 4, 4, 8 }, { 2, 0, 4, 4 }, { 0, 2, 2, 4 }, { 1, 0, 2, 2 }, { 0, 1, 1, 2 } };
    static const Pass whole = { 0, 0, 1, 1 };
    const Pass* passes = interlace ? adam7 : &whole;
    const int n_passes = interlace ? 7 : 1;
    auto pass_w = [&](const Pass& ps) { return w > ps.x0 ? (w - ps.x0 + ps.dx - 1) / ps.dx : 0u; };
    auto pass_h = [&](const Pass& ps) { return h > ps.y0 ? (h - ps.y0 + ps.dy - 1) / ps.dy : 0u; };
    size_t raw_size = 0;
    for (int i = 0; i < n_passes; ++i)
        if (pass_w(passes[i]) && pass_h(passes[i])) raw_size += ((pass_w(passes[i]) * bits_pp + 7) / 8 + 1) * pass_h(passes[i]);
    std::vector<unsigned char> raw(raw_size);
    {
        uLongf dst_len = (uLongf)raw.size();
        int rc = uncompress(raw.data(), &dst_len, idat.data(), (uLong)idat.size());
        if (rc != Z_OK || dst_len != raw.size()) { err = "PNG image data does not inflate to the image size"; return false; }
    }
    // to 8-bit samples (16-bit: the high byte; 1/2/4-bit grey: scaled to 0..255; palette indices: looked up)
    const bool has_trns = !trns.empty();
    int out_channels = ctype == 3 ? (has_trns ? 4 : 3) : src_channels + ((has_trns && (ctype == 0 || ctype == 2)) ? 1 : 0);
    std::vector<unsigned char> px((size_t)w * h * out_channels);
    auto sample = [&](const unsigned char* row, size_t index) -> unsigned {     // index counts samples within the row
        if (depth == 8) return row[index];
        if (depth == 16) return row[index * 2];
        size_t bit = index * depth;
        unsigned v = (row[bit >> 3] >> (8 - depth - (bit & 7))) & ((1u << depth) - 1u);
        return v;
    };
    auto sample16 = [&](const unsigned char* row, size_t index) -> unsigned { return depth == 16 ? (row[index * 2] << 8 | row[index * 2 + 1]) : sample(row, index); };
    const unsigned grey_scale = depth < 8 ? 255u / ((1u << depth) - 1u) : 1u;
    std::vector<unsigned char> lines;
    size_t raw_pos = 0;
    for (int pi = 0; pi < n_passes; ++pi)
    {
        const Pass& ps = passes[pi];
        const std::uint32_t pw = pass_w(ps), ph = pass_h(ps);
        if (!pw || !ph) continue;
        const size_t stride = (pw * bits_pp + 7) / 8;
        // un-filter (each row: filter byte + stride bytes)
        lines.assign(stride * ph, 0);
        for (std::uint32_t y = 0; y < ph; ++y)
        {
            const unsigned char* in = &raw[raw_pos + (stride + 1) * y];
            unsigned char* cur = &lines[stride * y];
            const unsigned char* up = y ? &lines[stride * (y - 1)] : nullptr;
            int filter = in[0];
            for (size_t x = 0; x < stride; ++x)
            {
                int a = x >= bpp ? cur[x - bpp] : 0, b = up ? up[x] : 0, c = (up && x >= bpp) ? up[x - bpp] : 0, v = in[1 + x];
                switch (filter)
                {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) >> 1; break;
                case 4: v += paeth(a, b, c); break;
                default: err = "bad PNG filter type"; return false;
                }
                cur[x] = (unsigned char)v;
            }
        }
        raw_pos += (stride + 1) * ph;
        for (std::uint32_t y = 0; y < ph; ++y)
        {
            const unsigned char* row = &lines[stride * y];
            for (std::uint32_t x = 0; x < pw; ++x)
            {
                unsigned char* o = &px[((size_t)(ps.y0 + y * ps.dy) * w + (ps.x0 + x * ps.dx)) * out_channels];
                if (ctype == 3)
                {
                    unsigned idx = sample(row, x);
                    if ((size_t)idx * 3 + 2 >= palette.size()) { err = "PNG palette index out of range"; return false; }
                    o[0] = palette[idx * 3]; o[1] = palette[idx * 3 + 1]; o[2] = palette[idx * 3 + 2];
                    if (has_trns) o[3] = idx < trns.size() ? trns[idx] : 255;
                }
                else
                {
                    for (int c = 0; c < src_channels; ++c)
                    {
                        unsigned v = sample(row, (size_t)x * src_channels + c);
                        o[c] = (unsigned char)(ctype == 0 && depth < 8 ? v * grey_scale : v);
                    }
                    if (has_trns && ctype == 0 && trns.size() >= 2)
                        o[1] = sample16(row, x) == (unsigned)((trns[0] << 8 | trns[1]) & (depth == 16 ? 0xFFFF : (1u << depth) - 1u)) ? 0 : 255;
                    if (has_trns && ctype == 2 && trns.size() >= 6)
                    {
                        bool key = true;
                        for (int c = 0; c < 3; ++c)
                            key = key && sample16(row, (size_t)x * 3 + c) == (unsigned)((trns[c * 2] << 8 | trns[c * 2 + 1]) & (depth == 16 ? 0xFFFF : 0xFF));
                        o[3] = key ? 0 : 255;
                    }
                }
            }
        }
    }
    pack(px, (int)w, (int)h, out_channels, out);
    return true;
}

bool load_tga(const std::vector<unsigned char>& f, TextureImage& out, std::string& err)
{
    if (f.size() < 18) { err = "truncated TGA header"; return false; }
    const int id_len = f[0], cmap_type = f[1], type = f[2];
    const int cmap_first = f[3] | f[4] << 8, cmap_len = f[5] | f[6] << 8, cmap_bits = f[7];
    const int w = f[12] | f[13] << 8, h = f[14] | f[15] << 8, bits = f[16], desc = f[17];
    const bool rle = type >= 8;
    const int base = type & 7;          // 1 colour-mapped, 2 true colour, 3 grey
    if (!(base == 1 || base == 2 || base == 3) || w <= 0 || h <= 0) { err = "unsupported TGA image type"; return false; }
    if (base == 1 && (cmap_type != 1 || (bits != 8 && bits != 16))) { err = "bad colour-mapped TGA"; return false; }
    const int px_bits = base == 1 ? cmap_bits : bits;
    int nc;
    if (base == 3) { if (bits != 8 && bits != 16) { err = "unsupported grey TGA depth"; return false; } nc = bits == 16 ? 2 : 1; }
    else if (px_bits == 15 || px_bits == 16) nc = 3;
    else if (px_bits == 24) nc = 3;
    else if (px_bits == 32) nc = 4;
    else { err = "unsupported TGA pixel depth"; return false; }
    size_t pos = 18 + (size_t)id_len;
    auto decode = [&](const unsigned char* p, int nbits, unsigned char* o) {     // one stored pixel -> nc bytes (RGB order)
        if (base == 3 && nbits == 8) { o[0] = p[0]; return; }
        if (base == 3) { o[0] = p[0]; o[1] = p[1]; return; }
        if (nbits == 15 || nbits == 16)
        {
            unsigned v = p[0] | p[1] << 8;
            unsigned r = (v >> 10) & 31, g = (v >> 5) & 31, b = v & 31;
            o[0] = (unsigned char)((r * 255) / 31); o[1] = (unsigned char)((g * 255) / 31); o[2] = (unsigned char)((b * 255) / 31);
            return;
        }
        o[0] = p[2]; o[1] = p[1]; o[2] = p[0];
        if (nbits == 32) o[3] = p[3];
    };
    std::vector<unsigned char> cmap;
    const int cmap_bytes = (cmap_bits + 7) / 8;
    if (cmap_type == 1)
    {
        size_t n = (size_t)cmap_len * cmap_bytes;
        if (pos + n > f.size()) { err = "truncated TGA colour map"; return false; }
        if (base == 1)
        {
            cmap.resize((size_t)cmap_len * nc);
            for (int i = 0; i < cmap_len; ++i) decode(&f[pos + (size_t)i * cmap_bytes], cmap_bits, &cmap[(size_t)i * nc]);
        }
        pos += n;
    }
    const int stored = (bits + 7) / 8;
    std::vector<unsigned char> px((size_t)w * h * nc);
    size_t count = (size_t)w * h, i = 0;
    auto emit = [&](const unsigned char* p) -> bool {
        unsigned char* o = &px[i * nc];
        if (base == 1)
        {
            int idx = (stored == 1 ? p[0] : (p[0] | p[1] << 8)) - cmap_first;
            if (idx < 0 || idx >= cmap_len) return false;
            memcpy(o, &cmap[(size_t)idx * nc], nc);
        }
        else decode(p, bits, o);
        ++i;
        return true;
    };
    while (i < count)
    {
        if (!rle)
        {
            if (pos + stored > f.size()) { err = "truncated TGA pixel data"; return false; }
            if (!emit(&f[pos])) { err = "TGA colour index out of range"; return false; }
            pos += stored;
        }
        else
        {
            if (pos + 1 + stored > f.size()) { err = "truncated TGA packet"; return false; }
            int hdr = f[pos++], n = (hdr & 127) + 1;
            if (hdr & 128)
            {
                for (int k = 0; k < n && i < count; ++k) if (!emit(&f[pos])) { err = "TGA colour index out of range"; return false; }
                pos += stored;
            }
            else
            {
                if (pos + (size_t)n * stored > f.size()) { err = "truncated TGA packet"; return false; }
                for (int k = 0; k < n && i < count; ++k, pos += stored) if (!emit(&f[pos])) { err = "TGA colour index out of range"; return false; }
            }
        }
    }
    if (!(desc & 0x20))      // stored bottom row first: flip to top-down
        for (int y = 0; y < h / 2; ++y)
            for (size_t x = 0; x < (size_t)w * nc; ++x) std::swap(px[(size_t)y * w * nc + x], px[(size_t)(h - 1 - y) * w * nc + x]);
    if (desc & 0x10)         // stored right to left
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w / 2; ++x)
                for (int c = 0; c < nc; ++c) std::swap(px[((size_t)y * w + x) * nc + c], px[((size_t)y * w + (w - 1 - x)) * nc + c]);
    pack(px, w, h, nc, out);
    return true;
}
} // namespace

bool LoadTextureImage(const char* filename, TextureImage& result, std::string& error)
{
    const char* ext = strrchr(filename, '.');
    if (!ext) { error = "Invalid texture extension"; return false; }       // scene.cpp:286-290
    std::string e(ext);
    for (char& c : e) c = (char)tolower((unsigned char)c);
    std::vector<unsigned char> file;
    if (e == ".png" || e == ".tga" || e == ".jpg")
    {
        // the reference hands all three extensions to one loader that recognises the format by content (scene.cpp:302,
        // image_loader.cpp:35): PNG signature, JPEG start-of-image, TGA otherwise
        if (!read_file(filename, file)) { error = "cannot read the file"; return false; }
        static const unsigned char png_sig[4] = { 137, 80, 78, 71 };
        if (file.size() >= 4 && memcmp(file.data(), png_sig, 4) == 0) return load_png(file, result, error);
        if (file.size() >= 2 && file[0] == 0xFF && file[1] == 0xD8)
        {
            int w = 0, h = 0, nc = 0;
            std::vector<unsigned char> px;
            if (!DecodeJpeg(file.data(), file.size(), w, h, nc, px, error)) return false;
            pack(px, w, h, nc, result);
            return true;
        }
        return load_tga(file, result, error);
    }
    error = "unsupported texture file type (the reference accepts .jpg, .tga and .png, scene.cpp:302)";
    return false;
}

} // namespace rt_host
