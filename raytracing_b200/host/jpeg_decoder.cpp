/*
 * jpeg_decoder.cpp — JPEG textures for the scene loader (baseline and progressive Huffman JPEG, 8-bit samples,
 * 1 / 3 / 4 components, restart intervals, any integer sub-sampling), written for this repository.
 *
 * The reference decodes ".jpg" textures with the stb_image v2.27 it vendors (3rdparty/stb/stb_image.h, called from
 * src/loaders/image_loader.cpp:30-63 with req_comp = 0).  A JPEG file fixes the entropy decoding but NOT the arithmetic of
 * the inverse DCT, of the chroma up-sampling or of YCbCr -> RGB, and a texel that differs by one step changes the rendered
 * image, so this decoder follows the arithmetic stb_image documents for those three stages:
 *   - inverse DCT: the 12-bit fixed-point "islow" factorisation, columns first keeping 2 extra bits ((x + 512) >> 10), then
 *     rows with rounding and the +128 level shift folded in ((x + 65536 + (128 << 17)) >> 17), clamped to 0..255;
 *   - up-sampling: 2x horizontally / vertically with the (3 near + 1 far + 2) >> 2 triangle filter, 2x2 with
 *     (3 * (3 near + far) + neighbour + 8) >> 4, every other ratio by sample replication; output rows pick the "near" chroma row
 *     with the reference's row stepping (the first output row of a pair uses the previous chroma row as "far");
 *   - colour: 20-bit fixed point, constants rounded to 12 bits, the green Cb term truncated to its upper 16 bits;
 *   - channel count handed back: 3 for colour files (also CMYK / YCCK, folded to RGB), 1 for greyscale.
 * tests/test_host.py::test_image_textures_match_the_reference_loader compares the result texel for texel with the reference's
 * own loader (oracle/_ref) where the reference tree exists, including the reference's assets/checker3.jpg.
 */
#include "reference_api.hpp"

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace rt_host
{
namespace
{

// position in the 8x8 block (row-major) of the k-th coefficient of the zig-zag scan (ITU T.81 figure A.6)
struct ZigZag
{
    unsigned char at[64];
    ZigZag()
    {
        int x = 0, y = 0;
        for (int k = 0; k < 64; ++k)
        {
            at[k] = (unsigned char)(y * 8 + x);
            if (((x + y) & 1) == 0)
            {   // moving up-right
                if (x == 7) ++y; else if (y == 0) ++x; else { ++x; --y; }
            }
            else
            {   // moving down-left
                if (y == 7) ++x; else if (x == 0) ++y; else { --x; ++y; }
            }
        }
    }
};
const ZigZag kZigZag;

struct HuffmanTable
{
    bool defined = false;
    unsigned char symbols[256];
    int first_code[17];        // code value of the first code of each length
    int first_index[17];       // index into symbols[] of that code
    int count[17];
    unsigned short fast[512];  // 9-bit prefix -> (length << 8) | symbol, 0 = longer code

    bool build(const unsigned char* counts16, const unsigned char* syms, int n)
    {
        memcpy(symbols, syms, (size_t)n);
        int code = 0, index = 0;
        memset(fast, 0, sizeof(fast));
        for (int len = 1; len <= 16; ++len)
        {
            count[len] = counts16[len - 1];
            first_code[len] = code;
            first_index[len] = index;
            if (code + count[len] > (1 << len)) return false;          // more codes of this length than the prefix code allows
            if (len <= 9)
                for (int i = 0; i < count[len]; ++i)
                {
                    int c = (code + i) << (9 - len);
                    for (int f = 0; f < (1 << (9 - len)); ++f) fast[c + f] = (unsigned short)((len << 8) | symbols[index + i]);
                }
            code = (code + count[len]) << 1;
            index += count[len];
        }
        defined = true;
        return true;
    }
};

// entropy-coded segment reader: MSB-first bits, 0xFF00 un-stuffing, stops (and feeds zero bits) at a marker
struct BitReader
{
    const unsigned char* p;
    const unsigned char* end;
    std::uint32_t acc = 0;     // bits left-aligned
    int n = 0;
    bool at_marker = false;

    void reset() { acc = 0; n = 0; at_marker = false; }
    void fill()
    {
        while (n <= 24)
        {
            unsigned b = 0;
            if (!at_marker)
            {
                if (p >= end) at_marker = true;
                else if (*p == 0xFF)
                {
                    const unsigned char* q = p + 1;
                    while (q < end && *q == 0xFF) ++q;                  // fill bytes before a marker
                    if (q < end && *q == 0x00) { b = 0xFF; p = q + 1; }     // a stuffed 0xFF data byte
                    else at_marker = true;
                }
                else b = *p++;
            }
            acc |= (std::uint32_t)b << (24 - n);
            n += 8;
        }
    }
    int peek(int k) { if (n < k) fill(); return (int)(acc >> (32 - k)); }
    void skip(int k) { acc <<= k; n -= k; }
    int bits(int k) { if (k == 0) return 0; int v = peek(k); skip(k); return v; }
    int bit() { return bits(1); }
    // magnitude-category decode (T.81 F.2.2.1 EXTEND)
    int receive_extend(int k)
    {
        if (k == 0) return 0;
        int v = bits(k);
        return v < (1 << (k - 1)) ? v - (1 << k) + 1 : v;
    }
    int symbol(const HuffmanTable& h)
    {
        int f = h.fast[peek(9)];
        if (f) { skip(f >> 8); return f & 0xFF; }
        int v = peek(16);
        for (int len = 10; len <= 16; ++len)
        {
            int code = v >> (16 - len);
            if (code - h.first_code[len] < h.count[len] && code >= h.first_code[len])
            {
                skip(len);
                return h.symbols[h.first_index[len] + code - h.first_code[len]];
            }
        }
        return -1;
    }
};

struct Component
{
    int id = 0, h = 1, v = 1, tq = 0, hd = 0, ha = 0;
    int x = 0, y = 0;              // samples that belong to the image
    int w2 = 0, h2 = 0;            // padded to whole MCUs
    int dc_pred = 0;
    std::vector<unsigned char> plane;       // w2 x h2 decoded samples
    std::vector<short> coeff;               // progressive: 64 per block, (w2/8) x (h2/8) blocks, row-major inside a block
};

inline unsigned char clamp255(int v) { return (unsigned char)(v < 0 ? 0 : v > 255 ? 255 : v); }

// 12-bit fixed-point constants of the inverse DCT: v * 4096 + 0.5 truncated TOWARDS ZERO, so the negative constants are one
// step smaller in magnitude than the negated positive ones (fx(-0.899976223) = -3685, not -3686)
constexpr int fx(double v) { return (int)(v * 4096 + 0.5); }

// one 8-point pass; 64-bit intermediates so that the coefficients of a corrupt file cannot overflow (a valid file stays far
// inside 32 bits, where the result is the same)
typedef long long wide;
struct Idct1D { wide x0, x1, x2, x3, t0, t1, t2, t3; };

inline Idct1D idct_1d(wide s0, wide s1, wide s2, wide s3, wide s4, wide s5, wide s6, wide s7)
{
    Idct1D r;
    // even part
    wide p1 = (s2 + s6) * fx(0.5411961f);
    wide e2 = p1 + s6 * fx(-1.847759065f);
    wide e3 = p1 + s2 * fx(0.765366865f);
    wide e0 = (s0 + s4) * 4096;
    wide e1 = (s0 - s4) * 4096;
    r.x0 = e0 + e3; r.x3 = e0 - e3; r.x1 = e1 + e2; r.x2 = e1 - e2;
    // odd part
    wide o0 = s7, o1 = s5, o2 = s3, o3 = s1;
    wide p3 = o0 + o2, p4 = o1 + o3;
    wide q1 = o0 + o3, q2 = o1 + o2;
    wide p5 = (p3 + p4) * fx(1.175875602f);
    o0 *= fx(0.298631336f); o1 *= fx(2.053119869f); o2 *= fx(3.072711026f); o3 *= fx(1.501321110f);
    q1 = p5 + q1 * fx(-0.899976223f);
    q2 = p5 + q2 * fx(-2.562915447f);
    p3 *= fx(-1.961570560f);
    p4 *= fx(-0.390180644f);
    r.t3 = o3 + q1 + p4; r.t2 = o2 + q2 + p3; r.t1 = o1 + q2 + p4; r.t0 = o0 + q1 + p3;
    return r;
}

inline unsigned char clamp255w(wide v) { return (unsigned char)(v < 0 ? 0 : v > 255 ? 255 : v); }

void idct_block(unsigned char* out, int stride, const short* d)
{
    int tmp[64];
    for (int c = 0; c < 8; ++c)
    {
        const short* s = d + c;
        int* v = tmp + c;
        if (!s[8] && !s[16] && !s[24] && !s[32] && !s[40] && !s[48] && !s[56])
        {
            int dc = s[0] * 4;
            for (int r = 0; r < 8; ++r) v[r * 8] = dc;
            continue;
        }
        Idct1D k = idct_1d(s[0], s[8], s[16], s[24], s[32], s[40], s[48], s[56]);
        const int rnd = 512;
        v[0] = (int)((k.x0 + rnd + k.t3) >> 10); v[56] = (int)((k.x0 + rnd - k.t3) >> 10);
        v[8] = (int)((k.x1 + rnd + k.t2) >> 10); v[48] = (int)((k.x1 + rnd - k.t2) >> 10);
        v[16] = (int)((k.x2 + rnd + k.t1) >> 10); v[40] = (int)((k.x2 + rnd - k.t1) >> 10);
        v[24] = (int)((k.x3 + rnd + k.t0) >> 10); v[32] = (int)((k.x3 + rnd - k.t0) >> 10);
    }
    for (int r = 0; r < 8; ++r, out += stride)
    {
        const int* v = tmp + r * 8;
        Idct1D k = idct_1d(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
        const int rnd = 65536 + (128 << 17);
        out[0] = clamp255w((k.x0 + rnd + k.t3) >> 17); out[7] = clamp255w((k.x0 + rnd - k.t3) >> 17);
        out[1] = clamp255w((k.x1 + rnd + k.t2) >> 17); out[6] = clamp255w((k.x1 + rnd - k.t2) >> 17);
        out[2] = clamp255w((k.x2 + rnd + k.t1) >> 17); out[5] = clamp255w((k.x2 + rnd - k.t1) >> 17);
        out[3] = clamp255w((k.x3 + rnd + k.t0) >> 17); out[4] = clamp255w((k.x3 + rnd - k.t0) >> 17);
    }
}

// ---- chroma up-sampling of one output row: `near` is the closer source row, `far` the other one
const unsigned char* upsample_row(unsigned char* out, const unsigned char* near, const unsigned char* far, int w, int hs, int vs)
{
    if (hs == 1 && vs == 1) return near;
    if (hs == 1 && vs == 2)
    {
        for (int i = 0; i < w; ++i) out[i] = (unsigned char)((3 * near[i] + far[i] + 2) >> 2);
        return out;
    }
    if (hs == 2 && vs == 1)
    {
        if (w == 1) { out[0] = out[1] = near[0]; return out; }
        out[0] = near[0];
        out[1] = (unsigned char)((near[0] * 3 + near[1] + 2) >> 2);
        for (int i = 1; i < w - 1; ++i)
        {
            int n = 3 * near[i] + 2;
            out[2 * i] = (unsigned char)((n + near[i - 1]) >> 2);
            out[2 * i + 1] = (unsigned char)((n + near[i + 1]) >> 2);
        }
        out[2 * (w - 1)] = (unsigned char)((near[w - 2] * 3 + near[w - 1] + 2) >> 2);
        out[2 * (w - 1) + 1] = near[w - 1];
        return out;
    }
    if (hs == 2 && vs == 2)
    {
        if (w == 1) { out[0] = out[1] = (unsigned char)((3 * near[0] + far[0] + 2) >> 2); return out; }
        int cur = 3 * near[0] + far[0];
        out[0] = (unsigned char)((cur + 2) >> 2);
        for (int i = 1; i < w; ++i)
        {
            int prev = cur;
            cur = 3 * near[i] + far[i];
            out[2 * i - 1] = (unsigned char)((3 * prev + cur + 8) >> 4);
            out[2 * i] = (unsigned char)((3 * cur + prev + 8) >> 4);
        }
        out[2 * w - 1] = (unsigned char)((cur + 2) >> 2);
        return out;
    }
    for (int i = 0; i < w; ++i)
        for (int j = 0; j < hs; ++j) out[i * hs + j] = near[i];
    return out;
}

// 12-bit constants, carried at 20 bits
constexpr int cfx(float v) { return ((int)(v * 4096.0f + 0.5f)) << 8; }

void ycc_to_rgb_row(unsigned char* out, const unsigned char* y, const unsigned char* cb_row, const unsigned char* cr_row, int count)
{
    for (int i = 0; i < count; ++i, out += 3)
    {
        int yf = (y[i] << 20) + (1 << 19);
        int cr = cr_row[i] - 128, cb = cb_row[i] - 128;
        int r = yf + cr * cfx(1.40200f);
        int g = yf + cr * -cfx(0.71414f) + (int)((unsigned)(cb * -cfx(0.34414f)) & 0xffff0000u);
        int b = yf + cb * cfx(1.77200f);
        out[0] = clamp255(r >> 20); out[1] = clamp255(g >> 20); out[2] = clamp255(b >> 20);
    }
}

// x * y / 255, rounded
inline unsigned char mul255(unsigned x, unsigned y) { unsigned t = x * y + 128; return (unsigned char)((t + (t >> 8)) >> 8); }

struct Decoder
{
    const unsigned char* data;
    size_t size;
    size_t pos = 0;
    std::string& err;

    int width = 0, height = 0, ncomp = 0;
    bool progressive = false, jfif = false;
    int adobe_transform = -1;
    int rgb_ids = 0;
    int h_max = 1, v_max = 1, mcu_x = 0, mcu_y = 0;
    int restart_interval = 0;
    Component comp[4];
    std::uint16_t quant[4][64];          // row-major position inside the block
    HuffmanTable dc_tab[4], ac_tab[4];
    // current scan
    int scan_n = 0, order[4] = {}, ss = 0, se = 63, ah = 0, al = 0;
    int eob_run = 0;
    bool frame_seen = false;

    Decoder(const unsigned char* d, size_t n, std::string& e) : data(d), size(n), err(e) { memset(quant, 0, sizeof(quant)); }

    bool fail(const char* what) { err = std::string("JPEG: ") + what; return false; }
    bool have(size_t n) const { return pos + n <= size; }
    int u8() { return pos < size ? data[pos++] : 0; }
    int u16() { int a = u8(); return (a << 8) | u8(); }

    // next marker code at or after pos (fill bytes and stray data skipped); 0 at the end of the file
    int next_marker()
    {
        while (pos + 1 < size)
        {
            if (data[pos] == 0xFF && data[pos + 1] != 0x00 && data[pos + 1] != 0xFF) { int m = data[pos + 1]; pos += 2; return m; }
            ++pos;
        }
        pos = size;
        return 0;
    }

    bool read_tables_or_misc(int m)
    {
        if (m == 0xDD)
        {
            if (!have(4) || u16() != 4) return fail("bad DRI segment");
            restart_interval = u16();
            return true;
        }
        if (!have(2)) return fail("truncated segment");
        int len = u16() - 2;
        if (len < 0 || !have((size_t)len)) return fail("truncated segment");
        size_t seg_end = pos + (size_t)len;
        if (m == 0xDB)
        {
            while (pos < seg_end)
            {
                int q = u8(), precision = q >> 4, t = q & 15;
                if (precision > 1 || t > 3) return fail("bad quantisation table header");
                if (pos + (precision ? 128u : 64u) > seg_end) return fail("bad DQT length");
                for (int k = 0; k < 64; ++k) quant[t][kZigZag.at[k]] = (std::uint16_t)(precision ? u16() : u8());
            }
        }
        else if (m == 0xC4)
        {
            while (pos < seg_end)
            {
                if (pos + 17 > seg_end) return fail("bad DHT length");
                int q = u8(), tc = q >> 4, th = q & 15;
                if (tc > 1 || th > 3) return fail("bad Huffman table header");
                unsigned char counts[16];
                int n = 0;
                for (int i = 0; i < 16; ++i) { counts[i] = (unsigned char)u8(); n += counts[i]; }
                if (n > 256 || pos + (size_t)n > seg_end) return fail("bad DHT length");
                if (!(tc ? ac_tab : dc_tab)[th].build(counts, data + pos, n)) return fail("bad Huffman code lengths");
                pos += (size_t)n;
            }
        }
        else if (m == 0xE0 && len >= 5) jfif = jfif || memcmp(data + pos, "JFIF\0", 5) == 0;
        else if (m == 0xEE && len >= 12 && memcmp(data + pos, "Adobe\0", 6) == 0) adobe_transform = data[pos + 11];
        else if (!((m >= 0xE0 && m <= 0xEF) || m == 0xFE)) return fail("unsupported marker (arithmetic coding, lossless and hierarchical JPEG are not decoded)");
        pos = seg_end;
        return true;
    }

    bool read_frame(int m)
    {
        if (frame_seen) return fail("more than one frame");
        frame_seen = true;
        progressive = m == 0xC2;
        if (!have(8)) return fail("truncated frame header");
        int len = u16();
        if (u8() != 8) return fail("only 8-bit samples are supported");
        height = u16(); width = u16(); ncomp = u8();
        if (height == 0 || width == 0) return fail("zero image size");
        if (width > (1 << 24) || height > (1 << 24)) return fail("image too large");
        if (ncomp != 1 && ncomp != 3 && ncomp != 4) return fail("bad component count");
        if (len != 8 + 3 * ncomp || !have((size_t)3 * ncomp)) return fail("bad frame header length");
        for (int i = 0; i < ncomp; ++i)
        {
            Component& c = comp[i];
            c.id = u8();
            if (ncomp == 3 && c.id == "RGB"[i]) ++rgb_ids;
            int q = u8();
            c.h = q >> 4; c.v = q & 15; c.tq = u8();
            if (c.h < 1 || c.h > 4 || c.v < 1 || c.v > 4 || c.tq > 3) return fail("bad sampling factors");
            if (c.h > h_max) h_max = c.h;
            if (c.v > v_max) v_max = c.v;
        }
        for (int i = 0; i < ncomp; ++i)
            if (h_max % comp[i].h != 0 || v_max % comp[i].v != 0) return fail("fractional sub-sampling ratios are not supported");
        mcu_x = (width + h_max * 8 - 1) / (h_max * 8);
        mcu_y = (height + v_max * 8 - 1) / (v_max * 8);
        for (int i = 0; i < ncomp; ++i)
        {
            Component& c = comp[i];
            c.x = (width * c.h + h_max - 1) / h_max;
            c.y = (height * c.v + v_max - 1) / v_max;
            c.w2 = mcu_x * c.h * 8; c.h2 = mcu_y * c.v * 8;
            c.plane.assign((size_t)c.w2 * c.h2, 0);
            if (progressive) c.coeff.assign((size_t)c.w2 * c.h2, 0);
        }
        return true;
    }

    bool read_scan_header()
    {
        if (!have(3)) return fail("truncated scan header");
        int len = u16();
        scan_n = u8();
        if (scan_n < 1 || scan_n > 4 || scan_n > ncomp) return fail("bad scan component count");
        if (len != 6 + 2 * scan_n || !have((size_t)2 * scan_n + 3)) return fail("bad scan header length");
        for (int i = 0; i < scan_n; ++i)
        {
            int id = u8(), q = u8(), which = 0;
            while (which < ncomp && comp[which].id != id) ++which;
            if (which == ncomp) return fail("scan names an unknown component");
            comp[which].hd = q >> 4; comp[which].ha = q & 15;
            if (comp[which].hd > 3 || comp[which].ha > 3) return fail("bad Huffman table index");
            order[i] = which;
        }
        ss = u8(); se = u8();
        int a = u8(); ah = a >> 4; al = a & 15;
        if (progressive)
        {
            if (ss > 63 || se > 63 || ss > se || ah > 13 || al > 13) return fail("bad progressive scan parameters");
        }
        else
        {
            if (ss != 0 || ah != 0 || al != 0) return fail("bad scan parameters");
            se = 63;
        }
        return true;
    }

    // ---- block decoders
    bool block_baseline(BitReader& br, Component& c, short* blk)
    {
        const HuffmanTable& hd = dc_tab[c.hd];
        const HuffmanTable& ha = ac_tab[c.ha];
        if (!hd.defined || !ha.defined) return fail("scan uses an undefined Huffman table");
        const std::uint16_t* dq = quant[c.tq];
        memset(blk, 0, 64 * sizeof(short));
        int t = br.symbol(hd);
        if (t < 0 || t > 15) return fail("bad Huffman code");
        c.dc_pred += br.receive_extend(t);
        blk[0] = (short)(c.dc_pred * dq[0]);
        for (int k = 1; k < 64;)
        {
            int rs = br.symbol(ha);
            if (rs < 0) return fail("bad Huffman code");
            int s = rs & 15, r = rs >> 4;
            if (s == 0)
            {
                if (rs != 0xF0) break;
                k += 16;
                continue;
            }
            k += r;
            if (k > 63) return fail("coefficient index out of range");
            int at = kZigZag.at[k++];
            blk[at] = (short)(br.receive_extend(s) * dq[at]);
        }
        return true;
    }

    bool block_dc_progressive(BitReader& br, Component& c, short* blk)
    {
        if (se != 0) return fail("progressive scan mixes DC and AC coefficients");
        if (ah == 0)
        {
            const HuffmanTable& hd = dc_tab[c.hd];
            if (!hd.defined) return fail("scan uses an undefined Huffman table");
            memset(blk, 0, 64 * sizeof(short));
            int t = br.symbol(hd);
            if (t < 0 || t > 15) return fail("bad Huffman code");
            c.dc_pred += br.receive_extend(t);
            blk[0] = (short)(c.dc_pred * (1 << al));
        }
        else if (br.bit()) blk[0] = (short)(blk[0] + (short)(1 << al));
        return true;
    }

    bool block_ac_progressive(BitReader& br, Component& c, short* blk)
    {
        if (ss == 0) return fail("progressive scan mixes DC and AC coefficients");
        const HuffmanTable& ha = ac_tab[c.ha];
        if (!ha.defined) return fail("scan uses an undefined Huffman table");
        if (ah == 0)
        {
            if (eob_run) { --eob_run; return true; }
            for (int k = ss; k <= se;)
            {
                int rs = br.symbol(ha);
                if (rs < 0) return fail("bad Huffman code");
                int s = rs & 15, r = rs >> 4;
                if (s == 0)
                {
                    if (r < 15)
                    {
                        eob_run = (1 << r) - 1;
                        if (r) eob_run += br.bits(r);
                        break;
                    }
                    k += 16;
                    continue;
                }
                k += r;
                if (k > 63) return fail("coefficient index out of range");
                blk[kZigZag.at[k++]] = (short)(br.receive_extend(s) * (1 << al));
            }
            return true;
        }
        // refinement: one more bit for every coefficient that is already non-zero, new +-1 coefficients in between
        const short bit = (short)(1 << al);
        auto refine = [&](short& p) {
            if (br.bit() && (p & bit) == 0) p = (short)(p > 0 ? p + bit : p - bit);
        };
        if (eob_run)
        {
            --eob_run;
            for (int k = ss; k <= se; ++k)
            {
                short& p = blk[kZigZag.at[k]];
                if (p != 0) refine(p);
            }
            return true;
        }
        int k = ss;
        do
        {
            int rs = br.symbol(ha);
            if (rs < 0) return fail("bad Huffman code");
            int s = rs & 15, r = rs >> 4;
            if (s == 0)
            {
                if (r < 15)
                {
                    eob_run = (1 << r) - 1;
                    if (r) eob_run += br.bits(r);
                    r = 64;                 // run to the end of the band
                }
            }
            else
            {
                if (s != 1) return fail("bad Huffman code");
                s = br.bit() ? bit : -bit;
            }
            while (k <= se)
            {
                short& p = blk[kZigZag.at[k++]];
                if (p != 0) refine(p);
                else
                {
                    if (r == 0) { p = (short)s; break; }
                    --r;
                }
            }
        } while (k <= se);
        return true;
    }

    bool decode_scan()
    {
        BitReader br;
        br.p = data + pos; br.end = data + size;
        for (int i = 0; i < 4; ++i) comp[i].dc_pred = 0;
        eob_run = 0;
        int todo = restart_interval ? restart_interval : 0x7fffffff;
        short local[64];

        // after `restart_interval` MCUs: byte-align, expect RSTn, clear the predictors; anything else ends the scan
        auto restart = [&]() -> bool {
            br.fill();
            const unsigned char* q = br.p;
            while (q + 1 < br.end && q[0] == 0xFF && q[1] == 0xFF) ++q;
            if (!(br.at_marker && q + 1 < br.end && q[0] == 0xFF && q[1] >= 0xD0 && q[1] <= 0xD7)) return false;
            br.p = q + 2;
            br.reset();
            for (int i = 0; i < 4; ++i) comp[i].dc_pred = 0;
            eob_run = 0;
            todo = restart_interval ? restart_interval : 0x7fffffff;
            return true;
        };
        auto one_block = [&](Component& c, int bx, int by) -> bool {
            if (!progressive)
            {
                if (!block_baseline(br, c, local)) return false;
                idct_block(&c.plane[(size_t)c.w2 * by * 8 + (size_t)bx * 8], c.w2, local);
                return true;
            }
            short* blk = &c.coeff[64 * ((size_t)bx + (size_t)by * (c.w2 / 8))];
            return ss == 0 ? block_dc_progressive(br, c, blk) : block_ac_progressive(br, c, blk);
        };

        bool stop = false;
        if (scan_n == 1)
        {
            Component& c = comp[order[0]];
            int bw = (c.x + 7) >> 3, bh = (c.y + 7) >> 3;
            for (int j = 0; j < bh && !stop; ++j)
                for (int i = 0; i < bw && !stop; ++i)
                {
                    if (!one_block(c, i, j)) return false;
                    if (--todo <= 0 && !restart()) stop = true;
                }
        }
        else
        {
            if (progressive && ss != 0) return fail("interleaved progressive scan with AC coefficients");
            for (int j = 0; j < mcu_y && !stop; ++j)
                for (int i = 0; i < mcu_x && !stop; ++i)
                {
                    for (int k = 0; k < scan_n; ++k)
                    {
                        Component& c = comp[order[k]];
                        for (int y = 0; y < c.v; ++y)
                            for (int x = 0; x < c.h; ++x)
                                if (!one_block(c, i * c.h + x, j * c.v + y)) return false;
                    }
                    if (--todo <= 0 && !restart()) stop = true;
                }
        }
        pos = (size_t)(br.p - data);
        return true;
    }

    void finish_progressive()
    {
        short blk[64];
        for (int n = 0; n < ncomp; ++n)
        {
            Component& c = comp[n];
            const std::uint16_t* dq = quant[c.tq];
            int bw = (c.x + 7) >> 3, bh = (c.y + 7) >> 3;
            for (int j = 0; j < bh; ++j)
                for (int i = 0; i < bw; ++i)
                {
                    const short* src = &c.coeff[64 * ((size_t)i + (size_t)j * (c.w2 / 8))];
                    for (int k = 0; k < 64; ++k) blk[k] = (short)(src[k] * dq[k]);
                    idct_block(&c.plane[(size_t)c.w2 * j * 8 + (size_t)i * 8], c.w2, blk);
                }
        }
    }

    bool decode_planes()
    {
        if (size < 2 || data[0] != 0xFF || data[1] != 0xD8) return fail("not a JPEG file");
        pos = 2;
        bool got_scan = false;
        for (;;)
        {
            int m = next_marker();
            if (m == 0) { if (got_scan) break; return fail("no image data"); }
            if (m == 0xD9) break;
            if (m == 0xC0 || m == 0xC1 || m == 0xC2) { if (!read_frame(m)) return false; }
            else if (m == 0xDA)
            {
                if (!frame_seen) return fail("scan before the frame header");
                if (!read_scan_header() || !decode_scan()) return false;
                got_scan = true;
            }
            else if (m == 0xDC)
            {
                if (!have(4) || u16() != 4) return fail("bad DNL segment");
                if (u16() != height) return fail("DNL height differs from the frame header");
            }
            else if (m >= 0xD0 && m <= 0xD7) continue;             // stray restart marker
            else if (!read_tables_or_misc(m)) return false;
        }
        if (!frame_seen || !got_scan) return fail("no image data");
        if (progressive) finish_progressive();
        return true;
    }

    bool decode(int& w, int& h, int& nc, std::vector<unsigned char>& px)
    {
        if (!decode_planes()) return false;
        w = width; h = height;
        nc = ncomp >= 3 ? 3 : 1;
        const bool is_rgb = ncomp == 3 && (rgb_ids == 3 || (adobe_transform == 0 && !jfif));
        px.assign((size_t)w * h * nc, 0);
        struct Up { int hs, vs, w_lores, ystep, ypos; size_t line0, line1; std::vector<unsigned char> buf; };
        Up up[4];
        for (int k = 0; k < ncomp; ++k)
        {
            up[k].hs = h_max / comp[k].h; up[k].vs = v_max / comp[k].v;
            up[k].ystep = up[k].vs >> 1;
            up[k].w_lores = (w + up[k].hs - 1) / up[k].hs;
            up[k].ypos = 0; up[k].line0 = up[k].line1 = 0;
            up[k].buf.resize((size_t)w + 8);
        }
        const unsigned char* row[4] = {};
        for (int j = 0; j < h; ++j)
        {
            unsigned char* out = &px[(size_t)j * w * nc];
            for (int k = 0; k < ncomp; ++k)
            {
                Up& u = up[k];
                const unsigned char* plane = comp[k].plane.data();
                bool bottom = u.ystep >= (u.vs >> 1);
                row[k] = upsample_row(u.buf.data(), plane + (bottom ? u.line1 : u.line0), plane + (bottom ? u.line0 : u.line1), u.w_lores, u.hs, u.vs);
                if (++u.ystep >= u.vs)
                {
                    u.ystep = 0;
                    u.line0 = u.line1;
                    if (++u.ypos < comp[k].y) u.line1 += (size_t)comp[k].w2;
                }
            }
            if (ncomp == 1) memcpy(out, row[0], (size_t)w);
            else if (ncomp == 3)
            {
                if (is_rgb) for (int i = 0; i < w; ++i) { out[3 * i] = row[0][i]; out[3 * i + 1] = row[1][i]; out[3 * i + 2] = row[2][i]; }
                else ycc_to_rgb_row(out, row[0], row[1], row[2], w);
            }
            else if (adobe_transform == 0)          // CMYK
                for (int i = 0; i < w; ++i)
                    for (int c = 0; c < 3; ++c) out[3 * i + c] = mul255(row[c][i], row[3][i]);
            else
            {
                ycc_to_rgb_row(out, row[0], row[1], row[2], w);
                if (adobe_transform == 2)           // YCCK
                    for (int i = 0; i < w; ++i)
                        for (int c = 0; c < 3; ++c) out[3 * i + c] = mul255(255u - out[3 * i + c], row[3][i]);
            }
        }
        return true;
    }
};

} // namespace

bool DecodeJpeg(const unsigned char* file, size_t size, int& w, int& h, int& channels, std::vector<unsigned char>& pixels, std::string& error)
{
    Decoder d(file, size, error);
    return d.decode(w, h, channels, pixels);
}

} // namespace rt_host
