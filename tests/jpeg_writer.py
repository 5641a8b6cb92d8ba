"""
tests/jpeg_writer.py — TEST INFRASTRUCTURE: a small baseline / progressive JPEG writer for the layouts Pillow's encoder
cannot produce, so that host/jpeg_decoder.cpp is exercised (against the reference's loader, tests/golden/make_texture_fixtures.py)
on every path it has: arbitrary sampling factors (4:4:0, 4:1:1, 4:1:0, luma sub-sampled below chroma), one scan per component,
restart intervals, 16-bit quantisation tables, component ids 'R' 'G' 'B', Adobe APP14 colour transforms (CMYK / YCCK / RGB),
fill bytes before markers, a DNL segment, spectral-selection + successive-approximation progressive scripts.

The sample data is arbitrary (smooth noise); nothing here tries to be a good encoder.  The Huffman tables are the ones Pillow
writes into a non-optimised file (the tables of ITU T.81 Annex K), read back from such a file.
"""
import io
import struct

import numpy as np
from scipy.fft import dctn

ZIGZAG = []
_x = _y = 0
for _k in range(64):
    ZIGZAG.append(_y * 8 + _x)
    if (_x + _y) % 2 == 0:
        if _x == 7: _y += 1
        elif _y == 0: _x += 1
        else: _x += 1; _y -= 1
    else:
        if _y == 7: _x += 1
        elif _x == 0: _y += 1
        else: _x -= 1; _y += 1
ZIGZAG = np.array(ZIGZAG)


def _standard_tables():
    """{(class, id): (counts[16], symbols)} parsed from a JPEG Pillow writes with its default tables."""
    from PIL import Image
    buf = io.BytesIO()
    Image.fromarray(np.zeros((8, 8, 3), dtype=np.uint8)).save(buf, "JPEG", quality=50)
    d = buf.getvalue()
    out, pos = {}, 2
    while pos < len(d):
        assert d[pos] == 0xFF
        m, ln = d[pos + 1], struct.unpack(">H", d[pos + 2: pos + 4])[0]
        if m == 0xC4:
            p, end = pos + 4, pos + 2 + ln
            while p < end:
                tc, th = d[p] >> 4, d[p] & 15
                counts = list(d[p + 1: p + 17])
                n = sum(counts)
                out[(tc, th)] = (counts, list(d[p + 17: p + 17 + n]))
                p += 17 + n
        if m == 0xDA:
            break
        pos += 2 + ln
    return out


class _Huff:
    def __init__(self, counts, symbols):
        self.counts, self.symbols = counts, symbols
        self.code = {}
        code = k = 0
        for length in range(1, 17):
            for _ in range(counts[length - 1]):
                self.code[symbols[k]] = (code, length)
                code += 1; k += 1
            code <<= 1


class _Bits:
    def __init__(self):
        self.out = bytearray(); self.acc = 0; self.n = 0

    def put(self, value, length):
        if length == 0:
            return
        self.acc = (self.acc << length) | (value & ((1 << length) - 1)); self.n += length
        while self.n >= 8:
            b = (self.acc >> (self.n - 8)) & 0xFF
            self.out.append(b)
            if b == 0xFF:
                self.out.append(0)
            self.n -= 8
        self.acc &= (1 << self.n) - 1

    def flush(self):
        if self.n:
            self.put((1 << (8 - self.n)) - 1, 8 - self.n)


def _magnitude(v):
    """(category, extra bits) of a coefficient (T.81 F.1.2.1)."""
    a = abs(int(v))
    s = a.bit_length()
    return s, (int(v) if v >= 0 else int(v) + (1 << s) - 1)


def _segment(marker, payload):
    return bytes([0xFF, marker]) + struct.pack(">H", len(payload) + 2) + payload


def smooth_plane(rng, h, w):
    yy, xx = np.mgrid[0:h, 0:w]
    a = 127 + 100 * np.sin(xx / rng.uniform(3, 9) + yy / rng.uniform(3, 9)) + rng.normal(0, 15, (h, w))
    return np.clip(a, 0, 255)


def write_jpeg(path, width, height, sampling, rng, *, quant_scale=8, sixteen_bit_quant=False, interleaved=True, restart_interval=0,
               ids=None, adobe_transform=None, jfif=True, fill_bytes=False, dnl=False, progressive=False):
    """sampling: [(h, v)] per component (1, 3 or 4 components).  progressive: DC first (Al = 1), AC bands 1-5 and 6-63 per
    component (Al = 1 for the low band), then the refinement scans."""
    nc = len(sampling)
    ids = ids or list(range(1, nc + 1))
    h_max, v_max = max(s[0] for s in sampling), max(s[1] for s in sampling)
    mcu_x, mcu_y = -(-width // (8 * h_max)), -(-height // (8 * v_max))
    tabs = _standard_tables()
    dc = _Huff(*tabs[(0, 0)])
    if progressive:
        # the Annex K table has no EOBn symbols: a table over every (run, size <= 10) pair, 16 codes of 5 bits + 160 of 10 bits
        symbols = [(r << 4) | sz for sz in range(11) for r in range(16)]
        tabs[(1, 0)] = ([0, 0, 0, 0, 16, 0, 0, 0, 0, 160, 0, 0, 0, 0, 0, 0], symbols)
    ac = _Huff(*tabs[(1, 0)])
    qt = np.clip((np.add.outer(np.arange(8), np.arange(8)) + 2) * quant_scale, 1, 65535 if sixteen_bit_quant else 255).astype(np.int64)
    if sixteen_bit_quant:
        qt[7, 7] = 300
    # quantised coefficients per component: (blocks_y, blocks_x, 64) in zig-zag order
    coeffs = []
    for (ch, cv) in sampling:
        bw, bh = mcu_x * ch, mcu_y * cv
        plane = smooth_plane(rng, bh * 8, bw * 8) - 128.0
        blocks = plane.reshape(bh, 8, bw, 8).transpose(0, 2, 1, 3)
        f = dctn(blocks, axes=(2, 3), norm="ortho")
        q = np.rint(f / qt).astype(np.int64).reshape(bh, bw, 64)[:, :, ZIGZAG]
        coeffs.append(q)

    out = bytearray(b"\xff\xd8")
    if jfif:
        out += _segment(0xE0, b"JFIF\0\x01\x01\0\0\x01\0\x01\0\0")
    if adobe_transform is not None:
        out += _segment(0xEE, b"Adobe\0" + bytes([100, 0, 0, 0, 0, adobe_transform]))
    out += _segment(0xFE, b"written by tests/jpeg_writer.py")
    zz = qt.reshape(64)[ZIGZAG]
    out += _segment(0xDB, (bytes([0x10]) + b"".join(struct.pack(">H", int(v)) for v in zz)) if sixteen_bit_quant else (bytes([0]) + bytes(int(v) for v in zz)))
    sof = struct.pack(">BHHB", 8, height, width, nc)          # (the reference's decoder refuses a zero height: with `dnl` the DNL segment repeats it)
    for i, (ch, cv) in enumerate(sampling):
        sof += bytes([ids[i], (ch << 4) | cv, 0])
    out += _segment(0xC2 if progressive else 0xC0, sof)
    for (tc, th) in ((0, 0), (1, 0)):
        counts, symbols = tabs[(tc, th)]
        out += _segment(0xC4, bytes([(tc << 4) | th]) + bytes(counts) + bytes(symbols))
    if restart_interval:
        out += _segment(0xDD, struct.pack(">H", restart_interval))

    def scan(comps, ss, se, ah, al):
        """One scan over `comps` (indices).  Baseline: ss, se, ah, al = 0, 63, 0, 0."""
        nonlocal out
        hdr = bytes([len(comps)]) + b"".join(bytes([ids[c], 0x00]) for c in comps) + bytes([ss, se, (ah << 4) | al])
        if fill_bytes:
            out += b"\xff\xff"
        out += _segment(0xDA, hdr)
        bits = _Bits()
        pred = [0] * nc
        state = {"eobrun": 0, "pending": []}
        units = []                         # list of MCUs, each a list of (component, by, bx)
        if len(comps) == 1:
            c = comps[0]
            ch, cv = sampling[c]
            bw = -(-(-(-width * ch // h_max)) // 8)
            bh = -(-(-(-height * cv // v_max)) // 8)
            units = [[(c, by, bx)] for by in range(bh) for bx in range(bw)]
        else:
            for my in range(mcu_y):
                for mx in range(mcu_x):
                    u = []
                    for c in comps:
                        ch, cv = sampling[c]
                        u += [(c, my * cv + y, mx * ch + x) for y in range(cv) for x in range(ch)]
                    units.append(u)

        def flush_eobrun():
            if state["eobrun"]:
                r = state["eobrun"].bit_length() - 1
                bits.put(*ac.code[r << 4])
                bits.put(state["eobrun"] - (1 << r), r)
                state["eobrun"] = 0
            for b in state["pending"]:
                bits.put(b, 1)
            state["pending"] = []

        def encode_block(c, blk):
            if not progressive:
                s, extra = _magnitude(blk[0] - pred[c]); pred[c] = int(blk[0])
                bits.put(*dc.code[s]); bits.put(extra, s)
                run = 0
                for k in range(1, 64):
                    v = int(blk[k])
                    if v == 0:
                        run += 1; continue
                    while run > 15:
                        bits.put(*ac.code[0xF0]); run -= 16
                    s, extra = _magnitude(v)
                    bits.put(*ac.code[(run << 4) | s]); bits.put(extra, s); run = 0
                if run:
                    bits.put(*ac.code[0x00])
                return
            if ss == 0:
                if ah == 0:
                    v = int(blk[0]) >> al
                    s, extra = _magnitude(v - pred[c]); pred[c] = v
                    bits.put(*dc.code[s]); bits.put(extra, s)
                else:
                    bits.put((int(blk[0]) >> al) & 1, 1)
                return
            # AC scans: point transform is a division towards zero
            def pt(v, shift):
                return (abs(int(v)) >> shift) * (1 if v >= 0 else -1)
            if ah == 0:
                run = 0
                for k in range(ss, se + 1):
                    v = pt(blk[k], al)
                    if v == 0:
                        run += 1; continue
                    flush_eobrun()
                    while run > 15:
                        bits.put(*ac.code[0xF0]); run -= 16
                    s, extra = _magnitude(v)
                    bits.put(*ac.code[(run << 4) | s]); bits.put(extra, s); run = 0
                if run:
                    state["eobrun"] += 1
                    if state["eobrun"] == 0x7FFF:
                        flush_eobrun()
                return
            # refinement (T.81 G.1.2.3)
            run = 0
            buffered = []
            vals = [pt(blk[k], al) for k in range(ss, se + 1)]
            last_new = max([i for i, v in enumerate(vals) if abs(v) == 1], default=-1)
            for i, v in enumerate(vals):
                if v == 0:
                    run += 1; continue
                while run > 15 and i <= last_new:
                    flush_eobrun()
                    bits.put(*ac.code[0xF0]); run -= 16
                    for b in buffered: bits.put(b, 1)
                    buffered = []
                if abs(v) > 1:
                    buffered.append(abs(v) & 1)
                    continue
                flush_eobrun()
                bits.put(*ac.code[(run << 4) | 1]); bits.put(1 if v > 0 else 0, 1)
                for b in buffered: bits.put(b, 1)
                buffered = []; run = 0
            if run or buffered:
                state["eobrun"] += 1
                state["pending"] += buffered
                if state["eobrun"] == 0x7FFF or len(state["pending"]) > 900:
                    flush_eobrun()

        rst = 0
        for n, u in enumerate(units):
            if restart_interval and n and n % restart_interval == 0:
                flush_eobrun(); bits.flush()
                if fill_bytes:
                    bits.out += b"\xff"
                bits.out += bytes([0xFF, 0xD0 + (rst & 7)]); rst += 1
                pred = [0] * nc
            for (c, by, bx) in u:
                encode_block(c, coeffs[c][by, bx])
        flush_eobrun(); bits.flush()
        out += bits.out

    if not progressive:
        if interleaved and nc > 1:
            scan(list(range(nc)), 0, 63, 0, 0)
        else:
            for c in range(nc):
                scan([c], 0, 63, 0, 0)
    else:
        scan(list(range(nc)) if (interleaved and nc > 1) else [0], 0, 0, 0, 1)
        if not (interleaved and nc > 1):
            for c in range(1, nc):
                scan([c], 0, 0, 0, 1)
        for c in range(nc):
            scan([c], 1, 5, 0, 1)
            scan([c], 6, 63, 0, 0)
        scan(list(range(nc)) if (interleaved and nc > 1) else [0], 0, 0, 1, 0)
        if not (interleaved and nc > 1):
            for c in range(1, nc):
                scan([c], 0, 0, 1, 0)
        for c in range(nc):
            scan([c], 1, 5, 1, 0)
    if dnl:
        out += _segment(0xDC, struct.pack(">H", height))
    out += b"\xff\xd9"
    with open(path, "wb") as f:
        f.write(bytes(out))
