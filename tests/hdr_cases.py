"""
tests/hdr_cases.py — TEST INFRASTRUCTURE: Radiance .hdr files for the environment-map reader parity tests: flat scanlines,
new-style (per-component) run-length coding, old-style (1,1,1,n) repeat pixels including two markers in a row, widths on both
sides of the 8..32767 window of the new coding, scanlines that merely start like a coded one, a longer header, a file cut short.
cases(seed) -> {name: file bytes}.
"""
import numpy as np


def _header(extra=b""):
    return b"#?RADIANCE\n" + extra + b"FORMAT=32-bit_rle_rgbe\n\n"


def _file(data, w, h, extra=b""):
    return _header(extra) + ("-Y %d +X %d\n" % (h, w)).encode() + data


def _new_rle(a):
    out = bytearray()
    h, w, _ = a.shape
    for y in range(h):
        out += bytes([2, 2, w >> 8, w & 255])
        for c in range(4):
            row = a[y, :, c]; x = 0
            while x < w:
                r = 1
                while x + r < w and r < 127 and row[x + r] == row[x]: r += 1
                if r >= 3:
                    out += bytes([128 + r, row[x]]); x += r
                else:
                    n = 1
                    while x + n < w and n < 128 and not (x + n + 2 < w and row[x + n] == row[x + n + 1] == row[x + n + 2]): n += 1
                    out += bytes([n]) + bytes(row[x:x + n]); x += n
    return bytes(out)


def _old_rle(a):
    out = bytearray()
    h, w, _ = a.shape
    for y in range(h):
        x = 0
        while x < w:
            out += bytes(a[y, x]); r = 1
            while x + r < w and r < 250 and (a[y, x + r] == a[y, x]).all(): r += 1
            if r > 2 and not (a[y, x, :3] == 1).all():
                out += bytes([1, 1, 1, r - 1]); x += r
            else:
                x += 1
    return bytes(out)


def cases(seed=4):
    rng = np.random.default_rng(seed)

    def rgbe(h, w):
        a = rng.integers(0, 256, (h, w, 4)).astype(np.uint8); a[..., 3] = rng.integers(100, 150, (h, w))
        return a

    def runs(h, w):
        a = rgbe(h, w)
        for y in range(h):
            x = 0
            while x < w:
                n = int(rng.integers(1, 12)); a[y, x:x + n] = a[y, x]; x += n
        return a
    out = {}
    for w, h in ((1, 3), (7, 4), (8, 3), (9, 5), (33, 6), (128, 4)):
        a = runs(h, w)
        out["flat_%dx%d" % (w, h)] = _file(a.tobytes(), w, h)
        if 8 <= w <= 32767:
            out["new_rle_%dx%d" % (w, h)] = _file(_new_rle(a), w, h)
        out["old_rle_%dx%d" % (w, h)] = _file(_old_rle(a), w, h)
    out["flat_33000_wide"] = _file(runs(2, 33000).tobytes(), 33000, 2)
    a = runs(6, 40)
    out["flat_cut_short"] = _file(a.tobytes()[:500], 40, 6)
    out["long_header"] = _file(_new_rle(a), 40, 6, extra=b"# comment line\nEXPOSURE=2.0\nSOFTWARE=x\n")
    # two repeat markers in a row: the format shifts the second count left by 8; the reference keeps counts in 8 bits
    a = np.zeros((2, 600, 4), np.uint8); a[...] = (90, 80, 70, 128); a[:, 300:] = (5, 6, 7, 130)
    d = bytearray()
    for y in range(2):
        d += bytes(a[y, 0]) + bytes([1, 1, 1, 43]) + bytes([1, 1, 1, 1]) + bytes(a[y, 300]) + bytes([1, 1, 1, 43]) + bytes([1, 1, 1, 1])
        d += rng.integers(2, 256, (600 - 88, 4)).astype(np.uint8).tobytes()
    out["old_rle_two_markers"] = _file(bytes(d), 600, 2)
    a = rgbe(2, 16); a[:, 0, 0] = 2; a[:, 0, 1] = 7
    out["flat_starting_with_2"] = _file(a.tobytes(), 16, 2)
    a = rgbe(2, 16); a[:, 0, 0] = 2; a[:, 0, 1] = 2; a[:, 0, 2] = 200
    out["flat_starting_2_2_200"] = _file(a.tobytes(), 16, 2)
    return out


def digest(env, width, height):
    import hashlib
    return "%dx%d:%s" % (width, height, hashlib.sha256(np.ascontiguousarray(env, dtype=np.float32).view(np.uint32).tobytes()).hexdigest())
