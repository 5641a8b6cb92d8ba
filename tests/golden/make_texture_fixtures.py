#!/usr/bin/env python3
"""
tests/golden/make_texture_fixtures.py — image-texture fixtures decoded by the REFERENCE's own loader.

    python tests/golden/make_texture_fixtures.py          (where /root/reference and oracle/_ref/libref.so exist)

Writes tests/golden/textures/: small JPEG files (baseline / progressive, 4:4:4 / 4:2:2 / 4:2:0, greyscale, CMYK, restart
intervals, optimised Huffman tables, 1-pixel-wide; and from tests/jpeg_writer.py the layouts Pillow does not write: 4:4:0,
4:1:1, 4:1:0, one scan per component, 16-bit quantisation tables, RGB component ids, Adobe colour transforms, YCCK, DNL,
successive-approximation progressive scripts) and Adam7-interlaced PNG files (every colour type, bit depths 1..16),
plus expected.npz: for every file the texel words the reference's Scene::LoadTexture -> LoadSTB (stb_image v2.27,
scene.cpp:276-323, image_loader.cpp:30-63) produced for it.  tests/test_host.py::test_texture_files_decode_like_the_reference
checks host/image_loader.cpp + host/jpeg_decoder.cpp against them wherever the tests run.

The ordinary JPEG files come from Pillow's encoder, the interlaced PNG files from the small writer below (Pillow cannot write
Adam7).
"""
import os
import struct
import sys
import zlib

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)

OUT = os.path.join(REPO, "tests", "golden", "textures")
ADAM7 = [(0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)]


def _chunk(tag, data):
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def _paeth(a, b, c):
    p = a + b - c
    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
    return a if pa <= pb and pa <= pc else (b if pb <= pc else c)


def _filter_rows(rows, bpp):
    """rows: list of bytes objects of one (sub-)image -> filtered stream, cycling through the five filter types."""
    out = bytearray()
    prev = None
    for y, row in enumerate(rows):
        f = y % 5
        out.append(f)
        for x, v in enumerate(row):
            a = row[x - bpp] if x >= bpp else 0
            b = prev[x] if prev is not None else 0
            c = prev[x - bpp] if prev is not None and x >= bpp else 0
            pred = (0, a, b, (a + b) >> 1, _paeth(a, b, c))[f]
            out.append((v - pred) & 0xFF)
        prev = row
    return bytes(out)


def _pack_row(samples, depth):
    """1-D array of sample values (< 2**depth) -> the row's bytes (MSB first; 16 bit big-endian)."""
    if depth == 8:
        return bytes(samples.astype(np.uint8))
    if depth == 16:
        return samples.astype(">u2").tobytes()
    per = 8 // depth
    pad = (-len(samples)) % per
    s = np.concatenate([samples, np.zeros(pad, dtype=samples.dtype)]).reshape(-1, per).astype(np.uint32)
    shifts = np.arange(per - 1, -1, -1, dtype=np.uint32) * depth
    return bytes((s << shifts).sum(axis=1).astype(np.uint8))


def write_png(path, samples, ctype, depth, interlace=True, palette=None, trns=None):
    """samples: (h, w, channels) integer array of sample values.  Writes a valid PNG, Adam7-interlaced on request."""
    h, w, ch = samples.shape
    bpp = max(1, ch * depth // 8)
    raw = bytearray()
    for (x0, y0, dx, dy) in (ADAM7 if interlace else [(0, 0, 1, 1)]):
        sub = samples[y0::dy, x0::dx]
        if sub.shape[0] == 0 or sub.shape[1] == 0:
            continue
        raw += _filter_rows([_pack_row(r.reshape(-1), depth) for r in sub], bpp)
    data = b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 1 if interlace else 0))
    if palette is not None:
        data += _chunk(b"PLTE", bytes(palette))
    if trns is not None:
        data += _chunk(b"tRNS", bytes(trns))
    z = zlib.compress(bytes(raw), 6)         # two IDAT chunks: the stream may be split anywhere
    data += _chunk(b"IDAT", z[: len(z) // 2]) + _chunk(b"IDAT", z[len(z) // 2:]) + _chunk(b"IEND", b"")
    with open(path, "wb") as f:
        f.write(data)


def smooth(rng, h, w, c):
    yy, xx = np.mgrid[0:h, 0:w]
    a = np.stack([127 + 120 * np.sin(xx / (7 + 3 * k) + yy / (5 + k)) for k in range(c)], -1)
    a = a + rng.normal(0, 12, a.shape)
    return np.clip(a, 0, 255).astype(np.uint8)


def make_files(out_dir, rng=None):
    """Writes the test images; returns their names (sorted, deterministic for a seed)."""
    from PIL import Image
    rng = rng or np.random.default_rng(2024)
    os.makedirs(out_dir, exist_ok=True)

    def jpg(name, arr, mode=None, **kw):
        Image.fromarray(arr, mode).save(os.path.join(out_dir, name), **kw)
    jpg("base_444.jpg", smooth(rng, 37, 53, 3), quality=90, subsampling=0)
    jpg("base_422.jpg", smooth(rng, 40, 61, 3), quality=75, subsampling=1)
    jpg("base_420.jpg", smooth(rng, 45, 67, 3), quality=60, subsampling=2)
    jpg("prog_420.jpg", smooth(rng, 64, 64, 3), quality=85, subsampling=2, progressive=True)
    jpg("prog_444.jpg", smooth(rng, 33, 17, 3), quality=95, subsampling=0, progressive=True)
    jpg("grey.jpg", smooth(rng, 29, 31, 1)[..., 0], "L", quality=80)
    jpg("prog_grey.jpg", smooth(rng, 50, 9, 1)[..., 0], "L", quality=80, progressive=True)
    jpg("noise_q100.jpg", rng.integers(0, 256, (24, 24, 3)).astype(np.uint8), quality=100, subsampling=2)
    jpg("noise_q5.jpg", rng.integers(0, 256, (40, 24, 3)).astype(np.uint8), quality=5, subsampling=1)
    jpg("one_pixel.jpg", smooth(rng, 1, 1, 3), quality=90)
    jpg("one_wide.jpg", smooth(rng, 17, 1, 3), quality=90, subsampling=2)
    jpg("optimised.jpg", smooth(rng, 48, 48, 3), quality=50, optimize=True)
    jpg("cmyk.jpg", smooth(rng, 20, 20, 4), "CMYK", quality=90)
    jpg("restart_420.jpg", smooth(rng, 40, 56, 3), quality=80, subsampling=2, restart_marker_blocks=2)
    jpg("restart_prog.jpg", smooth(rng, 35, 50, 3), quality=70, subsampling=1, progressive=True, restart_marker_rows=1)
    jpg("restart_grey.jpg", smooth(rng, 30, 30, 1)[..., 0], "L", quality=70, restart_marker_blocks=3)

    # layouts Pillow's encoder does not write (tests/jpeg_writer.py)
    from tests import jpeg_writer
    odd = [
        ("w_440", dict(sampling=[(1, 2), (1, 1), (1, 1)])),
        ("w_411", dict(sampling=[(4, 1), (1, 1), (1, 1)])),
        ("w_410", dict(sampling=[(4, 2), (1, 1), (1, 1)])),
        ("w_luma_low", dict(sampling=[(1, 1), (2, 2), (2, 1)])),
        ("w_141", dict(sampling=[(1, 4), (1, 1), (1, 2)])),
        ("w_scan_per_component", dict(sampling=[(2, 2), (1, 1), (1, 1)], interleaved=False)),
        ("w_scan_per_component_rst", dict(sampling=[(2, 1), (1, 1), (1, 1)], interleaved=False, restart_interval=3)),
        ("w_rst1_fill", dict(sampling=[(2, 2), (1, 1), (1, 1)], restart_interval=1, fill_bytes=True)),
        ("w_quant16", dict(sampling=[(1, 1), (1, 1), (1, 1)], sixteen_bit_quant=True)),
        ("w_rgb_ids", dict(sampling=[(1, 1), (1, 1), (1, 1)], ids=[ord("R"), ord("G"), ord("B")])),
        ("w_adobe_rgb", dict(sampling=[(1, 1), (1, 1), (1, 1)], adobe_transform=0, jfif=False)),
        ("w_adobe_rgb_jfif", dict(sampling=[(1, 1), (1, 1), (1, 1)], adobe_transform=0, jfif=True)),
        ("w_cmyk", dict(sampling=[(1, 1)] * 4, adobe_transform=0)),
        ("w_ycck", dict(sampling=[(2, 2), (1, 1), (1, 1), (2, 2)], adobe_transform=2)),
        ("w_four_plain", dict(sampling=[(1, 1)] * 4)),
        ("w_grey_2x2", dict(sampling=[(2, 2)])),
        ("w_dnl", dict(sampling=[(2, 1), (1, 1), (1, 1)], dnl=True)),
        ("w_prog", dict(sampling=[(2, 2), (1, 1), (1, 1)], progressive=True)),
        ("w_prog_scan_per_component_rst", dict(sampling=[(2, 1), (1, 2), (1, 1)], progressive=True, interleaved=False, restart_interval=2)),
        ("w_prog_grey_rst", dict(sampling=[(1, 1)], progressive=True, restart_interval=5)),
        ("w_prog_fine", dict(sampling=[(1, 1), (1, 1), (1, 1)], progressive=True, quant_scale=1)),
    ]
    for k, (name, kw) in enumerate(odd):
        w, h = ((37, 29), (16, 16), (5, 70))[k % 3]
        jpeg_writer.write_jpeg(os.path.join(out_dir, name + ".jpg"), w, h, rng=rng, **kw)

    def r(shape, hi):
        return rng.integers(0, hi, size=shape)
    write_png(os.path.join(out_dir, "a7_rgb8.png"), r((19, 23, 3), 256), 2, 8)
    write_png(os.path.join(out_dir, "a7_rgba8.png"), r((9, 10, 4), 256), 6, 8)
    write_png(os.path.join(out_dir, "a7_rgb16.png"), r((7, 11, 3), 65536), 2, 16)
    write_png(os.path.join(out_dir, "a7_grey1.png"), r((13, 21, 1), 2), 0, 1)
    write_png(os.path.join(out_dir, "a7_grey2.png"), r((10, 5, 1), 4), 0, 2)
    write_png(os.path.join(out_dir, "a7_grey4.png"), r((6, 15, 1), 16), 0, 4)
    write_png(os.path.join(out_dir, "a7_grey16.png"), r((5, 3, 1), 65536), 0, 16)
    write_png(os.path.join(out_dir, "a7_ga8.png"), r((12, 12, 2), 256), 4, 8)
    write_png(os.path.join(out_dir, "a7_pal4.png"), r((11, 14, 1), 16), 3, 4, palette=r(48, 256).astype(np.uint8))
    write_png(os.path.join(out_dir, "a7_pal8_trns.png"), r((8, 9, 1), 20), 3, 8, palette=r(60, 256).astype(np.uint8), trns=r(12, 256).astype(np.uint8))
    write_png(os.path.join(out_dir, "a7_tiny.png"), r((1, 1, 3), 256), 2, 8)          # six of the seven passes are empty
    write_png(os.path.join(out_dir, "a7_3x2.png"), r((2, 3, 3), 256), 2, 8)
    write_png(os.path.join(out_dir, "plain_rgb8.png"), r((9, 7, 3), 256), 2, 8, interlace=False)   # the writer itself, against Pillow
    return sorted(f for f in os.listdir(out_dir) if f.endswith((".jpg", ".png")))


def reference_texels(directory, files):
    """{file: (width, height, texel words)} from the reference's own Scene loader: one material per file."""
    from oracle.refbind import RefRenderer
    import tempfile
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for i in range(0, len(files), 200):                        # texture indices are 8 bit
            part = files[i: i + 200]
            with open(os.path.join(tmp, "t.mtl"), "w") as f:
                for k, name in enumerate(part):
                    f.write(f"newmtl m{k}\nKd 0.5 0.5 0.5\nmap_Kd {name}\n")       # looked up next to the OBJ (scene.cpp:161)
                    if not os.path.exists(os.path.join(tmp, name)):
                        os.symlink(os.path.join(directory, name), os.path.join(tmp, name))
            with open(os.path.join(tmp, "t.obj"), "w") as f:
                f.write("mtllib t.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nvt 0 0\n")
                for k in range(len(part)):
                    f.write(f"usemtl m{k}\nf 1/1/1 2/1/1 3/1/1\n")
            sc = RefRenderer().open_obj("/root/reference", os.path.join(tmp, "t.obj")).scene()
            assert len(sc["textures"]) == len(part)
            for k, name in enumerate(part):
                t = sc["textures"][k]
                n = int(t["width"]) * int(t["height"])
                out[name] = (int(t["width"]), int(t["height"]), sc["texels"][int(t["data_start"]): int(t["data_start"]) + n].copy())
    return out


def main():
    files = make_files(OUT)
    ref = reference_texels(OUT, files)
    arrays = {}
    for name in files:
        w, h, tex = ref[name]
        arrays[name + ":size"] = np.array([w, h], dtype=np.int32)
        arrays[name + ":texels"] = tex
        print(f"{name:20s} {w}x{h}")
    np.savez_compressed(os.path.join(OUT, "expected.npz"), **arrays)
    print("wrote", len(files), "files +", os.path.join(OUT, "expected.npz"))


if __name__ == "__main__":
    main()
