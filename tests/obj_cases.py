"""
tests/obj_cases.py — TEST INFRASTRUCTURE: OBJ / MTL inputs for the scene-loader parity tests (tests/test_host.py) and the
fixture script tests/golden/make_obj_fixtures.py: a seeded generator of small scenes (mixed number formats, quads, convex and
concave polygons, negative indices, groups, tabs, CRLF ...) and a list of hand-written files, one per decision the OBJ format
leaves to the loader (number syntax, material-name tokenising, texture options, which faces are dropped, line ends ...).
"""
import numpy as np


def _fmt(rng, v):
    style = rng.integers(0, 6)
    if style == 0: return f"{v:.6f}"
    if style == 1: return f"{v:.9f}"
    if style == 2: return f"{v:.4e}"
    if style == 3: return repr(float(v))
    if style == 4: return f"{v:.3f}"
    return f"{v:.12g}"


def random_scene(seed, features=()):
    """-> {"s.obj": text, "s.mtl": text}.  features: quads, polys, neg, groups, smooth, tabs, crlf, vcolor, badmtl, tex."""
    rng = np.random.default_rng(seed)
    nv, nn, nt = int(rng.integers(8, 40)), int(rng.integers(3, 10)), int(rng.integers(3, 10))
    V = rng.uniform(-2, 2, (nv, 3)); N = rng.normal(size=(nn, 3)); N /= np.linalg.norm(N, axis=1)[:, None]; T = rng.uniform(0, 1, (nt, 2))
    nm = int(rng.integers(1, 4))
    mtl = []
    for m in range(nm):
        mtl.append(f"newmtl mat{m}")
        mtl.append("Kd " + " ".join(_fmt(rng, x) for x in rng.uniform(0, 1, 3)))
        mtl.append("Ks " + " ".join(_fmt(rng, x) for x in rng.uniform(0, 1, 3)))
        if rng.random() < 0.5: mtl.append("Ke " + " ".join(_fmt(rng, x) for x in rng.uniform(0, 20, 3)))
        if rng.random() < 0.5: mtl.append("Ni " + _fmt(rng, rng.uniform(1, 3)))
        if rng.random() < 0.5: mtl.append("Pr " + _fmt(rng, rng.uniform(0, 1)))
        if rng.random() < 0.5: mtl.append("Pm " + _fmt(rng, rng.uniform(0, 1)))
        if rng.random() < 0.3: mtl.append("Tf " + " ".join(_fmt(rng, x) for x in rng.uniform(0, 1, 3)))
        if rng.random() < 0.3: mtl += ["Ns 10", "d 0.5", "illum 2"]
        if "tex" in features:          # image textures on random slots (write_case provides the three files)
            for key in ("map_Kd", "map_Ks", "map_Pr", "map_Pm", "map_Ke", "map_d"):
                if rng.random() < 0.35: mtl.append("%s %s" % (key, ("a.png", "b c.png", "d.tga")[int(rng.integers(0, 3))]))
    lines = ["mtllib s.mtl"]
    for v in V:
        l = "v " + " ".join(_fmt(rng, x) for x in v)
        if "vcolor" in features and rng.random() < 0.3: l += " 0.5 0.25 1"
        lines.append(l)
    for n in N: lines.append("vn " + " ".join(_fmt(rng, x) for x in n))
    for t in T:
        l = "vt " + " ".join(_fmt(rng, x) for x in t)
        if rng.random() < 0.2: l += " 0"
        lines.append(l)
    for i in range(int(rng.integers(10, 40))):
        r = rng.random()
        if r < 0.15: lines.append(f"usemtl mat{int(rng.integers(0, nm + (1 if 'badmtl' in features else 0)))}")
        if r > 0.9 and "groups" in features: lines.append(("g grp%d" % i) if rng.random() < 0.5 else ("o obj%d" % i))
        if r > 0.85 and "smooth" in features: lines.append("s %d" % int(rng.integers(0, 3)))
        k = 3
        if "quads" in features and rng.random() < 0.4: k = 4
        if "polys" in features and rng.random() < 0.3: k = int(rng.integers(5, 9))
        if k <= 4:
            idx = rng.choice(nv, k, replace=False)
        else:       # points around a centre in a random plane, radii equal (convex) or random (concave)
            c = rng.uniform(-1, 1, 3); a = rng.normal(size=3); a /= np.linalg.norm(a); b = np.cross(a, rng.normal(size=3)); b /= np.linalg.norm(b)
            ang = np.sort(rng.uniform(0, 2 * np.pi, k)); rad = rng.uniform(0.3, 1.0, k) if rng.random() < 0.5 else np.ones(k)
            for p in c + (np.cos(ang) * rad)[:, None] * a + (np.sin(ang) * rad)[:, None] * b:
                lines.append("v " + " ".join(_fmt(rng, x) for x in p))
            idx = np.arange(nv, nv + k); nv += k
        neg = "neg" in features and rng.random() < 0.3
        with_t = rng.random() < 0.7
        toks = []
        for j in idx:
            vi, ni, ti = int(j) + 1, int(rng.integers(1, nn + 1)), int(rng.integers(1, nt + 1))
            if neg: vi, ni, ti = vi - nv - 1, ni - nn - 1, ti - nt - 1
            toks.append(f"{vi}/{ti}/{ni}" if with_t else f"{vi}//{ni}")
        sep = "\t" if ("tabs" in features and rng.random() < 0.3) else " "
        lines.append("f" + sep + sep.join(toks) + ("  " if rng.random() < 0.2 else ""))
    eol = "\r\n" if "crlf" in features else "\n"
    return {"s.obj": eol.join(lines) + eol, "s.mtl": "\n".join(mtl) + "\n"}


RANDOM_CASES = [(s, f) for s in range(6) for f in ((), ("quads",), ("polys",), ("quads", "polys", "neg", "groups", "tabs"), ("tabs", "crlf", "smooth", "groups", "vcolor", "neg", "badmtl"))]

_V = "v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv 0.5 0.5 1\nvn 0 0 1\nvn 0 1 0\nvt 0 0\nvt 1 1\n"
_T = "f 1/1/1 2/2/1 3/1/2\nf 1//1 3//1 4//1\nf 1/2/1 2/2/2 5/1/1\n"
_M = "newmtl a\nKd 0.8 0.1 0.1\nnewmtl b\nKd 0.1 0.8 0.1\nKe 5 5 5\n"
_FLAT = "mtllib s.mtl\nvn 0 0 1\n"

_BOUNDARY = [("-2.834645152091980", "2.370645403862", "0.080899972468614578"), ("-2.797842383384705", "1.772440969944", "-2.4700402021408081"),
             ("-1.163677752017975", "1.424824655056", "2.4110091924667358"), ("-0.323213592171669", "-0.333979383111", "-1.6722009778022766"),
             ("1.766747772693634", "2.314248919487", "2.0358320474624634"), ("-2.206133961677551", "-1.941426217556", "0.86738529801368713")]

# name -> files ("a.png", "b c.png", "d.tga" are written by the user of this table: 4x4 images)
QUIRK_CASES = {
    "plain": {"s.obj": "mtllib s.mtl\n" + _V + "usemtl a\n" + _T + "usemtl b\n" + _T, "s.mtl": _M},
    "number_syntax": {"s.obj": "mtllib s.mtl\nv .5 5. +1\nv -.25 1e0 1E-2\nv 1.5e+1 abc 2\nv 1.5x 3 0x10\nv 1e 2 3\nv - + .\nv 1 2\nvn 0 0 1\nvt 0.5\nvt 1 1\n"
                               "f 1/1/1 2/2/1 3/1/1\nf 4/1/1 5/2/1 6/1/1\nf 7/1/1 1/2/1 2/1/1\n", "s.mtl": _M},
    "number_long_fractions": {"s.obj": "mtllib s.mtl\nv 0.1234567890123 -2.000000119 3.14159265358979\nv 1.00000011920928955 0.333333343267 1e-3\nv 123456.7890625 -0.000001 7.0e2\nvn 0 0 1\nusemtl a\nf 1//1 2//1 3//1\n", "s.mtl": _M},
    # decimal strings next to a float rounding boundary: a correctly rounding parser (strtod) yields the neighbouring float for each
    "number_rounding_boundaries": {"s.obj": "mtllib s.mtl\n" + "".join("v %s %s %s\n" % t for t in _BOUNDARY) + "vn 0 0 1\nusemtl a\n"
                                            + "".join("f %d//1 %d//1 %d//1\n" % (i + 1, (i + 1) % 6 + 1, (i + 2) % 6 + 1) for i in range(6)), "s.mtl": _M},
    "material_name_with_blanks": {"s.obj": "mtllib s.mtl\n" + _V + "usemtl my mat\n" + _T + "usemtl my\n" + _T + "usemtl b\n" + _T,
                                  "s.mtl": "newmtl my mat\nKd 1 0 0\nnewmtl my\nKd 0 1 0\nnewmtl b\nKd 0 0 1\n"},
    "material_duplicate_names": {"s.obj": "mtllib s.mtl\n" + _V + "usemtl a\n" + _T + "usemtl b\n" + _T, "s.mtl": "newmtl a\nKd 1 0 0\nnewmtl a\nKd 0 1 0\nnewmtl b\nKd 0 0 1\n"},
    "trailing_blanks": {"s.obj": "mtllib s.mtl  \n" + _V + "usemtl a  \n" + _T + "usemtl\tb\n" + _T, "s.mtl": "newmtl a  \t\nKd 1 0 0 \n  newmtl b\n\tKd 0 0 1\n"},
    "mtl_without_newmtl": {"s.obj": "mtllib s.mtl\n" + _V + "usemtl a\n" + _T, "s.mtl": "Kd 0.3 0.4 0.5\nKe 1 2 3\n"},
    "mtl_empty": {"s.obj": "mtllib s.mtl\n" + _V + _T, "s.mtl": ""},
    "mtl_keys_before_newmtl": {"s.obj": "mtllib s.mtl\n" + _V + "usemtl a\n" + _T, "s.mtl": "Kd 0.3 0.4 0.5\nnewmtl a\nKs 1 1 1\n"},
    "map_kd_default_grey": {"s.obj": "mtllib s.mtl\n" + _V + "usemtl a\n" + _T + "usemtl b\n" + _T + "usemtl c\n" + _T,
                            "s.mtl": "newmtl a\nmap_Kd a.png\nnewmtl b\nKd 0.2 0.2 0.2\nmap_Kd a.png\nnewmtl c\nmap_Kd d.tga\n"},
    "texture_options": {"s.obj": "mtllib s.mtl\n" + _V + "usemtl a\n" + _T + "usemtl b\n" + _T + "usemtl c\n" + _T,
                        "s.mtl": "newmtl a\nKd 1 1 1\nmap_Kd -s 1 1 1 -o 0 0 0 -blendu on -mm 0 1 a.png\nmap_Ks -clamp on -bm 2 b c.png\nnewmtl b\n"
                                 "map_Ke -texres 512 -imfchan r -type sphere -colorspace sRGB d.tga\nmap_Pr -boost 1 -t 0 0 0 a.png\nnewmtl c\nmap_d -blendv off a.png\nmap_Pm b c.png\n"},
    # two maps that are arguments of one call in the reference (scene.cpp:172-184): the later argument's file is loaded first
    "texture_load_order": {"s.obj": "mtllib s.mtl\n" + _V + "usemtl a\n" + _T + "usemtl b\n" + _T,
                           "s.mtl": "newmtl a\nKd 1 1 1\nmap_Pr a.png\nmap_Pm d.tga\nnewmtl b\nKd 1 1 1\nmap_Ke b c.png\nmap_d a.png\nmap_Kd d.tga\n"},
    "texture_name_with_blanks": {"s.obj": "mtllib s.mtl\n" + _V + "usemtl a\n" + _T, "s.mtl": "newmtl a\nmap_Kd b c.png   \n"},
    "two_mtllib_lines": {"s.obj": "mtllib s.mtl\nmtllib t.mtl\n" + _V + "usemtl a\n" + _T + "usemtl z\n" + _T, "s.mtl": _M, "t.mtl": "newmtl z\nKd 0.5 0.5 0\nnewmtl a\nKd 0 0 0.5\n"},
    "mtllib_several_names": {"s.obj": "mtllib missing.mtl t.mtl s.mtl\n" + _V + "usemtl a\n" + _T + "usemtl z\n" + _T, "s.mtl": _M, "t.mtl": "newmtl z\nKd 0.5 0.5 0\n"},
    "usemtl_before_mtllib": {"s.obj": _V + "usemtl a\n" + _T + "mtllib s.mtl\nusemtl a\n" + _T, "s.mtl": _M},
    "line_ends_cr": {"s.obj": ("mtllib s.mtl\n" + _V + "usemtl b\n" + _T).replace("\n", "\r"), "s.mtl": _M.replace("\n", "\r")},
    "line_ends_crlf": {"s.obj": ("mtllib s.mtl\n" + _V + "usemtl b\n" + _T).replace("\n", "\r\n"), "s.mtl": _M.replace("\n", "\r\n")},
    "no_final_newline": {"s.obj": "mtllib s.mtl\n" + _V + "usemtl b\n" + _T.rstrip("\n"), "s.mtl": _M.rstrip("\n")},
    "comments_blank_lines": {"s.obj": "# c\n\n  \nmtllib s.mtl\n" + _V + "  # indented comment\nusemtl b\n s 1\n" + _T + "\n\n", "s.mtl": "# x\n\n" + _M},
    "quad_with_missing_vertex": {"s.obj": "mtllib s.mtl\n" + _V + "usemtl a\nf 1//1 2//1 3//1 9//1\n" + _T, "s.mtl": _M},
    "quad_diagonals": {"s.obj": _FLAT + "v 0 0 0\nv 2 0 0\nv 2 1 0\nv 0 1 0\nv 0 0 1\nv 1 0 1\nv 3 1 1\nv 0.5 1 1\nv 0 0 2\nv 1 0 2\nv 1 1 2\nv 0 1 2\nusemtl a\nf 1//1 2//1 3//1 4//1\nf 5//1 6//1 7//1 8//1\nf 9//1 10//1 11//1 12//1\n", "s.mtl": _M},
    "forward_reference_open_group": {"s.obj": _FLAT + "v 0 0 0\nv 1 0 0\nv 1 1 0\nusemtl a\nf 1//1 2//1 3//1 4//1\nv 0 1 0\nf 1//1 2//1 3//1\n", "s.mtl": _M},
    "forward_reference_closed_group": {"s.obj": _FLAT + "v 0 0 0\nv 1 0 0\nv 1 1 0\nusemtl a\nf 1//1 2//1 3//1 4//1\ng next\nv 0 1 0\nf 1//1 2//1 3//1\n", "s.mtl": _M},
    "polygon_collinear": {"s.obj": _FLAT + "v 0 0 0\nv 1 0 0\nv 2 0 0\nv 3 0 0\nv 4 0 0\nv 2 1 0\nusemtl a\nf 1//1 2//1 3//1 4//1 5//1 6//1\n", "s.mtl": _M},
    "polygon_degenerate": {"s.obj": _FLAT + "v 0 0 0\nv 0 0 0\nv 0 0 0\nv 0 0 0\nv 0 0 0\nv 1 1 1\nv 1 0 0\nusemtl a\nf 1//1 2//1 3//1 4//1 5//1\nf 1//1 6//1 7//1\n", "s.mtl": _M},
    "polygon_self_crossing": {"s.obj": _FLAT + "v 0 0 0\nv 1 1 0\nv 1 0 0\nv 0 1 0\nv 0.5 2 0\nusemtl a\nf 1//1 2//1 3//1 4//1 5//1\n", "s.mtl": _M},
    "polygon_concave": {"s.obj": _FLAT + "v 0 0 0\nv 2 0 0\nv 2 2 0\nv 1 0.5 0\nv 0 2 0\nv 0 0 1\nv 0 2 1\nv 0 2 3\nv 0 1 1.5\nv 0 0 3\nv -1 1 2\nusemtl a\nf 1//1 2//1 3//1 4//1 5//1\nf 6//1 7//1 8//1 9//1 10//1 11//1\n", "s.mtl": _M},
    "polygon_repeated_corner": {"s.obj": _FLAT + "v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv -1 0.5 0\nusemtl a\nf 1//1 2//1 3//1 4//1 5//1\nf 1//1 2//1 3//1 4//1 5//1 3//1\n", "s.mtl": _M},
    "two_corner_face": {"s.obj": "mtllib s.mtl\n" + _V + "usemtl a\nf 1//1 2//1\n" + _T, "s.mtl": _M},
    "objects_and_groups": {"s.obj": "mtllib s.mtl\n" + _V + "o first\nusemtl a\n" + _T + "g g1 g2\n" + _T + "usemtl b\no second\n" + _T + "g\n" + _T, "s.mtl": _M},
    "vertex_colours_and_vt_w": {"s.obj": "mtllib s.mtl\nv 0 0 0 1 0 0\nv 1 0 0 0 1 0\nv 1 1 0 0 0 1\nvn 0 0 1\nvt 0.25 0.75 0\nusemtl a\nf 1/1/1 2/1/1 3/1/1\n", "s.mtl": _M},
    "negative_indices": {"s.obj": "mtllib s.mtl\n" + _V + "usemtl b\nf -5/-2/-2 -4/-1/-2 -3/-2/-1\nv 2 2 2\nf -1//-1 -6//-2 -5//-1\n", "s.mtl": _M},
}


def write_case(directory, files, rng=None):
    """Writes one case's files (+ the three small texture images every case may name) into `directory`."""
    import os
    from PIL import Image
    rng = rng or np.random.default_rng(1)
    os.makedirs(directory, exist_ok=True)
    for name, text in files.items():
        with open(os.path.join(directory, name), "w", newline="") as f:
            f.write(text)
    if any(("a.png" in t or "b c.png" in t or "d.tga" in t) for t in files.values()):
        Image.fromarray(rng.integers(0, 256, (4, 4, 3)).astype(np.uint8)).save(os.path.join(directory, "a.png"))
        Image.fromarray(rng.integers(0, 256, (5, 3, 3)).astype(np.uint8)).save(os.path.join(directory, "b c.png"))
        Image.fromarray(rng.integers(0, 256, (4, 4, 4)).astype(np.uint8)).save(os.path.join(directory, "d.tga"))
    return os.path.join(directory, "s.obj")


SOUP_CASES = [(mode, seed, n) for mode in ("uniform", "clusters", "duplicates", "grid", "line", "signed_zeros") for seed, n in ((1, 3000), (2, 20000))]


def triangle_soup(mode, seed, n):
    """OBJ text of n separate triangles whose centroids are spread uniformly / in tight clusters / on few repeated points (leaves
    of many identical centroids) / on a regular grid (ties in every SAH bucket) / along one axis / on a few values including
    both zeros: inputs for the BVH builder."""
    rng = np.random.default_rng(seed)
    if mode == "uniform":
        c = rng.uniform(-5, 5, (n, 3)); e = rng.normal(0, 0.2, (n, 3, 3))
    elif mode == "clusters":
        cc = rng.uniform(-5, 5, (20, 3)); c = cc[rng.integers(0, 20, n)] + rng.normal(0, 0.05, (n, 3)); e = rng.normal(0, 0.02, (n, 3, 3))
    elif mode == "duplicates":
        base = rng.uniform(-5, 5, (n // 50 + 1, 3)); c = base[rng.integers(0, len(base), n)]; e = np.tile(rng.normal(0, 0.2, (1, 3, 3)), (n, 1, 1))
    elif mode == "grid":
        g = int(np.ceil(n ** (1 / 3)))
        c = np.stack(np.unravel_index(np.arange(n), (g, g, g)), -1).astype(float)
        e = np.tile(np.array([[0, 0, 0], [0.5, 0, 0], [0, 0.5, 0]])[None], (n, 1, 1)) - 0.2
    elif mode == "line":
        c = np.stack([np.linspace(0, 100, n), np.zeros(n), np.zeros(n)], -1); e = rng.normal(0, 0.2, (n, 3, 3))
    else:       # coordinates from a small set that holds both zeros: the sign of a zero bound depends on the order of the unions
        values = np.array([-0.0, 0.0, -0.0, 0.0, 1.0, -1.0, 2.0, -2.0, 0.5])
        c = np.zeros((n, 3)); e = values[rng.integers(0, len(values), (n, 3, 3))]
    p = (c[:, None, :] + e).reshape(-1, 3)
    lines = ["mtllib s.mtl", "vn 0 0 1", "usemtl a"] + ["v %.6f %.6f %.6f" % tuple(v) for v in p]
    lines += ["f %d//1 %d//1 %d//1" % (3 * i + 1, 3 * i + 2, 3 * i + 3) for i in range(n)]
    return {"s.obj": "\n".join(lines) + "\n", "s.mtl": "newmtl a\nKd 0.5 0.5 0.5\n"}


def tree_digest(scene):
    """sha256 over the node array (bounds bits, offset, primitive count / axis) and the leaf-ordered vertex positions."""
    import hashlib
    h = hashlib.sha256()
    nodes, tris = scene["nodes"], scene["triangles"]
    for k in ("bounds_min", "bounds_max"):
        h.update(np.ascontiguousarray(nodes[k][:, :3]).view(np.uint32).tobytes())
    for k in ("offset", "num_primitives_axis"):
        h.update(np.ascontiguousarray(nodes[k]).tobytes())
    for v in ("v1", "v2", "v3"):
        h.update(np.ascontiguousarray(tris[v]["position"][:, :3]).view(np.uint32).tobytes())
    return "%d:%s" % (len(nodes), h.hexdigest())
