#!/usr/bin/env python3
"""
oracle/build_ref.py — TEST INFRASTRUCTURE.  Builds oracle/_ref/libref.so (and
libref_libm.so) from the reference's own sources where they lie under
/root/reference.  Outputs go ONLY to oracle/_ref/ (git-ignored; it travels to
the GPU box with the snapshot like any other built .so).  No reference source
is copied into the repository.

What is compiled, and how:

  host code, compiled IN PLACE (g++ on the reference files directly):
      src/bvh.cpp  src/mathlib/mathlib.cpp  src/loaders/hdr_loader.cpp
      src/loaders/image_loader.cpp  src/scene/scene.cpp  src/integrator/integrator.cpp
    with include-path stubs from oracle/ref_shim/stubs/ (an empty GL/gl.h and a
    portable utils/cl_exception.hpp — the reference's uses an MSVC-only
    std::exception constructor, cl_exception.hpp:113) and a prelude that pulls a
    few <cmath>/<cstring> names into scope that MSVC provides implicitly.

  device code (OpenCL C 1.2), compiled through oracle/ref_shim/clshim.h:
      src/kernels/cl/*.cl  src/kernels/common/*.h
    OpenCL C is not C++, so each file is passed through THREE purely syntactic
    rewrites on its way into oracle/_ref/gen/ (same relative path):
      1. vector literals   "(float3)(a, b, c)"  ->  "float3(a, b, c)"
         (in C++ the former is a cast of a comma expression);
      2. "#ifdef/#ifndef __cplusplus" -> "#ifdef/#ifndef RT_REF_HOST_VIEW"
         so that shared_structures.h presents its DEVICE view (OpenCL float3,
         Bounds3{pos[2]}) to the kernels even though g++ is the compiler.
      3. material.h:230 and :236 read
             bxdf = A * SampleSpecular(.., outgoing, pdf) * max(dot(OUT(outgoing), normal), 0.0f);
         where the right operand of the second `*` READS the direction the call in
         the left operand WRITES.  C leaves that order unspecified; OpenCL
         compilers (clang-based) evaluate left to right, g++ evaluates the right
         operand first and would read an uninitialised `outgoing`.  The statement
         is split into `bxdf = A * Sample(..); bxdf = bxdf * max(..);` — same
         association ((A*S)*m), same arithmetic, evaluation order pinned.
    Nothing else is touched: the arithmetic, control flow and data layout are
    the reference's.

The OpenCL runtime itself (context/queue/buffers) is replaced by
oracle/ref_harness.cpp, which drives the reference's Integrator::Integrate().

Usage: python oracle/build_ref.py [--reference /root/reference] [--force]
Exit code 0 and a line "ref: unavailable (...)" if the reference tree is absent.
"""
import argparse
import hashlib
import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
OUT = os.path.join(HERE, "_ref")

VEC_LITERAL = re.compile(r"\((float[234]|int[234]|uint[234])\)\s*\(")
# material.h:230,236 — `bxdf = A * SampleX(.., outgoing, pdf) * max(dot(OUT(outgoing), normal), 0.0f);`
EVAL_ORDER = re.compile(r"(bxdf = [^;\n]*?Sample(?:Specular|Diffuse)\([^;\n]*?\)) \* (max\(dot\(OUT\(outgoing\), normal\), 0\.0f\));")

HOST_SOURCES = [
    "src/bvh.cpp",
    "src/mathlib/mathlib.cpp",
    "src/loaders/hdr_loader.cpp",
    "src/loaders/image_loader.cpp",
    "src/scene/scene.cpp",
    "src/integrator/integrator.cpp",
]

# (kernel file, RT_KERNEL_ID, suffix, extra -D)
KERNEL_TUS = [
    ("raygeneration.cl", 1, "_std", []),
    ("trace_bvh.cl", 2, "_std", []),
    ("trace_bvh.cl", 2, "_shadow", ["SHADOW_RAYS"]),
    ("miss.cl", 3, "_std", []),
    ("miss.cl", 3, "_wf", ["ENABLE_WHITE_FURNACE"]),
    ("hit_surface.cl", 4, "_std", []),
    ("hit_surface.cl", 4, "_wf", ["ENABLE_WHITE_FURNACE"]),
    ("hit_surface.cl", 4, "_bn", ["BLUE_NOISE_SAMPLER"]),
    ("hit_surface.cl", 4, "_wfbn", ["ENABLE_WHITE_FURNACE", "BLUE_NOISE_SAMPLER"]),
    ("accumulate_direct_samples.cl", 5, "_std", []),
    ("clear_counter.cl", 6, "_std", []),
    ("increment_counter.cl", 7, "_std", []),
    ("reset_radiance.cl", 8, "_std", []),
    ("aov.cl", 9, "_std", []),
    ("denoiser.cl", 10, "_std", []),
    ("resolve_radiance.cl", 11, "_std", []),
    ("resolve_radiance.cl", 11, "_dn", ["ENABLE_DENOISER"]),
]


def rewrite_device_source(text: str) -> str:
    text = VEC_LITERAL.sub(lambda m: m.group(1) + "(", text)
    text = EVAL_ORDER.sub(lambda m: m.group(1) + "; bxdf = bxdf * " + m.group(2) + ";", text)
    text = text.replace("#ifdef __cplusplus", "#ifdef RT_REF_HOST_VIEW")
    text = text.replace("#ifndef __cplusplus", "#ifndef RT_REF_HOST_VIEW")
    return text


def run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise SystemExit(1)


def build(reference: str, force: bool = False) -> bool:
    if not os.path.isdir(os.path.join(reference, "src", "kernels", "cl")):
        print(f"ref: unavailable (no reference tree at {reference}); using prebuilt oracle/_ref if present")
        return False
    gen = os.path.join(OUT, "gen")
    obj = os.path.join(OUT, "obj")
    os.makedirs(gen, exist_ok=True)
    os.makedirs(obj, exist_ok=True)

    # stamp: rebuild only when inputs change
    h = hashlib.sha256()
    inputs = [os.path.join(HERE, "build_ref.py"), os.path.join(HERE, "ref_harness.cpp"),
              os.path.join(HERE, "ref_shim", "clshim.h"), os.path.join(HERE, "ref_shim", "kernel_tu.cpp"),
              os.path.join(HERE, "ref_shim", "host_prelude.h"), os.path.join(HERE, "ref_shim", "cl_stubs.c"), os.path.join(REPO, "include", "rt_math.h")]
    for d in ("src/kernels/cl", "src/kernels/common"):
        for f in sorted(os.listdir(os.path.join(reference, d))):
            inputs.append(os.path.join(reference, d, f))
    inputs += [os.path.join(reference, s) for s in HOST_SOURCES]
    for p in inputs:
        with open(p, "rb") as fh:
            h.update(fh.read())
    stamp = os.path.join(OUT, "stamp")
    libs = [os.path.join(OUT, "libref.so"), os.path.join(OUT, "libref_libm.so")]
    if not force and os.path.exists(stamp) and open(stamp).read() == h.hexdigest() and all(os.path.exists(l) for l in libs):
        print("ref: up to date")
        return True

    # 1. device sources -> gen/ (syntactic rewrites only)
    for d in ("src/kernels/cl", "src/kernels/common"):
        os.makedirs(os.path.join(gen, d), exist_ok=True)
        for f in sorted(os.listdir(os.path.join(reference, d))):
            with open(os.path.join(reference, d, f), "r", encoding="utf-8", errors="replace") as fh:
                text = fh.read()
            with open(os.path.join(gen, d, f), "w") as fh:
                fh.write(rewrite_device_source(text))

    common = ["g++", "-std=c++17", "-O2", "-fPIC", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-w"]
    shim = os.path.join(HERE, "ref_shim")
    jobs = []

    # 2. host sources, in place
    host_flags = ["-include", os.path.join(shim, "host_prelude.h"), "-DCL_TARGET_OPENCL_VERSION=120", "-DGLEW_NO_GLU",
                  "-I" + os.path.join(shim, "stubs"), "-I" + os.path.join(reference, "src"),
                  "-I" + os.path.join(reference, "3rdparty/OCL_SDK_Light/include"),
                  "-I" + os.path.join(reference, "3rdparty/glew-2.1.0/include"),
                  "-I" + os.path.join(reference, "3rdparty/tinyobjloader"),
                  "-I" + os.path.join(reference, "3rdparty/stb")]
    host_objs = []
    for s in HOST_SOURCES:
        o = os.path.join(obj, "host_" + os.path.basename(s).replace(".cpp", ".o"))
        host_objs.append(o)
        jobs.append(common + host_flags + ["-c", os.path.join(reference, s), "-o", o])

    # 3. kernels + harness, once per math-library variant
    variant_objs = {0: [], 1: []}
    for libm in (0, 1):
        tag = "libm" if libm else "rt"
        for (kfile, kid, suffix, defs) in KERNEL_TUS:
            o = os.path.join(obj, f"k_{tag}_{kfile.replace('.cl', '')}{suffix}.o")
            variant_objs[libm].append(o)
            jobs.append(common + ["-I" + shim, "-I" + gen, "-I" + os.path.join(REPO, "include"),
                                  f"-DRT_REF_LIBM={libm}", f'-DRT_KERNEL_FILE="src/kernels/cl/{kfile}"',
                                  f"-DRT_KERNEL_ID={kid}", f"-DRT_SUFFIX={suffix}",
                                  f"-DRT_NS=k_{kfile.replace('.cl', '')}{suffix}"] + ["-D" + d for d in defs] +
                        ["-c", os.path.join(shim, "kernel_tu.cpp"), "-o", o])
        o = os.path.join(obj, f"harness_{tag}.o")
        variant_objs[libm].append(o)
        jobs.append(common + host_flags + [f"-DRT_REF_LIBM={libm}", "-c", os.path.join(HERE, "ref_harness.cpp"), "-o", o])

    stubs_o = os.path.join(obj, "cl_stubs.o")
    jobs.append(["gcc", "-O2", "-fPIC", "-c", os.path.join(shim, "cl_stubs.c"), "-o", stubs_o])
    host_objs.append(stubs_o)

    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        list(ex.map(run, jobs))

    for libm, lib in ((0, libs[0]), (1, libs[1])):
        run(["g++", "-shared", "-fopenmp", "-o", lib] + host_objs + variant_objs[libm])
    with open(stamp, "w") as fh:
        fh.write(h.hexdigest())
    print("ref: built", ", ".join(os.path.relpath(l, REPO) for l in libs))
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    build(a.reference, a.force)
