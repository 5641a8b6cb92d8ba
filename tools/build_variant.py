"""Builds a tuning variant of the CUDA library with extra -D flags: raytracing_b200/variants/librt_b200_<name>.so
(select it with RT_B200_LIB=<path>).  usage: build_variant.py name [-DX=Y ...]"""
import os, subprocess, sys
HERE = os.path.dirname(os.path.abspath(__file__)); REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)
import __graft_entry__ as g
name, flags = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(REPO, "raytracing_b200", "variants"); os.makedirs(out_dir, exist_ok=True)
out = os.path.join(out_dir, f"librt_b200_{name}.so")
cmd = [g.NVCC] + g.NVCC_FLAGS + flags + ["-Xptxas", "-v", "-o", out, os.path.join(REPO, "raytracing_b200", "csrc", "rt_kernels.cu")]
r = subprocess.run(cmd, capture_output=True, text=True)
if r.returncode:
    sys.exit(r.stdout + r.stderr)
lines = r.stderr.splitlines()
for i, l in enumerate(lines):
    if "Compiling entry function" in l and ("k_frame" in l or "k_trace_both" in l or "k_shade_queues" in l):
        print(l.split("'")[1][52:90], "|", lines[i + 2].strip() if i + 2 < len(lines) else "", "|", lines[i + 3].strip()[:60] if i + 3 < len(lines) else "")
print(out)
