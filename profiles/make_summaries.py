#!/usr/bin/env python3
"""Turn the ncu reports of one profiling run (gpurun_out/prof_{trace,shade,shadow}_<tag>.ncu-rep + launches_<tag>.csv,
written by profiles/run_ncu.sh) into the committed text summaries and profiles/traffic.json.
usage: profiles/make_summaries.py <tag> [gpurun_out]"""
import collections
import csv
import json
import os
import subprocess
import sys

tag = sys.argv[1]
src = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out"
HERE = os.path.dirname(os.path.abspath(__file__))


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def fnum(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


traffic = {}
for k, cls in (("trace", "trace_closest"), ("shade", "shade_queues"), ("shadow", "shadow_accumulate"), ("both", "trace_both")):
    rep = os.path.join(src, f"prof_{k}_{tag}.ncu-rep")
    if not os.path.exists(rep):
        continue
    s1 = subprocess.run([sys.executable, os.path.join(HERE, "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    s2 = subprocess.run([sys.executable, os.path.join(HERE, "ncu_lines.py"), rep, ":::2", "25"], capture_output=True, text=True).stdout
    hdr, units, rows = raw(rep)
    col = lambda name: hdr.index(name)
    scale = lambda name: {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}[units[col(name)]]
    rd = [fnum(r[col("dram__bytes_read.sum")]) * scale("dram__bytes_read.sum") for r in rows]
    wr = [fnum(r[col("dram__bytes_write.sum")]) * scale("dram__bytes_write.sum") for r in rows]
    du = [fnum(r[col("gpu__time_duration.sum")]) * scale("gpu__time_duration.sum") for r in rows]
    per_launch = (sum(rd) + sum(wr)) / len(rows)
    traffic[cls] = {"launches": len(rows), "dram_bytes_per_launch": per_launch,
                    "dram_GBs_under_ncu": (sum(rd) + sum(wr)) / sum(du) / 1e9, "duration_us_under_ncu": [round(d * 1e6, 1) for d in du]}
    with open(os.path.join(HERE, f"{tag}_{k}_summary.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none, kernel class {cls}, bench.py CornellBox 1920x1080 x 8 bounces, tag {tag}\n")
        f.write("# one column per launch = the launches of this kernel in one frame (cold-cache, serialised by the profiler: compare shares, not absolutes)\n")
        if k in ("trace", "shadow"):
            f.write("# captured with bench.py --overlap 0 (each pass its own kernel: the configuration bench.py's roofline region times);\n"
                    "# in the default schedule the shadow pass of bounce b runs inside the traversal kernel of bounce b+1 (k_trace_both)\n")
        f.write("\n".join(l[:400] for l in s1.splitlines()) + "\n\n")
        f.write(f"DRAM traffic per launch (read+write, mean over {len(rows)} launches): {per_launch / 1e6:.1f} MB; "
                f"achieved DRAM bandwidth under ncu: {traffic[cls]['dram_GBs_under_ncu']:.0f} GB/s "
                f"(= {100 * traffic[cls]['dram_GBs_under_ncu'] / 8000:.1f} % of the 8 TB/s nominal chip peak)\n\n")
        f.write("## per source line, bounce 1 (launch 2)\n" + "\n".join(l[:170] for l in s2.splitlines()) + "\n")

lc = os.path.join(src, f"launches_{tag}.csv")
if os.path.exists(lc):
    lines = [l for l in open(lc) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = fnum(row["Metric Value"])
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        agg.setdefault(row["Kernel Name"].split("(")[0][-44:], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(os.path.join(HERE, f"{tag}_launch_shares.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none, bench.py --steps 2 --warmup 3 (+1 counter frame, +e2e frames), tag {tag}\n")
        f.write("# per-launch times are cold-cache and serialised: the SHARES are what must agree with bench.py's CUDA-event split\n")
        for k, v in agg.items():
            f.write(f"{k:46s} launches {len(v):4d}  total {sum(v):10.1f} us  share {100 * sum(v) / tot:5.1f}%\n")
    with open(os.path.join(HERE, f"{tag}_launches.csv"), "w") as f:
        f.writelines(lines)

json.dump({"tag": tag, "workload": "CornellBox 1920x1080 1spp 8-bounce", "kernels": traffic}, open(os.path.join(HERE, "traffic.json"), "w"), indent=1)
print(json.dumps(traffic, indent=1))
