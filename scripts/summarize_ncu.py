#!/usr/bin/env python
"""Summarise ncu output brought back from the GPU box into small text files under profiles/.

    python scripts/summarize_ncu.py <tag> <gpurun_out/dir> [workload ...]

Reads <dir>/launches.csv (the `--metrics gpu__time_duration.sum` launch list) and
<dir>/prof_<workload>.ncu-rep (one `--set full` capture of each kernel of a step), writes
profiles/<tag>_launches.csv (verbatim), profiles/<tag>_summary.md and
profiles/traffic_<workload>.json (dram bytes per launch, read by bench.py for roofline.traffic).
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    ("gpu__time_duration.sum", "time_us", 1.0),
    ("dram__bytes_read.sum", "dram_read", 1.0),
    ("dram__bytes_write.sum", "dram_write", 1.0),
    ("launch__grid_size", "grid", 1.0),
    ("launch__block_size", "block", 1.0),
    ("launch__registers_per_thread", "regs", 1.0),
    ("launch__waves_per_multiprocessor", "waves", 1.0),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct", 1.0),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct", 1.0),
    ("smsp__inst_executed.sum", "warp_insts", 1.0),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct", 1.0),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct", 1.0),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex_pct", 1.0),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct", 1.0),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct", 1.0),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_pct", 1.0),
]
SHORT = {"advectKernel": "advect", "advectCellsKernel": "advect", "advectParticlesKernel": "advect_particles",
         "geometricKernel": "geometric", "divergenceFFTKernel": "divergence_fft",
         "tridiagonalKernel": "tridiagonal", "inverseFFTGradientKernel": "inverse_fft_gradient"}


def short(name):
    for k, v in SHORT.items():
        if k in name:
            return v
    return name.split("(")[0][-40:]


def to_bytes(value, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    return float(value.replace(",", "")) * scale


def main():
    tag, src = sys.argv[1], sys.argv[2]
    workloads = sys.argv[3:] or ["c2"]
    out = [f"# ncu summary {tag}\n"]
    lpath = os.path.join(src, "launches.csv")
    if os.path.exists(lpath):
        shutil.copy(lpath, os.path.join(ROOT, "profiles", f"{tag}_launches.csv"))
        rows = [r for r in csv.reader(open(lpath)) if len(r) > 5]
        hdr = rows[0]
        ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
        d = collections.OrderedDict()
        for r in rows[1:]:
            try:
                d.setdefault(short(r[ki]), []).append(float(r[vi].replace(",", "")))
            except ValueError:
                pass
        tot = sum(sum(v) for v in d.values())
        out.append("## launch list (C2; cold-cache, serialised under ncu: shares only)\n")
        out.append("| kernel | launches | mean us | share of step |\n|---|---|---|---|")
        for k, v in d.items():
            out.append(f"| {k} | {len(v)} | {sum(v)/len(v)/1e3:.2f} | {sum(v)/tot:.3f} |")
        out.append("")
    for w in workloads:
        rep = os.path.join(src, f"prof_{w}.ncu-rep")
        if not os.path.exists(rep):
            continue
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        out.append(f"## full capture, workload {w} (one launch of each kernel, `--set full --clock-control none`)\n")
        cols = [k for k, _, _ in KEYS if k in hdr]
        out.append("| kernel | " + " | ".join(n for k, n, _ in KEYS if k in hdr) + " |")
        out.append("|---|" + "---|" * len(cols))
        traffic = {}
        for r in rows[2:]:
            name = short(r[hdr.index("Kernel Name")])
            vals = []
            for k in cols:
                i = hdr.index(k)
                v = r[i]
                if k.startswith("dram__bytes"):
                    b = to_bytes(v, units[i])
                    traffic[name] = traffic.get(name, 0) + b
                    v = f"{b/1e6:.2f} MB"
                elif k == "gpu__time_duration.sum":
                    v = f"{float(v.replace(',', '')) * ({'ns': 1e-3, 'us': 1, 'ms': 1e3}.get(units[i], 1)):.2f}"
                else:
                    try:
                        v = f"{float(v.replace(',', '')):.4g}"
                    except ValueError:
                        pass
                vals.append(v)
            out.append(f"| {name} | " + " | ".join(vals) + " |")
        out.append("")
        json.dump(traffic, open(os.path.join(ROOT, "profiles", f"traffic_{w}.json"), "w"), indent=1)
    open(os.path.join(ROOT, "profiles", f"{tag}_summary.md"), "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main()
