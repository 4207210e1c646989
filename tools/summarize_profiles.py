#!/usr/bin/env python3
"""Turn gpurun_out/{launches.csv, prof_*.ncu-rep} into small text summaries under profiles/ (run in the dev container:
ncu can read reports without a GPU).   python tools/summarize_profiles.py r01"""
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "rXX"
out_dir = os.path.join(ROOT, "profiles")

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
]


def launches():
    path = os.path.join(ROOT, "gpurun_out", "launches.csv")
    if not os.path.isfile(path):
        return
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    lines, total, agg = [], 0.0, {}
    for r in rows:
        name = r[4].split("(")[0].replace("void ", "").replace("vnect::", "")
        t = float(r[-1])
        t = t / 1000 if r[-2] in ("ns", "nsecond") else t * 1000 if r[-2] in ("ms", "msecond") else t
        total += t
        agg.setdefault(name, [0.0, 0])
        agg[name][0] += t
        agg[name][1] += 1
        lines.append(f"{int(r[0]):4d} {t:9.1f} us  block {r[7]:>14s} grid {r[8]:>14s}  {name}")
    with open(os.path.join(out_dir, f"{tag}_ncu_launches.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none: every launch of ONE bench step\n"
                "# (64 frames, 2 scales = 128 forwards; cold-cache, serialised: compare shares, not absolutes)\n")
        f.write(f"# total {total:.1f} us over {len(rows)} launches\n\n## by kernel\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            f.write(f"{v[0]:9.1f} us {v[1]:3d}x {100 * v[0] / total:5.1f}%  {k}\n")
        f.write("\n## in launch order\n" + "\n".join(lines) + "\n")
    print("wrote launches:", total, "us")


def full(rep, label):
    """rep: one report, or a glob of single-launch reports (gpurun_out/ is capped at 64 MiB, a full-set launch is ~6 MB)"""
    import glob
    paths = sorted(glob.glob(os.path.join(ROOT, "gpurun_out", rep)))
    if not paths:
        return
    hdr = units = None
    rows = [None, None]
    for path in paths:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        part = list(csv.reader(raw.splitlines()))
        if hdr is None:
            hdr, units = part[0], part[1]
            rows = [hdr, units]
        rows += [r for r in part[2:] if len(r) == len(hdr)]
    idx = {m: hdr.index(m) for m in METRICS if m in hdr}
    kn = hdr.index("Kernel Name")
    with open(os.path.join(out_dir, f"{tag}_ncu_full_{label}.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on ({rep}); one block per captured launch\n")
        for r in rows[2:]:
            f.write("\n" + r[kn].replace("CUtensorMap_st, ", "")[:110] + "\n")
            for m, i in idx.items():
                f.write(f"    {m:75s} {r[i]:>16s} {units[i]}\n")
            rd, wr = float(r[idx["dram__bytes_read.sum"]]), float(r[idx["dram__bytes_write.sum"]])
            u = units[idx["dram__bytes_read.sum"]]
            f.write(f"    {'traffic = dram read + write':75s} {rd + wr:16.3f} {u}\n")
    print("wrote", label, len(rows) - 2, "kernels")


def traffic():
    """gpurun_out/step_traffic.csv (ncu metrics pass over one step) -> profiles/<tag>_step_traffic.txt + conv_traffic.json"""
    import collections
    import json
    path = os.path.join(ROOT, "gpurun_out", "step_traffic.csv")
    if not os.path.isfile(path):
        return
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    per = collections.OrderedDict()
    for r in rows:
        d = per.setdefault(int(r[0]), {"name": r[4].split("(")[0].replace("void ", "").replace("vnect::", "")})
        m, unit, val = r[-3], r[-2], float(r[-1].replace(",", ""))
        if m.startswith("dram__bytes") or m.startswith("l1tex"):
            val *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        if m == "gpu__time_duration.sum":
            val *= {"ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}[unit]
        d[m] = val
    conv = [d for d in per.values() if any(k in d["name"] for k in ("conv_gemm", "block_tail", "stem_roll"))]
    ct = sum(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in conv)
    tt = sum(d["gpu__time_duration.sum"] for d in conv)
    out = ["# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,"
           "sm__pipe_tensor_cycles_active...,l1tex__m_xbar2l1tex_read_bytes.sum --clock-control none",
           "# one bench step (64 frames, 2 scales = 128 forwards), every launch",
           f"# implicit-GEMM family ({len(conv)} launches): {tt:.1f} us, DRAM traffic {ct / 1e9:.3f} GB = "
           f"{ct / len(conv) / 1e6:.1f} MB per launch on average, {ct / tt / 1e6:.2f} TB/s", "",
           f"{'#':>3} {'us':>8} {'rd MB':>9} {'wr MB':>9} {'TB/s':>6} {'tensor%':>8} {'L2->SM MB':>10} {'L2->SM TB/s':>11}  kernel"]
    for lid, d in per.items():
        t, rd, wr = d["gpu__time_duration.sum"], d["dram__bytes_read.sum"], d["dram__bytes_write.sum"]
        l2 = d.get("l1tex__m_xbar2l1tex_read_bytes.sum", 0)
        out.append(f"{lid:3d} {t:8.1f} {rd / 1e6:9.1f} {wr / 1e6:9.1f} {(rd + wr) / t / 1e6:6.2f} "
                   f"{d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0):8.1f} {l2 / 1e6:10.1f} "
                   f"{l2 / t / 1e6:11.2f}  {d['name']}")
    open(os.path.join(out_dir, f"{tag}_step_traffic.txt"), "w").write("\n".join(out) + "\n")
    json.dump({"source": f"profiles/{tag}_step_traffic.txt (ncu, one step of 64 frames x 2 scales)",
               "conv_family_launches": len(conv), "conv_family_dram_bytes_per_step": ct,
               "conv_family_dram_bytes_per_launch_avg": ct / len(conv), "conv_family_time_us_under_ncu": tt,
               "frames_per_step": 64,
               "captured_at_commit": subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True,
                                                    text=True).stdout.strip()},
              open(os.path.join(out_dir, "conv_traffic.json"), "w"), indent=1)
    print("wrote traffic:", ct / 1e9, "GB")


def full_csv():
    """gpurun_out/<tag>_full_all.csv (raw page of an --set full capture of every launch of one step) -> one block per launch"""
    path = os.path.join(ROOT, "gpurun_out", f"{tag}_full_all.csv")
    if not os.path.isfile(path):
        return
    part = [r for r in csv.reader(open(path)) if len(r) > 20]
    hdr, units, rows = part[0], part[1], part[2:]
    idx = {m: hdr.index(m) for m in METRICS if m in hdr}
    kn = hdr.index("Kernel Name")
    with open(os.path.join(out_dir, f"{tag}_ncu_full_step.txt"), "w") as f:
        f.write("# ncu --set full --clock-control none over every launch of one bench step (64 frames x 2 scales)\n")
        for n, r in enumerate(rows):
            f.write(f"\n[{n}] " + r[kn].replace("CUtensorMap_st, ", "")[:120] + "\n")
            for m, i in idx.items():
                f.write(f"    {m:75s} {r[i]:>16s} {units[i]}\n")
    print("wrote full step:", len(rows), "launches")


launches()
traffic()
full_csv()
full("prof_stem.ncu-rep", "stem")
full("prof_prepost.ncu-rep", "prepost")
