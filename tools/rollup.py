#!/usr/bin/env python
"""tools/rollup.py -- roll run logs up into the reference's CSV schema.

The reference ships `data/approach2/approach2/per_run.csv` and `avg_by_batch.csv` but not the script that made them.
This reproduces those files from the stdout logs of the two programs -- the reference's own logs under data/ or the
logs of this repo's CLIs, whose report keeps the same section structure -- so B200 results line up column for column
with the published ones.  Device columns: `cpu_*` = the first device block of the report (section 2), `gpu_*` = the
second (section 3); `n_devices` and `extra_devices_total_ms` carry the rest when more than two GPUs ran.

    python tools/rollup.py LOG [LOG ...] [--per-run per_run.csv] [--avg avg_by_batch.csv]
File names like `35_run_1.txt` provide batch_size_file / run, as in the reference's data directory.
"""
import argparse
import csv
import os
import re
import statistics
import sys

COLUMNS = ["batch_size_file", "run", "file", "mode", "gpu_ratio_cfg", "cpu_ratio_cfg", "images", "batches", "img_w", "img_h",
           "wg_w", "wg_h", "wall_ms", "cpu_images", "cpu_total_ms", "cpu_in_ms", "cpu_kernel_ms", "cpu_out_ms",
           "cpu_ms_per_img", "gpu_images", "gpu_total_ms", "gpu_in_ms", "gpu_kernel_ms", "gpu_out_ms", "gpu_ms_per_img",
           "speedup_gpu_vs_cpu", "imbalance_pct", "bottleneck", "bottleneck_delta_ms", "mpix_per_sec", "img_per_sec",
           "recommended_gpu_ratio", "batch_size_log"]
EXTRA = ["n_devices", "extra_devices_total_ms"]


def _f(pattern, text, cast=float, default=""):
    m = re.search(pattern, text)
    return cast(m.group(1)) if m else default


def parse_log(path):
    text = open(path, errors="replace").read()
    name = os.path.basename(path)
    row = dict.fromkeys(COLUMNS + EXTRA, "")
    m = re.match(r"(?:\d+_)?(\d+)_run_(\d+)\.txt$", name)
    if m:
        row["batch_size_file"], row["run"] = int(m.group(1)), int(m.group(2))
    row["file"] = name
    mode = re.search(r"^Mode: (.+)$", text, re.M)
    row["mode"] = mode.group(1).strip() if mode else ""
    g = re.search(r"GPU ratio: ([\d.]+)% GPU, ([\d.]+)% CPU", text)
    if g:
        row["gpu_ratio_cfg"], row["cpu_ratio_cfg"] = float(g.group(1)) / 100, float(g.group(2)) / 100
    row["images"] = _f(r"Number of images in stream: (\d+)", text, int)
    row["batches"] = _f(r"Number of batches: (\d+)", text, int)
    m = re.search(r"Original image loaded: (\d+)x(\d+)", text)
    if m:
        row["img_w"], row["img_h"] = int(m.group(1)), int(m.group(2))
    m = re.search(r"Work-group size: (\d+)x(\d+)", text)
    if m:
        row["wg_w"], row["wg_h"] = int(m.group(1)), int(m.group(2))
    row["wall_ms"] = _f(r"Total wall-clock time: ([\d.]+) ms", text)
    if row["wall_ms"] == "":
        row["wall_ms"] = _f(r"Device-resident kernel time[^:]*: ([\d.]+) ms", text)
    # device blocks: "N. <NAME> DEVICE (processed M images ...)" followed by totals
    blocks = list(re.finditer(r"^\d+\. (.+?) DEVICE \(processed (\d+) images[^)]*\)\s*\n"
                              r"\s*Total .*? time:\s+([\d.]+) ms\s*\n"
                              r"\s*- Transfer IN:\s+([\d.]+) ms.*\n"
                              r"\s*- Kernel execution:\s+([\d.]+) ms.*\n"
                              r"\s*- Transfer OUT:\s+([\d.]+) ms.*\n"
                              r"(?:\s*Average per image:\s+([\d.]+) ms)?", text, re.M))
    for prefix, b in zip(("cpu", "gpu"), blocks[:2]):
        row[prefix + "_images"] = int(b.group(2))
        row[prefix + "_total_ms"] = float(b.group(3))
        row[prefix + "_in_ms"] = float(b.group(4))
        row[prefix + "_kernel_ms"] = float(b.group(5))
        row[prefix + "_out_ms"] = float(b.group(6))
        row[prefix + "_ms_per_img"] = float(b.group(7)) if b.group(7) else ""
    row["n_devices"] = len(blocks)
    if len(blocks) > 2:
        row["extra_devices_total_ms"] = ";".join(b.group(3) for b in blocks[2:])
    m = re.search(r"GPU(?: \d+)? is ([\d.]+)x FASTER than (?:CPU|GPU \d+)", text)
    row["speedup_gpu_vs_cpu"] = float(m.group(1)) if m else ""
    row["imbalance_pct"] = _f(r"Workload imbalance: ([\d.]+)%", text)
    m = re.search(r"(CPU|GPU(?: \d+)?) is the BOTTLENECK \(([\d.]+) ms slower\)", text)
    if m:
        row["bottleneck"], row["bottleneck_delta_ms"] = m.group(1), float(m.group(2))
    row["mpix_per_sec"] = _f(r"Overall throughput: ([\d.]+) Megapixels/sec", text)
    row["img_per_sec"] = _f(r"Images per second: ([\d.]+)", text)
    row["recommended_gpu_ratio"] = _f(r"Run with: \./\w+(?: both)? ([\d.]+)", text)
    row["batch_size_log"] = _f(r"BATCH SIZE\s*:\s*(\d+)", text, int)
    if row["batch_size_log"] == "":
        row["batch_size_log"] = _f(r"Batch size: (\d+) images", text, int)
    return row


def average_by_batch(rows):
    out = []
    keys = sorted({r["batch_size_file"] for r in rows if r["batch_size_file"] != ""})
    for k in keys:
        grp = [r for r in rows if r["batch_size_file"] == k]
        avg = dict.fromkeys(COLUMNS + EXTRA, "")
        avg["batch_size_file"] = k
        for c in COLUMNS + EXTRA:
            if c in ("batch_size_file", "run", "file"):
                continue
            vals = [r[c] for r in grp if r[c] != ""]
            if vals and all(isinstance(v, (int, float)) for v in vals):
                avg[c] = round(statistics.mean(vals), 6)
            elif vals and all(v == vals[0] for v in vals):
                avg[c] = vals[0]
        avg["run"] = len(grp)
        out.append(avg)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("logs", nargs="+")
    ap.add_argument("--per-run", default="-")
    ap.add_argument("--avg", default="")
    a = ap.parse_args()
    rows = [parse_log(p) for p in a.logs]
    rows.sort(key=lambda r: (r["batch_size_file"] if r["batch_size_file"] != "" else 0, r["run"] if r["run"] != "" else 0))
    f = sys.stdout if a.per_run == "-" else open(a.per_run, "w", newline="")
    w = csv.DictWriter(f, fieldnames=COLUMNS + EXTRA)
    w.writeheader()
    w.writerows(rows)
    if a.avg:
        with open(a.avg, "w", newline="") as g:
            cols = [c if c != "run" else "runs" for c in COLUMNS + EXTRA if c != "file"]
            w2 = csv.writer(g)
            w2.writerow(cols)
            for r in average_by_batch(rows):
                w2.writerow([r["run"] if c == "runs" else r[c] for c in cols])


if __name__ == "__main__":
    main()
