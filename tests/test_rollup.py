"""tools/rollup.py must reproduce the reference's own CSV roll-up (data/approach2/approach2/per_run.csv) from the
reference's own run logs -- the generator script is not shipped with the reference.  Runs only
where the reference tree is mounted; the parser is also exercised on a log in this repo's CLI format."""
import csv
import glob
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import rollup  # noqa: E402

REF = "/root/reference/data/approach2"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")
def test_reproduces_the_reference_per_run_csv():
    with open(os.path.join(REF, "approach2", "per_run.csv")) as f:
        want = {r["file"]: r for r in csv.DictReader(f)}
    logs = [p for p in glob.glob(os.path.join(REF, "*_run_*.txt")) if os.path.basename(p) in want]
    assert len(logs) >= 20
    numeric = ["images", "batches", "img_w", "img_h", "wg_w", "wg_h", "wall_ms", "cpu_images", "cpu_total_ms", "cpu_in_ms",
               "cpu_kernel_ms", "cpu_out_ms", "gpu_images", "gpu_total_ms", "gpu_in_ms", "gpu_kernel_ms", "gpu_out_ms",
               "imbalance_pct", "bottleneck_delta_ms", "mpix_per_sec", "img_per_sec", "recommended_gpu_ratio",
               "batch_size_log", "batch_size_file", "run"]
    for p in logs:
        got = rollup.parse_log(p)
        ref = want[os.path.basename(p)]
        for c in numeric:
            assert ref[c] != "", (c, p)
            assert float(got[c]) == pytest.approx(float(ref[c]), abs=1e-9), (c, os.path.basename(p))
        assert got["bottleneck"] == ref["bottleneck"]
        if ref["speedup_gpu_vs_cpu"]:
            assert float(got["speedup_gpu_vs_cpu"]) == pytest.approx(float(ref["speedup_gpu_vs_cpu"]))
    rows = [rollup.parse_log(p) for p in logs]
    avg = {r["batch_size_file"]: r for r in rollup.average_by_batch(rows)}
    with open(os.path.join(REF, "approach2", "avg_by_batch.csv")) as f:
        for r in csv.DictReader(f):
            mine = avg[int(float(r["batch_size_file"]))]
            assert float(mine["img_per_sec"]) == pytest.approx(float(r["img_per_sec"]), rel=1e-3)
            assert float(mine["wall_ms"]) == pytest.approx(float(r["wall_ms"]), rel=1e-3)


def test_parses_this_repos_cli_report(tmp_path):
    log = tmp_path / "35_run_1.txt"
    log.write_text("""Mode: HETEROGENEOUS (CPU + GPU)
GPU ratio: 72.8% GPU, 27.2% CPU
Number of images in stream: 5000
Batch size: 35 images
Number of batches: 143
Work-group size: 16x16
Original image loaded: 320x240, 3 channels
1. OVERALL EXECUTION TIME
   Total wall-clock time: 35.45 ms (0.04 seconds)
   Total images processed: 5000

2. GPU 0 DEVICE (processed 2500 images)
   Total GPU time:        21.00 ms
   - Transfer IN:         10.00 ms (47.6%)
   - Kernel execution:    1.00 ms (4.8%)
   - Transfer OUT:        10.00 ms (47.6%)
   Average per image:     0.00840 ms

3. GPU 1 DEVICE (processed 2500 images)
   Total GPU time:        22.00 ms
   - Transfer IN:         10.50 ms (47.7%)
   - Kernel execution:    1.00 ms (4.5%)
   - Transfer OUT:        10.50 ms (47.7%)
   Average per image:     0.00880 ms

4. DEVICE COMPARISON
   GPU 0 is 1.05x FASTER than GPU 1
5. WORKLOAD BALANCE
   Workload imbalance: 4.5%
   GPU 1 is the BOTTLENECK (1.00 ms slower)
7. THROUGHPUT
   Overall throughput: 10831.43 Megapixels/sec
   Images per second: 141039.06
   Run with: ./heterogeneous_blur both 0.728   (shares are even by construction)
""")
    r = rollup.parse_log(str(log))
    assert (r["batch_size_file"], r["run"], r["images"], r["batches"], r["n_devices"]) == (35, 1, 5000, 143, 2)
    assert r["cpu_total_ms"] == 21.0 and r["gpu_out_ms"] == 10.5 and r["img_per_sec"] == 141039.06
    assert r["bottleneck"] == "GPU 1" and r["speedup_gpu_vs_cpu"] == 1.05 and r["recommended_gpu_ratio"] == 0.728
