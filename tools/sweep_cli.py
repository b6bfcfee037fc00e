#!/usr/bin/env python
"""tools/sweep_cli.py -- the reference authors' methodology on B200: for each batch size in {35,50,100,200,500,800,1200}
run the CLI three times, keep the stdout logs as `<batch>_run_<i>.txt` (the naming of the reference's data/ directory),
then roll them up with tools/rollup.py into per_run.csv and avg_by_batch.csv (the reference's CSV schema).

    python tools/sweep_cli.py approach1|approach2 OUT_DIR [--gpus G] [extra CLI flags...]
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "heterogeneous-opencl-image-processing-engine_b200", "bin")
BATCHES = (35, 50, 100, 200, 500, 800, 1200)


def main():
    which, out_dir = sys.argv[1], sys.argv[2]
    extra = sys.argv[3:]
    os.makedirs(out_dir, exist_ok=True)
    logs = []
    for b in BATCHES:
        for run in (1, 2, 3):
            if which == "approach1":
                cmd = [os.path.join(BIN, "heterogeneous_blur"), "both", "0.728", str(b), "--quiet"] + extra
            else:
                cmd = [os.path.join(BIN, "split_image_blur"), "0.837", str(b), "--quiet"] + extra
            out = subprocess.run(cmd, capture_output=True, text=True, cwd=out_dir)
            if out.returncode != 0:
                raise SystemExit(f"{' '.join(cmd)} failed:\n{out.stdout[-2000:]}{out.stderr[-2000:]}")
            path = os.path.join(out_dir, f"{b}_run_{run}.txt")
            with open(path, "w") as f:
                f.write(out.stdout)
            logs.append(path)
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "rollup.py"), *logs, "--per-run",
                    os.path.join(out_dir, "per_run.csv"), "--avg", os.path.join(out_dir, "avg_by_batch.csv")], check=True)
    print(open(os.path.join(out_dir, "avg_by_batch.csv")).read())


if __name__ == "__main__":
    main()
