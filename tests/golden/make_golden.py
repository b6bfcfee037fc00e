"""tests/golden/make_golden.py -- regenerates the committed golden vectors.

Run in the build container (where /root/reference exists):  python tests/golden/make_golden.py
Every OUTPUT in these fixtures is produced by the reference's own kernel source, gaussian_kernel.cl, compiled
unmodified behind oracle/cl_shim.h (oracle/_ref/libgaussian_ref.so); the distribution known-answers are copied
from the reference's run logs with their file:line.  The fixtures are what pins oracle/gaussian_oracle.c and the
CUDA path on machines where the reference tree does not exist (the GPU box).

  vectors.npz     small images (known-answer impulses, edge-case shapes, C in {1,3,4}, a crop of the reference's
                  own test photo) -> in_<name>, out_<name>
  checksums.json  full-size seeded synthetic batches (inputs regenerated from the seed) -> sha256 of the output
  distribution.json  Approach-1 / Approach-2 partition known-answers from data/**.txt
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402

REF = "/root/reference"


def synth(seed: int, n: int, h: int, w: int, c: int = 3) -> np.ndarray:
    """The synthetic stream used everywhere: uniform uint8 per byte, PCG64(seed)."""
    return np.random.default_rng(seed).integers(0, 256, size=(n, h, w, c), dtype=np.uint8)


def main() -> None:
    oracle.build()
    assert oracle.have_ref(), "needs the reference tree to build oracle/_ref"
    vec = {}

    def add(name, img):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        vec["in_" + name] = img
        vec["out_" + name] = oracle.ref_blur(img)

    # known-answer impulses
    for name, (r, c) in {"impulse_centre": (2, 2), "impulse_corner": (0, 0), "impulse_top_edge": (0, 2),
                         "impulse_bottom_right": (4, 4), "impulse_left_edge": (2, 0)}.items():
        img = np.zeros((5, 5, 3), np.uint8)
        img[r, c, :] = 255
        add(name, img)
    for k in (0, 1, 127, 255):
        add(f"const_{k}", np.full((7, 9, 3), k, np.uint8))
    # edge-case shapes, C = 3
    rng = np.random.default_rng(20261018)
    for (h, w) in [(1, 1), (1, 9), (9, 1), (2, 2), (3, 3), (17, 33), (16, 16), (5, 16), (33, 64), (48, 80),
                   (2, 48), (31, 100)]:
        add(f"rand_{h}x{w}x3", rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8))
    # other channel counts (the kernel loops over `channels`, gaussian_kernel.cl:44)
    for c in (1, 2, 4):
        add(f"rand_12x20x{c}", rng.integers(0, 256, size=(12, 20, c), dtype=np.uint8))
        add(f"rand_7x16x{c}", rng.integers(0, 256, size=(7, 16, c), dtype=np.uint8))
    # saturated / checkerboard patterns: exercise the truncation at the top of the range
    cb = np.indices((16, 32)).sum(axis=0) % 2 * 255
    add("checker_16x32x3", np.repeat(cb[:, :, None], 3, axis=2))
    # a crop of the reference's own input photo (decoded with Pillow; only a 64x96 window is stored)
    try:
        from PIL import Image
        photo = np.asarray(Image.open(os.path.join(REF, "image_320x240.jpg")).convert("RGB"))
        add("photo_crop_64x96x3", photo[100:164, 120:216, :])
        full = oracle.ref_blur(np.ascontiguousarray(photo))
        photo_sha = {"in_sha256": hashlib.sha256(np.ascontiguousarray(photo).tobytes()).hexdigest(),
                     "out_sha256": hashlib.sha256(full.tobytes()).hexdigest(), "shape": list(photo.shape)}
    except Exception as e:  # pragma: no cover
        print("photo crop skipped:", e)
        photo_sha = None
    np.savez_compressed(os.path.join(HERE, "vectors.npz"), **vec)

    sums = {"generator": "np.random.default_rng(seed).integers(0,256,(n,h,w,c),uint8)", "cases": []}
    for (seed, n, h, w, c) in [(1, 8, 256, 256, 3), (2, 8, 240, 320, 3), (3, 35, 240, 320, 3), (4, 3, 100, 52, 3),
                               (5, 2, 512, 1024, 3), (6, 4, 64, 64, 4), (7, 4, 33, 48, 1)]:
        x = synth(seed, n, h, w, c)
        y = oracle.ref_blur_batch(x)
        sums["cases"].append({"seed": seed, "n": n, "h": h, "w": w, "c": c,
                              "in_sha256": hashlib.sha256(x.tobytes()).hexdigest(),
                              "out_sha256": hashlib.sha256(y.tobytes()).hexdigest()})
    if photo_sha:
        sums["reference_photo_320x240"] = photo_sha
    with open(os.path.join(HERE, "checksums.json"), "w") as f:
        json.dump(sums, f, indent=1)

    # (batch_size, gpu_ratio) -> totals printed by the reference's own runs (ratio as passed on the command line,
    # BASELINE.md section 1; the logs print it with one decimal).
    a1_logs = [(35, 0.728, 143, 1429, 3571, "data/approach1/35_run_1.txt:50,:57"),
               (50, 0.728, 100, 1400, 3600, "data/approach1/50_run_1.txt:50,:57"),
               (100, 0.814, 50, 950, 4050, "data/approach1/100_run_1.txt:50,:57"),
               (200, 0.814, 25, 950, 4050, "data/approach1/200_run_1.txt:50,:57"),
               (500, 0.833, 10, 840, 4160, "data/approach1/500_run_1.txt:50,:57"),
               (800, 0.837, 7, 819, 4181, "data/approach1/800_run_1.txt:50,:57"),
               (1200, 0.834, 5, 834, 4166, "data/approach1/1200_run_1.txt:50,:57")]
    a2_logs = [(240, 0.837, 39, "data/approach2/35_run_1.txt:16-18"),
               (240, 0.836, 39, "data/approach2/50_run_1.txt:16-18"),
               (240, 0.889, 26, "data/approach2/100_run_1.txt:16-18"),
               (240, 0.886, 27, "data/approach2/200_run_1.txt:16-18"),
               (240, 0.885, 27, "data/approach2/500_run_1.txt:16-18"),
               (240, 0.5, 120, "data/approach1/run_1.txt:16-18 (mis-filed Approach-2 log)")]
    dist = {
        "a1": [{"num_images": 5000, "batch_size": b, "gpu_ratio": r, "mode": 0, "num_batches": nb,
                "total_cpu": c, "total_gpu": g, "source": src} for (b, r, nb, c, g, src) in a1_logs],
        "a2": [{"height": h, "gpu_ratio": r, "split_row": s, "cpu_input_rows": s + 1, "cpu_output_rows": s,
                "gpu_input_rows": h - s + 1, "gpu_output_rows": h - s, "source": src}
               for (h, r, s, src) in a2_logs],
    }
    with open(os.path.join(HERE, "distribution.json"), "w") as f:
        json.dump(dist, f, indent=1)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
