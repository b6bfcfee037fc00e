#!/usr/bin/env python
"""tests/golden/make_jpeg_golden.py -- writes the JPEG fixtures under tests/golden/jpeg/ and jpeg/expected.json.

Each fixture is a small baseline JPEG written by Pillow (libjpeg-turbo) from a seeded synthetic image; expected.json
holds the sha256 of the pixels libjpeg-turbo DECODES from it (Pillow's decoder: integer slow IDCT, fancy upsampling --
the same defaults CImg/libjpeg use in the reference, heterogeneous_blur.c:106).  host/jpeg_decode.hpp must reproduce
those bytes exactly (tests/test_jpeg_cpu.py).  One crop of the reference's own photo is included when the reference tree
is present (re-encoded, so no reference file is copied).  Run from the repo root: python tests/golden/make_jpeg_golden.py
"""
import hashlib
import json
import os

import numpy as np
from PIL import Image

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "jpeg")


def synth(seed, h, w, gray=False):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    base = np.stack([(x * 255 // max(w - 1, 1)), (y * 255 // max(h - 1, 1)), ((x * 3 + y * 5) % 256)], -1).astype(np.int32)
    img = np.clip(base + rng.integers(-40, 41, size=(h, w, 3)), 0, 255).astype(np.uint8)
    return img[..., 0] if gray else img


CASES = [  # name, (h, w), save options
    ("420_q90_64x48", (48, 64), dict(quality=90, subsampling=2)),
    ("420_q100_37x21", (21, 37), dict(quality=100, subsampling=2)),
    ("420_q75_odd_101x67", (67, 101), dict(quality=75, subsampling=2, optimize=True)),
    ("422_q85_50x33", (33, 50), dict(quality=85, subsampling=1)),
    ("444_q95_31x29", (29, 31), dict(quality=95, subsampling=0)),
    ("gray_q80_45x40", (40, 45), dict(quality=80)),
    ("420_tiny_1x1", (1, 1), dict(quality=90, subsampling=2)),
    ("420_tiny_2x2", (2, 2), dict(quality=90, subsampling=2)),
    ("420_narrow_3x40", (40, 3), dict(quality=90, subsampling=2)),
    ("420_narrow_5x17", (17, 5), dict(quality=60, subsampling=2)),
    ("422_narrow_4x9", (9, 4), dict(quality=90, subsampling=1)),
    ("420_q30_96x80", (80, 96), dict(quality=30, subsampling=2)),
]


def main():
    os.makedirs(HERE, exist_ok=True)
    expected = {}
    for i, (name, (h, w), opts) in enumerate(CASES):
        gray = name.startswith("gray")
        img = Image.fromarray(synth(100 + i, h, w, gray))
        path = os.path.join(HERE, name + ".jpg")
        img.save(path, "JPEG", **opts)
        expected[name + ".jpg"] = None
    # restart intervals (Pillow >= 10.2 exposes them)
    try:
        Image.fromarray(synth(300, 40, 72)).save(os.path.join(HERE, "420_restart_72x40.jpg"), "JPEG", quality=88, subsampling=2,
                                                 restart_marker_blocks=3)
        expected["420_restart_72x40.jpg"] = None
    except Exception as e:  # pragma: no cover
        print("no restart-interval fixture:", e)
    ref = "/root/reference/image_320x240.jpg"
    if os.path.exists(ref):
        crop = Image.open(ref).convert("RGB").crop((100, 60, 180, 110))
        crop.save(os.path.join(HERE, "420_photo_crop_80x50.jpg"), "JPEG", quality=100, subsampling=2)
        expected["420_photo_crop_80x50.jpg"] = None
    Image.fromarray(synth(400, 24, 32)).save(os.path.join(HERE, "progressive_32x24.jpg"), "JPEG", quality=80, progressive=True)
    for name in sorted(expected):
        im = Image.open(os.path.join(HERE, name))
        arr = np.array(im if im.mode == "L" else im.convert("RGB"))
        expected[name] = {"width": arr.shape[1], "height": arr.shape[0], "channels": 1 if arr.ndim == 2 else 3,
                          "sha256": hashlib.sha256(arr.tobytes()).hexdigest()}
    with open(os.path.join(HERE, "expected.json"), "w") as f:
        json.dump({"decoder": "libjpeg-turbo via Pillow %s" % __import__("PIL").__version__, "unsupported": ["progressive_32x24.jpg"],
                   "files": expected}, f, indent=1, sort_keys=True)
    print(len(expected), "fixtures")


if __name__ == "__main__":
    main()
