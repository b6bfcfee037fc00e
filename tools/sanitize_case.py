#!/usr/bin/env python
"""tools/sanitize_case.py -- a small, fast exercise of every kernel (streamed full-width, streamed column blocks,
strips, generic, band with halo pointers) for `compute-sanitizer --tool memcheck|racecheck`; checks parity too."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "heterogeneous-opencl-image-processing-engine_b200")):
    sys.path.insert(0, p)
import numpy as np
import b200blur
from oracle import oracle

rng = np.random.default_rng(9)
ctx = b200blur.Context(0, 4)
for variant in (0, 1):
    ctx.set_kernel_variant(variant)
    for shape in [(6, 40, 320, 3), (3, 17, 2048, 3), (4, 9, 16, 3), (3, 7, 21, 3), (2, 33, 100, 4), (5, 64, 256, 1)]:
        x = rng.integers(0, 256, size=shape, dtype=np.uint8)
        assert np.array_equal(ctx.blur_numpy(x), oracle.c_blur_batch(x)), (variant, shape)
    # bands with halo pointers
    n, h, w, c = 4, 48, 128, 3
    P = w * c
    x = rng.integers(0, 256, size=(n, h, w, c), dtype=np.uint8)
    out = np.zeros_like(x)
    d_in, d_out = ctx.dev_alloc(x.nbytes), ctx.dev_alloc(x.nbytes)
    ctx.enqueue_write(0, d_in, x, x.nbytes)
    for r0, rows in ((0, 16), (16, 16), (32, 16)):
        top = 1 if r0 > 0 else 0
        in_rows = rows + top + (1 if r0 + rows < h else 0)
        ctx.enqueue_blur(0, ctx.launch_rows(d_in + (r0 - top) * P, d_out + r0 * P, w, in_rows, c, top, rows, n, P * h, P * h))
    ctx.enqueue_read(0, out, d_out, x.nbytes)
    ctx.finish()
    ctx.dev_free(d_in); ctx.dev_free(d_out)
    assert np.array_equal(out, oracle.c_blur_batch(x)), ("bands", variant)
ctx.close()
print("sanitize_case ok")
