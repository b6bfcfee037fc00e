#!/usr/bin/env python
"""tools/tight_bench.py -- kernel-only timing of the streamed kernel on odd-width images: tight unaligned input rows
(TIGHT input form) vs the same rows re-pitched to a multiple of 16 bytes, both with a 16-byte-pitched output."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "heterogeneous-opencl-image-processing-engine_b200")):
    sys.path.insert(0, p)
import torch, b200blur
ctx = b200blur.Context(0, 4)
for (n, h, w) in [(5000, 250, 250), (2000, 480, 642)]:
    c = 3; P = w * c; pitch = (P + 15) // 16 * 16
    x = torch.randint(0, 256, (n, h, w, c), dtype=torch.uint8, device="cuda")
    y = torch.zeros((n, h, pitch), dtype=torch.uint8, device="cuda")
    xp = torch.zeros((n, h, pitch), dtype=torch.uint8, device="cuda")
    xp[:, :, :P] = x.view(n, h, P)
    torch.cuda.synchronize()
    lt = ctx.launch_rows(x, y, w, h, c, 0, h, n, in_row_pitch=0, out_row_pitch=pitch, in_image_stride=h * P, out_image_stride=h * pitch)
    lp = ctx.launch_rows(xp, y, w, h, c, 0, h, n, in_row_pitch=pitch, out_row_pitch=pitch)
    for name, l in (("tight-in", lt), ("pitched-in", lp)):
        for _ in range(3): ctx.enqueue_blur(0, l)
        e0 = ctx.enqueue_marker(0)
        for _ in range(10): ctx.enqueue_blur(0, l)
        e1 = ctx.enqueue_marker(0)
        ms = ctx.elapsed_ms(e0, e1) / 10
        print(json.dumps({"shape": [n, h, w], "kernel": name, "ms": round(ms, 4), "GBps": round(2 * n * h * P / ms / 1e6, 1)}), flush=True)
