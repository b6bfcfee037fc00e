#!/usr/bin/env python
"""tools/prof.py -- small, ncu-friendly driver of the device-resident hot path (no CPU baseline, no e2e).

    python tools/prof.py [--images 5000] [--height 240] [--width 320] [--iters 5] [--variant 0] [--batch 35]
                         [--per-batch] [--check]
Prints one line per configuration: images/s and algorithmic GB/s from CUDA events on the launching queue."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "heterogeneous-opencl-image-processing-engine_b200")):
    sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=5000)
    ap.add_argument("--height", type=int, default=240)
    ap.add_argument("--width", type=int, default=320)
    ap.add_argument("--channels", type=int, default=3)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--variant", type=int, nargs="*", default=[0])
    ap.add_argument("--batch", type=int, default=35)
    ap.add_argument("--per-batch", action="store_true", help="same as --coalesce 0")
    ap.add_argument("--coalesce", type=int, nargs="*", default=None,
                    help="1 = batches fused (default), 0 = one work descriptor per batch (feed kernel), 2 = one launch per batch")
    ap.add_argument("--back-to-back", type=int, default=0, help="also time this many passes enqueued back to back (one event pair)")
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--copy", action="store_true", help="also time torch's d_out.copy_(d_in) on the same buffers (practical ceiling)")
    a = ap.parse_args()
    import torch
    import b200blur
    n, h, w, c = a.images, a.height, a.width, a.channels
    ctx = b200blur.Context(0, 4)
    g = torch.Generator(device="cuda").manual_seed(1)
    d_in = torch.randint(0, 256, (n, h, w, c), dtype=torch.uint8, device="cuda", generator=g)
    d_out = torch.zeros_like(d_in)
    torch.cuda.synchronize()
    if a.copy:
        for _ in range(3):
            d_out.copy_(d_in)
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(a.iters):
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            d_out.copy_(d_in)
            t1.record()
            torch.cuda.synchronize()
            best = min(best, t0.elapsed_time(t1))
        print(json.dumps({"copy_ceiling_best_ms": round(best, 4), "GBps_best": round(2.0 * n * h * w * c / 1e9 / best * 1e3, 1)}), flush=True)
    modes = a.coalesce if a.coalesce is not None else [0 if a.per_batch else 1]
    for v, mode in [(v, m) for v in a.variant for m in modes]:
        ctx.set_kernel_variant(v)
        d_out.zero_()
        for _ in range(a.warmup):
            ctx.run_resident(d_in, d_out, w, h, c, n, a.batch, mode, stats=False)
        ctx.finish()
        best = 1e30
        tot = 0.0
        for _ in range(a.iters):
            e0 = ctx.enqueue_marker(0)
            ctx.run_resident(d_in, d_out, w, h, c, n, a.batch, mode, stats=False)
            e1 = ctx.enqueue_marker(0)
            ms = ctx.elapsed_ms(e0, e1)
            best = min(best, ms)
            tot += ms
        avg = tot / a.iters
        b2b = None
        if a.back_to_back:
            e0 = ctx.enqueue_marker(0)
            for _ in range(a.back_to_back):
                ctx.run_resident(d_in, d_out, w, h, c, n, a.batch, mode, stats=False)
            e1 = ctx.enqueue_marker(0)
            b2b = ctx.elapsed_ms(e0, e1) / a.back_to_back
        gb = 2.0 * n * h * w * c / 1e9
        ok = None
        if a.check:
            from oracle import oracle
            idx = list(range(0, n, max(1, n // 16)))[:16]
            ok = bool((d_out[idx].cpu().numpy() == oracle.c_blur_batch(d_in[idx].cpu().numpy(), integer=True)).all())
        print(json.dumps({"variant": v, "shape": [n, h, w, c], "batch": a.batch, "coalesce": mode, "avg_ms": round(avg, 4),
                          "back_to_back_ms": None if b2b is None else round(b2b, 4),
                          "back_to_back_GBps": None if b2b is None else round(gb / b2b * 1e3, 1),
                          "best_ms": round(best, 4), "img_per_s_avg": round(n / avg * 1e3),
                          "GBps_avg": round(gb / avg * 1e3, 1), "GBps_best": round(gb / best * 1e3, 1), "parity": ok}),
              flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
