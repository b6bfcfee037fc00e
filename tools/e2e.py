#!/usr/bin/env python
"""tools/e2e.py -- times b200blur_run_host (pinned host buffers -> H2D -> blur -> D2H) on the bench workload."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "heterogeneous-opencl-image-processing-engine_b200")):
    sys.path.insert(0, p)
import torch, b200blur
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 35
w = int(sys.argv[2]) if len(sys.argv) > 2 else 320
h = int(sys.argv[3]) if len(sys.argv) > 3 else 240
n, c = 5000, 3
ctx = b200blur.Context(0, 4)
h_in = torch.randint(0, 256, (n, h, w, c), dtype=torch.uint8).pin_memory()
h_out = torch.empty_like(h_in).pin_memory()
for _ in range(2):
    ctx.run_host(h_in, h_out, w, h, c, n, batch, stats=False)
best = 1e9
for _ in range(5):
    e0 = ctx.enqueue_marker(0)
    st = ctx.run_host(h_in, h_out, w, h, c, n, batch)
    e1 = ctx.enqueue_marker(2)
    ctx.finish()
    best = min(best, ctx.elapsed_ms(e0, e1))
print(json.dumps({"batch": batch, "shape": [n, h, w, c], "ring": os.environ.get("B200BLUR_RING"), "fuse": os.environ.get("B200BLUR_E2E_FUSE"),
                  "best_ms": round(best, 3), "img_per_s": round(n / best * 1e3), "GBps_each_way": round(n * h * w * c / best / 1e6, 2),
                  "h2d_ms": round(st.h2d_ms, 2), "d2h_ms": round(st.d2h_ms, 2), "kernel_ms": round(st.kernel_ms, 2)}))
