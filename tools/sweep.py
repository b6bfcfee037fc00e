#!/usr/bin/env python
"""tools/sweep.py -- BASELINE.json configs[3]: batch-size sweep {1,35,50,100,200,500,800,1200} on 50k x 256x256 RGB,
device-resident -- batches fused (coalesce=1), one work descriptor per batch through the feed kernel (coalesce=0), one
kernel launch per batch (coalesce=2) -- vs end-to-end (pinned host buffers, H2D+D2H in the timed region).
Prints CSV: batch_size, mode, ms, img_per_sec, algorithmic_GBps (resident) or host_link_GBps_each_way (e2e), launches."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "heterogeneous-opencl-image-processing-engine_b200")):
    sys.path.insert(0, p)
import torch, b200blur

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
    h = w = 256; c = 3
    ctx = b200blur.Context(0, 4)
    g = torch.Generator(device="cuda").manual_seed(4)
    d_in = torch.randint(0, 256, (n, h, w, c), dtype=torch.uint8, device="cuda", generator=g)
    d_out = torch.zeros_like(d_in)
    n_e2e = min(n, 10000)
    h_in = torch.empty((n_e2e, h, w, c), dtype=torch.uint8).pin_memory(); h_in.copy_(d_in[:n_e2e])
    h_out = torch.empty_like(h_in).pin_memory()
    torch.cuda.synchronize()
    img_bytes = h * w * c
    print("batch_size,mode,ms,img_per_sec,GBps,launches")
    for b in (1, 35, 50, 100, 200, 500, 800, 1200):
        for mode in ("resident_coalesced", "resident_per_batch_descriptor", "resident_per_batch_launch", "end_to_end"):
            if mode == "resident_per_batch_launch" and b == 1 and n > 5000:
                nn = 5000  # 50k single-image launches would only measure the host launch rate for longer
            else:
                nn = n if mode != "end_to_end" else n_e2e
            best = 1e30; launches = 0
            for it in range(3):
                before = ctx.launch_count
                if mode == "end_to_end":
                    e0 = ctx.enqueue_marker(0)
                    ctx.run_host(h_in, h_out, w, h, c, nn, b, stats=False)
                    e1 = ctx.enqueue_marker(2)
                else:
                    e0 = ctx.enqueue_marker(0)
                    ctx.run_resident(d_in, d_out, w, h, c, nn, b,
                                     {"resident_coalesced": 1, "resident_per_batch_descriptor": 0, "resident_per_batch_launch": 2}[mode],
                                     stats=False)
                    e1 = ctx.enqueue_marker(0)
                ctx.finish()
                ms = ctx.elapsed_ms(e0, e1)
                launches = ctx.launch_count - before
                if it > 0: best = min(best, ms)
            gb = (2.0 if mode != "end_to_end" else 1.0) * nn * img_bytes / 1e9
            print(f"{b},{mode},{best:.4f},{nn / best * 1e3:.0f},{gb / best * 1e3:.1f},{launches}", flush=True)

if __name__ == "__main__":
    main()
