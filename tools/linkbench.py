#!/usr/bin/env python
"""tools/linkbench.py -- host-link ceiling for the end-to-end mode: pinned H2D only, D2H only, both at once, at several
chunk sizes (torch is only the copy plumbing here).  The e2e roofline in bench.py is judged against these numbers."""
import json
import sys
import torch

def run(total_mb=1152, chunk_mb=8, mode="both", iters=3):
    n = total_mb * 1024 * 1024
    c = chunk_mb * 1024 * 1024
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    best = 1e9
    for _ in range(iters):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s1.wait_event(e0); s2.wait_event(e0)
        for off in range(0, n, c):
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_a[off:off + c].copy_(h_in[off:off + c], non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_out[off:off + c].copy_(d_b[off:off + c], non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return {"mode": mode, "chunk_mb": chunk_mb, "ms": round(best, 3), "GBps_each_way": round(n / best / 1e6, 2)}

if __name__ == "__main__":
    for chunk in (1, 8, 64, 1152):
        for mode in ("h2d", "d2h", "both"):
            print(json.dumps(run(chunk_mb=chunk, mode=mode)), flush=True)
