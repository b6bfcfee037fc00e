#!/usr/bin/env python
"""tools/linkbench_multi.py -- what the box's host fabric gives G GPUs at once (no kernels): the ceiling of every
end-to-end number.  One host thread per GPU in ONE process, pinned buffers, 64 MB linear copies through the C ABI
(b200blur_enqueue_write / _read), CUDA events per GPU, all GPUs released together by a barrier.

    python tools/linkbench_multi.py [--gpus 8] [--mb 1024] [--reps 3] [--json out.json]

Scenarios (GB/s per direction, summed over the active GPUs; `min`/`max` = slowest/fastest GPU):
  both      every GPU copies host->device and device->host at the same time      (what the pipeline does)
  h2d, d2h  one direction only
  split     even GPUs host->device only, odd GPUs device->host only               (alternating phases across GPUs)
for G in 1, 2, 4, ... up to --gpus, with cudaHostAlloc buffers; `both` again with buffers from an anonymous mmap
advised MADV_HUGEPAGE and cudaHostRegister-ed (fewer IOMMU/TLB entries per byte), if the box allows it.
"""
import argparse
import ctypes
import json
import mmap
import os
import sys
import threading

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "heterogeneous-opencl-image-processing-engine_b200")):
    sys.path.insert(0, p)

CHUNK = 64 << 20


def run_scenario(b200blur, ctxs, bufs, nbytes, mode, reps):
    g = len(ctxs)
    barrier = threading.Barrier(g)
    best = [None] * g

    def worker(k):
        ctx, (h_in, h_out, d_in, d_out) = ctxs[k], bufs[k]
        do_in = mode in ("both", "h2d") or (mode == "split" and k % 2 == 0)
        do_out = mode in ("both", "d2h") or (mode == "split" and k % 2 == 1)
        for _ in range(reps):
            ctx.finish()
            barrier.wait()
            e0 = ctx.enqueue_marker(0)
            ctx.enqueue_wait(2, e0)
            for off in range(0, nbytes, CHUNK):
                m = min(CHUNK, nbytes - off)
                if do_in:
                    ctx.enqueue_write(0, d_in + off, h_in + off, m)
                if do_out:
                    ctx.enqueue_read(2, h_out + off, d_out + off, m)
            e2 = ctx.enqueue_marker(2)
            ctx.enqueue_wait(0, e2)
            e1 = ctx.enqueue_marker(0)
            ctx.finish()
            ms = ctx.elapsed_ms(e0, e1)
            ctx._lib.b200blur_event_release(ctx.handle, e2)
            best[k] = ms if best[k] is None else min(best[k], ms)
            barrier.wait()

    ts = [threading.Thread(target=worker, args=(k,)) for k in range(g)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    per_gpu = [nbytes / (ms * 1e-3) / 1e9 for ms in best]
    slowest = max(best)
    n_in = sum(1 for k in range(g) if mode in ("both", "h2d") or (mode == "split" and k % 2 == 0))
    n_out = sum(1 for k in range(g) if mode in ("both", "d2h") or (mode == "split" and k % 2 == 1))
    return {"gpus": g, "mode": mode, "h2d_GBps_total": n_in * nbytes / (slowest * 1e-3) / 1e9,
            "d2h_GBps_total": n_out * nbytes / (slowest * 1e-3) / 1e9,
            "per_gpu_min": round(min(per_gpu), 2), "per_gpu_max": round(max(per_gpu), 2)}


def huge_buffers(b200blur, nbytes):
    """anonymous mmap + MADV_HUGEPAGE + cudaHostRegister -> (address, mmap object) or None"""
    try:
        size = (nbytes + (2 << 20) - 1) // (2 << 20) * (2 << 20)
        m = mmap.mmap(-1, size, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
        if hasattr(mmap, "MADV_HUGEPAGE"):
            m.madvise(mmap.MADV_HUGEPAGE)
        addr = ctypes.addressof(ctypes.c_char.from_buffer(m))
        ctypes.memset(addr, 1, size)          # touch: pages (huge if THP allows) exist before they are pinned
        rc = b200blur.load().b200blur_host_register(ctypes.c_void_p(addr), ctypes.c_size_t(size))
        if rc != 0:
            return None
        return addr, m
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=0)
    ap.add_argument("--mb", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--json", default="")
    a = ap.parse_args()
    import b200blur
    n_dev = b200blur.device_count()
    g_max = min(a.gpus or n_dev, n_dev)
    nbytes = a.mb << 20
    ctxs = [b200blur.Context(k, 3) for k in range(g_max)]
    bufs = []
    for ctx in ctxs:
        bufs.append((ctx.host_alloc(nbytes), ctx.host_alloc(nbytes), ctx.dev_alloc(nbytes), ctx.dev_alloc(nbytes)))
        ctypes.memset(bufs[-1][0], 1, nbytes)
        ctypes.memset(bufs[-1][1], 0, nbytes)
    results = []
    g = 1
    sizes = []
    while g < g_max:
        sizes.append(g)
        g *= 2
    sizes.append(g_max)
    for g in sizes:
        for mode in ("both", "h2d", "d2h") + (("split",) if g > 1 else ()):
            r = run_scenario(b200blur, ctxs[:g], bufs[:g], nbytes, mode, a.reps)
            r["memory"] = "cudaHostAlloc"
            results.append(r)
            print(json.dumps(r), flush=True)
    huge = []
    for k in range(g_max):
        hi, ho = huge_buffers(b200blur, nbytes), huge_buffers(b200blur, nbytes)
        if hi is None or ho is None:
            huge = None
            break
        huge.append((hi, ho))
    if huge:
        hb = [(hi[0], ho[0], bufs[k][2], bufs[k][3]) for k, (hi, ho) in enumerate(huge)]
        thp = "unknown"
        try:
            thp = open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip()
        except Exception:
            pass
        for g in sizes:
            r = run_scenario(b200blur, ctxs[:g], hb[:g], nbytes, "both", a.reps)
            r["memory"] = f"mmap + MADV_HUGEPAGE + cudaHostRegister (THP: {thp})"
            results.append(r)
            print(json.dumps(r), flush=True)
    else:
        print(json.dumps({"memory": "hugepage-advised registered buffers unavailable on this box"}), flush=True)
    if a.json:
        with open(a.json, "w") as f:
            json.dump({"bytes_per_gpu_per_direction": nbytes, "host_cores": os.cpu_count(), "results": results}, f, indent=1)
    os._exit(0)   # registered mmaps and contexts are torn down with the process


if __name__ == "__main__":
    main()
