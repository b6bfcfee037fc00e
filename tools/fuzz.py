#!/usr/bin/env python
"""tools/fuzz.py -- time-boxed randomized differential test of every compute entry point of the C ABI against the oracle
(test infrastructure, like tests/: the oracle is the checker, never the thing measured).

    python tools/fuzz.py [--seconds 60] [--seed 1] [--json out.json]

Every case draws a geometry (channels 1..4; row lengths that hit the streamed kernel in its aligned, pitched, tight-input
and tight-in/out forms, the column-block form, the strip kernel and the generic kernel), random pixels, and one of:

  rows      b200blur_enqueue_blur on a random row range of a taller buffer (neighbour rows act as halos), at a random
            byte misalignment of input and output, tight or pitched rows
  resident  b200blur_run_resident in its three batch modes with a random batch size
  host      b200blur_run_host (pinned host buffers, the pipelined path) with a random batch size
  batches   b200blur_enqueue_blur_batches on a random list of launches (equal or mixed geometry, with and without halo rows)
  feed      b200blur_feed_*: random batch sizes through a small descriptor ring, waited for out of order

Outputs sit between guard bands (0xA5) that must come back untouched; a mismatch prints the case and exits 1.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "heterogeneous-opencl-image-processing-engine_b200")):
    sys.path.insert(0, p)

GUARD = 256
FILL = 0xA5


def draw_geometry(rng):
    c = int(rng.choice([1, 2, 3, 3, 3, 4]))
    kind = int(rng.integers(0, 6))
    if kind == 0:       # aligned streamed rows: 256..4096 bytes, multiple of 16
        rb = 16 * int(rng.integers(16, 257))
        w = max(1, rb // c)
        if (w * c) % 16:
            w = (w * c // 16 * 16) // c if c in (1, 2, 4) else 16 * max(6, w // 16)
    elif kind == 1:     # tight rows of any length >= 256 bytes
        w = int(rng.integers(256 // c + 1, 1400))
    elif kind == 2:     # column blocks: rows wider than 4 KB
        w = int(rng.integers(4096 // c + 1, 4096 // c + 1500))
    elif kind == 3:     # strips: short rows, multiple of 16 bytes
        w = 16 * int(rng.integers(1, 5))
    elif kind == 4:     # generic: anything small
        w = int(rng.integers(1, 90))
    else:               # the BASELINE shapes
        w = int(rng.choice([320, 256]))
        c = 3
    h = int(rng.integers(1, 97)) if kind != 5 else int(rng.choice([240, 256]))
    return w, h, c


class Dev:
    """device buffer with guard bands; `off` = byte misalignment of the payload"""

    def __init__(self, ctx, nbytes, off=0):
        self.ctx, self.nbytes, self.off = ctx, nbytes, off
        self.total = GUARD + off + nbytes + GUARD + 16
        self.base = ctx.dev_alloc(self.total)
        self.ptr = self.base + GUARD + off
        self.host = np.full(self.total, FILL, np.uint8)

    def upload(self, payload=None):
        if payload is not None:
            self.host[GUARD + self.off:GUARD + self.off + self.nbytes] = np.frombuffer(payload.tobytes(), np.uint8)
        self.ctx.enqueue_write(0, self.base, self.host, self.total)
        self.ctx.finish(0)

    def download(self):
        back = np.empty(self.total, np.uint8)
        self.ctx.enqueue_read(0, back, self.base, self.total)
        self.ctx.finish(0)
        lo, hi = GUARD + self.off, GUARD + self.off + self.nbytes
        guards_ok = bool((back[:lo] == FILL).all() and (back[hi:] == FILL).all())
        return back[lo:hi], guards_ok

    def free(self):
        self.ctx.dev_free(self.base)


def case_rows(ctx, rng, oracle):
    w, h, c = draw_geometry(rng)
    n = int(rng.integers(1, 6))
    x = rng.integers(0, 256, size=(n, h, w, c), dtype=np.uint8)
    want = oracle.c_blur_batch(x, integer=True)
    rb = w * c
    pitched = rng.random() < 0.3
    in_pitch = (rb + 15) // 16 * 16 + 16 * int(rng.integers(0, 3)) if pitched else rb
    out_pitch = (rb + 15) // 16 * 16 + 16 * int(rng.integers(0, 3)) if pitched else rb
    r0 = int(rng.integers(0, h))
    nr = int(rng.integers(1, h - r0 + 1))
    off_in = 0 if rng.random() < 0.5 else int(rng.integers(0, 16))
    off_out = 0 if rng.random() < 0.5 else int(rng.integers(0, 16))
    if pitched:
        off_in = off_out = 0
    xin = np.full((n, h, in_pitch), 0x3C, np.uint8)
    xin[:, :, :rb] = x.reshape(n, h, rb)
    d_in, d_out = Dev(ctx, xin.nbytes, off_in), Dev(ctx, n * nr * out_pitch, off_out)
    d_in.upload(xin)
    d_out.upload()
    l = ctx.launch_rows(d_in.ptr, d_out.ptr, w, h, c, r0, nr, n, in_pitch * h, out_pitch * nr,
                        in_pitch if pitched else 0, out_pitch if pitched else 0)
    ctx.enqueue_blur(int(rng.integers(0, 4)), l)
    ctx.finish()
    got, guards = d_out.download()
    got = got.reshape(n, nr, out_pitch)
    ok = guards and np.array_equal(got[:, :, :rb], want[:, r0:r0 + nr].reshape(n, nr, rb))
    d_in.free()
    d_out.free()
    return ok, dict(kind="rows", w=w, h=h, c=c, n=n, r0=r0, nr=nr, in_pitch=in_pitch, out_pitch=out_pitch, off_in=off_in,
                    off_out=off_out, guards=guards)


def case_resident(ctx, rng, oracle):
    w, h, c = draw_geometry(rng)
    n = int(rng.integers(1, 200))
    while n * h * w * c > 48 << 20:
        n = max(1, n // 2)
    batch = int(rng.integers(1, n + 8))
    mode = int(rng.integers(0, 3))
    x = rng.integers(0, 256, size=(n, h, w, c), dtype=np.uint8)
    want = oracle.c_blur_batch(x, integer=True)
    d_in, d_out = Dev(ctx, x.nbytes), Dev(ctx, x.nbytes)
    d_in.upload(x)
    d_out.upload()
    for _ in range(int(rng.integers(1, 3))):      # twice: cached tables / graphs are re-used
        ctx.run_resident(d_in.ptr, d_out.ptr, w, h, c, n, batch, mode)
    got, guards = d_out.download()
    ok = guards and np.array_equal(got, want.reshape(-1))
    d_in.free()
    d_out.free()
    return ok, dict(kind="resident", w=w, h=h, c=c, n=n, batch=batch, mode=mode, guards=guards)


def case_host(ctx, rng, oracle):
    w, h, c = draw_geometry(rng)
    n = int(rng.integers(1, 300))
    while n * h * w * c > 64 << 20:
        n = max(1, n // 2)
    batch = int(rng.integers(1, n + 8))
    x = rng.integers(0, 256, size=(n, h, w, c), dtype=np.uint8)
    want = oracle.c_blur_batch(x, integer=True)
    nbytes = x.nbytes
    h_in, h_out = ctx.host_alloc(nbytes + 64), ctx.host_alloc(nbytes + 64)
    import ctypes
    ctypes.memmove(h_in, x.ctypes.data, nbytes)
    ctypes.memset(h_out, FILL, nbytes + 64)
    ctx.run_host(h_in, h_out, w, h, c, n, batch)
    got = np.ctypeslib.as_array((ctypes.c_uint8 * (nbytes + 64)).from_address(h_out)).copy()
    ok = bool((got[nbytes:] == FILL).all()) and np.array_equal(got[:nbytes], want.reshape(-1))
    ctx.host_free(h_in)
    ctx.host_free(h_out)
    return ok, dict(kind="host", w=w, h=h, c=c, n=n, batch=batch)


def case_batches(ctx, rng, oracle):
    w, h, c = draw_geometry(rng)
    mixed = rng.random() < 0.2
    k = int(rng.integers(2, 9))
    bands = rng.random() < 0.4 and h >= 6
    sets, launches, checks = [], [], []
    for i in range(k):
        wi, hi, ci = (draw_geometry(rng) if mixed and i == k - 1 else (w, h, c))
        n = int(rng.integers(1, 12))
        x = rng.integers(0, 256, size=(n, hi, wi, ci), dtype=np.uint8)
        want = oracle.c_blur_batch(x, integer=True)
        d_in, d_out = Dev(ctx, x.nbytes), Dev(ctx, x.nbytes)
        d_in.upload(x)
        d_out.upload()
        P = wi * ci
        if bands and not (mixed and i == k - 1):
            rows = hi // 3       # the middle band of three: both halo rows present
            r0 = rows
            launches.append(ctx.launch_rows(d_in.ptr, d_out.ptr + r0 * P, wi, hi, ci, r0, rows, n, hi * P, hi * P))
            checks.append((d_out, want, (r0, rows), x.shape))
        else:
            launches.append(ctx.launch_rows(d_in.ptr, d_out.ptr, wi, hi, ci, 0, hi, n))
            checks.append((d_out, want, None, x.shape))
        sets.append((d_in, d_out))
    q = int(rng.integers(0, 4))
    for _ in range(int(rng.integers(1, 3))):
        ctx.enqueue_blur_batches(q, launches)
    ctx.finish()
    ok = True
    for d_out, want, band, shape in checks:
        got, guards = d_out.download()
        got = got.reshape(shape)
        if band is None:
            ok = ok and guards and np.array_equal(got, want)
        else:
            r0, rows = band
            ok = ok and guards and np.array_equal(got[:, r0:r0 + rows], want[:, r0:r0 + rows])
            ok = ok and bool((got[:, :r0] == FILL).all() and (got[:, r0 + rows:] == FILL).all())
    for d_in, d_out in sets:
        d_in.free()
        d_out.free()
    return ok, dict(kind="batches", w=w, h=h, c=c, k=k, mixed=mixed, bands=bands, queue=q)


def case_feed(ctx, rng, oracle):
    # geometries a feed accepts: tight rows, multiple of 16 bytes, >= 256 bytes
    c = int(rng.choice([1, 2, 3, 4]))
    rb = 16 * int(rng.integers(16, 200))
    while rb % c:
        rb += 16
    w, h = rb // c, int(rng.integers(1, 80))
    max_batch = int(rng.integers(1, 40))
    k = int(rng.integers(1, 20))
    cap = int(rng.choice([0, 2, 3, 8]))
    items = []
    for _ in range(k):
        n = int(rng.integers(1, max_batch + 1))
        x = rng.integers(0, 256, size=(n, h, w, c), dtype=np.uint8)
        d_in, d_out = Dev(ctx, x.nbytes), Dev(ctx, x.nbytes)
        d_in.upload(x)
        d_out.upload()
        items.append((x, d_in, d_out))
    ok = True
    with ctx.feed(w, h, c, max_batch, cap) as f:
        f.start()
        tickets = []
        for x, d_in, d_out in items:
            tickets.append(f.submit(d_in.ptr, d_out.ptr, x.shape[0]))
            if rng.random() < 0.5:
                f.flush()
        f.flush()
        for t in rng.permutation(len(tickets)):
            f.wait(tickets[int(t)])
        f.stop()
    for x, d_in, d_out in items:
        got, guards = d_out.download()
        ok = ok and guards and np.array_equal(got, oracle.c_blur_batch(x, integer=True).reshape(-1))
        d_in.free()
        d_out.free()
    return ok, dict(kind="feed", w=w, h=h, c=c, max_batch=max_batch, k=k, cap=cap)


CASES = {"rows": case_rows, "resident": case_resident, "host": case_host, "batches": case_batches, "feed": case_feed}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60.0)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--only", default="")
    ap.add_argument("--json", default="")
    a = ap.parse_args()
    import b200blur
    from oracle import oracle
    oracle.build()
    rng = np.random.default_rng(a.seed)
    ctx = b200blur.Context(0, 4)
    names = [s for s in a.only.split(",") if s] or ["rows", "rows", "rows", "resident", "host", "batches", "feed"]
    counts, t0, it = {}, time.time(), 0
    failure = None
    while time.time() - t0 < a.seconds:
        name = names[it % len(names)]
        it += 1
        try:
            ok, info = CASES[name](ctx, rng, oracle)
        except b200blur.BlurError as e:
            ok, info = False, dict(kind=name, error=str(e))
        counts[name] = counts.get(name, 0) + 1
        if not ok:
            failure = dict(info, case_index=it, seed=a.seed)
            print("MISMATCH", json.dumps(failure), flush=True)
            break
    ctx.close()
    summary = {"seed": a.seed, "seconds": round(time.time() - t0, 1), "cases": counts, "total": sum(counts.values()),
               "failure": failure}
    print(json.dumps(summary), flush=True)
    if a.json:
        with open(a.json, "w") as f:
            json.dump(summary, f, indent=1)
    sys.exit(1 if failure else 0)


if __name__ == "__main__":
    main()
