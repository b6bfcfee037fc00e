"""GPU tests of the feed (b200blur_feed_*): the reference's batch loop (heterogeneous_blur.c:418-539 -- stage a batch,
enqueue it, wait for it) as ONE resident kernel that pulls per-batch descriptors from a ring the host appends to.
Every result is compared bit-exactly with the oracle."""
import os

import numpy as np
import pytest

import b200blur
from oracle import oracle

pytestmark = pytest.mark.gpu


def synth(seed, n, h, w, c=3):
    return np.random.default_rng(seed).integers(0, 256, size=(n, h, w, c), dtype=np.uint8)


@pytest.fixture(scope="module")
def ctx():
    oracle.build()
    c = b200blur.Context(0, 4)
    yield c
    c.close()


def _upload(ctx, x):
    d_in, d_out = ctx.dev_alloc(x.nbytes), ctx.dev_alloc(x.nbytes)
    ctx.enqueue_write(0, d_in, x, x.nbytes)
    ctx.enqueue_write(0, d_out, np.zeros_like(x), x.nbytes)
    ctx.finish(0)
    return d_in, d_out


def _download(ctx, d, like):
    out = np.empty_like(like)
    ctx.enqueue_read(0, out, d, like.nbytes)
    ctx.finish(0)
    return out


@pytest.mark.parametrize("n,h,w,c,batch", [(70, 24, 320, 3, 35), (37, 40, 256, 3, 5), (9, 33, 128, 3, 1), (50, 16, 96, 3, 7),
                                           (12, 64, 64, 4, 4), (11, 9, 512, 1, 3), (6, 12, 1408, 3, 2)])
def test_feed_batches_match_oracle(ctx, n, h, w, c, batch):
    """Batches of different sizes (the last one short), descriptors published in several flushes, tickets waited for
    out of order; the output of a completed batch is read back while later batches are still running."""
    x = synth(n * 3 + w, n, h, w, c)
    want = oracle.c_blur_batch(x, integer=True)
    d_in, d_out = _upload(ctx, x)
    img = h * w * c
    with ctx.feed(w, h, c, batch, capacity=4) as feed:       # small ring: submit must recycle descriptor slots
        feed.start()
        tickets = []
        for i0 in range(0, n, batch):
            m = min(batch, n - i0)
            tickets.append((feed.submit(d_in + i0 * img, d_out + i0 * img, m), i0, m))
            if len(tickets) % 3 == 0:
                feed.flush()
        assert [t for t, _, _ in tickets] == list(range(len(tickets)))
        first = tickets[0]
        feed.wait(first[0])                                    # flushes what is still staged
        got0 = np.empty((first[2], h, w, c), np.uint8)
        ctx.enqueue_read(1, got0, d_out + first[1] * img, got0.nbytes)
        ctx.finish(1)
        assert np.array_equal(got0, want[first[1]:first[1] + first[2]])
        for t, _, _ in reversed(tickets):
            feed.wait(t)
            assert feed.completed(t)
        feed.stop()
    assert np.array_equal(_download(ctx, d_out, x), want)
    ctx.dev_free(d_in)
    ctx.dev_free(d_out)


def test_feed_restart_and_distinct_buffers(ctx):
    """Two runs of one feed (start/stop twice); every batch has its own device allocations, like the reference's
    per-batch malloc (heterogeneous_blur.c:431-432) -- allocated BEFORE the feed starts (cudaMalloc/cudaFree wait for
    the device to go idle, which a resident kernel never does); uploads and downloads run while the kernel is resident."""
    h, w, c, batch = 30, 320, 3, 8
    sizes = [batch if b % 3 else batch - 3 for b in range(10)]
    bufs = [(ctx.dev_alloc(m * h * w * c), ctx.dev_alloc(m * h * w * c)) for m in sizes]
    with ctx.feed(w, h, c, batch, capacity=16) as feed:
        for run in range(2):
            feed.start()
            work = []
            for b, m in enumerate(sizes):
                x = synth(100 * run + b, m, h, w, c)
                d_in, d_out = bufs[b]
                ctx.enqueue_write(0, d_in, x, x.nbytes)
                ctx.finish(0)                                   # the batch's input is in place before it is submitted
                work.append((feed.submit(d_in, d_out, m), x, d_out))
            feed.flush()
            for t, x, d_out in work:
                feed.wait(t)
                assert np.array_equal(_download(ctx, d_out, x), oracle.c_blur_batch(x, integer=True))
            feed.stop()
        assert feed.submitted == 20
    for d_in, d_out in bufs:
        ctx.dev_free(d_in)
        ctx.dev_free(d_out)


def test_feed_rejects_what_it_cannot_run(ctx):
    with pytest.raises(b200blur.BlurError):
        ctx.feed(250, 37, 3, 8)            # rows of 750 bytes: not a multiple of 16
    with pytest.raises(b200blur.BlurError):
        ctx.feed(16, 16, 3, 8)             # rows shorter than 256 bytes
    with ctx.feed(320, 8, 3, 4) as feed:
        d = ctx.dev_alloc(4 * 8 * 320 * 3 * 2)
        with pytest.raises(b200blur.BlurError):
            feed.submit(d, d, 4)           # in place
        with pytest.raises(b200blur.BlurError):
            feed.submit(d, d + 4 * 8 * 960, 5)   # more images than max_batch
        with pytest.raises(b200blur.BlurError):
            feed.flush()                   # not running
        t = feed.submit(d, d + 4 * 8 * 960, 4)
        with pytest.raises(b200blur.BlurError):
            feed.wait(t)                   # not running: cannot complete
        feed.start()
        feed.wait(t)
        feed.stop()
        ctx.dev_free(d)


def test_feed_watchdog_stops_an_abandoned_kernel(ctx):
    """A resident kernel whose host stops feeding it (and never stops it) gives up after the timeout instead of
    occupying the GPU for ever; the feed then reports the failure."""
    os.environ["B200BLUR_FEED_TIMEOUT_MS"] = "300"
    try:
        feed = ctx.feed(320, 8, 3, 4)
    finally:
        del os.environ["B200BLUR_FEED_TIMEOUT_MS"]
    import time
    feed.start()
    time.sleep(1.0)
    d = ctx.dev_alloc(2 * 4 * 8 * 960)
    t = feed.submit(d, d + 4 * 8 * 960, 4)
    with pytest.raises(b200blur.BlurError):
        feed.wait(t)
    with pytest.raises(b200blur.BlurError):
        feed.submit(d, d + 4 * 8 * 960, 4)
    feed.close()
    ctx.dev_free(d)
    # the context is still healthy
    x = synth(1, 2, 8, 320, 3)
    assert np.array_equal(ctx.blur_numpy(x), oracle.c_blur_batch(x))


@pytest.mark.parametrize("n,batch", [(5000, 35), (300, 1), (1000, 500), (64, 64), (65, 64)])
def test_run_resident_per_batch_descriptors(ctx, n, batch):
    """b200blur_run_resident(coalesce=0): the batches go through the feed kernel as a descriptor table (one launch per
    call), repeated calls re-use the table; == fused batches == one launch per batch (coalesce=2) == oracle sample."""
    import torch
    h, w, c = 240, 320, 3
    g = torch.Generator(device="cuda").manual_seed(n + batch)
    d_in = torch.randint(0, 256, (n, h, w, c), dtype=torch.uint8, device="cuda", generator=g)
    fused, per_batch, per_launch = torch.zeros_like(d_in), torch.zeros_like(d_in), torch.zeros_like(d_in)
    torch.cuda.synchronize()
    ctx.run_resident(d_in, fused, w, h, c, n, batch, 1)
    for _ in range(3):                                           # 2nd and 3rd call: table already on the device
        per_batch.zero_()
        torch.cuda.synchronize()
        st = ctx.run_resident(d_in, per_batch, w, h, c, n, batch, 0)
        assert st.launches == (1 if n > batch else st.launches)
        assert torch.equal(per_batch, fused)
    ctx.run_resident(d_in, per_launch, w, h, c, n, batch, 2)
    torch.cuda.synchronize()
    assert torch.equal(per_launch, fused)
    idx = sorted({0, n // 2, n - 1})
    assert np.array_equal(fused[idx].cpu().numpy(), oracle.c_blur_batch(d_in[idx].cpu().numpy(), integer=True))


def test_enqueue_blur_batches_one_launch_many_descriptors(ctx):
    """b200blur_enqueue_blur_batches: independent launches of one geometry run as ONE kernel launch (per-batch
    descriptors with their own pointers and image counts); a list of mixed geometries runs launch by launch."""
    h, w, c = 40, 320, 3
    sizes = [7, 3, 7, 1, 5]
    xs = [synth(500 + i, m, h, w, c) for i, m in enumerate(sizes)]
    bufs = [_upload(ctx, x) for x in xs]
    launches = [ctx.launch_rows(d_in, d_out, w, h, c, 0, h, x.shape[0]) for x, (d_in, d_out) in zip(xs, bufs)]
    before = ctx.launch_count
    ev = ctx.enqueue_blur_batches(0, launches, want_event=True)
    assert ctx.event_ms(ev) > 0
    assert ctx.launch_count - before == 1
    for x, (d_in, d_out) in zip(xs, bufs):
        assert np.array_equal(_download(ctx, d_out, x), oracle.c_blur_batch(x, integer=True))
    # mixed geometries: not one launch, same results
    y = synth(9, 4, 24, 256, 3)
    d_y = _upload(ctx, y)
    ctx.enqueue_write(0, bufs[0][1], np.zeros_like(xs[0]), xs[0].nbytes)
    before = ctx.launch_count
    ctx.enqueue_blur_batches(0, [launches[0], ctx.launch_rows(d_y[0], d_y[1], 256, 24, 3, 0, 24, 4)])
    ctx.finish(0)
    assert ctx.launch_count - before == 2
    assert np.array_equal(_download(ctx, bufs[0][1], xs[0]), oracle.c_blur_batch(xs[0], integer=True))
    assert np.array_equal(_download(ctx, d_y[1], y), oracle.c_blur_batch(y, integer=True))
    for d_in, d_out in bufs + [d_y]:
        ctx.dev_free(d_in)
        ctx.dev_free(d_out)


@pytest.mark.parametrize("h,w,g,n", [(64, 256, 4, 9), (96, 320, 6, 4)])
def test_enqueue_blur_batches_with_halo_pointers(ctx, h, w, g, n):
    """Approach 2 bands as batches of one launch: the interior bands (both halo rows present, equal heights) of several
    independent image sets go through the feed kernel with per-batch halo pointers; every band equals the oracle's rows."""
    c = 3
    P = w * c
    rows = h // g
    sets = [synth(40 + k, n, h, w, c) for k in range(3)]
    bufs = [_upload(ctx, x) for x in sets]
    launches, where = [], []
    for k, (x, (d_in, d_out)) in enumerate(zip(sets, bufs)):
        for band in range(1, g - 1):
            launches.append(ctx.launch_rows(d_in, d_out + band * rows * P, w, h, c, band * rows, rows, n, h * P, h * P))
            where.append((k, band))
    assert all(l.halo_top and l.halo_bottom for l in launches)
    before = ctx.launch_count
    for _ in range(2):                                   # twice: the descriptor table is rewritten between calls
        ctx.enqueue_blur_batches(1, launches)
    ctx.finish(1)
    assert ctx.launch_count - before == 2
    for k, band in where:
        got = _download(ctx, bufs[k][1], sets[k])
        want = oracle.c_blur_batch(sets[k], integer=True)
        assert np.array_equal(got[:, band * rows:(band + 1) * rows], want[:, band * rows:(band + 1) * rows])
    for d_in, d_out in bufs:
        ctx.dev_free(d_in)
        ctx.dev_free(d_out)


def test_enqueue_blur_batches_from_alternating_queues(ctx):
    """The descriptor table and its work counters belong to the context, not to a queue: calls that alternate between
    queues (nothing but the library orders them) must not overwrite the table under the previous call's kernel.  Two
    different lists of the same geometry, 12 calls round-robin over the queues, every output checked."""
    h, w, c = 48, 320, 3
    lists = []
    for k in range(2):
        xs = [synth(700 + 10 * k + i, m, h, w, c) for i, m in enumerate([40, 33, 40])]
        bufs = [_upload(ctx, x) for x in xs]
        lists.append((xs, bufs, [ctx.launch_rows(d_in, d_out, w, h, c, 0, h, x.shape[0]) for x, (d_in, d_out) in zip(xs, bufs)]))
    ctx.finish()
    for i in range(12):
        ctx.enqueue_blur_batches(i % 4, lists[i % 2][2])        # the fixture context has 4 queues
    ctx.finish()
    for xs, bufs, _ in lists:
        for x, (d_in, d_out) in zip(xs, bufs):
            assert np.array_equal(_download(ctx, d_out, x), oracle.c_blur_batch(x, integer=True))
            ctx.dev_free(d_in)
            ctx.dev_free(d_out)
