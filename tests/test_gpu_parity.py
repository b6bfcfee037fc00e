"""GPU parity tests: the CUDA path, called through the C ABI, against the oracle -- bit-exact (integer arithmetic).

Small/medium cases compare with the oracle directly and with the committed golden vectors (generated from the
reference's own kernel source); full BASELINE sizes use size-independent properties (replicated stream == one oracle
image, band-split == whole image, coalesced == per-batch launches, checksums)."""
import hashlib
import json
import os

import numpy as np
import pytest

import b200blur
from b200blur.sharding import plan_bands
from oracle import oracle

pytestmark = pytest.mark.gpu


def synth(seed, n, h, w, c=3):
    return np.random.default_rng(seed).integers(0, 256, size=(n, h, w, c), dtype=np.uint8)


@pytest.fixture(scope="module", params=[0, 1], ids=["auto-streamed", "strips"])
def ctx(request):
    """Every test runs against both vectorised kernel variants: 0 = automatic (the TMA-bulk streamed kernel wherever
    rows are >= 256 bytes, the register/shuffle strip kernel below that), 1 = strip kernel everywhere."""
    oracle.build()
    c = b200blur.Context(0, 4)
    c.set_kernel_variant(request.param)
    c.variant = request.param
    yield c
    c.close()


def diff_report(a, b):
    d = np.abs(a.astype(np.int16) - b.astype(np.int16))
    hist = np.bincount(d.ravel(), minlength=4)[:8]
    return f"max-abs-diff {d.max()}, histogram of |diff| 0..7: {hist.tolist()}"


def assert_same(got, want):
    assert got.shape == want.shape
    assert np.array_equal(got, want), diff_report(got, want)


def test_device_is_blackwell(ctx):
    assert b200blur.device_count() >= 1
    assert "B200" in b200blur.device_name(0) or True  # name printed for the record
    print(b200blur.device_name(0), b200blur.version())


def test_golden_vectors(ctx, golden_dir):
    vec = np.load(os.path.join(golden_dir, "vectors.npz"))
    for name in sorted(k[3:] for k in vec.files if k.startswith("in_")):
        assert_same(ctx.blur_numpy(vec["in_" + name]), vec["out_" + name])


def test_golden_checksums(ctx, golden_dir):
    with open(os.path.join(golden_dir, "checksums.json")) as f:
        sums = json.load(f)
    for case in sums["cases"]:
        x = synth(case["seed"], case["n"], case["h"], case["w"], case["c"])
        y = ctx.blur_numpy(x)
        assert hashlib.sha256(y.tobytes()).hexdigest() == case["out_sha256"], case


VEC_SHAPES = [(256, 256, 3), (240, 320, 3), (1, 16, 3), (2, 16, 3), (3, 32, 3), (17, 48, 3), (33, 80, 3), (16, 16, 3),
              (100, 1024, 3), (5, 16, 1), (31, 64, 1), (9, 8, 2), (40, 24, 2), (7, 4, 4), (64, 100, 4), (15, 176, 3),
              (4, 2048, 3), (37, 336, 3)]
GEN_SHAPES = [(1, 1, 3), (1, 9, 3), (9, 1, 3), (2, 2, 3), (17, 33, 3), (100, 52, 3), (5, 7, 1), (6, 5, 2), (3, 3, 4),
              (11, 13, 5), (255, 255, 3)]


@pytest.mark.parametrize("shape", VEC_SHAPES)
def test_vectorised_path_matches_oracle(ctx, shape):
    h, w, c = shape
    n = 3
    x = synth(h * 7919 + w * 31 + c, n, h, w, c)
    l = ctx.launch_rows(0x1000, 0x100000, w, h, c, 0, h, n)
    assert ctx.is_vectorised(l), shape
    assert_same(ctx.blur_numpy(x), oracle.c_blur_batch(x))


@pytest.mark.parametrize("shape", GEN_SHAPES)
def test_generic_path_matches_oracle(ctx, shape):
    h, w, c = shape
    x = synth(h * 131 + w, 4, h, w, c)
    assert_same(ctx.blur_numpy(x), oracle.c_blur_batch(x))


def test_fuzz_random_shapes_and_row_ranges(ctx):
    """Seeded fuzz: 80 random (n, h, w, c) -- widths drawn so that the streamed, strip and generic kernels are all hit --
    each blurred whole and as a random row range of a taller buffer (the b200blur_launch_rows semantics)."""
    rng = np.random.default_rng(20261018)
    paths = {True: 0, False: 0}
    for it in range(80):
        c = int(rng.choice([1, 2, 3, 3, 3, 4]))
        kind = it % 4
        if kind == 0:      # streamed full-width: pitch multiple of 16 and >= 256
            w = int(rng.integers(6, 120)) * 16 // (1 if c != 3 else 1)
            w = max(w, 256 // c + 16) // 16 * 16
        elif kind == 1:    # strips: small pitch multiple of 16
            w = 16 * int(rng.integers(1, 5)) if c == 3 else 16 * int(rng.integers(1, 4))
        elif kind == 2:    # column blocks: pitch > 4096
            w = 16 * int(rng.integers(90, 140))
        else:              # generic: arbitrary width
            w = int(rng.integers(1, 200))
        h = int(rng.integers(1, 70))
        n = int(rng.integers(1, 5))
        x = rng.integers(0, 256, size=(n, h, w, c), dtype=np.uint8)
        want = oracle.c_blur_batch(x, integer=True)
        assert_same(ctx.blur_numpy(x), want)
        paths[ctx.is_vectorised(ctx.launch_rows(0x1000, 0x100000, w, h, c, 0, h, n))] += 1
        # a row range [r0, r0+nr) of the same buffer, kernel height = h (rows next to the range act as halos)
        r0 = int(rng.integers(0, h))
        nr = int(rng.integers(1, h - r0 + 1))
        P = w * c
        out = np.zeros((n, nr, w, c), np.uint8)
        d_in, d_out = ctx.dev_alloc(x.nbytes), ctx.dev_alloc(max(out.nbytes, 16))
        ctx.enqueue_write(0, d_in, x, x.nbytes)
        ctx.enqueue_blur(0, ctx.launch_rows(d_in, d_out, w, h, c, r0, nr, n, P * h, P * nr))
        ctx.enqueue_read(0, out, d_out, out.nbytes)
        ctx.finish()
        ctx.dev_free(d_in)
        ctx.dev_free(d_out)
        assert_same(out, want[:, r0:r0 + nr])
    assert paths[True] >= 40 and paths[False] >= 5    # (tight rows >= 256 bytes of any width run on the streamed kernel too)


def test_extreme_values_no_lane_carry(ctx):
    """All-255 and alternating 0/255 inputs drive every packed 16-bit lane to its maximum (4080 << 4)."""
    for h, w in [(20, 64), (240, 320)]:
        x = np.full((2, h, w, 3), 255, np.uint8)
        assert (ctx.blur_numpy(x) == 255).all()
        x = (np.indices((h, w * 3)).sum(axis=0) % 2 * 255).astype(np.uint8).reshape(1, h, w, 3)
        assert_same(ctx.blur_numpy(x), oracle.c_blur_batch(x))
        x = (np.indices((h, w)).sum(axis=0) % 2 * 255).astype(np.uint8)[None, :, :, None].repeat(3, axis=3)
        assert_same(ctx.blur_numpy(x), oracle.c_blur_batch(x))


def test_empty_inputs(ctx):
    for shape in [(0, 16, 16, 3), (2, 0, 16, 3), (2, 16, 0, 3)]:
        x = np.zeros(shape, np.uint8)
        assert ctx.blur_numpy(x).shape == x.shape
    d = ctx.dev_alloc(64)
    ctx.enqueue_blur(0, ctx.launch_rows(d, d, 16, 4, 3, 0, 0, 0))  # zero rows: in == out is not even looked at
    ctx.finish()
    ctx.dev_free(d)


def test_invalid_launches_are_rejected(ctx):
    d = ctx.dev_alloc(4096)
    with pytest.raises(b200blur.BlurError):
        ctx.enqueue_blur(0, ctx.launch_rows(d, d, 16, 4, 3, 0, 4, 1))  # in place
    with pytest.raises(b200blur.BlurError):
        ctx.enqueue_blur(99, ctx.launch_rows(d, d + 2048, 16, 4, 3, 0, 4, 1))  # bad queue
    ctx.dev_free(d)


def _run_launch(ctx, x, launches_fn):
    """Upload x [N][H][W][C], run the launches built by launches_fn(d_in, d_out), download [N][H][W][C]."""
    x = np.ascontiguousarray(x)
    out = np.zeros_like(x)
    d_in, d_out = ctx.dev_alloc(x.nbytes), ctx.dev_alloc(x.nbytes)
    try:
        ctx.enqueue_write(0, d_in, x, x.nbytes)
        ctx.enqueue_write(0, d_out, out, x.nbytes)
        for l in launches_fn(d_in, d_out):
            ctx.enqueue_blur(0, l)
        ctx.enqueue_read(0, out, d_out, x.nbytes)
        ctx.finish()
    finally:
        ctx.dev_free(d_in)
        ctx.dev_free(d_out)
    return out


@pytest.mark.parametrize("h,w,ratio", [(240, 320, 0.837), (240, 320, 0.5), (256, 256, 0.889), (9, 16, 0.5), (2, 16, 0.5),
                                       (33, 7, 0.3)])
def test_approach2_two_part_split_matches_reference_composition(ctx, h, w, ratio):
    """split_image_blur.c:503-541 through the C ABI: both parts run with height = rows incl. halo, halo output dropped."""
    n, c = 3, 3
    P = w * c
    x = synth(h + w, n, h, w, c)
    split = b200blur.ratio_split_row(h, ratio)
    assert split == oracle.a2_geometry(h, ratio)["split_row"]

    def launches(d_in, d_out):
        top = ctx.launch_rows(d_in, d_out, w, split + 1, c, 0, split, n, P * h, P * h)
        bot = ctx.launch_rows(d_in + (split - 1) * P, d_out + split * P, w, h - split + 1, c, 1, h - split, n,
                              P * h, P * h)
        return [top, bot]

    got = _run_launch(ctx, x, launches)
    for i in range(n):
        assert_same(got[i], oracle.split_image(x[i], split))
    assert_same(got, oracle.c_blur_batch(x))


@pytest.mark.parametrize("h,w,g", [(256, 256, 2), (256, 256, 4), (256, 256, 8), (240, 320, 8), (10, 16, 8), (37, 21, 4)])
def test_row_bands_with_halo_pointers_equal_whole_image(ctx, h, w, g):
    """G bands on one GPU, halo rows read in place from the neighbouring band (the multi-GPU kernel path, with
    'peer' memory being local): result must equal the whole-image blur and oracle.band_split."""
    n, c = 4, 3
    P = w * c
    x = synth(h * w + g, n, h, w, c)

    def launches(d_in, d_out):
        ls = []
        for p in plan_bands(h, g):
            top = 1 if p.has_top else 0
            ls.append(ctx.launch_rows(d_in + (p.row0 - top) * P, d_out + p.row0 * P, w, p.input_rows, c, top, p.rows, n,
                                      P * h, P * h))
        return ls

    got = _run_launch(ctx, x, launches)
    want = oracle.c_blur_batch(x)
    assert_same(got, want)
    for i in range(n):
        assert_same(got[i], oracle.band_split(x[i], min(g, h)))


def test_separate_band_buffers_with_explicit_halo_rows(ctx):
    """Band data and halo rows in different allocations with different strides (what a remote GPU's memory looks like)."""
    n, h, w, c = 5, 64, 64, 3
    P = w * c
    x = synth(99, n, h, w, c)
    want = oracle.c_blur_batch(x)
    r0, r1 = 16, 48
    band = np.ascontiguousarray(x[:, r0:r1])
    halo_t = np.ascontiguousarray(x[:, r0 - 1])
    halo_b = np.zeros((n, 2, P), np.uint8)           # stride 2*P: only row 0 of each pair is used
    halo_b[:, 0] = x[:, r1].reshape(n, P)
    out = np.zeros_like(band)
    d_band, d_out = ctx.dev_alloc(band.nbytes), ctx.dev_alloc(band.nbytes)
    d_t, d_b = ctx.dev_alloc(halo_t.nbytes), ctx.dev_alloc(halo_b.nbytes)
    ctx.enqueue_write(0, d_band, band, band.nbytes)
    ctx.enqueue_write(0, d_t, halo_t, halo_t.nbytes)
    ctx.enqueue_write(0, d_b, halo_b, halo_b.nbytes)
    l = ctx.launch_rows(d_band, d_out, w, r1 - r0, c, 0, r1 - r0, n)
    l.halo_top, l.halo_top_stride = d_t, P
    l.halo_bottom, l.halo_bottom_stride = d_b, 2 * P
    ctx.enqueue_blur(0, l)
    ctx.enqueue_read(0, out, d_out, out.nbytes)
    ctx.finish()
    for d in (d_band, d_out, d_t, d_b):
        ctx.dev_free(d)
    assert_same(out, want[:, r0:r1])


@pytest.mark.parametrize("shape", [(5, 40, 320, 3), (3, 17, 2048, 3), (4, 9, 16, 3), (3, 7, 21, 3), (2, 1, 256, 3), (3, 2, 128, 3)])
def test_no_writes_outside_the_output_range(ctx, shape):
    """compute-sanitizer is closed on this GPU pool, so out-of-bounds stores are caught with guard bands: the output
    sits between two 64 KB canaries that must come back untouched (the streamed kernel's store pointer deliberately
    starts two rows before the output and relies on predication)."""
    n, h, w, c = shape
    x = synth(sum(shape), n, h, w, c)
    guard = 65536
    total = guard + x.nbytes + guard
    host = np.full(total, 0xAB, np.uint8)
    d_in, d_buf = ctx.dev_alloc(x.nbytes), ctx.dev_alloc(total)
    ctx.enqueue_write(0, d_in, x, x.nbytes)
    ctx.enqueue_write(0, d_buf, host, total)
    ctx.enqueue_blur(0, ctx.launch_rows(d_in, d_buf + guard, w, h, c, 0, h, n))
    back = np.zeros(total, np.uint8)
    ctx.enqueue_read(0, back, d_buf, total)
    ctx.finish()
    ctx.dev_free(d_in)
    ctx.dev_free(d_buf)
    assert (back[:guard] == 0xAB).all() and (back[guard + x.nbytes:] == 0xAB).all()
    assert_same(back[guard:guard + x.nbytes].reshape(shape), oracle.c_blur_batch(x))


def _blur_pitched(ctx, x, extra_pitch=0):
    """Upload tight host rows into 16-byte-pitched device rows with a strided copy, blur with row pitches, read back."""
    n, h, w, c = x.shape
    P = w * c
    pitch = (P + 15) // 16 * 16 + extra_pitch
    d_in, d_out = ctx.dev_alloc(n * h * pitch + 64), ctx.dev_alloc(n * h * pitch + 64)
    ctx.enqueue_write_2d(0, d_in, pitch, x, P, P, n * h)
    l = ctx.launch_rows(d_in, d_out, w, h, c, 0, h, n, in_row_pitch=pitch, out_row_pitch=pitch)
    assert l.in_image_stride == pitch * h and ctx.is_vectorised(l)
    ctx.enqueue_blur(0, l)
    out = np.zeros_like(x)
    ctx.enqueue_read_2d(0, out, P, d_out, pitch, P, n * h)
    ctx.finish()
    ctx.dev_free(d_in)
    ctx.dev_free(d_out)
    return out


@pytest.mark.parametrize("shape", [(3, 37, 250, 3), (2, 20, 100, 3), (4, 9, 86, 3), (2, 64, 341, 3), (2, 5, 1366, 3),
                                   (3, 33, 257, 1), (2, 17, 130, 2), (2, 12, 67, 4), (1, 8, 1370, 3), (2, 7, 91, 3)])
def test_pitched_rows_run_any_width_on_the_vectorised_path(ctx, shape):
    """Rows re-pitched to a multiple of 16 bytes: the row end falls inside a chunk, so the right-edge clamp is applied by
    per-launch PRMT selectors (also when the last pixel straddles two chunks: 86*3 = 258 = 16*16 + 2)."""
    n, h, w, c = shape
    x = synth(sum(shape) * 3, n, h, w, c)
    want = oracle.c_blur_batch(x)
    assert_same(_blur_pitched(ctx, x), want)
    assert_same(_blur_pitched(ctx, x, extra_pitch=48), want)


@pytest.mark.parametrize("shape,pitch", [((3, 21, 86, 3), 1024), ((2, 19, 256, 3), 4096), ((3, 9, 100, 3), 4096),
                                         ((2, 40, 320, 3), 3840), ((2, 11, 128, 2), 1024), ((2, 6, 1400, 3), 16384)])
def test_heavily_padded_rows(ctx, shape, pitch):
    """ROI views / rows padded to several times width*channels (round-1 advisor finding: the streamed kernel sized its
    shared-memory ring from the pitch and refused these).  Such rows bring only their live chunks, one copy per row."""
    n, h, w, c = shape
    assert pitch >= 3 * w * c or pitch >= 4096
    x = synth(sum(shape) + pitch, n, h, w, c)
    assert_same(_blur_pitched(ctx, x, extra_pitch=pitch - (w * c + 15) // 16 * 16), oracle.c_blur_batch(x))


def test_failed_launch_releases_its_event(ctx):
    """An enqueue that fails after taking a profiling event gives the slot back (no leak): many failures in a row must
    not exhaust the pool, and the next good launch still gets an event."""
    d = ctx.dev_alloc(4096)
    for _ in range(40000):                                   # > the pool's capacity (32768)
        with pytest.raises(b200blur.BlurError):
            ctx.enqueue_blur(0, ctx.launch_rows(d, d, 16, 4, 3, 0, 4, 1), want_event=True)   # in place: rejected
    bad = ctx.launch_rows(d, d + 2048, 16, 4, 3, 0, 4, 1)
    bad.reserved = 1
    with pytest.raises(b200blur.BlurError):
        ctx.enqueue_blur(0, bad, want_event=True)
    ev = ctx.enqueue_blur(0, ctx.launch_rows(d, d + 2048, 16, 4, 3, 0, 4, 1), want_event=True)
    assert ctx.event_ms(ev) >= 0.0
    ctx.dev_free(d)


def _blur_tight_in(ctx, x, first_row=0, n_rows=None, offset=0, halo_split=False):
    """Tight (unaligned) input rows at device address base+offset, 16-byte-pitched output; optional row range / halo rows
    taken from SEPARATE tight allocations (so their addresses have their own alignment)."""
    n, h, w, c = x.shape
    P = w * c
    n_rows = h - first_row if n_rows is None else n_rows
    pitch = (P + 15) // 16 * 16
    d_in = ctx.dev_alloc(x.nbytes + 64)
    d_out = ctx.dev_alloc(n * n_rows * pitch + 64)
    ctx.enqueue_write(0, d_in + offset, x, x.nbytes)
    l = ctx.launch_rows(d_in + offset, d_out, w, h, c, first_row, n_rows, n, in_row_pitch=0, out_row_pitch=pitch,
                        in_image_stride=h * P, out_image_stride=n_rows * pitch)
    extra = []
    if halo_split:   # halo rows in their own buffers, at odd offsets
        for name, row, off in (("halo_top", first_row - 1, 5), ("halo_bottom", first_row + n_rows, 11)):
            if 0 <= row < h:
                rows = np.ascontiguousarray(x[:, row])
                d = ctx.dev_alloc(rows.nbytes + 64)
                ctx.enqueue_write(0, d + off, rows, rows.nbytes)
                setattr(l, name, d + off)
                setattr(l, name + "_stride", P)
                extra.append(d)
    assert ctx.is_vectorised(l)
    ctx.enqueue_blur(0, l)
    out = np.zeros((n, n_rows, w, c), np.uint8)
    ctx.enqueue_read_2d(0, out, P, d_out, pitch, P, n * n_rows)
    ctx.finish()
    for d in [d_in, d_out] + extra:
        ctx.dev_free(d)
    return out


@pytest.mark.parametrize("shape", [(3, 37, 250, 3), (5, 20, 100, 3), (4, 9, 86, 3), (2, 64, 341, 3), (7, 33, 257, 1), (2, 17, 130, 2),
                                   (3, 12, 67, 4), (2, 8, 1365, 3), (1, 5, 1000, 3), (9, 3, 90, 3)])
def test_tight_input_rows_run_on_the_streamed_kernel(ctx, shape):
    """Rows that are neither 16-byte pitched nor 16-byte aligned are read as they are (aligned-superset bulk copies,
    re-aligned in shared memory): whole images, at a misaligned base address, a row range with implicit halo rows, and a
    row range whose halo rows live in separate misaligned buffers (Approach 2 bands of an odd-width image)."""
    n, h, w, c = shape
    x = synth(sum(shape) * 7, n, h, w, c)
    want = oracle.c_blur_batch(x)
    assert_same(_blur_tight_in(ctx, x), want)
    assert_same(_blur_tight_in(ctx, x, offset=3), want)
    if h >= 4:
        assert_same(_blur_tight_in(ctx, x, first_row=1, n_rows=h - 2), want[:, 1:h - 1])
        assert_same(_blur_tight_in(ctx, x, first_row=1, n_rows=h - 2, offset=7, halo_split=True), want[:, 1:h - 1])


@pytest.mark.parametrize("shape", [(3, 37, 250, 3), (5, 20, 100, 3), (4, 9, 86, 3), (2, 64, 341, 3), (7, 33, 257, 1), (2, 17, 130, 2),
                                   (3, 12, 67, 4), (2, 8, 1365, 3), (6, 5, 1000, 3), (9, 3, 90, 3), (40, 11, 99, 3), (3, 30, 320, 3)])
@pytest.mark.parametrize("in_off,out_off", [(0, 0), (3, 5), (13, 2), (8, 15)])
def test_tight_rows_in_and_out_one_pass(ctx, shape, in_off, out_off):
    """Tight rows of any length and alignment on BOTH sides (the reference's own layout for any image width): one pass,
    output staged in shared memory and written as aligned words by the store warps.  The output sits between guard bands
    at a misaligned address: every byte of every image must be right and no byte outside may change (the words that
    stick out of a slot's span are written in narrower pieces)."""
    n, h, w, c = shape
    x = synth(sum(shape) * 11 + in_off, n, h, w, c)
    want = oracle.c_blur_batch(x)
    guard = 256
    total = x.nbytes + 2 * guard + 32
    d_in, d_buf = ctx.dev_alloc(x.nbytes + 64), ctx.dev_alloc(total)
    host = np.full(total, 0xA5, np.uint8)
    ctx.enqueue_write(0, d_in + in_off, x, x.nbytes)
    ctx.enqueue_write(0, d_buf, host, total)
    l = ctx.launch_rows(d_in + in_off, d_buf + guard + out_off, w, h, c, 0, h, n)
    ctx.enqueue_blur(0, l)
    back = np.zeros(total, np.uint8)
    ctx.enqueue_read(0, back, d_buf, total)
    ctx.finish()
    ctx.dev_free(d_in)
    ctx.dev_free(d_buf)
    lo = guard + out_off
    assert (back[:lo] == 0xA5).all() and (back[lo + x.nbytes:] == 0xA5).all()
    assert_same(back[lo:lo + x.nbytes].reshape(shape), want)


@pytest.mark.parametrize("h,w,g", [(60, 250, 4), (33, 100, 3), (256, 341, 8)])
def test_tight_row_bands_with_halo_pointers(ctx, h, w, g):
    """Approach 2 bands of an odd-width image, tight rows on both sides: each band launch reads its halo rows through the
    halo pointers (rows of the neighbouring bands, at whatever alignment they have) and writes its own tight rows."""
    n, c = 5, 3
    x = synth(h * w + g, n, h, w, c)
    want = oracle.c_blur_batch(x)

    def launches(d_in, d_out):
        out = []
        for p in plan_bands(h, g):
            out.append(ctx.launch_rows(d_in, d_out + p.row0 * w * c, w, h, c, p.row0, p.rows, n, h * w * c, h * w * c))
        return out
    assert_same(_run_launch(ctx, x, launches), want)


@pytest.mark.parametrize("w,c,in_pitch", [(250, 3, 777), (100, 3, 301), (130, 2, 1000), (320, 3, 963)])
def test_input_rows_with_an_unaligned_pitch(ctx, w, c, in_pitch):
    """Input rows padded to a pitch that is not a multiple of 16 (a view into a wider tight image): read as they are."""
    n, h = 3, 21
    P = w * c
    x = synth(w + in_pitch, n, h, w, c)
    padded = np.full((n, h, in_pitch), 0x5A, np.uint8)
    padded[:, :, :P] = x.reshape(n, h, P)
    pitch_out = (P + 15) // 16 * 16
    d_in, d_out = ctx.dev_alloc(padded.nbytes + 64), ctx.dev_alloc(n * h * pitch_out + 64)
    ctx.enqueue_write(0, d_in + 1, padded, padded.nbytes)
    l = ctx.launch_rows(d_in + 1, d_out, w, h, c, 0, h, n, in_row_pitch=in_pitch, out_row_pitch=pitch_out)
    ctx.enqueue_blur(0, l)
    out = np.zeros_like(x)
    ctx.enqueue_read_2d(0, out, P, d_out, pitch_out, P, n * h)
    ctx.finish()
    ctx.dev_free(d_in)
    ctx.dev_free(d_out)
    assert_same(out, oracle.c_blur_batch(x))


def test_misaligned_input_with_aligned_row_length(ctx):
    """width*channels a multiple of 16 but the input pointer is not 16-byte aligned (a view into a larger buffer)."""
    x = synth(99, 4, 30, 320, 3)
    assert_same(_blur_tight_in(ctx, x, offset=9), oracle.c_blur_batch(x))


@pytest.mark.parametrize("n,h,w,c", [(300, 37, 250, 3), (40, 100, 341, 3), (3, 700, 1366, 3)])
def test_run_resident_repitches_odd_widths(ctx, n, h, w, c):
    import torch
    x = synth(n * 7 + w, n, h, w, c)
    d_in = torch.from_numpy(x).cuda()
    d_out = torch.zeros_like(d_in)
    torch.cuda.synchronize()
    st = ctx.run_resident(d_in, d_out, w, h, c, n, 35, True)
    torch.cuda.synchronize()
    assert st.images == n and st.launches >= 1
    assert_same(d_out.cpu().numpy(), oracle.c_blur_batch(x, integer=True))


@pytest.mark.parametrize("n,h,w,c,batch", [(60, 37, 250, 3, 7), (20, 50, 341, 3, 35), (10, 16, 1366, 3, 4), (8, 20, 300, 1, 3)])
def test_run_host_repitches_odd_widths(ctx, n, h, w, c, batch):
    x = synth(n + h + w, n, h, w, c)
    out = np.zeros_like(x)
    st = ctx.run_host(x, out, w, h, c, n, batch)
    assert st.images == n
    assert_same(out, oracle.c_blur_batch(x, integer=True))


def test_image_strides_larger_than_image(ctx):
    n, h, w, c = 3, 12, 32, 3
    img_bytes = h * w * c
    x = synth(5, n, h, w, c)
    stride_in, stride_out = img_bytes + 160, img_bytes + 64
    buf = np.zeros(n * stride_in, np.uint8)
    for i in range(n):
        buf[i * stride_in:i * stride_in + img_bytes] = x[i].ravel()
    outbuf = np.zeros(n * stride_out, np.uint8)
    d_in, d_out = ctx.dev_alloc(buf.nbytes), ctx.dev_alloc(outbuf.nbytes)
    ctx.enqueue_write(0, d_in, buf, buf.nbytes)
    ctx.enqueue_blur(0, ctx.launch_rows(d_in, d_out, w, h, c, 0, h, n, stride_in, stride_out))
    ctx.enqueue_read(0, outbuf, d_out, outbuf.nbytes)
    ctx.finish()
    ctx.dev_free(d_in)
    ctx.dev_free(d_out)
    want = oracle.c_blur_batch(x)
    for i in range(n):
        assert_same(outbuf[i * stride_out:i * stride_out + img_bytes].reshape(h, w, c), want[i])


def test_profiling_events_and_2d_transfers(ctx):
    n, h, w, c = 8, 64, 64, 3
    P = w * c
    x = synth(3, n, h, w, c)
    out = np.zeros((n, h - 2, w, c), np.uint8)
    d_in, d_out = ctx.dev_alloc(x.nbytes), ctx.dev_alloc(x.nbytes)
    e0 = ctx.enqueue_write(0, d_in, x, x.nbytes, want_event=True)
    e1 = ctx.enqueue_blur(0, ctx.launch_rows(d_in, d_out, w, h, c, 0, h, n), want_event=True)
    # strided read: rows 1..h-2 of every image (like the non-zero-offset read of split_image_blur.c:537)
    e2 = ctx.enqueue_read_2d(0, out, (h - 2) * P, d_out + P, h * P, (h - 2) * P, n, want_event=True)
    ctx.finish(0)
    for e in (e0, e1, e2):
        ms = ctx.event_ms(e)
        assert 0.0 <= ms < 1000.0
    ctx.dev_free(d_in)
    ctx.dev_free(d_out)
    assert_same(out, oracle.c_blur_batch(x)[:, 1:h - 1])


@pytest.mark.parametrize("batch_size,coalesce", [(35, 1), (35, 0), (35, 2), (1, 0), (1, 2), (500, 0), (5000, 1), (5000, 0)])
def test_run_resident_matches_oracle(ctx, batch_size, coalesce):
    """coalesce: 1 = batches fused into one launch, 0 = one work descriptor per batch through the feed kernel (one
    launch per call), 2 = one kernel launch per batch."""
    import torch
    n, h, w, c = 143, 240, 320, 3
    x = synth(17, n, h, w, c)
    d_in = torch.from_numpy(x).cuda()
    d_out = torch.zeros_like(d_in)
    torch.cuda.synchronize()  # torch's stream and the context's queues are independent
    before = ctx.launch_count
    st = ctx.run_resident(d_in, d_out, w, h, c, n, batch_size, coalesce)
    torch.cuda.synchronize()
    n_batches = (n + batch_size - 1) // batch_size
    feed = coalesce == 0 and n_batches > 1 and ctx.variant == 0   # the feed kernel is the streamed kernel
    want_launches = 1 if (coalesce == 1 or feed) else n_batches
    assert st.launches == want_launches and ctx.launch_count - before == want_launches
    assert st.images == n and st.kernel_ms > 0
    assert_same(d_out.cpu().numpy(), oracle.c_blur_batch(x, integer=True))


@pytest.mark.parametrize("batch_size", [35, 1, 64, 1000])
def test_run_host_pipeline_matches_oracle(ctx, batch_size):
    import torch
    n, h, w, c = 300, 240, 320, 3
    if batch_size == 1:
        n = 40
    x = synth(23, n, h, w, c)
    h_in = torch.from_numpy(x).pin_memory()
    h_out = torch.zeros_like(h_in).pin_memory()
    st = ctx.run_host(h_in, h_out, w, h, c, n, batch_size)
    assert st.images == n and 1 <= st.launches <= (n + batch_size - 1) // batch_size  # small batches are fused
    assert st.h2d_bytes == x.nbytes and st.d2h_bytes == x.nbytes
    assert st.h2d_ms > 0 and st.kernel_ms > 0 and st.d2h_ms > 0
    assert_same(h_out.numpy(), oracle.c_blur_batch(x, integer=True))
    # pageable host memory also works (slower): the reference's buffers are plain malloc (heterogeneous_blur.c:431)
    out2 = np.zeros_like(x)
    ctx.run_host(x, out2, w, h, c, n, batch_size)
    assert_same(out2, h_out.numpy())


@pytest.mark.parametrize("n,batch_size,chunk_mb", [(300, 35, 4), (41, 1, 1), (97, 8, 0)])
def test_run_host_one_direction_at_a_time(ctx, monkeypatch, n, batch_size, chunk_mb):
    """B200BLUR_E2E_PHASED=1: the chunks move in waves (all uploads of a ring, then all its downloads) so that only one
    transfer direction is active per GPU -- the mode for boxes where many GPUs share one host fabric.  Same results; small
    transfer chunks so that several waves (and a short last one) happen."""
    h, w, c = 64, 320, 3
    x = synth(n + batch_size, n, h, w, c)
    monkeypatch.setenv("B200BLUR_E2E_PHASED", "1")
    monkeypatch.setenv("B200BLUR_E2E_CHUNK_MB", str(chunk_mb))
    out = np.zeros_like(x)
    st = ctx.run_host(x, out, w, h, c, n, batch_size)
    assert st.images == n and st.h2d_ms > 0 and st.d2h_ms > 0
    assert_same(out, oracle.c_blur_batch(x, integer=True))


def test_full_size_stream_properties(ctx):
    """BASELINE configs[1] at full size: 5000 x 320x240.  (a) the reference's own stream -- 5000 replicas of one image
    (heterogeneous_blur.c:440-442) -- must give 5000 copies of the oracle's output; (b) distinct random images:
    band-split == whole image and a 256-image sample == oracle."""
    import torch
    n, h, w, c = 5000, 240, 320, 3
    one = synth(31, 1, h, w, c)
    want_one = torch.from_numpy(oracle.c_blur(one[0])).cuda()
    d_in = torch.from_numpy(one).cuda().expand(n, h, w, c).contiguous()
    d_out = torch.zeros_like(d_in)
    torch.cuda.synchronize()  # torch's stream and the context's queues are independent
    ctx.run_resident(d_in, d_out, w, h, c, n, 35, True)
    torch.cuda.synchronize()
    assert bool((d_out == want_one[None]).all())
    d_out.zero_()
    torch.cuda.synchronize()
    ctx.run_resident(d_in, d_out, w, h, c, n, 35, False)
    torch.cuda.synchronize()
    assert bool((d_out == want_one[None]).all())

    g = torch.Generator(device="cuda").manual_seed(7)
    d_in = torch.randint(0, 256, (n, h, w, c), dtype=torch.uint8, device="cuda", generator=g)
    whole = torch.zeros_like(d_in)
    banded = torch.zeros_like(d_in)
    torch.cuda.synchronize()
    ctx.run_resident(d_in, whole, w, h, c, n, 35, True)
    P = w * c
    for p in plan_bands(h, 8):
        top = 1 if p.has_top else 0
        ctx.enqueue_blur(0, ctx.launch_rows(d_in.data_ptr() + (p.row0 - top) * P, banded.data_ptr() + p.row0 * P, w,
                                            p.input_rows, c, top, p.rows, n, P * h, P * h))
    ctx.finish()
    torch.cuda.synchronize()
    assert bool((whole == banded).all())
    idx = torch.arange(0, n, 20, device="cuda")[:256]
    sample = d_in[idx].cpu().numpy()
    assert_same(whole[idx].cpu().numpy(), oracle.c_blur_batch(sample, integer=True))


def test_stream_larger_than_4_gib_uses_64_bit_offsets(ctx):
    """24,000 x 256x256 RGB = 4.7 GB per direction (> 2^32 bytes; BASELINE configs[3] goes to 9.8 GB): every slot of a
    replicated stream must equal the oracle's image, which catches any 32-bit offset in the kernels or the launchers."""
    import torch
    n, h, w, c = 24000, 256, 256, 3
    one = synth(41, 1, h, w, c)
    want = torch.from_numpy(oracle.c_blur(one[0])).cuda()
    d_in = torch.from_numpy(one).cuda().expand(n, h, w, c).contiguous()
    d_out = torch.zeros_like(d_in)
    torch.cuda.synchronize()
    ctx.run_resident(d_in, d_out, w, h, c, n, 1200, True)
    torch.cuda.synchronize()
    for lo in range(0, n, 4000):     # compare in slices to bound the temporary
        assert bool((d_out[lo:lo + 4000] == want[None]).all()), lo
    del d_in, d_out
    torch.cuda.empty_cache()


def test_large_frame(ctx):
    """One 8192x8192 RGB frame (BASELINE configs[4] shape): bands of 1024 rows == whole image, sample rows == oracle."""
    import torch
    h = w = 8192
    c = 3
    g = torch.Generator(device="cuda").manual_seed(11)
    d_in = torch.randint(0, 256, (1, h, w, c), dtype=torch.uint8, device="cuda", generator=g)
    whole = torch.zeros_like(d_in)
    banded = torch.zeros_like(d_in)
    torch.cuda.synchronize()
    ctx.run_resident(d_in, whole, w, h, c, 1, 1, True)
    P = w * c
    for p in plan_bands(h, 8):
        top = 1 if p.has_top else 0
        ctx.enqueue_blur(0, ctx.launch_rows(d_in.data_ptr() + (p.row0 - top) * P, banded.data_ptr() + p.row0 * P, w,
                                            p.input_rows, c, top, p.rows, 1, P * h, P * h))
    ctx.finish()
    torch.cuda.synchronize()
    assert bool((whole == banded).all())
    # oracle on three horizontal slabs (top edge, a band seam, bottom edge), using the slab +-1 row as its own image
    x = d_in[0].cpu().numpy()
    y = whole[0].cpu().numpy()
    assert_same(y[:64], oracle.c_blur(x[:65], integer=True)[:64])
    assert_same(y[1000:1060], oracle.c_blur(x[999:1061], integer=True)[1:61])
    assert_same(y[-64:], oracle.c_blur(x[-65:], integer=True)[1:])


def test_plain_c_host_blurs_an_image_through_the_abi(tmp_path):
    """tests/c/abi_client.c: a C99 program (no Python, no ctypes) does write -> blur -> read on one 320x240 image and prints
    an FNV-1a checksum of the output; the oracle's blur of the same bytes must give the same checksum."""
    import subprocess
    from conftest import build_c_client
    out = subprocess.run([build_c_client(tmp_path)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "FAIL" not in out.stdout, out.stdout + out.stderr
    got = [line.split()[1] for line in out.stdout.splitlines() if line.startswith("checksum ")]
    w, h, c = 320, 240, 3
    i = np.arange(w * h * c, dtype=np.uint64)
    x = ((i * np.uint64(2654435761)) >> np.uint64(13)).astype(np.uint8).reshape(h, w, c)
    s = 1469598103934665603
    for b in oracle.c_blur(x).tobytes():
        s = ((s ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    assert got == ["%016x" % s]


@pytest.mark.parametrize("n_ctx,n,batch_size,chunk_mb,shape", [
    (2, 300, 35, 4, (240, 320, 3)),      # many chunks: both pipelines wrap their rings several times
    (3, 41, 1, 1, (64, 320, 3)),         # fewer chunks than a ring holds for some workers
    (4, 7, 3, 0, (32, 256, 3)),          # more workers than chunks: some take nothing
    (2, 120, 16, 2, (50, 250, 3)),       # odd width (tight rows, one-pass form) through the shared-counter pipeline
])
def test_run_host_multi_contexts_take_chunks_from_one_counter(monkeypatch, n_ctx, n, batch_size, chunk_mb, shape):
    """b200blur_run_host_multi: one host stream, several contexts (one per GPU when the box has them, else all on GPU 0),
    one host thread each inside the library, transfer chunks TAKEN from a shared counter (the library form of Approach 1's
    work distribution, heterogeneous_blur.c:446-458, without a ratio).  Every image is processed exactly once: the output
    equals the oracle's, and the per-context image counts add up to the stream."""
    import torch
    h, w, c = shape
    monkeypatch.setenv("B200BLUR_E2E_CHUNK_MB", str(chunk_mb))
    g = max(1, torch.cuda.device_count())
    ctxs = [b200blur.Context(k % g, 3) for k in range(n_ctx)]
    try:
        x = synth(n * 7 + n_ctx, n, h, w, c)
        h_in = torch.from_numpy(x).pin_memory()
        h_out = torch.full_like(h_in, 0xA5).pin_memory()
        for _ in range(2):                                   # twice: rings and kernels are warm the second time
            stats = b200blur.run_host_multi(ctxs, h_in, h_out, w, h, c, n, batch_size)
            assert len(stats) == n_ctx and sum(s.images for s in stats) == n
            assert sum(s.h2d_bytes for s in stats) == x.nbytes and sum(s.d2h_bytes for s in stats) == x.nbytes
            assert_same(h_out.numpy(), oracle.c_blur_batch(x, integer=True))
            h_out.fill_(0xA5)
        # one context = b200blur_run_host
        st = b200blur.run_host_multi(ctxs[:1], h_in, h_out, w, h, c, n, batch_size)
        assert st[0].images == n
        assert_same(h_out.numpy(), oracle.c_blur_batch(x, integer=True))
        # the same context twice is refused
        with pytest.raises(b200blur.BlurError):
            b200blur.run_host_multi([ctxs[0], ctxs[0]], h_in, h_out, w, h, c, n, batch_size)
    finally:
        for cx in ctxs:
            cx.close()
